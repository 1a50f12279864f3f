#!/bin/bash
# Development helper: per-launch durations (ncu, serialised) of one bench_trace configuration.  VARIANTS as for bench_trace.py.
tag=${1:-launches}; cfg=${2:-cfg3}; spp=${3:-16}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/${tag}.csv python tools/bench_trace.py $cfg $spp > gpurun_out/${tag}.log 2>&1
python - "$tag" <<'PY'
import csv, sys, collections
rows = list(csv.reader(open(f"gpurun_out/{sys.argv[1]}.csv")))
hdr = None; per = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r: hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr): continue
    k = (r[hdr["ID"]], r[hdr["Kernel Name"]][:60]); per.setdefault(k, {})[r[hdr["Metric Name"]]] = r[hdr["Metric Value"]]
items = list(per.items())[-40:]
for (i, k), m in items:
    t = float(m.get("gpu__time_duration.sum", "0").replace(",", "")) / 1e6
    print(f"{i:>5} {k:60s} {t:9.3f} ms  lanes {m.get('smsp__thread_inst_executed_per_inst_executed.ratio', '')}  winst {m.get('smsp__inst_executed.sum', '')}")
PY
