#!/bin/bash
# After `bash tools/gpu_round.sh <tag>` has come back: the tracked summaries for profiles/ from gpurun_out/<tag>_* (run here,
# no GPU): bench / reference lines, test and smoke logs, launch list + shares, ncu details / raw / regions of the walk
# kernels, raw page of the downscale kernels, SASS of the two pool kernels.
#   bash tools/round_profiles.sh r08
tag=${1:-r08}; g=gpurun_out; p=profiles
for f in bench_n1.json reference_n1.json gpu_tests.log smoke.log launches.csv; do cp $g/${tag}_$f $p/${tag}_$f; done
ncu -i $g/${tag}_walk.ncu-rep --page details > $p/${tag}_walk_kernels_details.txt 2>&1
ncu -i $g/${tag}_walk.ncu-rep --page raw --csv > $p/${tag}_walk_kernels_raw.csv 2>/dev/null
ncu -i $g/${tag}_downscale.ncu-rep --page raw --csv > $p/${tag}_downscale_raw.csv 2>/dev/null
for k in trace_kernel_pool shade_kernel shadow_kernel_pool; do python tools/ncu_regions.py $g/${tag}_walk.ncu-rep $k 40 > $p/${tag}_${k}_regions.txt 2>&1; done
cuobjdump -sass moonrtx_b200/libmoonb200.so 2>/dev/null | awk -v p=$p -v t=$tag '/Function : /{name=$3; f=""; if (name ~ /trace_kernel_poolILb1E/) f=p"/"t"_trace_kernel_pool_i16.sass"; if (name ~ /shadow_kernel_poolILb1ELb0E/) f=p"/"t"_shadow_kernel_pool_i16.sass"} f!=""{print > f}'
python - "$tag" <<'PY'
import csv, collections, sys
tag = sys.argv[1]
rows = list(csv.reader(l for l in open(f"gpurun_out/{tag}_launches.csv") if l.startswith('"')))
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
tot = {}; cnt = collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); u = r[iu]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    tot[r[ik]] = tot.get(r[ik], 0.0) + ms; cnt[r[ik]] += 1
allms = sum(tot.values())
step_k = ("trace_kernel_pool", "trace_kernel_fast", "shadow_kernel", "shade_kernel", "trace_kernel_referee", "cull_kernel", "resolve_kernel", "fold_kernel", "referee_hard")
out = ["ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-downscale",
       "(whole process: scene set-up + 24 frames; per-launch times are cold-cache and serialised: shares, not absolutes)"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    out.append(f"{v:10.3f} ms  x{cnt[k]:4d}  share {100 * v / allms:6.2f}%  {k[:110]}")
st = {k: v for k, v in tot.items() if any(s in k for s in step_k)}
sms = sum(st.values())
out += ["", "step kernels only:"]
for k, v in sorted(st.items(), key=lambda kv: -kv[1]):
    out.append(f"{v:10.3f} ms  x{cnt[k]:4d}  share of step {100 * v / sms:6.2f}%  {k[:110]}")
open(f"profiles/{tag}_launch_shares.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out[-10:-5]))
PY
ls $p | grep "^${tag}_"
