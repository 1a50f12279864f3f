"""Development helper: table of a light `ncu --metrics ... --csv` log, one line per launch (last N launches)."""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
last = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
d = OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ix['ID']], r[ix['Kernel Name']]), {})[r[ix['Metric Name']]] = r[ix['Metric Value']]
def f(v, k, scale=1.0, fmt="{:8.2f}"):
    try: return fmt.format(float(v[k].replace(',', '')) * scale)
    except Exception: return "       -"
print(f"{'id':>4} {'kernel':44s} {'ms':>8} {'Ginst':>8} {'lanes':>8} {'ipc':>8} {'occ%':>8} {'L1hit':>8} {'L2hit':>8} {'rdGB':>8} {'wrGB':>8}")
for (i, name), v in list(d.items())[-last:]:
    n = name.replace('void <unnamed>::', '').replace('(<unnamed>::RenderArgs)', '')[:44]
    print(f"{i:>4} {n:44s} " + " ".join([
        f(v, 'gpu__time_duration.sum', 1e-6), f(v, 'smsp__inst_executed.sum', 1e-9), f(v, 'smsp__thread_inst_executed_per_inst_executed.ratio'),
        f(v, 'sm__inst_executed.avg.per_cycle_active'), f(v, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        f(v, 'l1tex__t_sector_hit_rate.pct'), f(v, 'lts__t_sector_hit_rate.pct'), f(v, 'dram__bytes_read.sum', 1e-9), f(v, 'dram__bytes_write.sum', 1e-9)]))
