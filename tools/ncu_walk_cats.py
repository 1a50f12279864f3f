"""Development helper: instruction / lane / stall breakdown of trace_kernel_walk by code region (source page of an ncu report)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass", shell=True, capture_output=True, text=True).stdout
def cat(f, ln):
    if f == "trace_fast.cuh":
        if ln < 112: return "misc_fast"
        if ln < 137: return "local_f"
        if ln < 152: return "load_patch"
        if ln < 320: return "fast_test"
        if ln < 355: return "walk_setup"
        if ln < 392: return "walk_begin"
        if ln < 399: return "lat_side"
        if ln < 423: return "lat_cross"
        if ln < 439: return "walk_advance"
        if ln < 468: return "step:load+decode"
        if ln < 480: return "step:shell"
        if ln < 495: return "step:lon wall"
        if ln < 516: return "step:lat wall"
        if ln < 528: return "step:overlap"
        if ln < 545: return "step:child"
        return "other_fast"
    if f == "trace.cu":
        if 842 <= ln < 870: return "k:refill"
        if 870 <= ln < 883: return "k:test glue"
        if 883 <= ln < 888: return "k:walk glue"
        if 888 <= ln < 915: return "k:finish"
        if 811 <= ln < 842: return "k:loop head"
        return "k:other"
    return f
fname = ""; func = ""; hdr = None
res = {}
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name": func = r[1]; hdr = None; continue
    if "Instructions Executed" in r: hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) != len(hdr) or r[0] == "": continue
    try:
        wi = int(r[ix["Instructions Executed"]]); ti = int(r[ix["Thread Instructions Executed"]]); sm = int(r[ix["# Samples"]])
    except ValueError: continue
    k = "shadow" if "(bool)1>" in func.split(",")[-1] else "primary"
    a = res.setdefault(k, {}).setdefault(cat(fname, int(r[0])), [0, 0, 0])
    a[0] += wi; a[1] += ti; a[2] += sm
for k, cats in res.items():
    tw = sum(v[0] for v in cats.values()); tt = sum(v[1] for v in cats.values()); ts = sum(v[2] for v in cats.values())
    print(f"\n{k}: warp-instr {tw:.3e}  thread-instr {tt:.3e}  lanes/instr {tt / tw:.2f}")
    print(f"{'region':20s} {'instr%':>7} {'lanes':>6} {'thread%':>8} {'stall%':>7}")
    for c, (wi, ti, sm) in sorted(cats.items(), key=lambda kv: -kv[1][0]):
        print(f"{c:20s} {100 * wi / tw:7.1f} {ti / max(wi, 1):6.1f} {100 * ti / tt:8.1f} {100 * sm / max(ts, 1):7.1f}")
