"""Development helper: per-source-line summary of an ncu report (instructions, active lanes, stall samples)."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass", shell=True, capture_output=True, text=True).stdout
hdr = None; fname = ""
items = []; tot_i = tot_t = tot_s = 0
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if "Instructions Executed" in r:
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue            # SASS rows have an empty line number
    try:
        wi = int(r[ix["Instructions Executed"]]); ti = int(r[ix["Thread Instructions Executed"]]); sm = int(r[ix["# Samples"]])
    except ValueError:
        continue
    tot_i += wi; tot_t += ti; tot_s += sm
    items.append((wi, ti, sm, fname, r[0], r[1].strip()[:100]))
print(f"warp-instr {tot_i:.3e} thread-instr {tot_t:.3e} lanes/instr {tot_t / max(tot_i, 1):.2f} samples {tot_s}")
items.sort(key=lambda t: -t[0])
print(" share  lanes  stall%  where")
for wi, ti, sm, f, ln, src in items[:top]:
    print(f"{wi / tot_i:6.3f} {ti / max(wi, 1):6.1f} {sm / max(tot_s, 1):7.3f}  {f}:{ln}  {src}")

# coarse categories (line ranges of trace_fast.cuh / trace.cu at the time of the capture)
if len(sys.argv) > 3:
    import re
    cats = {}
    for wi, ti, sm, f, ln, src in items:
        ln = int(ln)
        if f == "trace_fast.cuh":
            c = "fast_test/local_f" if ln < 330 else ("walk_begin" if ln < 373 else "walk_step")
        elif f == "trace.cu":
            c = "raygen/rng" if ln < 500 else ("shade" if ln < 596 else "kernel loop")
        else:
            c = f
        a = cats.setdefault(c, [0, 0, 0]); a[0] += wi; a[1] += ti; a[2] += sm
    print("\ncategory            share  lanes  stall%")
    for c, (wi, ti, sm) in sorted(cats.items(), key=lambda kv: -kv[1][0]):
        print(f"{c:20s} {wi / tot_i:5.3f} {ti / max(wi, 1):6.1f} {sm / max(tot_s, 1):7.3f}")
