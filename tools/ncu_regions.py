"""Development helper: warp instructions, active lanes and stall samples of one kernel of an ncu report by enclosing
source function (and the hottest lines).   python tools/ncu_regions.py report.ncu-rep [kernel-substring] [nlines]"""
import csv, os, re, subprocess, sys
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else "trace_kernel_fast"; nlines = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass", shell=True, capture_output=True, text=True).stdout
fn_re = re.compile(r"^(?:template.*>\s*)?(?:MRTX_HD|__device__|__global__|static|inline|__forceinline__|__noinline__|\s)+[\w:<>\*&\s]*?\b(\w+)\s*\(")
funcs = {}
def enclosing(fname, ln):
    if fname not in funcs:
        path = None
        for d in ("moonrtx_b200/csrc", "include"):
            q = os.path.join(ROOT, d, fname)
            if os.path.exists(q): path = q
        table = []
        if path:
            for i, line in enumerate(open(path), 1):
                if line[:1] in (" ", "\t", "/", "#", "}", "\n"): continue
                m = fn_re.match(line)
                if m: table.append((i, m.group(1)))
        funcs[fname] = table
    name = "?"
    for i, n in funcs[fname]:
        if i <= ln: name = n
        else: break
    return name
fname = ""; func = ""; hdr = None; res = {}; lines = {}; src = {}
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name": func = r[1]; hdr = None; continue
    if "Instructions Executed" in r: hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) != len(hdr) or r[0] == "" or want not in func: continue
    try:
        wi = int(r[ix["Instructions Executed"]]); ti = int(r[ix["Thread Instructions Executed"]]); sm = int(r[ix["# Samples"]])
    except ValueError: continue
    key = f"{fname}:{enclosing(fname, int(r[0]))}"
    a = res.setdefault(key, [0, 0, 0]); a[0] += wi; a[1] += ti; a[2] += sm
    b = lines.setdefault((fname, int(r[0])), [0, 0, 0]); b[0] += wi; b[1] += ti; b[2] += sm
    src[(fname, int(r[0]))] = r[ix["Source"]] if "Source" in ix else ""
tw = sum(v[0] for v in res.values()); tt = sum(v[1] for v in res.values()); ts = sum(v[2] for v in res.values())
print(f"{want}: warp-instr {tw:.3e}  thread-instr {tt:.3e}  lanes/instr {tt / max(tw, 1):.2f}  samples {ts}")
print(f"{'function':44s} {'instr%':>7} {'lanes':>6} {'thread%':>8} {'stall%':>7}")
for c, (wi, ti, sm) in sorted(res.items(), key=lambda kv: -kv[1][0]):
    if wi * 1000 < tw: continue
    print(f"{c:44s} {100 * wi / tw:7.1f} {ti / max(wi, 1):6.1f} {100 * ti / tt:8.1f} {100 * sm / max(ts, 1):7.1f}")
print(f"\n{'share':>6} {'lanes':>6} {'stall%':>7}  line")
for (f, ln), (wi, ti, sm) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:nlines]:
    print(f"{wi / tw:6.3f} {ti / max(wi, 1):6.1f} {sm / max(ts, 1):7.3f}  {f}:{ln}  {src[(f, ln)].strip()[:110]}")
