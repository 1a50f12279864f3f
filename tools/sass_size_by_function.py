"""Development helper: static code size of one kernel by (inlined) source function, from nvdisasm -g -gi of a cubin.

    cuobjdump -xelf all moonrtx_b200/_obj/trace.o && nvdisasm -g -gi -c trace.sm_100a.cubin > trace.dis
    python tools/sass_size_by_function.py trace.dis trace_kernel_fastILb1ELb1E
"""
import bisect, collections, os, re, sys
dis, key = sys.argv[1], sys.argv[2]
txt = open(dis).read().split('\n')
start = [i for i, l in enumerate(txt) if l.startswith('.text.') and key in l][0]
end = [i for i, l in enumerate(txt) if i > start and l.startswith('//---------------------')]
end = end[0] if end else len(txt)
CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'moonrtx_b200', 'csrc')
def fnmap(path):
    out = []
    for i, l in enumerate(open(path).read().split('\n'), 1):
        m = re.match(r'^(?:MRTX_HD|__device__|static|__global__|inline|template <[^>]*>\s*(?:MRTX_HD|__device__|static|__global__))[^(]*?\b(\w+)\s*\(', l)
        if m: out.append((i, m.group(1)))
        elif re.match(r'^(trace_kernel_fast|shadow_kernel|beam_kernel|trace_kernel_referee)\(', l): out.append((i, l.split('(')[0]))
    return out
maps = {f: fnmap(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.cu'))}
inner = None; fresh = True
agg = collections.Counter(); outer = collections.Counter(); n = 0; chain = []
for l in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: chain = []; fresh = False
        chain.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        n += 1; fresh = True
        def name(k):
            f, ln = k
            if f in maps and maps[f]:
                mm = maps[f]; idx = bisect.bisect_right([a for a, _ in mm], ln) - 1
                return f + ':' + (mm[idx][1] if idx >= 0 else '?')
            return f
        if chain:
            # innermost function of OUR sources
            ours = [c for c in chain if c[0] in maps]
            agg[name(ours[0]) if ours else chain[0][0]] += 1
            outer[f'{chain[-1][0]}:{chain[-1][1]}'] += 1
        else: agg['?'] += 1
print(f'{key}: {n} instructions = {n * 16 / 1024:.1f} KB')
for k, c in agg.most_common(40): print(f'{c:6d} {c * 16 / 1024:6.1f} KB  {k}')
print('by outermost call site (kernel source line):')
for k, c in outer.most_common(25): print(f'{c:6d} {c * 16 / 1024:6.1f} KB  {k}')
