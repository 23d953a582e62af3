"""Join an ncu report's per-SASS-instruction metrics with nvdisasm line info of the built library.
usage: python tools/ncu_lines.py gpurun_out/k1.ncu-rep [kernel-mangled-substring] [min_pct]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else 'k_render_pathsILb0'; minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
lib = os.environ.get('UECRT_LIB', 'uecraytracing_b200/libuecrt.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if 'kernels' in f][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
start = next(i for i, l in enumerate(dis) if ('.text._ZN5uecrt' + kern) in l and l.startswith('//---'))
cur = None; seq = []
for l in dis[start + 1:]:
    if l.startswith('//---'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)))
    m2 = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m2: seq.append((cur, m2.group(2).strip()))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split('\n')))
h = rows[1]; data = [r for r in rows[2:] if len(r) > 5]
ii, si, ti = h.index('Instructions Executed'), h.index('# Samples'), h.index('Thread Instructions Executed')
assert len(data) == len(seq), (len(data), len(seq))
tot_i = sum(int(r[ii]) for r in data); tot_s = sum(int(r[si]) for r in data)
by = collections.defaultdict(lambda: [0, 0, 0])
for (loc, txt), r in zip(seq, data):
    by[loc][0] += int(r[ii]); by[loc][1] += int(r[si]); by[loc][2] += int(r[ti])
src = {}
for f in ('kernels.cu', 'closest_hit.cuh', 'materials.cuh', 'philox.cuh', 'tc_scan.cuh', 'tc_collect.cuh', 'grid_walk.cuh'):
    src[f] = open('uecraytracing_b200/csrc/' + f).read().split('\n')
agg = collections.defaultdict(lambda: [0, 0, 0])
for loc, v in by.items():
    if loc: 
        for k in range(3): agg[loc[0]][k] += v[k]
print(f'total warp-inst {tot_i:.4e}')
for f, (ni, ns, nt) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f'{f:28s} {ni / tot_i * 100:6.2f}% inst {ns / tot_s * 100:6.2f}% samples  avg lanes {nt / max(ni, 1):5.1f}')
print('--- lines')
for (loc), (ni, ns, nt) in sorted(by.items(), key=lambda x: -x[1][0]):
    if loc and (ni / tot_i * 100 > minp or ns / tot_s * 100 > minp):
        f, l = loc
        text = src[f][l - 1].strip()[:88] if f in src and l - 1 < len(src[f]) else ''
        print(f'{f:16s}:{l:4d} {ni / tot_i * 100:6.2f}%i {ns / tot_s * 100:6.2f}%s lanes {nt / max(ni, 1):5.1f}  {text}')
