"""K1 time of the swept tensor scan on the dense scene: python tools/tc_sweep_perf.py [spp] [n]   (UECRT_LIB selects the build)"""
import sys
sys.path.insert(0, '.')
import uecraytracing_b200 as u
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
sc, cam = u.Scene.named('dense', seed=1, n=n), u.Camera.named('dense')
ts = []
for _ in range(4):
    _, _, st = u.render(sc, cam, 1920, spp, 50, seed=3, scan="tensor")
    ts.append(st['kernel_ms'])
print(f"n={n} spp={spp} K1 ms {min(ts[1:]):.2f} (all {['%.2f' % t for t in ts]}) rays {st['rays_traced']} threads {st['k1_threads']} smem {st['k1_smem_bytes']}")
