import sys, time
sys.path.insert(0, '.')
import uecraytracing_b200 as u
sc, cam = u.Scene.named('final', seed=1), u.Camera.named('final')
for spp in (1, 20, 100, 500, 500, 500):
    t0 = time.perf_counter(); rgb, _, st = u.render(sc, cam, 1200, spp, 50, seed=1); t1 = time.perf_counter()
    print(f"spp {spp}: wall {1e3*(t1-t0):.1f} ms  K1 {st['kernel_ms']:.1f} resolve {st['resolve_ms']:.3f} d2h {st['d2h_ms']:.3f}")
