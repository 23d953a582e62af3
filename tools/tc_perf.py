"""K1 timing of one scan mode on the final scene (1200x675): python tools/tc_perf.py [spp] [mode] [reps]
UECRT_LIB=build/variants/libuecrt_X.so selects a tuning build."""
import sys
sys.path.insert(0, '.')
import uecraytracing_b200 as u
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 100
mode = sys.argv[2] if len(sys.argv) > 2 else "tensor"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sc, cam = u.Scene.named('final', seed=1), u.Camera.named('final')
for _ in range(reps):
    rgb, acc, st = u.render(sc, cam, 1200, spp, 50, seed=3, return_accum=True, scan=mode)
    n = 1200 * 675 * spp
    tf = 18.0 * st['sphere_tests'] / (st['kernel_ms'] * 1e-3) / 1e12
    print(f"{mode:6s}: K1 {st['kernel_ms']:.2f} ms {n / st['kernel_ms'] * 1e3:.4e} samples/s {tf:.2f} alg TF ({tf / 74.45 * 100:.1f}% FP32 peak) "
          f"regs {st['k1_regs']} thr {st['k1_threads']} mean {acc[..., :3].mean() / spp:.6f} rays {st['rays_traced']}", flush=True)
