import sys
sys.path.insert(0, '.')
import uecraytracing_b200 as u
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sc, cam = u.Scene.named('dense', seed=1, n=n), u.Camera.named('dense')
for _ in range(2):
    rgb, acc, st = u.render(sc, cam, 1920, spp, 50, seed=3, return_accum=True)
    ns = 1920 * 1080 * spp
    tf = 18.0 * st['sphere_tests'] / (st['kernel_ms'] * 1e-3) / 1e12
    print(f"dense n={sc.n} 1920x1080x{spp}: K1 {st['kernel_ms']:.1f} ms {ns / st['kernel_ms'] * 1e3:.4e} samples/s {tf:.2f} TF = {tf / 74.45 * 100:.1f}% peak; "
          f"blocks {st['k1_blocks']} smem {st['k1_smem_bytes']} rays/sample {st['rays_traced'] / ns:.3f} mean {acc[..., :3].mean() / spp:.4f}")
u.write_png('gpurun_out/dense.png', rgb)
