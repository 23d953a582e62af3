import sys, time
sys.path.insert(0, '.')
import torch
x = torch.zeros(675, 1200, 4, device='cuda'); torch.cuda.synchronize()
import uecraytracing_b200 as u
sc, cam = u.Scene.named('final', seed=1), u.Camera.named('final')
def run(tag, n=3, spp=500):
    for _ in range(n):
        t0 = time.perf_counter(); rgb, _, st = u.render(sc, cam, 1200, spp, 50, seed=1); t1 = time.perf_counter()
        print(f"{tag} spp {spp}: wall {1e3*(t1-t0):.1f} ms  K1 {st['kernel_ms']:.1f} resolve {st['resolve_ms']:.3f} d2h {st['d2h_ms']:.3f}", flush=True)
run('torch-imported')
ctx = u.Context(sc, cam)
run('with-live-ctx')
ctx.accumulate(x.data_ptr(), 1200, 675, 500, 50, 1, 0, 500, torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
print(ctx.stats()['kernel_ms'])
run('after-accumulate')
import subprocess
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.DEVNULL)
time.sleep(0.5)
run('with-nvidia-smi-polling')
p.terminate(); p.wait()
run('after-nvidia-smi')
