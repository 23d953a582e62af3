#!/usr/bin/env python
"""bench.py -- path samples/s of the path-tracing hot path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload final|ref4|dense|final4k|final4096]
  N>1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU)
       python bench.py --gpus N --single-process                                      (one process, uecrt_render(num_gpus=N))

A "step" is one whole frame of the workload: every pixel x every sample-per-pixel path.  The frame's
(pixel, sample) work items are split in N equal ranges (rank g renders range g: "a few whole samples of
every pixel and part of one more"), then ONE NCCL reduce of the float4 accumulation buffer to rank 0 and
the K3 resolve there.  Workload at N=1 is BASELINE configs[1]: RTIOW final scene (484 spheres),
1200x675, 500 spp, depth 50.  Scaling is "strong": the frame is fixed as N grows.

  value : whole-job samples/s with the scene resident in HBM (device-timed, max over ranks)
  e2e   : the same through the host-buffer C ABI call uecrt_render (scene + camera H2D, render, reduce,
          resolve, RGB D2H inside the timed region).  N=1: on rank 0's GPU.  N>1: rank 0 calls
          uecrt_render(num_gpus=N) -- the library's own single-process multi-GPU path (one host thread
          per GPU, ncclCommInitAll, one ncclReduce) -- while the other ranks idle.
  roofline : the dominant kernel K1.  Its candidate scan runs on the tensor cores (tcgen05 TF32, 3-way
          split, K = 32 slots), so bound = "tensor": achieved = executed TF32 FLOP/s (2 x 128 x rows x 32
          per 128-ray tile; tiles >= rays / 128), peak = the tcgen05 TF32 rate measured on this pool by
          tools/microbench/tc_probe.cu (profiles/tf32_peak.json).  `fp32_algorithmic` keeps SURVEY 8(d)'s
          figure (18 FLOP per ray-sphere test x spheres x rays over the K1 time) as a plain rate, without
          a fraction: the pair tests do not run on the FP32 pipe.  With --scan fp32 (or below 96 spheres)
          K1 is the FP32-pipe kernel and bound = "fp32" against 2 x 128 lanes x SMs x clock.
  exact : the same step with UECRT_FLAG_EXACT_HITS (FP64 resolve in the reference's operation order:
          every closest hit is the double-precision reference's sphere)
  cpu_baseline : the oracle (the unmodified reference headers when oracle/_ref is prebuilt, else the
          restatement) rendering a bounded crop of the same workload on all host threads
--impl reference times that CPU path alone, same metric/config, without loading libuecrt.so.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene, camera, width, spp, depth, description)
    "final": ("final", "final", 1200, 500, 50, "RTIOW final random scene (484 spheres) 1200x675, 500 spp, depth 50"),
    "ref4": ("ref4", "ref", 400, 100, 50, "reference 4-sphere scene 400x225, 100 spp, depth 50"),
    "dense": ("dense", "dense", 1920, 256, 50, "synthetic dense scene, 10k spheres, 1920x1080, 256 spp, depth 50"),
    "final4k": ("final", "final", 3840, 1024, 50, "RTIOW final scene 3840x2160, 1024 spp, depth 50"),
    "final4096": ("final", "final", 1200, 4096, 50, "RTIOW final scene 1200x675, 4096 spp, depth 50 (convergence sweep)"),
}
FLOP_PER_TEST = 18.0


def image_height(width):
    return int(width / (16.0 / 9.0))                            # uint32(width / (16/9)), source.cpp:59-62


def load_json(rel):
    try:
        with open(os.path.join(ROOT, rel)) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        busy = [v for v in sm if v > 0.5 * (mx[0] if mx else 1)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_arm(scene_name, cam_name, W, H, spp, depth, budget_s=18.0):
    """The reference's CPU path on all host threads, on a bounded crop of the same workload.  Scene and camera
    come from the committed fixture (tests/golden/bench_scenes.npz): nothing of the product is loaded."""
    from oracle import pyoracle                      # checker / baseline only
    orc = pyoracle.load("best")
    scene, cam = pyoracle.bench_scene(scene_name, cam_name)
    threads = os.cpu_count() or 1
    # calibrate on one row, then size the sample (rows x spp_sample) for ~budget_s of wall time
    y0 = H // 2
    spp_cal = min(spp, 2)
    t0 = time.perf_counter()
    orc.render_rows(scene, cam, W, H, spp_cal, depth, 1, y0, y0 + 1)
    per_sample = (time.perf_counter() - t0) / (W * spp_cal)          # seconds per path sample, 1 thread
    target = budget_s * threads / per_sample                            # samples we can afford
    spp_s = int(max(1, min(spp, target // (W * H))))
    rows = int(max(threads, min(H, target // (W * spp_s))))
    ya = max(0, (H - rows) // 2)
    from concurrent.futures import ThreadPoolExecutor
    bounds = [ya + rows * k // threads for k in range(threads + 1)]      # rows centred in the frame, split over the threads
    jobs = [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]

    def step():
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            res = list(ex.map(lambda ab: orc.render_rows(scene, cam, W, H, spp_s, depth, 1, ab[0], ab[1]), jobs))
        dt = time.perf_counter() - t0
        return dt, sum(r[1] for r in res)
    return orc.kind, threads, rows, spp_s, step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="final", choices=sorted(WORKLOADS))
    ap.add_argument("--exact-hits", action="store_true")
    ap.add_argument("--scan", default="auto", choices=["auto", "fp32", "tensor", "grid"],
                    help="K1 closest-hit search (auto: brute force on the tensor cores when the scene fits; grid: opt-in spatial index, reported separately)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-scans", action="store_true", help="skip the side-by-side FP32-scan and exact-resolve timings after the timed region")
    ap.add_argument("--single-process", action="store_true",
                    help="N GPUs from ONE process: every step is uecrt_render(num_gpus=N) (no torch.distributed)")
    ap.add_argument("--psnr", action="store_true",
                    help="final scene at 1200x675 only: PSNR of the rendered frame against the committed 1000-spp reference render")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    scene_name, cam_name, W, spp, depth, desc = WORKLOADS[args.workload]
    H = image_height(W)
    samples_per_step = W * H * spp
    n_gpus = args.gpus if args.single_process else max(world, 1)
    config = {"workload": desc, "scene": scene_name, "width": W, "height": H, "spp": spp, "max_depth": depth,
              "parallelism": f"work-item shard x{n_gpus}", "l2": "accumulation buffer re-zeroed and scene re-read every step; "
              "per-step working set (framebuffer) is rewritten, inputs (16 B/sphere) live in shared memory by design"}

    # ------------------------------------------------------------------ reference arm (CPU): no product code is loaded
    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, args.steps), max(0, args.warmup)
        # bounded sample: the whole --steps K --warmup W run is sized to ~150 s of wall time
        kind, threads, rows, spp_s, step = cpu_reference_arm(scene_name, cam_name, W, H, spp, depth,
                                                             budget_s=max(0.5, min(20.0, 150.0 / (steps + warm))))
        for _ in range(warm):
            step()
        tot = 0.0
        for _ in range(steps):
            dt, _ = step()
            tot += dt
        n = rows * W * spp_s
        v = n * steps / tot
        sample = f"{rows} centre rows x {W} px x {spp_s} spp of the frame per step ({n} path samples)"
        print(json.dumps({"metric": "path samples/sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps,
                          "warmup": warm, "ms_per_step": tot / steps * 1e3, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                          "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ our arm (B200)
    # rank 0 times the CPU baseline FIRST, before any process group exists: the other ranks then wait in the TCP
    # rendezvous (asleep) instead of spinning their GPUs in an NCCL barrier while the CPU render runs
    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        kind, threads, rows, spp_s, cstep = cpu_reference_arm(scene_name, cam_name, W, H, spp, depth)
        dt, _ = cstep()
        n = rows * W * spp_s
        cpu_baseline = {"value": n / dt, "unit": "samples/s", "cores": threads, "kind": kind,
                        "sample": f"{rows} centre rows x {W} px x {spp_s} spp of the frame ({n} path samples, {dt:.1f} s)"}

    import numpy as np
    import torch
    import uecraytracing_b200 as u
    if not torch.cuda.is_available() or u.device_count() < 1:
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback")
    scene = u.Scene.named(scene_name, seed=1, n=10000)
    cam = u.Camera.named(cam_name)
    assert u.image_height(W) == H
    scan_flag = {"auto": 0, "fp32": u.FLAG_SCAN_FP32, "tensor": u.FLAG_SCAN_TENSOR, "grid": u.FLAG_SCAN_GRID}[args.scan]
    flags = (u.FLAG_EXACT_HITS if args.exact_hits else 0) | scan_flag
    u.set_scan_mode(args.scan)                                           # uecrt_render (e2e) follows the same choice
    h2d = scene.geom.nbytes + scene.mat_type.nbytes + scene.mat_param.nbytes + 76
    d2h = W * H * 3

    if args.single_process:
        return single_process(args, u, scene, cam, W, H, spp, depth, samples_per_step, config, cpu_baseline, h2d, d2h)

    import torch.distributed as dist
    from uecraytracing_b200.distributed import item_shard, reduce_accum
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = u.Context(scene, cam)
    shard = item_shard(rank, world)
    accum = torch.zeros(H, W, 4, device=dev)
    rgb = torch.zeros(H, W, 3, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def k1(fl=flags):
        ctx.accumulate(accum.data_ptr(), W, H, spp, depth, 2026, 0, 0, stream.cuda_stream, fl, shard)

    def step():
        accum.zero_()                                                        # memset kernel (torch)
        k1()                                                                 # K1
        reduce_accum(accum, 0)                                               # NCCL reduce (world > 1)
        if rank == 0:
            u.resolve_device(accum.data_ptr(), W, H, spp, rgb.data_ptr(), stream.cuda_stream)              # K3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    st = ctx.stats()                                                         # last step's K1 (events on the launch stream)
    clocks = sampler.stop() if rank == 0 else None
    psnr = None
    if args.psnr and rank == 0:
        psnr = psnr_vs_reference(rgb.cpu().numpy(), W)

    # K1 per-launch duration: re-time each launch with its own events (same launches as in the timed region)
    def time_k1(fl, reps):
        ms, rays, tests = [], 0, 0
        for _ in range(reps):
            accum.zero_()
            k1(fl)
            s_ = ctx.stats()
            ms.append(s_["kernel_ms"]); rays += s_["rays_traced"]; tests += s_["sphere_tests"]
        return ms, rays, tests
    barrier()
    ke, rays, tests = time_k1(flags, args.steps)
    k1_avg = sum(ke) / len(ke) if ke else 0.0
    few = max(2, min(args.steps, 3))
    # side by side, same launches: the FP32-pipe scan (the design north_star names) and the FP64-exact resolve
    fp32_ms = time_k1((flags & ~u.FLAG_SCAN_TENSOR) | u.FLAG_SCAN_FP32, few)[0][1:] if st["scan_mode"] == 2 and not args.no_side_scans else []
    exact_ms = time_k1(flags | u.FLAG_EXACT_HITS, few)[0][1:] if not (args.exact_hits or args.no_side_scans) else []
    fp32_k1 = sum(fp32_ms) / len(fp32_ms) if fp32_ms else 0.0
    exact_k1 = sum(exact_ms) / len(exact_ms) if exact_ms else 0.0
    t = torch.tensor([total_ms, k1_avg, float(rays), float(tests), fp32_k1, exact_k1], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, k1_avg, fp32_k1, exact_k1 = float(tmax[0]), float(tmax[1]), float(tmax[4]), float(tmax[5])
        rays, tests = float(tsum[2]), float(tsum[3])
    ms_per_step = total_ms / args.steps
    value = samples_per_step / (ms_per_step * 1e-3)

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region, through uecrt_render
    e2e_path = "uecrt_render(num_gpus=1)"
    if world > 1:
        # the other ranks go idle (no NCCL barrier spinning their GPUs): they wait on the store for rank 0's e2e
        ctx.close(); del accum, rgb
        torch.cuda.synchronize(); torch.cuda.empty_cache()
        dist.barrier()
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            e2e_s, e2e_path = e2e_render(u, scene, cam, W, spp, depth, args.steps, world), f"uecrt_render(num_gpus={world}), single process"
            store.set("uecrt_e2e_done", "1")
        else:
            import datetime
            store.wait(["uecrt_e2e_done"], datetime.timedelta(seconds=1200))   # blocks in the TCP store: CPU and GPU idle
    else:
        e2e_s = e2e_render(u, scene, cam, W, spp, depth, args.steps, 1)

    if rank == 0:
        out = report(args, u, scene, st, world, value, ms_per_step, k1_avg, rays, tests, fp32_k1, exact_k1, samples_per_step, e2e_s, e2e_path,
                     h2d, d2h, config, clocks, cpu_baseline, W, H, launches=args.steps * 3)
        if psnr is not None:
            out["psnr_vs_reference_1000spp_db"] = psnr
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def e2e_render(u, scene, cam, W, spp, depth, steps, gpus):
    import torch
    for _ in range(2):
        u.render(scene, cam, W, spp, depth, seed=2026, gpus=gpus)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        u.render(scene, cam, W, spp, depth, seed=2026, gpus=gpus)          # uecrt_render: H2D scene, K1 (+ reduce), K3, D2H rgb
    return (time.perf_counter() - t0) / steps


def psnr_vs_reference(rgb, W):
    import numpy as np
    p = os.path.join(ROOT, "tests", "golden", "ref_final_1200x675_1000spp.png")
    if W != 1200 or not os.path.exists(p):
        return None
    from PIL import Image
    ref = np.asarray(Image.open(p).convert("RGB"), np.float64)
    mse = np.mean((rgb.astype(np.float64) - ref) ** 2)
    return float(10 * np.log10(255.0 ** 2 / mse))


def single_process(args, u, scene, cam, W, H, spp, depth, samples_per_step, config, cpu_baseline, h2d, d2h):
    """N GPUs from one process: value and e2e are both uecrt_render(num_gpus=N); K1 / reduce / K3 times from uecrt_stats."""
    import torch
    n = args.gpus
    sampler = ClockSampler(0)
    for _ in range(max(args.warmup, 3)):
        rgb, _, st = u.render(scene, cam, W, spp, depth, seed=2026, gpus=n)
    torch.cuda.synchronize()
    sampler.start()
    k1, red, res, rays, tests = [], [], [], 0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rgb, _, st = u.render(scene, cam, W, spp, depth, seed=2026, gpus=n)
        k1.append(st["kernel_ms"]); red.append(st["reduce_ms"]); res.append(st["resolve_ms"])
        rays += st["rays_traced"]; tests += st["sphere_tests"]
    e2e_s = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    k1_avg = sum(k1) / len(k1)
    device_ms = k1_avg + sum(red) / len(red) + sum(res) / len(res)          # device-side time of a step (max K1 over devices + reduce + K3)
    out = report(args, u, scene, st, n, samples_per_step / (device_ms * 1e-3), device_ms, k1_avg, rays, tests, 0.0, 0.0, samples_per_step, e2e_s,
                 f"uecrt_render(num_gpus={n}), single process", h2d, d2h, config, clocks, cpu_baseline, W, H, launches=args.steps * (2 * n + 1))
    out["single_process"] = {"k1_ms": k1_avg, "reduce_ms": sum(red) / len(red), "resolve_ms": sum(res) / len(res),
                             "note": "value = samples / (max K1 over devices + reduce + K3), from uecrt_stats CUDA events"}
    if args.psnr:
        out["psnr_vs_reference_1000spp_db"] = psnr_vs_reference(rgb, W)
    print(json.dumps(out))


def report(args, u, scene, st, world, value, ms_per_step, k1_avg, rays, tests, fp32_k1, exact_k1, samples_per_step, e2e_s, e2e_path,
           h2d, d2h, config, clocks, cpu_baseline, W, H, launches):
    peaks = load_json("MEASURED_PEAKS.json")
    traffic_db = load_json(os.path.join("profiles", "k1_traffic.json"))       # one ncu --set full capture per kernel, per launch
    tf32_db = load_json(os.path.join("profiles", "tf32_peak.json"))           # tcgen05 TF32 rate measured by tools/microbench/tc_probe.cu
    sm_max = peaks.get("sm_max_mhz", 1965.0)
    fp32_peak = 2 * 128 * st["sm_count"] * sm_max * 1e6 / 1e12
    rays_per_launch = rays / max(args.steps, 1) / max(world, 1)
    tests_per_launch = tests / max(args.steps, 1) / max(world, 1)
    alg = FLOP_PER_TEST * scene.n * rays_per_launch / (k1_avg * 1e-3) / 1e12 if k1_avg > 0 else 0.0
    tensor_scan, grid_scan = st["scan_mode"] == 2, st["scan_mode"] == 3
    swept = tensor_scan and scene.n > u.tensor_scan_resident_spheres()
    self_scan = tensor_scan and st["k1_threads"] == 896      # csrc/tc_self.cuh: 6 shade groups + the issuer warpgroup
    kernel_name = "k_render_paths_ts" if self_scan else "k_render_paths_tc" if tensor_scan else "k_render_paths_grid" if grid_scan else "k_render_paths"
    scan_desc = ("tensor cores (tcgen05 TF32 x3 split, TMEM accumulators), sphere operand swept through shared memory in 512-row slices" if swept else
                 "tensor cores (tcgen05 TF32 x3 split, TMEM accumulators), self-scan pipeline: every shade warp reads its own rays' accumulators" if self_scan else
                 "tensor cores (tcgen05 TF32 x3 split, TMEM accumulators), scan-warp pipeline" if tensor_scan else
                 "uniform grid (opt-in spatial index; NOT the brute-force 18 x N x rays figure)" if grid_scan else
                 "FP32 pipe (packed FFMA2 from shared memory)")
    traffic = None                                                            # DRAM bytes per launch exist only for captured (kernel, workload) pairs
    if args.workload == "final":
        traffic = (traffic_db.get("self_scan_r2" if self_scan else "tensor_scan_r1", traffic_db) if tensor_scan else traffic_db.get("fp32_scan", {}) if not grid_scan else {}).get("dram_bytes_per_launch")
    elif args.workload == "dense" and swept:
        traffic = traffic_db.get("swept_scan", {}).get("dram_bytes_per_launch")
    common = {"kernel": kernel_name, "scan": scan_desc, "k1_ms_per_launch": k1_avg, "rays_per_launch": rays_per_launch,
              "rays_per_sample": rays / args.steps / samples_per_step, "traffic": traffic,
              "hbm_framebuffer_gbs": (W * H * 16 * 2 + W * H * 3) / (ms_per_step * 1e-3) / 1e9}
    if tensor_scan:
        # executed tensor work: every 128-ray tile x 128-sphere chunk is four M=128, N=128, K=8 TF32 MMAs
        chunks = (scene.n + 127) // 128
        rows = 128 * (min((chunks + 3) // 4 * 4, (chunks + 2) // 3 * 3) if swept else chunks)      # kernels.cu tc_sweep_rows: slices of 4 or 3 chunks
        tiles = rays_per_launch / 128.0                     # lower bound: tiles also carry rows without a live ray
        tf32 = 2.0 * 128 * rows * 32 * tiles / (k1_avg * 1e-3) / 1e12 if k1_avg > 0 else 0.0
        tf32_peak = tf32_db.get("tf32_tflops") or (peaks.get("bf16_tflops") or 2250.0) / 2.0
        roof = {"bound": "tensor", "achieved": tf32, "peak": tf32_peak, "unit": "TFLOP/s", "frac": tf32 / tf32_peak,
                "peak_source": tf32_db.get("how", "half the dense bf16 rate of MEASURED_PEAKS.json (profiles/tf32_peak.json missing)"),
                "algorithmic": "executed TF32 FLOPs: 2 x 128 rays x padded sphere rows x 32 K-slots (10-term lifted product x 3 TF32 splits + P) per tile, tiles = rays / 128",
                "fp32_algorithmic": {"achieved": alg, "unit": "TFLOP/s", "fp32_pipe_peak": fp32_peak,
                                     "note": "SURVEY 8(d): 18 FLOP per ray-sphere test x spheres x rays traced (kernel counter) over the K1 time; "
                                             "the tests run on the tensor cores, so this is a rate, not a fraction of the FP32 pipe"}}
    else:
        ach = FLOP_PER_TEST * tests_per_launch / (k1_avg * 1e-3) / 1e12 if (grid_scan and k1_avg > 0) else alg
        roof = {"bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak if fp32_peak else None,
                "peak_source": f"2 FLOP x 128 lanes x {st['sm_count']} SMs x {sm_max:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz); "
                               "FFMA microbenchmark measured 73.3 TFLOP/s (profiles/ubench_r1.log)",
                "algorithmic": "18 FLOP per ray-sphere test x spheres x rays traced (kernel counter)"}
        if grid_scan:
            roof["algorithmic"] = "18 FLOP per EXECUTED ray-sphere test (kernel counter); the grid walk is latency-bound, not FP32-bound"
            roof["tests_per_ray"] = tests / rays if rays else None
            roof["rays_per_second"] = rays_per_launch * max(world, 1) / (k1_avg * 1e-3) if k1_avg > 0 else None
    roof.update(common)
    out = {"metric": "path samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
           "e2e": {"value": samples_per_step / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "path": e2e_path},
           "gpu_launches": launches, "roofline": roof,
           "kernel": {"regs": st["k1_regs"], "blocks": st["k1_blocks"], "threads": st["k1_threads"], "smem_bytes": st["k1_smem_bytes"], "spheres": scene.n},
           "clocks": clocks}
    if fp32_k1 > 0:
        a32 = FLOP_PER_TEST * scene.n * rays_per_launch / (fp32_k1 * 1e-3) / 1e12
        out["fp32_scan"] = {"k1_ms_per_launch": fp32_k1, "value": samples_per_step / ((ms_per_step - k1_avg + fp32_k1) * 1e-3), "unit": "samples/s",
                            "roofline": {"bound": "fp32", "kernel": "k_render_paths", "achieved": a32, "peak": fp32_peak, "unit": "TFLOP/s", "frac": a32 / fp32_peak}}
    if exact_k1 > 0:
        out["exact"] = {"k1_ms_per_launch": exact_k1, "value": samples_per_step / ((ms_per_step - k1_avg + exact_k1) * 1e-3), "unit": "samples/s",
                        "slowdown": exact_k1 / k1_avg if k1_avg else None,
                        "note": "same step with UECRT_FLAG_EXACT_HITS: FP64 resolve in the reference's operation order, every closest hit is the double-precision reference's sphere"}
    if cpu_baseline:
        out["cpu_baseline"] = cpu_baseline
    return out


if __name__ == "__main__":
    main()
