#!/usr/bin/env python
"""bench.py -- path samples/s of the path-tracing hot path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload final|ref4|dense]
  N>1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU)

A "step" is one whole frame of the workload: every pixel x every sample-per-pixel path, spp-sharded
over the ranks (rank g renders global sample ids [g*S/N, (g+1)*S/N) of every pixel), then ONE NCCL
reduce of the float4 accumulation buffer to rank 0 and the K3 resolve there.  Workload at N=1 is
BASELINE configs[1]: RTIOW final scene (484 spheres), 1200x675, 500 spp, depth 50.  Scaling is
"strong": the frame is fixed as N grows.

  value : whole-job samples/s with the scene resident in HBM (device-timed, max over ranks)
  e2e   : the same through the host-buffer C ABI call uecrt_render (scene + camera H2D, render,
          resolve, RGB + accumulation buffer D2H inside the timed region), on rank 0's GPU when N=1
          and through the per-rank path + pinned-host copies when N>1
  roofline : FP32 pipe (the roofline north_star names). algorithmic FLOPs = 18 per ray-sphere test
          (sphere.hpp:27-31) x N_spheres x rays traced (counted by the kernel), over the K1 duration from
          CUDA events.  K1's default candidate scan runs on the TENSOR cores (tcgen05 TF32 x3, TMEM), so the
          algorithmic rate can exceed the FP32-pipe peak; `roofline.tensor_pipe` gives the executed TF32
          rate against the tensor peak and `fp32_scan` the same step with the FP32-pipe scan (--scan fp32
          makes that the measured path)
  cpu_baseline : the oracle (the unmodified reference headers when oracle/_ref is prebuilt, else the
          restatement) rendering a bounded crop of the same workload on all host threads
--impl reference times that CPU path alone, same metric/config.
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
}
FP32_PEAK_TFLOPS_NOMINAL = 2 * 128 * 148 * 1.965e9 / 1e12     # 74.45: 2 FLOP x 128 lanes x 148 SMs x sm_max_mhz
FLOP_PER_TEST = 18.0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
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


def cpu_reference_arm(args, scene, cam, W, H, spp, depth, budget_s=18.0):
    """The reference's CPU path on all host threads, on a bounded crop of the same workload."""
    import numpy as np  # noqa: F401
    from oracle import pyoracle                      # checker / baseline only
    orc = pyoracle.load("best")
    threads = os.cpu_count() or 1
    # calibrate on a few rows, then size the sample (rows x spp_sample) for ~budget_s of wall time
    y0 = H // 2
    spp_cal = min(spp, 2)
    t0 = time.perf_counter()
    orc.render_rows(scene, cam, W, H, spp_cal, depth, 1, y0, y0 + 1)
    per_sample = (time.perf_counter() - t0) / (W * spp_cal)          # seconds per path sample, 1 thread
    target = budget_s * threads / per_sample                            # samples we can afford
    spp_s = int(max(1, min(spp, target // (W * H))))
    rows = int(max(threads, min(H, target // (W * spp_s))))
    ya = max(0, (H - rows) // 2)
    # rows centred in the frame, split over the threads
    from concurrent.futures import ThreadPoolExecutor
    bounds = [ya + rows * k // threads for k in range(threads + 1)]
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
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    scene_name, cam_name, W, spp, depth, desc = WORKLOADS[args.workload]

    import numpy as np
    import uecraytracing_b200 as u
    scene = u.Scene.named(scene_name, seed=1, n=10000)
    cam = u.Camera.named(cam_name)
    H = u.image_height(W)
    samples_per_step = W * H * spp
    config = {"workload": desc, "scene": scene_name, "spheres": scene.n, "width": W, "height": H, "spp": spp, "max_depth": depth,
              "parallelism": f"spp-shard x{max(world, 1)}", "l2": "accumulation buffer re-zeroed and scene re-read every step; "
              "per-step working set (framebuffer) is rewritten, inputs (16 B/sphere) live in shared memory by design"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, args.steps), max(0, args.warmup)
        # bounded sample: the whole --steps K --warmup W run is sized to ~150 s of wall time
        kind, threads, rows, spp_s, step = cpu_reference_arm(args, scene, cam, W, H, spp, depth,
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
    import torch
    if not torch.cuda.is_available() or u.device_count() < 1:
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback")
    import torch.distributed as dist
    from uecraytracing_b200.distributed import reduce_accum, sample_shard
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = u.Context(scene, cam)
    lo, cnt = sample_shard(spp, rank, world)
    accum = torch.zeros(H, W, 4, device=dev)
    rgb = torch.zeros(H, W, 3, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    scan_flag = {"auto": 0, "fp32": u.FLAG_SCAN_FP32, "tensor": u.FLAG_SCAN_TENSOR, "grid": u.FLAG_SCAN_GRID}[args.scan]
    flags = (u.FLAG_EXACT_HITS if args.exact_hits else 0) | scan_flag
    u.set_scan_mode(args.scan)                                           # uecrt_render (e2e) follows the same choice
    pin_rgb = torch.empty(H, W, 3, dtype=torch.uint8).pin_memory()
    launches_per_step = 0

    def step(count_launch=False):
        accum.zero_()                                                        # memset kernel (torch)
        if cnt:
            ctx.accumulate(accum.data_ptr(), W, H, spp, depth, 2026, lo, cnt, stream.cuda_stream, flags)   # K1
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
    k1_ms, rays, tests = [], 0, 0
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
    # K1 per-launch duration over the timed region: re-time each launch with its own events
    ke = []
    barrier()
    for _ in range(args.steps):
        accum.zero_()
        if cnt:
            ctx.accumulate(accum.data_ptr(), W, H, spp, depth, 2026, lo, cnt, stream.cuda_stream, flags)
            s_ = ctx.stats()
            ke.append(s_["kernel_ms"]); rays += s_["rays_traced"]; tests += s_["sphere_tests"]
    torch.cuda.synchronize()
    k1_avg = sum(ke) / len(ke) if ke else 0.0
    # the same K1 launches with the FP32-pipe scan (the design north_star names), for the side-by-side figure
    alt_ms = []
    if cnt and st["scan_mode"] == 2:
        for _ in range(max(2, min(args.steps, 3))):
            accum.zero_()
            ctx.accumulate(accum.data_ptr(), W, H, spp, depth, 2026, lo, cnt, stream.cuda_stream,
                           (flags & ~u.FLAG_SCAN_TENSOR) | u.FLAG_SCAN_FP32)
            alt_ms.append(ctx.stats()["kernel_ms"])
        alt_ms = alt_ms[1:]
    fp32_k1 = sum(alt_ms) / len(alt_ms) if alt_ms else 0.0
    t = torch.tensor([total_ms, k1_avg, float(rays), float(tests), fp32_k1], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, k1_avg, fp32_k1 = float(tmax[0]), float(tmax[1]), float(tmax[4])
        rays, tests = float(tsum[2]), float(tsum[3])
    ms_per_step = total_ms / args.steps
    value = samples_per_step / (ms_per_step * 1e-3)

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region
    h2d = scene.geom.nbytes + scene.mat_type.nbytes + scene.mat_param.nbytes + 76
    d2h = W * H * 3
    if world == 1:
        for _ in range(2):
            u.render(scene, cam, W, spp, depth, seed=2026)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            u.render(scene, cam, W, spp, depth, seed=2026)                  # uecrt_render: H2D scene, K1, K3, D2H rgb
        e2e_s = (time.perf_counter() - t0) / args.steps
    else:
        def e2e_step():
            ctx.set_scene(scene, cam)                                        # H2D of scene + camera (every rank)
            accum.zero_()
            if cnt:
                ctx.accumulate(accum.data_ptr(), W, H, spp, depth, 2026, lo, cnt, stream.cuda_stream, flags)
            reduce_accum(accum, 0)
            if rank == 0:
                u.resolve_device(accum.data_ptr(), W, H, spp, rgb.data_ptr(), stream.cuda_stream)
                pin_rgb.copy_(rgb, non_blocking=True)                        # D2H of the frame to pinned host memory
            torch.cuda.synchronize()
        e2e_step(); barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_t = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_s = float(e2e_t[0])

    if rank == 0:
        peaks = measured_peaks()
        try:
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
                traffic_db = json.load(f)                                # one ncu --set full capture per kernel, per launch
        except Exception:
            traffic_db = {}
        traffic = None
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        peak = 2 * 128 * st["sm_count"] * sm_max * 1e6 / 1e12
        rays_per_launch = rays / max(args.steps, 1) / max(world, 1)
        achieved = FLOP_PER_TEST * scene.n * rays_per_launch / (k1_avg * 1e-3) / 1e12 if k1_avg > 0 else 0.0
        tensor_scan, grid_scan = st["scan_mode"] == 2, st["scan_mode"] == 3
        swept = tensor_scan and scene.n > u.tensor_scan_resident_spheres()
        config["scan"] = ("tensor cores (tcgen05 TF32 x3 split, TMEM accumulators), sphere operand swept through shared memory in 512-row slices" if swept else
                          "tensor cores (tcgen05 TF32 x3 split, TMEM accumulators)" if tensor_scan else
                          "uniform grid (opt-in spatial index; NOT the brute-force 18 x N x rays figure)" if grid_scan else
                          "FP32 pipe (packed FFMA2 from shared memory)")
        kernel_name = "k_render_paths_tc" if tensor_scan else "k_render_paths_grid" if grid_scan else "k_render_paths"
        # DRAM bytes per launch exist only for the (kernel, workload) pairs that were captured; null otherwise
        if args.workload == "final":
            traffic = (traffic_db if tensor_scan else traffic_db.get("fp32_scan", {}) if not grid_scan else {}).get("dram_bytes_per_launch")
        elif args.workload == "dense" and swept:
            traffic = traffic_db.get("swept_scan", {}).get("dram_bytes_per_launch")
        if grid_scan:       # executed tests, not N per ray
            achieved = FLOP_PER_TEST * (tests / max(args.steps, 1) / max(world, 1)) / (k1_avg * 1e-3) / 1e12 if k1_avg > 0 else 0.0
        out = {"metric": "path samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
               "e2e": {"value": samples_per_step / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": args.steps * (2 if rank == 0 else 1),
               "roofline": {"bound": "fp32", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak if peak else None, "traffic": traffic,
                            "peak_source": f"2 FLOP x 128 lanes x {st['sm_count']} SMs x {sm_max:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz); "
                                           "FFMA microbenchmark measured 73.3 TFLOP/s (profiles/ubench_r1.log)",
                            "algorithmic": "18 FLOP per ray-sphere test x spheres x rays traced (kernel counter)",
                            "k1_ms_per_launch": k1_avg, "rays_per_launch": rays_per_launch, "rays_per_sample": rays / args.steps / samples_per_step,
                            "hbm_framebuffer_gbs": (W * H * 16 * 2 + W * H * 3) / (ms_per_step * 1e-3) / 1e9},
               "kernel": {"regs": st["k1_regs"], "blocks": st["k1_blocks"], "threads": st["k1_threads"], "smem_bytes": st["k1_smem_bytes"]},
               "clocks": clocks}
        if grid_scan:
            out["roofline"]["algorithmic"] = "18 FLOP per EXECUTED ray-sphere test (kernel counter); the grid walk is latency-bound, not FP32-bound"
            out["roofline"]["tests_per_ray"] = tests / rays if rays else None
            out["roofline"]["rays_per_second"] = rays_per_launch * max(world, 1) / (k1_avg * 1e-3) if k1_avg > 0 else None
        if tensor_scan:
            # executed tensor work: every 128-ray tile x 128-sphere chunk is four M=128, N=128, K=8 TF32 MMAs
            chunks = (scene.n + 127) // 128
            rows = 128 * (min((chunks + 3) // 4 * 4, (chunks + 2) // 3 * 3) if swept else chunks)      # kernels.cu tc_sweep_rows: slices of 4 or 3 chunks
            tiles = rays_per_launch / 128.0                     # lower bound: tiles also carry dead rows
            tf32 = 2.0 * 128 * rows * 32 * tiles / (k1_avg * 1e-3) / 1e12 if k1_avg > 0 else 0.0
            bf16_peak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 2250.0
            out["roofline"]["tensor_pipe"] = {"bound": "tensor", "achieved": tf32, "peak": bf16_peak / 2.0, "unit": "TFLOP/s",
                                              "frac": tf32 / (bf16_peak / 2.0),
                                              "note": "executed TF32 FLOPs (K = 32 slots: 10-term lifted product x 3 TF32 splits + P); "
                                                      "peak = half the measured dense bf16 rate of MEASURED_PEAKS.json"}
            out["roofline"]["note"] = ("the pair tests run on the tensor cores, so the algorithmic 18-FLOP-per-test rate is not capped by "
                                       "the FP32 pipe; fp32_scan is the same step with the FP32-pipe scan")
            if fp32_k1 > 0:
                a32 = FLOP_PER_TEST * scene.n * rays_per_launch / (fp32_k1 * 1e-3) / 1e12
                out["fp32_scan"] = {"k1_ms_per_launch": fp32_k1, "value": samples_per_step / ((ms_per_step - k1_avg + fp32_k1) * 1e-3),
                                    "unit": "samples/s", "roofline": {"bound": "fp32", "kernel": "k_render_paths", "achieved": a32,
                                                                      "peak": peak, "unit": "TFLOP/s", "frac": a32 / peak}}
        if not args.no_cpu_baseline:
            kind, threads, rows, spp_s, cstep = cpu_reference_arm(args, scene, cam, W, H, spp, depth)
            dt, _ = cstep()
            n = rows * W * spp_s
            out["cpu_baseline"] = {"value": n / dt, "unit": "samples/s", "cores": threads, "kind": kind,
                                   "sample": f"{rows} centre rows x {W} px x {spp_s} spp of the frame ({n} path samples, {dt:.1f} s)"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
