#!/bin/bash
# Round-2 multi-GPU evidence on ONE 8-GPU box (gpurun --gpus 8): strong scaling 1/2/4/8 of BASELINE configs[1], configs[2]
# (3840x2160 x 1024 spp) and configs[4] (4096 spp, PSNR vs the committed reference image) on the default kernel, one process
# per GPU over NCCL; the single-process ABI path uecrt_render(num_gpus=N); the 2-GPU test.  Everything lands in gpurun_out/.
O=gpurun_out
run() {  # run N out-file bench-args...
  local n=$1 out=$2; shift 2
  if [ "$n" = 1 ]; then timeout 300 python bench.py --gpus 1 "$@" > $O/$out.json 2> $O/$out.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > $O/$out.json 2> $O/$out.err; fi
  tail -c 300 $O/$out.json | tr -d '\n' | cut -c1-10 > /dev/null
}
nvidia-smi -L | wc -l > $O/scale_r2_gpus.txt
timeout 300 python -m pytest tests/test_gpu_render.py -q -k "multi_gpu_single_process" > $O/multigpu_test_r2.log 2>&1; tail -2 $O/multigpu_test_r2.log
for N in 1 2 4 8; do
  run $N scale_r2_n$N --steps 10 --warmup 3 --no-cpu-baseline --no-side-scans
  [ -n "$SCALE_ONLY_FINAL" ] && continue
  run $N scale_r2_4k_n$N --workload final4k --steps 2 --warmup 3 --no-cpu-baseline --no-side-scans
  run $N scale_r2_conv4096_n$N --workload final4096 --steps 2 --warmup 3 --no-cpu-baseline --no-side-scans --psnr
done
[ -n "$SCALE_ONLY_FINAL" ] && { timeout 300 python bench.py --gpus 8 --single-process --steps 5 --warmup 3 --no-cpu-baseline > $O/scale_r2_sp_n8.json 2> $O/scale_r2_sp_n8.err; exit 0; }
for N in 2 8; do
  timeout 300 python bench.py --gpus $N --single-process --steps 5 --warmup 3 --no-cpu-baseline > $O/scale_r2_sp_n$N.json 2> $O/scale_r2_sp_n$N.err
done
timeout 300 python bench.py --gpus 8 --single-process --workload final4k --steps 2 --warmup 3 --no-cpu-baseline > $O/scale_r2_sp_4k_n8.json 2> $O/scale_r2_sp_4k_n8.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/scale_r2_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], d['n_gpus'], round(d['ms_per_step'], 3), f"{d['value']:.4e}", f"{d['e2e']['value']:.4e}", d.get('psnr_vs_reference_1000spp_db'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
