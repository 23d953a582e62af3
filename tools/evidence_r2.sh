#!/bin/bash
# Round-2 evidence on ONE B200 (run through gpurun): bench lines, ncu launch list + full capture of the default K1, PSNR gate,
# convergence sweep, accuracy of the FAST resolve, frame-queue overlap.  Everything lands in gpurun_out/.
set -x
O=gpurun_out
python bench.py --steps 10 --warmup 3 > $O/bench_r2_n1.json 2> $O/bench_r2_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_reference_arm.json 2> $O/bench_r2_reference_arm.err
python bench.py --workload dense --steps 3 --no-cpu-baseline > $O/bench_r2_dense_auto.json 2> $O/bench_r2_dense_auto.err
python bench.py --workload final4k --steps 3 --no-cpu-baseline > $O/bench_r2_final4k.json 2> $O/bench_r2_final4k.err
python bench.py --workload ref4 --steps 20 --no-cpu-baseline > $O/bench_r2_ref4.json 2> $O/bench_r2_ref4.err
python tools/psnr_gate.py > $O/psnr_gate_r2.json 2> $O/psnr_gate_r2.err
python tools/convergence_sweep.py > $O/convergence_r2.log 2>&1
python tools/fast_accuracy.py > $O/fast_accuracy_r2.log 2>&1
./uecraytracing_b200/bin/raytrace --stats --scene final --width 3840 --spp 64 --seed 3 --frames 8 --orbit 3 /tmp/seq.png > $O/frames_4k_r2.log 2>&1
./uecraytracing_b200/bin/raytrace --stats --scene final --width 3840 --spp 64 --seed 3 /tmp/one.png >> $O/frames_4k_r2.log 2>&1
# profiler passes LAST (numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launch_r2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_render_paths_ts -s 2 -c 1 -o $O/k1_ts_r2 -f python tools/profile_k1.py tensor > $O/ncu_k1_ts_r2.log 2>&1
tail -2 $O/ncu_k1_ts_r2.log
python -c "
import json
for f in ('bench_r2_n1','bench_r2_dense_auto','bench_r2_final4k','bench_r2_ref4'):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])
"
