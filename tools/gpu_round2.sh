#!/bin/bash
# Round-2 evidence bundle in ONE GPU visit (1 GPU): parity tests, benches, ncu launch list + full captures of the hot
# kernels, pipeline trace, encoder-tail accuracy.  Usage (repo root, GPU box): bash tools/gpu_round2.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_c2.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"
for wl in c4 c3 c3d512 c1 train; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-extras > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "bench $wl rc=$?"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_c2.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 3 -c 1 -o $OUT/screen_c4 -f \
    python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_c4.log 2>&1; echo "ncu full c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 3 -c 1 -o $OUT/screen_c2 -f \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_c2full.log 2>&1; echo "ncu full c2 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"cm4_kernel|rows4_kernel|rowsw_kernel" -s 6 -c 6 -o $OUT/stream_c2 -f \
    python tools/ncu_stream.py c2 > $OUT/ncu_stream_c2.log 2>&1; echo "ncu stream c2 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:encoder_tail_kernel -s 2 -c 1 -o $OUT/encoder_tail -f \
    python tools/time_tail.py > $OUT/ncu_tail.log 2>&1; echo "ncu tail rc=$?"
python tools/trace_screen.py c2 > $OUT/trace_c2.txt 2>&1
python tools/tail_accuracy.py > $OUT/tail_accuracy.txt 2>&1
python tools/time_tail.py > $OUT/time_tail.txt 2>&1
python tools/tau_cost.py c2 > $OUT/tau_c2.txt 2>&1
python tools/tau_cost.py c3 > $OUT/tau_c3.txt 2>&1
python tools/diag_forward.py c2 > $OUT/diag_c2.txt 2>&1
for wl in c3d512 c3 c4; do python tools/time_stream_8x8.py $wl 2>&1 | tail -1; done > $OUT/stream_8x8.txt
for plan in 128,2,2 128,3,2; do CCVSQ_SCREEN_PLAN=$plan python tools/time_screen.py c2 2>&1 | tail -1; CCVSQ_SCREEN_PLAN=$plan python tools/ab_plan.py c2 2>&1 | tail -1; done > $OUT/plan_ab.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hardening.py tests/test_gpu_encoder_tail.py tests/test_gpu_train_graph.py -m gpu -x -q > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"
tail -3 $OUT/memcheck.log
python tools/show_bench.py $OUT/bench_c2.json $OUT/bench_c4.json $OUT/bench_c3.json $OUT/bench_c3d512.json $OUT/bench_c1.json $OUT/bench_train.json 2>&1 | tail -40
