#!/bin/bash
# multi-GPU visit: NCCL tests, default bench at N ranks (with the train sub-record), H2D probe
N=$1; O=gpurun_out/$2; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/pytest_multi.log 2>&1; tail -3 $O/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $O/bench_c2_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/h2d_probe.py > $O/h2d_${N}rank.json 2>> $O/h2d.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/h2d_probe.py --bind > $O/h2d_${N}rank_bound.json 2>> $O/h2d.err
tail -2 $O/bench_${N}gpu.err
