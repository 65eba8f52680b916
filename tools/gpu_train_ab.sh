#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
run() { env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/diag_train_multi.py 2>&1 | grep "overlapped\|serial \|no all-reduce\|phases" | sed "s/^/[$*] /"; }
run A=0
run CCVSQ_EMA_COPY_COUNTS=1
run CCVSQ_EMA_PERSISTENT_BUF=1
run CCVSQ_EMA_COPY_COUNTS=1 CCVSQ_EMA_PERSISTENT_BUF=1
run NCCL_NVLS_ENABLE=0
run NCCL_ALGO=Ring
