#!/bin/bash
# 8-GPU all-reduce sweep of the 89.6 MB gradient arena: bucket counts x NCCL settings (each setting needs its own communicator)
run() { env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29517 tools/ar_bench.py 2>&1 | grep "buckets="; }
run TAG=default
run TAG=minch32 NCCL_MIN_NCHANNELS=32
run TAG=nvls NCCL_ALGO=NVLS
run TAG=tree NCCL_ALGO=Tree
run TAG=ring_simple NCCL_ALGO=Ring NCCL_PROTO=Simple
