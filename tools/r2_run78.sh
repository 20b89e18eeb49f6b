#!/bin/bash
O=gpurun_out/r2_78; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533"
timeout 500 $T bench.py --gpus 4 > $O/bench_n4.json 2> $O/bench_n4.err
