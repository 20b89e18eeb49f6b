#!/bin/bash
O=gpurun_out/r2_72; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
STEPS=8 timeout 240 $T tools/p2p_check.py > $O/p2p_check.log 2>&1
timeout 600 $T bench.py --gpus 8 > $O/bench_n8.json 2> $O/bench_n8.err
timeout 120 $T tools/mgpu_timeline.py > $O/timeline_n8.txt 2>&1
