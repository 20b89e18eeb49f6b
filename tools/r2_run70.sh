#!/bin/bash
O=gpurun_out/r2_70; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 240 $T tools/p2p_check.py > $O/p2p_check.log 2>&1
timeout 600 $T bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 300 $T bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
timeout 120 $T tools/mgpu_timeline.py > $O/timeline_n2.txt 2>&1
