#!/bin/bash
O=gpurun_out/r2_44; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 200 $T tools/xchg_phases.py > $O/xchg_phases.txt 2>&1
