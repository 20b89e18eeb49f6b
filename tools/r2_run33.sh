#!/bin/bash
O=gpurun_out/r2_33; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $O/tests.log
timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_timeline.py > $O/timeline_n1.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_timeline.py --batch 2048 > $O/timeline_n1_b2048.txt 2>&1
