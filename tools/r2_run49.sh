#!/bin/bash
O=gpurun_out/r2_49; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/tests.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err
DQRM_MLP_TC_MIN_DIM=1 timeout 600 python bench.py --no-cpu-baseline > $O/bench_alltc.json 2> $O/bench_alltc.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_timeline.py --batch 2048 > $O/timeline_n1_b2048.txt 2>&1
