#!/bin/bash
O=gpurun_out/r2_38; mkdir -p $O
timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err
DQRM_SCAN_IN_GRAPH=1 timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_scaningraph.json 2> $O/bench_scaningraph.err
DQRM_SCAN_IN_GRAPH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_timeline.py > $O/timeline_n1_scaningraph.txt 2>&1
