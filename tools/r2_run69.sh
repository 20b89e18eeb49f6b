#!/bin/bash
O=gpurun_out/r2_69; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_timeline.py --batch 2048 > $O/timeline_n1_b2048.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29534 tools/mgpu_timeline.py --batch 8192 --workload terabyte > $O/timeline_n1_terabyte.txt 2>&1
timeout 900 python bench.py --no-cpu-baseline --only c2,c3 > $O/bench.json 2> $O/bench.err
