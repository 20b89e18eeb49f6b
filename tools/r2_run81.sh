#!/bin/bash
O=gpurun_out/r2_81; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_timeline.py --batch 2048 > $O/timeline_n1_b2048.txt 2>&1
