#!/bin/bash
O=gpurun_out/r2_28; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/tests.log
for C in 8 7 6 5 4 3; do
  DQRM_SCAN_CTAS_PER_SM=$C timeout 300 python bench.py --no-cpu-baseline --no-extras > $O/bench_c${C}.json 2>/dev/null
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_timeline.py > $O/timeline_n1.txt 2>&1
