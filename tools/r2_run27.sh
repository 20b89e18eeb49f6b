#!/bin/bash
O=gpurun_out/r2_27; mkdir -p $O
for P in 0 1; do for C in 8 4 2 1; do
  DQRM_GEMM_PREFETCH=$P DQRM_MLP_MAX_CLUSTER=$C timeout 300 python tools/kernel_times.py 2>&1 | grep -v Warn | head -12 > $O/ktimes_p${P}_c${C}.txt
  DQRM_GEMM_PREFETCH=$P DQRM_MLP_MAX_CLUSTER=$C timeout 300 python bench.py --no-cpu-baseline --no-extras > $O/bench_p${P}_c${C}.json 2>/dev/null
done; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_timeline.py > $O/timeline_n1.txt 2>&1
