#!/bin/bash
O=gpurun_out/r2_30; mkdir -p $O
timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err
DQRM_OVERLAP_BOTTOM=0 timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_nooverlap.json 2> $O/bench_nooverlap.err
DQRM_OVERLAP_BOTTOM=0 DQRM_GEMM_BK=16 timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_nooverlap_bk16.json 2> $O/bench_nooverlap.err
DQRM_OVERLAP_BOTTOM=0 DQRM_SIDE_BACKWARD=0 timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_nooverlap_noside.json 2> $O/bench_nooverlap.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_timeline.py > $O/timeline_n1.txt 2>&1
