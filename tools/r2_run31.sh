#!/bin/bash
O=gpurun_out/r2_31; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/tests.log
timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err
DQRM_GEMM_BK=16 timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_bk16.json 2> $O/bench_bk16.err
DQRM_SIDE_BACKWARD=0 timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_noside.json 2> $O/bench_noside.err
DQRM_FUSE_LOCAL_DENSE=0 timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_nofuse.json 2> $O/bench_nofuse.err
DQRM_SCAN_CTAS_PER_SM=8 timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_scan8.json 2> $O/bench_scan8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_timeline.py > $O/timeline_n1.txt 2>&1
