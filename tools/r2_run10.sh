#!/bin/bash
O=gpurun_out/r2_10; mkdir -p $O
python tools/tc_bench.py --shapes 8192x512x512 --iters 2 --paths tc --modes fwd > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 3 -c 1 -o $O/prof_fwd python tools/tc_bench.py --shapes 8192x512x512 --iters 2 --paths tc --modes fwd > $O/ncu.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15 > $O/tests.log
