#!/bin/bash
O=gpurun_out/r2_7; mkdir -p $O
timeout 300 python tools/tc_bench.py > $O/tc_bench.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py tests/test_gpu_full_config.py -x -q 2>&1 | tail -15 > $O/tests.log
python tools/tc_bench.py --shapes 8192x512x512 --iters 2 --paths tc > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 9 -c 3 -o $O/prof_tc python tools/tc_bench.py --shapes 8192x512x512 --iters 2 --paths tc > $O/ncu.log 2>&1
