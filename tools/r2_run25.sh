#!/bin/bash
O=gpurun_out/r2_25; mkdir -p $O
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_s2.json 2> $O/bench_s2.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_gemm -s 40 -c 20 -o $O/prof_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu.log 2>&1
