#!/bin/bash
O=gpurun_out/r2_63; mkdir -p $O
DQRM_SCAN_IN_GRAPH=1 timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_scaningraph.json 2> $O/bench_scaningraph.err
timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_default.json 2> $O/bench_default.err
DQRM_SCAN_IN_GRAPH=1 timeout 600 python bench.py --no-cpu-baseline --no-extras --batch 2048 > $O/bench_scaningraph_b2048.json 2> $O/bench_scaningraph_b2048.err
timeout 600 python bench.py --no-cpu-baseline --no-extras --batch 2048 > $O/bench_default_b2048.json 2> $O/bench_default_b2048.err
