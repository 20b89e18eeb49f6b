#!/bin/bash
O=gpurun_out/r2_12; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/tests.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err
