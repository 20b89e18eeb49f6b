#!/bin/bash
O=gpurun_out/r2_26; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err
timeout 300 python tools/kernel_times.py > $O/ktimes_kaggle128.txt 2>&1
