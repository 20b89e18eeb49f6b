#!/bin/bash
O=gpurun_out/r2_50; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/tests.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python tools/timeline.py --policy full > $O/timeline_n1.txt 2>&1
