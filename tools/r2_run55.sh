#!/bin/bash
O=gpurun_out/r2_55; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/tests.log
timeout 200 python tools/bwd_profile.py --fused > $O/prof_fused.txt 2>&1
timeout 1500 python bench_sweep.py --cpu --out $O/sweep.jsonl > $O/sweep.log 2>&1
