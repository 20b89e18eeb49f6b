#!/bin/bash
O=gpurun_out/r2_77; mkdir -p $O
timeout 1500 python bench_sweep.py --cpu --out $O/sweep.jsonl > $O/sweep.log 2>&1
