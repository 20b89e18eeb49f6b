#!/bin/bash
O=gpurun_out/r2_83; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu.log 2>&1
