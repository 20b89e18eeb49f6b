#!/bin/bash
O=gpurun_out/r2_86; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:table_absmax -s 4 -c 2 -o $O/prof_scan -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_scan.log 2>&1
