#!/bin/bash
O=gpurun_out/r2_19; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/tests.log
timeout 120 python tools/kernel_times.py --batch 128 > $O/kt_128.txt 2>&1
timeout 120 python tools/kernel_times.py --batch 2048 > $O/kt_2048.txt 2>&1
timeout 200 python tools/kernel_times.py --workload terabyte --batch 8192 --steps 4 > $O/kt_tb8192.txt 2>&1
for cfg in "10000000 64 16 65536" "1000000 16 1 65536" "40000000 128 64 65536"; do set -- $cfg; timeout 200 python tools/bwd_profile.py --rows $1 --dim $2 --pooling $3 --batch $4 >> $O/bwd_profile.txt 2>&1; done
