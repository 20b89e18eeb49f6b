#!/bin/bash
O=gpurun_out/r2_6; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_linear_tc.py -q 2>&1 | tail -15 > $O/tc.log
timeout 120 python tools/kernel_times.py --batch 2048 > $O/kt_2048_tc.txt 2>&1
timeout 120 python tools/kernel_times.py --batch 512 > $O/kt_512_tc.txt 2>&1
timeout 200 python tools/kernel_times.py --workload terabyte --batch 8192 --steps 4 > $O/kt_tb8192.txt 2>&1
timeout 200 python tools/kernel_times.py --workload terabyte --batch 1024 --steps 4 > $O/kt_tb1024.txt 2>&1
