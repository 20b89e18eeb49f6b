#!/bin/bash
O=gpurun_out/r2_68; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > $O/tests.log
timeout 200 python tools/bwd_profile.py > $O/prof_unfused.txt 2>&1
timeout 200 python tools/bwd_profile.py --dim 128 > $O/prof_unfused_d128.txt 2>&1
