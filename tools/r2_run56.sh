#!/bin/bash
O=gpurun_out/r2_56; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -15 > $O/tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:embbag_bwd_sort -s 3 -c 1 -o $O/prof_sort_fused -f python tools/bwd_profile.py --fused > $O/ncu_fused.log 2>&1
