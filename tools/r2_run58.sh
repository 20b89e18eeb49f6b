#!/bin/bash
O=gpurun_out/r2_58; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tracker.py -m gpu -q -x -k "large or bwd_sgd or max_cta or fused" 2>&1 | tail -5 > $O/tests.log
timeout 200 python tools/bwd_profile.py > $O/prof_unfused.txt 2>&1
timeout 200 python tools/bwd_profile.py --fused > $O/prof_fused.txt 2>&1
timeout 200 python tools/bwd_profile.py --fused --dim 16 --pooling 16 > $O/prof_fused_d16p16.txt 2>&1
timeout 200 python tools/bwd_profile.py --fused --dim 128 --pooling 64 --rows 40000000 > $O/prof_fused_d128.txt 2>&1
