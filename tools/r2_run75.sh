#!/bin/bash
O=gpurun_out/r2_75; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tracker.py -m gpu -q -x -k "large or bwd_sgd or max_cta or fused" 2>&1 | tail -5 > $O/tests.log
timeout 200 python tools/bwd_profile.py --fused > $O/prof_fused_10m.txt 2>&1
timeout 200 python tools/bwd_profile.py --fused --dim 16 > $O/prof_fused_10m_d16.txt 2>&1
timeout 200 python tools/bwd_profile.py --fused --rows 40000000 --dim 128 --pooling 64 > $O/prof_fused_40m_d128.txt 2>&1
