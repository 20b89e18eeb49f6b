#!/bin/bash
O=gpurun_out/r2_73; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tracker.py -m gpu -q -x -k "large or bwd_sgd or max_cta or fused" 2>&1 | tail -5 > $O/tests.log
timeout 200 python tools/bwd_profile.py --fused > $O/prof_fused_10m.txt 2>&1
timeout 200 python tools/bwd_profile.py --fused --rows 1000000 > $O/prof_fused_1m.txt 2>&1
timeout 200 python tools/bwd_profile.py --fused --rows 40000000 > $O/prof_fused_40m.txt 2>&1
DQRM_BWD_RADIX_BITS=8 timeout 200 python tools/bwd_profile.py --fused --rows 1000000 > $O/prof_fused_1m_r8.txt 2>&1
DQRM_BWD_RADIX_BITS=8 timeout 200 python tools/bwd_profile.py --fused --rows 40000000 > $O/prof_fused_40m_r8.txt 2>&1
