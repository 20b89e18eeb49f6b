#!/bin/bash
O=gpurun_out/r2_79; mkdir -p $O
for S in 4 8 16; do
DQRM_BWD_SHORT_ROW=$S timeout 200 python tools/bwd_profile.py --fused --rows 1000000 --pooling 64 > $O/heavy_s$S.txt 2>&1
DQRM_BWD_SHORT_ROW=$S timeout 200 python tools/bwd_profile.py --fused > $O/normal_s$S.txt 2>&1
done
DQRM_BWD_SHORT_ROW=8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large or bwd_sgd" 2>&1 | tail -3 > $O/tests_s8.log
