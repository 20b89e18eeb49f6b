#!/bin/bash
O=gpurun_out/r2_80; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $O/tests.log
for S in 32 64; do
DQRM_BWD_SHORT_ROW=$S timeout 200 python tools/bwd_profile.py --fused --rows 1000000 --pooling 64 > $O/heavy_s$S.txt 2>&1
done
timeout 200 python tools/bwd_profile.py --fused --rows 1000000 --pooling 64 --dim 128 > $O/heavy_d128_s16.txt 2>&1
timeout 200 python tools/bwd_profile.py --rows 1000000 --pooling 64 > $O/heavy_unfused_s16.txt 2>&1
