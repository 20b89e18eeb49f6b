#!/bin/bash
O=gpurun_out/r2_36; mkdir -p $O
python tools/bwd_profile.py --rows 10000000 --dim 64 --pooling 16 --batch 65536 > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'embbag_fwd_kernel|embbag_bwd_sort|grad_pack|grad_merge_apply' -s 12 -c 4 -o $O/prof_sweep_point python tools/bwd_profile.py --rows 10000000 --dim 64 --pooling 16 --batch 65536 > $O/ncu.log 2>&1
