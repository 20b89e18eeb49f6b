#!/bin/bash
O=gpurun_out/r2_20; mkdir -p $O
python tools/bwd_profile.py --rows 10000000 --dim 64 --pooling 16 --batch 65536 > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:embbag_bwd_sort -s 3 -c 1 -o $O/prof_sort python tools/bwd_profile.py --rows 10000000 --dim 64 --pooling 16 --batch 65536 > $O/ncu.log 2>&1
