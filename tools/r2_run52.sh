#!/bin/bash
O=gpurun_out/r2_52; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:embbag_bwd_sort -s 3 -c 1 -o $O/prof_sort_fused -f python tools/bwd_profile.py --fused > $O/ncu_fused.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:embbag_bwd_sort -s 3 -c 1 -o $O/prof_sort_unfused -f python tools/bwd_profile.py > $O/ncu_unfused.log 2>&1
