#!/bin/bash
O=gpurun_out/r2_76; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'embbag_fwd_kernel|embbag_bwd_sort|grad_pack|grad_merge_apply' -s 12 -c 4 -o $O/prof_sweep_point -f python tools/bwd_profile.py --rows 10000000 --dim 64 --pooling 16 --batch 65536 > $O/ncu_sweep_point.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'embbag_bwd_cta' -s 6 -c 1 -o $O/prof_bwd_cta_b2048 -f python tools/timeline.py --policy full --batch 2048 > $O/ncu_bwd_cta.log 2>&1
timeout 200 python tools/bwd_profile.py --fused > $O/prof_fused.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large or bwd_sgd" 2>&1 | tail -3 > $O/tests.log
