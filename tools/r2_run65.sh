#!/bin/bash
O=gpurun_out/r2_65; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/tests.log
timeout 1500 python bench_sweep.py --cpu --out $O/sweep.jsonl > $O/sweep.log 2>&1
timeout 200 python tools/bwd_profile.py --fused > $O/prof_fused.txt 2>&1
timeout 200 python tools/bwd_profile.py > $O/prof_unfused.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:embbag_bwd_sort -s 3 -c 1 -o $O/prof_sort_fused -f python tools/bwd_profile.py --fused > $O/ncu_fused.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
