#!/bin/bash
O=gpurun_out/r2_21; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shadow.py -x -q 2>&1 | tail -5 > $O/tests.log
for cfg in "10000000 64 16 65536" "40000000 128 64 65536"; do set -- $cfg; timeout 200 python tools/bwd_profile.py --rows $1 --dim $2 --pooling $3 --batch $4 >> $O/bwd_profile.txt 2>&1; done
timeout 1500 python bench_sweep.py --cpu --out $O/sweep.jsonl > $O/sweep.log 2>&1
