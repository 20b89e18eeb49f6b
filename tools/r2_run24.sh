#!/bin/bash
O=gpurun_out/r2_24; mkdir -p $O
timeout 1200 python bench_sweep.py --cpu --out $O/sweep.jsonl > $O/sweep.log 2>&1
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_s2.json 2> $O/bench_s2.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu.log 2>&1
