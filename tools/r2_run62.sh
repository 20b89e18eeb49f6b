#!/bin/bash
O=gpurun_out/r2_62; mkdir -p $O
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
