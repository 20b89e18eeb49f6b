#!/bin/bash
O=gpurun_out/r2_40; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $O/tests.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
