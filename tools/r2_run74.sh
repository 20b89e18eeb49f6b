#!/bin/bash
O=gpurun_out/r2_74; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 200 python tools/bwd_profile.py --fused > $O/prof_fused.txt 2>&1
