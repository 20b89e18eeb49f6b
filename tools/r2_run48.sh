#!/bin/bash
O=gpurun_out/r2_48; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/tests.log
timeout 240 $T tools/p2p_check.py > $O/p2p_check.log 2>&1
timeout 400 $T bench.py --gpus 2 --no-extras > $O/bench_n2.json 2> $O/bench_n2.err
DQRM_DENSE_XCHG_EARLY=0 timeout 400 $T bench.py --gpus 2 --no-extras > $O/bench_n2_noearly.json 2> $O/bench_n2_noearly.err
timeout 120 $T tools/mgpu_timeline.py > $O/timeline_n2.txt 2>&1
