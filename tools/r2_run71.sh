#!/bin/bash
O=gpurun_out/r2_71; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $T bench.py --gpus 2 --only c3 --no-cpu-baseline > $O/c3_default.json 2> $O/c3_default.err
DQRM_FUSE_DENSE_XCHG=0 timeout 400 $T bench.py --gpus 2 --only c3 --no-cpu-baseline > $O/c3_nofuse.json 2> $O/c3_nofuse.err
DQRM_BWD_CTA_SORT=bitonic timeout 400 $T bench.py --gpus 2 --only c3 --no-cpu-baseline > $O/c3_bitonic.json 2> $O/c3_bitonic.err
timeout 300 $T bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
