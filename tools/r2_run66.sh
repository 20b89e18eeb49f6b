#!/bin/bash
O=gpurun_out/r2_66; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bwd_sgd_fused" 2>&1 | tail -8 > $O/tests.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bwd_sgd_fused and (ragged_multihot or dim8 or zipf_onehot) and False" > $O/memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/memcheck.log
