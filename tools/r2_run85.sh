#!/bin/bash
O=gpurun_out/r2_85; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "reference_optimizer_golden or uncoalesced" 2>&1 | tail -15 > $O/tests.log
