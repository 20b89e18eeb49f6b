#!/bin/bash
O=gpurun_out/r2_82; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "uncoalesced" 2>&1 | tail -15 > $O/tests.log
