#!/bin/bash
O=gpurun_out/r2_41; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests.log
