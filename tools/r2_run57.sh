#!/bin/bash
O=gpurun_out/r2_57; mkdir -p $O
timeout 200 python tools/bwd_profile.py --fused > $O/prof_fused_ns3.txt 2>&1
timeout 200 python tools/bwd_profile.py --fused > $O/prof_fused_ns2.txt 2>&1
timeout 200 python tools/bwd_profile.py > $O/prof_unfused_ns2.txt 2>&1
timeout 200 python tools/bwd_profile.py --fused --dim 128 --pooling 64 --rows 40000000 > $O/prof_fused_d128_ns2.txt 2>&1
timeout 200 python tools/bwd_profile.py --fused --dim 128 --pooling 64 --rows 40000000 > $O/prof_fused_d128_ns3.txt 2>&1
