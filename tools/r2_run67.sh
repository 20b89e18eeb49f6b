#!/bin/bash
O=gpurun_out/r2_67; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "train_cli" 2>&1 | tail -30 > $O/tests.log
timeout 300 python -m deep_quantized_recommendation_model_dqrm_b200.dlrm_s_pytorch_comm_grad --arch-sparse-feature-size=16 --arch-embedding-size=10000-10000-10000-10000-10000-10000-10000-10000 --arch-mlp-bot=13-512-256-64-16 --arch-mlp-top=512-256-1 --data-generation=random --mini-batch-size=128 --num-batches=20 --num-indices-per-lookup=10 --quantization_flag --embedding_bit=4 --weight_bit=4 --linear_channel --quantize_act_and_lin --loss-function=bce --learning-rate=0.1 --print-freq=5 > $O/cli.log 2>&1
