#!/bin/bash
# round-2 first GPU trip: validate the tree as committed + the items NOTES_NEXT_ROUND lists as unvalidated
mkdir -p gpurun_out/r2_0
O=gpurun_out/r2_0
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/tests.log
DQRM_GEMM_PREFETCH=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/tests_prefetch.log
python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err
DQRM_GEMM_PREFETCH=1 python bench.py --no-extras --no-cpu-baseline > $O/bench_prefetch.json 2> $O/bench_prefetch.err
python tools/kernel_times.py > $O/ktimes_kaggle128.txt 2>&1
DQRM_GEMM_PREFETCH=1 python tools/kernel_times.py > $O/ktimes_kaggle128_prefetch.txt 2>&1
python tools/kernel_times.py --batch 2048 > $O/ktimes_kaggle2048.txt 2>&1
python -m deep_quantized_recommendation_model_dqrm_b200.dlrm_s_pytorch_comm_grad --arch-sparse-feature-size=16 \
  --arch-embedding-size=10000-10000-10000-10000-10000-10000-10000-10000 --arch-mlp-bot=13-512-256-64-16 \
  --arch-mlp-top=512-256-1 --data-generation=random --mini-batch-size=128 --num-batches=20 --num-indices-per-lookup=10 \
  --quantization_flag --embedding_bit=4 --weight_bit=4 --linear_channel --quantize_act_and_lin --loss-function=bce \
  --learning-rate=0.1 --print-freq=5 --quantize_embedding_bag_gradient --embedding_bag_gradient_bit_num=8 > $O/train_c1.log 2>&1
python bench.py --workload terabyte --batch 8192 --no-extras --no-cpu-baseline --steps 10 > $O/bench_tb.json 2> $O/bench_tb.err
python tools/kernel_times.py --workload terabyte --batch 8192 --steps 4 > $O/ktimes_tb8192.txt 2>&1
nvidia-smi > $O/smi.txt
