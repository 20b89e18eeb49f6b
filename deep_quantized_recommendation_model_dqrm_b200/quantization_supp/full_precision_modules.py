"""Type names that the reference's DP optimizer checks with isinstance
(quantization_supp/full_precision_modules.py:11-59; sgd_quantized_gradients_parallel_comm.py:340).
FP32 modules carrying gradient-compression buffers; only the names matter on the hot path."""
import torch.nn as nn


class EmbeddingBagCompressedGrad(nn.EmbeddingBag):
    pass


class LinearCompressedGrad(nn.Linear):
    pass
