"""dqrm-b200: the data-parallel hot path of DQRM (INT4 QAT EmbeddingBag fwd / sparse bwd + SGD /
quantised sparse gradient exchange) as hand-written sm_100a CUDA behind the reference's Python surface.

Sub-modules mirror the reference tree: ``quantization_supp.quant_modules`` (+ ``_not_quantize_grad``),
``quantization_supp.quant_utils``, ``sgd_quantized_gradients_parallel_comm``, ``extend_distributed``,
``dlrm_s_pytorch_comm_grad``.  The CUDA library is loaded on first use and there is no CPU fallback.
"""
__version__ = "0.1.0"
