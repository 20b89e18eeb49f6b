"""The reference imports these names from ``sgd_quantized_gradients`` and then shadows them with
``sgd_quantized_gradients_parallel_comm`` (dlrm_s_pytorch_comm_grad.py:113-126); same objects here."""
from .sgd_quantized_gradients_parallel_comm import *  # noqa: F401,F403
