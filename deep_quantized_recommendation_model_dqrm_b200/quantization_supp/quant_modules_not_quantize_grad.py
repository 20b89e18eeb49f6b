"""The reference drivers import QuantEmbeddingBagTwo/QuantLinear/QuantAct from ``quant_modules`` and then
re-import the same names from this module, which shadows them (dlrm_s_pytorch_comm_grad.py:101-108).
Both names resolve to the same implementation here."""
from .quant_modules import *  # noqa: F401,F403
from .quant_modules import QuantAct, QuantEmbeddingBagTwo, QuantLinear  # noqa: F401
