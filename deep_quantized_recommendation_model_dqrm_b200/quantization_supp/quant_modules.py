"""Drop-in for quantization_supp/quant_modules_not_quantize_grad.py (the module the reference
drivers actually use): QuantEmbeddingBagTwo, QuantLinear, QuantAct.

Same constructor arguments, attributes and return conventions; forward/backward run
in libdqrm_b200 kernels.  CUDA only -- a module on the CPU raises (no fallback).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function
from torch.nn import Module, Parameter

from .. import _lib
from ..tables import EmbeddingTableGroup
from .quant_utils import *  # noqa: F401,F403  (the reference star-imports quant_utils here)

__all__ = ["QuantEmbeddingBagTwo", "QuantLinear", "QuantAct", "EmbBagGroupFunction"]


# --------------------------------------------------------------------------
# QAT EmbeddingBag
# --------------------------------------------------------------------------
class EmbBagGroupFunction(Function):
    """Autograd node of the fused multi-table QAT EmbeddingBag.

    forward : (a3)  out[T,B,D] = dequantised fake-quantised sum pooling
    backward: (a4 bwd, a5, a7-1/2) fills the group's de-duplicated row gradients;
              no dense or per-lookup gradient is ever materialised.  The table
              weights receive ``None`` (or, with ``group.materialize_grads``, the
              coalesced sparse COO gradient the reference would hold).
    """

    @staticmethod
    def forward(ctx, group, indices, offsets, idx_begin, bags, full_precision, *weights):
        out = group.forward(indices, offsets, idx_begin, bags, full_precision)
        ctx.group = group
        ctx.last = group.last            # this forward's inputs (another forward may run before backward)
        return out

    @staticmethod
    def backward(ctx, dout):
        g = ctx.group
        if dout.stride(2) != 1 or dout.stride(0) % 4 or dout.stride(1) % 4 or dout.data_ptr() % 16:
            dout = dout.contiguous()
        ste_done = getattr(g, "ste_done_for", None) == dout.data_ptr()      # fused into the interaction backward
        g.ste_done_for = None
        fu = g.fused_update
        if fu is not None and g.dp_world == 1 and not g.materialize_grads:
            # single process, un-quantised: sort + de-duplicate + row update in one call (north-star kernel 3)
            g.backward_sgd(dout, fu["lr"], momentum=fu["momentum"], eps=fu["eps"], last=ctx.last, ste_done=ste_done)
            g.applied_fused = True
            return (None, None, None, None, None, None) + (None,) * g.T
        if g.side_backward:
            # de-duplicating backward (+ the exchange / pack) on the group's side stream: the bottom-MLP backward that
            # autograd runs next only needs the interaction's OTHER output; finish_exchange() joins
            g.backward_async(dout, last=ctx.last, ste_done=ste_done)
        else:
            g.backward(dout, world=g.dp_world, last=ctx.last, ste_done=ste_done)
            if g.eager_exchange and g.dp_world > 1:
                g.start_exchange()           # overlaps with the bottom-MLP backward that autograd runs next
        if g.materialize_grads:
            grads = tuple(g.sparse_grad(t) for t in range(g.T))
        else:
            grads = (None,) * g.T
        return (None, None, None, None, None, None) + grads


def _new_group(weights, embedding_bit):
    g = EmbeddingTableGroup(weights, embedding_bit=embedding_bit)
    g.dp_world, g.dp_rank, g.process_group = 1, 0, None
    g.materialize_grads = False
    g.modules = None
    return g


class QuantEmbeddingBagTwo(Module):
    """INT-k quantisation-aware EmbeddingBag (quant_modules_not_quantize_grad.py:220-398).

    fp32 master rows; every training forward rescans the table for its scale
    (a1), pools in fp32, fake-quantises the POOLED vector and dequantises.
    ``_weight`` (extension) adopts an existing [N, D] CUDA tensor -- e.g. a slice
    of a table arena -- instead of drawing the reference's numpy uniform init on
    the host (qm:273-275), which does not scale to 10M-row tables.
    """

    def __init__(self, num_embeddings, embedding_dim, embedding_bit=4, full_precision_flag=False,
                 quant_mode="symmetric", fix_flag=False, weight_percentile=0, embedding_id=None, _weight=None):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.embedding_bit = embedding_bit
        self.full_precision_flag = full_precision_flag
        self.quant_mode = quant_mode
        self.fix_flag = fix_flag
        self.weight_percentile = weight_percentile
        self.batch_size = 128
        self.register_buffer("eb_scaling_factor", torch.zeros(self.batch_size, 1), persistent=True)
        self.register_buffer("embedding_bound", torch.sqrt(torch.tensor(1 / self.num_embeddings)) * 4.0, persistent=False)
        self.register_buffer("now_iteration", torch.zeros(1), persistent=True)
        self.register_buffer("iteration_bound", torch.zeros(1), persistent=True)
        self.register_buffer("iteration_nt", torch.zeros(1), persistent=True)
        self.embedding_id = embedding_id
        self.register_buffer("emb_scaling_factor", torch.zeros(1), persistent=True)
        self.register_buffer("gradient_bit_width", torch.zeros(1), persistent=True)
        if _weight is None:
            W = np.random.uniform(low=-np.sqrt(1 / num_embeddings), high=np.sqrt(1 / num_embeddings),
                                  size=(num_embeddings, embedding_dim)).astype(np.float32)
            _weight = torch.tensor(W)
        self.embedding_bag = nn.EmbeddingBag(num_embeddings, embedding_dim, mode="sum", sparse=True, _weight=_weight)
        self._solo = None             # single-table group used when the module is called directly
        self._group = None            # group (solo or the model's joint group) that ran this table last
        self._group_index = 0

    def __repr__(self):
        s = super().__repr__()
        return "(" + s + " embedding_bit = {}, full_precision_flag = {}, quant_mode = {})".format(
            self.embedding_bit, self.full_precision_flag, self.quant_mode)

    def fix(self):
        self.fix_flag = True

    def unfix(self):
        self.fix_flag = False

    def set_iteration_bound(self):
        """Periodic-rescan schedule of the reference (quant_modules_not_quantize_grad.py:303-315). Its only caller
        there is commented out (:348-361), so the counters never move and every training forward rescans; kept
        for API parity."""
        if self.iteration_nt == 1 and self.iteration_bound == 0:
            self.iteration_bound += 1000
            print("bound increasing to {}".format(self.iteration_bound.item()))

    # -- group plumbing -----------------------------------------------------
    def _own_group(self):
        w = self.embedding_bag.weight
        g = self._solo
        if g is None or g.weights[0] is not w or g.device != w.device or g.embedding_bit != self.embedding_bit:
            g = _new_group([w], self.embedding_bit)
            g.modules = [self]
            self._solo = g
        self._group, self._group_index = g, 0
        return g

    @property
    def output_integer(self):
        """Integer codes of the last forward as fp32 [B, D] (the reference keeps
        the fp32 code tensor, qm:378); stored here as int8/int16."""
        g = self._group
        if g is None or getattr(g, "codes", None) is None:
            return torch.zeros((1, 16))
        return g.codes[self._group_index].float()

    def forward(self, input, offsets=None, per_sample_weights=None, full_precision_flag=False, test_mode=False):
        full_precision_flag = full_precision_flag or self.full_precision_flag
        if self.quant_mode not in ("symmetric", "speed_symmetric", "asymmetric"):
            raise ValueError("unknown quant mode: {}".format(self.quant_mode))
        if self.quant_mode != "symmetric" and not full_precision_flag:
            raise Exception("for embedding weights, we only support symmetric quantization")
        if per_sample_weights is not None:
            print("Warning: Embedding Table Assumes per_sample_weights to be None but it is not")
        g = self._own_group()
        if input.dim() == 2:                      # nn.EmbeddingBag 2-D input: fixed-length bags
            offsets = torch.arange(0, input.numel(), input.shape[1], device=input.device)
            input = input.reshape(-1)
        idx, off, idx_begin, bags = EmbeddingTableGroup.pack_inputs([input], [offsets], g.device)
        if (not full_precision_flag and not test_mode) or not g.scale_valid:
            g.scan_scales()
            self.eb_scaling_factor = g.scale[0]
        out = EmbBagGroupFunction.apply(g, idx, off, idx_begin, bags, full_precision_flag, self.embedding_bag.weight)
        return out[0]


# --------------------------------------------------------------------------
# QAT Linear
# --------------------------------------------------------------------------
class _QuantLinearFunction(Function):
    """y = (x W_int^t + b_int) * s_row with W_int, b_int from dqrm_linear_fakequant; backward is
    the straight-through estimator of SymmetricQuantFunction (grad / s_row)."""

    @staticmethod
    def forward(ctx, x, weight, bias, bits, module):
        lib = _lib.load()
        out_f, in_f = weight.shape
        w = weight.detach()
        W_int = torch.empty_like(w)
        s_row = torch.empty(out_f, dtype=torch.float32, device=w.device)
        b_int = torch.empty_like(bias.detach()) if bias is not None else None
        rc = lib.dqrm_linear_fakequant(w.data_ptr(), _lib.ptr(bias.detach() if bias is not None else None), out_f, in_f,
                                       int(bits), W_int.data_ptr(), _lib.ptr(b_int), s_row.data_ptr(),
                                       _lib.stream_ptr())
        _lib.check(rc, "dqrm_linear_fakequant")
        y = F.linear(x, W_int, b_int)
        y.mul_(s_row.view(1, -1))
        ctx.save_for_backward(x, W_int, s_row)
        ctx.has_bias = bias is not None
        module.fc_scaling_factor = s_row
        module.weight_integer = W_int
        module.bias_integer = b_int
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W_int, s_row = ctx.saved_tensors
        g = dy * s_row.view(1, -1)
        dx = g.mm(W_int) if ctx.needs_input_grad[0] else None
        dW = g.t().mm(x)
        dW.div_(s_row.view(-1, 1))
        db = None
        if ctx.has_bias:
            db = g.sum(0)
            db.div_(s_row)
        return dx, dW, db, None, None


class _FusedQuantLinearFunction(Function):
    """One kernel forward (GEMM + bias + per-row scale + activation), two backward (dx; dW/db accumulated
    straight into the gradient arena) on the weights fake-quantised by DenseArena.fakequant_all()."""

    @staticmethod
    def forward(ctx, x, weight, bias, module, act):
        lib = _lib.load()
        x = x.contiguous()
        B = x.shape[0]
        out_f, in_f = weight.shape
        out = torch.empty((B, out_f), dtype=torch.float32, device=x.device)
        rc = lib.dqrm_linear_fwd(x.data_ptr(), module._w_int.data_ptr(), _lib.ptr(module._b_int),
                                 module._fc_scale.data_ptr(), B, out_f, in_f, act, out.data_ptr(),
                                 _lib.linear_path if module.fwd_path is None else module.fwd_path, _lib.stream_ptr())
        _lib.check(rc, "dqrm_linear_fwd")
        ctx.save_for_backward(x, out)
        ctx.module, ctx.act = module, act
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x, out = ctx.saved_tensors
        m = ctx.module
        dout = dout.contiguous()
        B = x.shape[0]
        out_f, in_f = m.weight.shape
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        # grads cleared since the last write (clear_gradients) are overwritten, otherwise accumulated
        accumulate = 1 if getattr(m, "_grad_dirty", True) else 0
        args = (x.data_ptr(), m._w_int.data_ptr(), m._fc_scale.data_ptr(), dout.data_ptr(), out.data_ptr(), B, out_f,
                in_f, ctx.act)
        wg, bg = m.weight.grad.data_ptr(), _lib.ptr(m.bias.grad if m.bias is not None else None)
        arena = getattr(m, "_arena", None)
        side = arena.next_side_stream() if arena is not None else None
        if side is None:
            rc = lib.dqrm_linear_bwd(*args, _lib.ptr(dx), wg, bg, accumulate, _lib.linear_path, _lib.stream_ptr())
            _lib.check(rc, "dqrm_linear_bwd")
        else:
            # dx is on the critical path of the backward chain; dW/db only feed the optimizer, so they run
            # on the arena's side stream and are joined in DenseArena.join() before the gradients are used
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            if dx is not None:
                _lib.check(lib.dqrm_linear_bwd(*args, dx.data_ptr(), None, None, 0, _lib.linear_path, main.cuda_stream),
                           "dqrm_linear_bwd")
            _lib.check(lib.dqrm_linear_bwd(*args, None, wg, bg, accumulate, _lib.linear_path, side.cuda_stream),
                       "dqrm_linear_bwd")
            arena.keepalive.append((x, out, dout))        # these must outlive the side-stream kernel
            arena.after_dw(m)
        m._grad_dirty = True
        return dx, None, None, None, None


class QuantLinear(Module):
    """Per-channel INT-k weight + bias QAT linear layer (quant_modules_not_quantize_grad.py:20-211).
    Returns a tuple ``(y, None)`` like the reference (quantize_activation=False path)."""

    fwd_path = None      # forward contraction engine of the fused kernel (None: _lib.linear_path); graph_step sets
                         # LINEAR_FFMA_SERIAL on the bottom MLP when it runs beside the table scan (same bits)

    def __init__(self, weight_bit=4, bias_bit=None, full_precision_flag=False, quant_mode="symmetric",
                 per_channel=False, fix_flag=False, weight_percentile=0, quantize_activation=False):
        super().__init__()
        self.full_precision_flag = full_precision_flag
        self.weight_bit = weight_bit
        self.quant_mode = quant_mode
        self.per_channel = per_channel
        self.fix_flag = fix_flag
        self.weight_percentile = weight_percentile
        self.bias_bit = bias_bit
        self.quantize_bias = bias_bit is not None
        self.counter = 0
        self.quantize_activation = quantize_activation

    def __repr__(self):
        s = super().__repr__()
        return "(" + s + " weight_bit={}, full_precision_flag={}, quantize_fn={})".format(
            self.weight_bit, self.full_precision_flag, self.quant_mode)

    def set_param(self, linear):
        self.in_features = linear.in_features
        self.out_features = linear.out_features
        self.register_buffer("fc_scaling_factor", torch.zeros(self.out_features))
        self.register_buffer("correct_output_scale", torch.ones(self.out_features))
        self.weight = Parameter(linear.weight.data.clone())
        self.register_buffer("weight_integer", torch.zeros_like(self.weight), persistent=False)
        self.register_buffer("bias_integer", torch.zeros_like(linear.bias), persistent=False)
        self.register_buffer("weight_scaling_factor", torch.zeros(self.out_features))
        try:
            self.bias = Parameter(linear.bias.data.clone())
            self.register_buffer("bias_scaling_factor", torch.zeros(self.out_features))
        except AttributeError:
            self.bias = None

    def fix(self):
        self.fix_flag = True

    def unfix(self):
        self.fix_flag = False

    def forward_fused(self, x, act):
        """Fused layer + activation (act: 0 none, 1 relu, 2 sigmoid).  Requires the model's DenseArena
        (weights / grads are arena views) and a DenseArena.fakequant_all() since the last weight update."""
        return _FusedQuantLinearFunction.apply(x, self.weight, self.bias, self, act)

    def forward(self, x, prev_act_scaling_factor=None):
        if self.full_precision_flag:
            return F.linear(x, weight=self.weight, bias=self.bias), None
        if type(x) is tuple:
            prev_act_scaling_factor = x[1]
            x = x[0]
        if self.quant_mode == "asymmetric":
            raise Exception("For weight, we only support symmetric quantization.")
        if self.quant_mode != "symmetric":
            raise ValueError("unknown quant mode: {}".format(self.quant_mode))
        if prev_act_scaling_factor is not None or self.quantize_activation:
            raise NotImplementedError("activation quantisation (QuantAct chain) is outside the hot path: the DQRM "
                                      "scripts run --linear_channel, which forces quantize_activation=False "
                                      "(dlrm_s_pytorch_comm_grad.py:1155-1156)")
        if not self.per_channel:
            raise NotImplementedError("per-tensor QuantLinear is not on the hot path (scripts pass --linear_channel)")
        if not self.weight.is_cuda:
            raise _lib.DqrmLibraryError("QuantLinear: CUDA only (no CPU fallback)")
        if self.quantize_bias and self.bias_bit != self.weight_bit:
            raise NotImplementedError("bias_bit != weight_bit is never used by the reference (dlrm_s_pytorch_comm_grad.py:318-319)")
        y = _QuantLinearFunction.apply(x, self.weight, self.bias if self.quantize_bias else None, self.weight_bit, self)
        if self.bias is not None and not self.quantize_bias:
            raise NotImplementedError("un-quantised bias with quantised weight is not on the hot path")
        return y, None


class QuantAct(Module):
    """Placeholder for the activation quantiser (quant_modules_not_quantize_grad.py QuantAct).
    DLRM_Net constructs two of these (dlrm_s_pytorch_comm_grad.py:454-455) but the hot
    configuration (--linear_channel => quantize_activation=False) never calls them."""

    def __init__(self, activation_bit=8, act_range_momentum=0.95, full_precision_flag=False, running_stat=True,
                 quant_mode="symmetric", fix_flag=False, act_percentile=0, fixed_point_quantization=False):
        super().__init__()
        self.activation_bit = activation_bit
        self.act_range_momentum = act_range_momentum
        self.full_precision_flag = full_precision_flag
        self.quant_mode = quant_mode
        self.fixed_point_quantization = fixed_point_quantization

    def forward(self, *a, **k):
        raise NotImplementedError("QuantAct is out of scope: the DQRM hot path runs with quantize_activation=False")
