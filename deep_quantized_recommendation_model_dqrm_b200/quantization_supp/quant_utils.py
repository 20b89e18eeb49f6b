"""Drop-in for quantization_supp/quant_utils.py (the functions on the hot path).

Same names, argument meaning and error behaviour as the reference; the
arithmetic runs in libdqrm_b200 kernels on the current CUDA stream.  Scales stay
on the device (the reference's Python ``max()`` of two device scalars costs a
host sync per call, quant_utils.py:191).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib

__all__ = ["linear_quantize", "linear_dequantize", "finding_range_for_gradient",
           "symmetric_linear_quantization_param_two", "symmetric_linear_quantization_params",
           "ste_round", "SymmetricQuantFunction"]


def _need_cuda(t, what):
    if not t.is_cuda:
        raise _lib.DqrmLibraryError(f"{what}: expects a CUDA tensor (B200 path only, no CPU fallback)")


def _absmax(weight: torch.Tensor) -> torch.Tensor:
    """max|W| of one 2-D tensor via the table-scan kernel. Returns dev fp32 [1]."""
    _need_cuda(weight, "absmax")
    lib = _lib.load()
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    if w.dim() != 2:
        w = w.reshape(1, -1) if w.dim() < 2 else w.reshape(w.shape[0], -1)
    rows, dim = int(w.shape[0]), int(w.shape[1])
    out = torch.empty(1, dtype=torch.float32, device=w.device)
    ws = torch.zeros(int(lib.dqrm_scan_workspace_bytes(1)), dtype=torch.uint8, device=w.device)
    rc = lib.dqrm_table_absmax_scale(1, _lib.ptr_array([w]), _lib.i64_array([rows]), dim, 0, 0, 1,
                                     out.data_ptr(), None, None, ws.data_ptr(), _lib.stream_ptr())
    _lib.check(rc, "dqrm_table_absmax_scale")
    return out


def finding_range_for_gradient(embedding_bag):
    """max(|min W|, |max W|) (quant_utils.py:130-139). 0-dim device tensor."""
    weight = embedding_bag.weight.data if isinstance(embedding_bag, torch.nn.Module) else embedding_bag
    return _absmax(weight).view(())


def symmetric_linear_quantization_param_two(num_bits, embedding_bag, embedding_bound, num_embeddings, embedding_id):
    """Per-table scale = clamp(max|W|, 1e-8) / (2^(bits-1)-1) (quant_utils.py:141-194).
    Accepts a module (uses ``.weight.data``) or a raw 2-D tensor; the last three
    arguments are unused, as in the reference.  Returns a 0-dim fp32 device tensor."""
    weight = embedding_bag.weight.data if isinstance(embedding_bag, torch.nn.Module) else embedding_bag
    am = _absmax(weight)
    s = torch.empty_like(am)
    rc = _lib.load().dqrm_scale_from_absmax(1, am.data_ptr(), int(num_bits), s.data_ptr(), None, _lib.stream_ptr())
    _lib.check(rc, "dqrm_scale_from_absmax")
    return s.view(())


def symmetric_linear_quantization_params(num_bits, saturation_min, saturation_max, per_channel=False):
    """Scale from a (min, max) range (quant_utils.py:196-220); per_channel works on vectors."""
    with torch.no_grad():
        _need_cuda(saturation_min, "symmetric_linear_quantization_params")
        am = torch.maximum(saturation_min.abs(), saturation_max.abs()).float().contiguous().view(-1)
        s = torch.empty_like(am)
        rc = _lib.load().dqrm_scale_from_absmax(int(am.numel()), am.data_ptr(), int(num_bits), s.data_ptr(), None,
                                                _lib.stream_ptr())
        _lib.check(rc, "dqrm_scale_from_absmax")
        return s if per_channel else s.view(saturation_min.shape)


def _fake_quant(x, k, scale):
    _need_cuda(x, "SymmetricQuantFunction")
    xc = x.detach().float().contiguous()
    sc = scale.detach().float().contiguous().view(-1)
    if xc.dim() == 2:
        rows, cols = xc.shape
    else:
        rows, cols = 1, xc.numel()
    per_row = 0
    if sc.numel() != 1:
        if xc.dim() == 2 and sc.numel() == rows:
            per_row = 1
        elif xc.dim() == 1 and sc.numel() == xc.numel():
            rows, cols, per_row = xc.numel(), 1, 1         # element-wise scale (QuantLinear bias, qm:153-154)
        else:
            raise ValueError(f"scale with {sc.numel()} entries does not broadcast over input {tuple(x.shape)}")
    q = torch.empty_like(xc)
    rc = _lib.load().dqrm_fake_quant(xc.data_ptr(), rows, cols, sc.data_ptr(), per_row, int(k), q.data_ptr(), None,
                                     _lib.stream_ptr())
    _lib.check(rc, "dqrm_fake_quant")
    return q.view(x.shape)


def linear_quantize(input, scale, zero_point, inplace=False):
    """round(1/scale * input + zero_point) without the clamp (quant_utils.py:75-101).
    Only the symmetric zero_point == 0 form exists on the hot path."""
    if inplace:
        raise NotImplementedError("linear_quantize(inplace=True) is the unused backwardpass branch (quant_utils.py:337-340)")
    return torch.round(1.0 / _bcast(scale, input) * input + zero_point)


def _bcast(scale, x):
    if x.dim() == 4:
        return scale.view(-1, 1, 1, 1)
    if x.dim() == 2:
        if scale.dim() != 1 or scale.shape[0] != 1:
            return scale.view(-1, 1)
        return scale
    return scale.view(-1)


def linear_dequantize(input_q, scale, zero_point, inplace=False):
    """(input_q - zero_point) * scale (quant_utils.py:103-128)."""
    if input_q.dim() == 4:
        scale, zero_point = scale.view(-1, 1, 1, 1), zero_point.view(-1, 1, 1, 1)
    elif input_q.dim() == 2:
        scale, zero_point = scale.view(-1, 1), zero_point.view(-1, 1)
    else:
        scale, zero_point = scale.view(-1), zero_point.view(-1)
    if inplace:
        input_q.sub_(zero_point).mul_(scale)
        return input_q
    return (input_q - zero_point) * scale


class ste_round(Function):
    """Straight-through round (quant_utils.py:284-299)."""

    @staticmethod
    def forward(ctx, x):
        with torch.no_grad():
            return torch.round(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output


class SymmetricQuantFunction(Function):
    """clamp(round(x / scale), -2^(k-1), 2^(k-1)-1) with a straight-through backward
    ``grad / scale`` and no clipping mask (quant_utils.py:316-363)."""

    @staticmethod
    def forward(ctx, x, k, specified_scale=None, backwardpass=False):
        if specified_scale is None:
            raise ValueError("The SymmetricQuantFunction requires a pre-calculated scaling factor")
        if backwardpass:
            raise NotImplementedError("backwardpass=True is never used by the hot path (quant_utils.py:337-340)")
        ctx.scale = specified_scale
        return _fake_quant(x, k, specified_scale)

    @staticmethod
    def backward(ctx, grad_output):
        scale = ctx.scale
        if grad_output.dim() == 4:
            scale = scale.view(-1, 1, 1, 1)
        elif grad_output.dim() == 2:
            scale = scale.view(-1, 1)
        else:
            scale = scale.view(-1)
        return grad_output / scale, None, None, None
