"""GPU: the tensor-core (tcgen05 kind::tf32, split-TF32) QuantLinear kernels against an fp64 reference and against
the fp32 FFMA kernels, through the C ABI (dqrm_linear_fwd / dqrm_linear_bwd with path = DQRM_LINEAR_TC / _FFMA).
Reference op: F.linear on the integer-valued weights + autograd, quant_modules_not_quantize_grad.py:209.
Tolerance: 1e-5 relative (north_star), measured against the largest magnitude of each output tensor because the
entries are sums with cancellation."""
import numpy as np
import pytest
import torch

from deep_quantized_recommendation_model_dqrm_b200 import _lib

pytestmark = pytest.mark.gpu

SHAPES = [  # (batch, out_f, in_f, act)
    (128, 512, 13, 1), (128, 256, 512, 1), (129, 16, 64, 1), (300, 1, 256, 2), (1024, 512, 367, 1),
    (2048, 512, 512, 1), (2048, 64, 256, 0), (5000, 512, 415, 1), (8192, 256, 512, 1), (8192, 1, 256, 2),
    (37, 40, 24, 1),
]


def _run(path, x, Wi, bi, s, dout, act, accumulate_into=None):
    lib = _lib.load()
    B, in_f = x.shape
    out_f = Wi.shape[0]
    st = _lib.stream_ptr()
    out = torch.empty((B, out_f), device="cuda")
    _lib.check(lib.dqrm_linear_fwd(x.data_ptr(), Wi.data_ptr(), bi.data_ptr(), s.data_ptr(), B, out_f, in_f, act,
                                   out.data_ptr(), path, st), "fwd")
    dx = torch.empty_like(x)
    dW = torch.zeros_like(Wi) if accumulate_into is None else accumulate_into[0].clone()
    db = torch.zeros_like(bi) if accumulate_into is None else accumulate_into[1].clone()
    _lib.check(lib.dqrm_linear_bwd(x.data_ptr(), Wi.data_ptr(), s.data_ptr(), dout.data_ptr(), out.data_ptr(), B, out_f,
                                   in_f, act, dx.data_ptr(), dW.data_ptr(), db.data_ptr(),
                                   0 if accumulate_into is None else 1, path, st), "bwd")
    torch.cuda.synchronize()
    return out, dx, dW, db


def _reference(x, Wi, bi, s, dout, act):
    x64, W64, b64, s64, d64 = (t.double() for t in (x, Wi, bi, s, dout))
    z = (x64 @ W64.t() + b64) * s64
    out = torch.relu(z) if act == 1 else (torch.sigmoid(z) if act == 2 else z)
    gact = d64 * ((out > 0).double() if act == 1 else ((1 - out) * out if act == 2 else 1.0))
    g = gact * s64
    return out, g @ W64, (g.t() @ x64) / s64[:, None], g.sum(0) / s64


def _close(got, want, tol, what):
    scale = float(want.abs().max()) + 1e-30
    err = float((got.double() - want).abs().max()) / scale
    assert err <= tol, f"{what}: max error {err:.3e} of the tensor's largest magnitude (tolerance {tol:.1e})"
    return err


@pytest.mark.parametrize("shape", SHAPES)
def test_tensor_core_linear_vs_fp64_and_ffma(shape):
    B, out_f, in_f, act = shape
    g = torch.Generator(device="cuda").manual_seed(B * 7 + out_f * 3 + in_f)
    x = torch.randn((B, in_f), device="cuda", generator=g) * 2.0
    x[x.abs() < 0.3] = 0.0                                                   # ReLU-like sparsity
    Wi = torch.randint(-8, 8, (out_f, in_f), device="cuda", generator=g).float()
    bi = torch.randint(-8, 8, (out_f,), device="cuda", generator=g).float()
    s = torch.rand((out_f,), device="cuda", generator=g) * 0.05 + 0.005
    dout = torch.randn((B, out_f), device="cuda", generator=g) * 0.01
    want = _reference(x, Wi, bi, s, dout, act)
    tc = _run(_lib.LINEAR_TC, x, Wi, bi, s, dout, act)
    ff = _run(_lib.LINEAR_FFMA, x, Wi, bi, s, dout, act)
    for name, a, b, w in zip(("out", "dx", "dW", "db"), tc, ff, want):
        e_tc = _close(a, w, 1e-5, f"tensor-core {name} {shape}")
        e_ff = _close(b, w, 1e-5, f"ffma {name} {shape}")
        assert e_tc <= max(4 * e_ff, 4e-6), f"{name} {shape}: tensor-core error {e_tc:.2e} vs FFMA {e_ff:.2e}"
    # accumulate = 1 adds to what is there (AccumulateGrad semantics)
    base = (torch.randn_like(Wi), torch.randn_like(bi))
    acc = _run(_lib.LINEAR_TC, x, Wi, bi, s, dout, act, accumulate_into=base)
    torch.testing.assert_close(acc[2], base[0] + tc[2], rtol=0, atol=1e-6 * float(tc[2].abs().max() + 1))
    torch.testing.assert_close(acc[3], base[1] + tc[3], rtol=0, atol=1e-6 * float(tc[3].abs().max() + 1))
    # deterministic: a second launch reproduces the bits
    again = _run(_lib.LINEAR_TC, x, Wi, bi, s, dout, act)
    for a, b in zip(tc, again):
        assert torch.equal(a, b)


@pytest.mark.parametrize("shape", [(128, 512, 13, 1), (128, 256, 512, 1), (128, 64, 256, 1), (128, 16, 64, 1),
                                   (128, 512, 367, 1), (100, 1, 256, 2), (37, 40, 24, 0)])
def test_serial_slice_forward_has_the_bits_of_the_cluster_kernel(shape):
    """DQRM_LINEAR_FFMA_SERIAL (one CTA walks the K-slices a cluster would have split, slice sums added in rank order:
    the variant that can run beside the table scan) must equal the cluster split-K kernel bit for bit."""
    B, out_f, in_f, act = shape
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(B + out_f + in_f)
    x = torch.randn((B, in_f), device="cuda", generator=g)
    Wi = torch.randint(-8, 8, (out_f, in_f), device="cuda", generator=g).float()
    bi = torch.randint(-8, 8, (out_f,), device="cuda", generator=g).float()
    s = torch.rand((out_f,), device="cuda", generator=g) * 0.05 + 0.01
    outs = []
    for path in (_lib.LINEAR_FFMA, _lib.LINEAR_FFMA_SERIAL):
        out = torch.empty((B, out_f), device="cuda")
        _lib.check(lib.dqrm_linear_fwd(x.data_ptr(), Wi.data_ptr(), bi.data_ptr(), s.data_ptr(), B, out_f, in_f, act,
                                       out.data_ptr(), path, _lib.stream_ptr()), "fwd")
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
