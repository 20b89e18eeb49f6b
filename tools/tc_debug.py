"""Diagnostics for the tensor-core linear kernels: per shape / mode error statistics against fp64.
    [DQRM_MLP_MAX_CLUSTER=1] python tools/tc_debug.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_quantized_recommendation_model_dqrm_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)

def run(path, x, Wi, bi, s, dout, act):
    B, in_f = x.shape; out_f = Wi.shape[0]; st = _lib.stream_ptr()
    out = torch.full((B, out_f), 7.0, device="cuda")
    _lib.check(lib.dqrm_linear_fwd(x.data_ptr(), Wi.data_ptr(), bi.data_ptr(), s.data_ptr(), B, out_f, in_f, act, out.data_ptr(), path, st), "fwd")
    torch.cuda.synchronize()
    dx = torch.full_like(x, 7.0); dW = torch.full_like(Wi, 7.0); db = torch.full_like(bi, 7.0)
    _lib.check(lib.dqrm_linear_bwd(x.data_ptr(), Wi.data_ptr(), s.data_ptr(), dout.data_ptr(), out.data_ptr(), B, out_f, in_f, act, dx.data_ptr(), None, None, 0, path, st), "dx")
    torch.cuda.synchronize()
    _lib.check(lib.dqrm_linear_bwd(x.data_ptr(), Wi.data_ptr(), s.data_ptr(), dout.data_ptr(), out.data_ptr(), B, out_f, in_f, act, None, dW.data_ptr(), db.data_ptr(), 0, path, st), "dw")
    torch.cuda.synchronize()
    return out, dx, dW, db

def ref(x, Wi, bi, s, dout, act, out_used):
    x64, W64, b64, s64, d64 = (t.double() for t in (x, Wi, bi, s, dout))
    z = (x64 @ W64.t() + b64) * s64
    out = torch.relu(z) if act == 1 else (torch.sigmoid(z) if act == 2 else z)
    o = out_used.double()
    gact = d64 * ((o > 0).double() if act == 1 else ((1 - o) * o if act == 2 else 1.0))
    g = gact * s64
    return out, g @ W64, (g.t() @ x64) / s64[:, None], g.sum(0) / s64

shapes = [(8, 8, 8, 0), (128, 64, 32, 0), (128, 64, 64, 0), (128, 128, 96, 1), (37, 40, 24, 1), (128, 512, 13, 1), (256, 256, 512, 1),
          (2048, 512, 512, 1), (1024, 512, 367, 1), (300, 1, 256, 2), (8192, 256, 512, 1)]
for (B, o, i, act) in shapes:
    x = torch.randn(B, i, device="cuda"); Wi = torch.randint(-8, 8, (o, i), device="cuda").float()
    bi = torch.randint(-8, 8, (o,), device="cuda").float(); s = torch.rand(o, device="cuda") * 0.05 + 0.005
    dout = torch.randn(B, o, device="cuda") * 0.01
    try:
        got = run(_lib.LINEAR_TC, x, Wi, bi, s, dout, act)
    except Exception as e:
        print((B, o, i, act), "EXC", e); break
    want = ref(x, Wi, bi, s, dout, act, got[0])
    msg = []
    for name, a, w in zip(("out", "dx", "dW", "db"), got, want):
        sc = float(w.abs().max()) + 1e-30
        d = (a.double() - w).abs()
        msg.append(f"{name}: err {float(d.max()) / sc:.2e} bad {float((d > 1e-4 * sc).double().mean()):.3f} zeros {float((a == 0).double().mean()):.2f} sevens {float((a == 7).double().mean()):.2f}")
    print((B, o, i, act), " | ".join(msg), flush=True)
    if B <= 8:
        for name, a, w in zip(("out", "dx", "dW"), got, want):
            print(name, "got\n", a[:8, :8].cpu().numpy().round(4), "\nwant\n", w[:8, :8].cpu().numpy().round(4))
