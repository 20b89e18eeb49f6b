"""Times of the fused QuantLinear kernels alone (CUDA events), tensor-core vs FFMA path.
    python tools/tc_bench.py [--shapes 8192x512x512,2048x512x512] [--iters 20] [--paths tc,ffma]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_quantized_recommendation_model_dqrm_b200 import _lib
p = argparse.ArgumentParser()
p.add_argument("--shapes", default="8192x512x512,8192x512x415,8192x256x512,2048x512x512,2048x512x367,1024x512x512,512x512x512,256x512x512")
p.add_argument("--iters", type=int, default=20); p.add_argument("--paths", default="tc,ffma"); p.add_argument("--modes", default="fwd,dx,dw")
a = p.parse_args()
lib = _lib.load(); st = _lib.stream_ptr()
for shp in a.shapes.split(","):
    B, o, i = (int(v) for v in shp.split("x"))
    x = torch.randn(B, i, device="cuda"); Wi = torch.randint(-8, 8, (o, i), device="cuda").float()
    bi = torch.randint(-8, 8, (o,), device="cuda").float(); s = torch.rand(o, device="cuda") * 0.05 + 0.005
    dout = torch.randn(B, o, device="cuda") * 0.01; out = torch.empty(B, o, device="cuda")
    dx = torch.empty_like(x); dW = torch.empty_like(Wi); db = torch.empty_like(bi)
    for path in a.paths.split(","):
        pid = {"tc": _lib.LINEAR_TC, "ffma": _lib.LINEAR_FFMA}[path]
        calls = {"fwd": lambda: lib.dqrm_linear_fwd(x.data_ptr(), Wi.data_ptr(), bi.data_ptr(), s.data_ptr(), B, o, i, 1, out.data_ptr(), pid, st),
                 "dx": lambda: lib.dqrm_linear_bwd(x.data_ptr(), Wi.data_ptr(), s.data_ptr(), dout.data_ptr(), out.data_ptr(), B, o, i, 1, dx.data_ptr(), None, None, 0, pid, st),
                 "dw": lambda: lib.dqrm_linear_bwd(x.data_ptr(), Wi.data_ptr(), s.data_ptr(), dout.data_ptr(), out.data_ptr(), B, o, i, 1, None, dW.data_ptr(), db.data_ptr(), 0, pid, st)}
        res = []
        for m in a.modes.split(","):
            f = calls[m]
            for _ in range(3): f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters): f()
            e1.record(); torch.cuda.synchronize()
            us = 1000 * e0.elapsed_time(e1) / a.iters
            res.append(f"{m} {us:7.1f} us {2 * B * o * i / us / 1e6:7.1f} TFLOP/s")
        print(f"{shp:>14} {path:>4}: " + " | ".join(res), flush=True)
