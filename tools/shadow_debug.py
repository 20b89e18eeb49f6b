import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_quantized_recommendation_model_dqrm_b200 import synthetic, tables
N, D, B = 10_000_000, 64, 65536
W = torch.empty((N, D), device="cuda"); synthetic.table_weights_(W, 0, 99)
g = tables.EmbeddingTableGroup([W], embedding_bit=4)
idx = torch.randint(0, N, (B,), device="cuda"); off = torch.arange(B, device="cuda").view(1, B)
out = torch.empty((1, B, D), device="cuda")
g.scan_scales(); g.enable_shadow()
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
print("refresh us", t(g._shadow_refresh), "flags", g._shadow_flags.tolist(), "scale", g.scale.tolist(), g.shadow_scale.tolist())
print("fwd shadow us", t(lambda: g.forward(idx, off, [0, B], B, out=out)))
sh = g.shadow; g.shadow = None
print("fwd fp32 us", t(lambda: g.forward(idx, off, [0, B], B, out=out)))
g.shadow = sh
print("scan us", t(g.scan_scales))
print("scan+fwd shadow us", t(lambda: (g.scan_scales(), g.forward(idx, off, [0, B], B, out=out))), "flags", g._shadow_flags.tolist())
