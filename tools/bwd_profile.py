"""Kernel-level times of the single-table fwd / bwd / pack / merge path at one sweep point (torch.profiler).
    python tools/bwd_profile.py --rows 10000000 --dim 64 --pooling 16 --batch 65536"""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from deep_quantized_recommendation_model_dqrm_b200 import synthetic, tables
p = argparse.ArgumentParser()
p.add_argument("--rows", type=int, default=10_000_000); p.add_argument("--dim", type=int, default=64)
p.add_argument("--pooling", type=int, default=16); p.add_argument("--batch", type=int, default=65536)
p.add_argument("--fused", action="store_true", help="dqrm_embbag_bwd_sgd instead of bwd + pack + merge_apply")
a = p.parse_args()
N, D, P, B = a.rows, a.dim, a.pooling, a.batch
W = torch.empty((N, D), device="cuda"); synthetic.table_weights_(W, 0, 99)
g = tables.EmbeddingTableGroup([W], embedding_bit=4)
gen = torch.Generator(device="cuda").manual_seed(5)
idx = torch.randint(0, N, (B * P,), device="cuda", generator=gen, dtype=torch.int64)
off = (torch.arange(B, device="cuda", dtype=torch.int64) * P).view(1, B)
dout = torch.randn((1, B, D), device="cuda", generator=gen) * 1e-4
out = torch.empty((1, B, D), device="cuda")
def step():
    g.scan_scales(); g.forward(idx, off, [0, B * P], B, out=out)
    if a.fused: g.backward_sgd(dout, 0.01)
    else: g.backward(dout, world=1); g.exchange(world=1, rank=0); g.merge_apply(0.01)
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for e in prof.events():
    if e.device_type.name != "CUDA": continue
    k = agg.setdefault(e.name[:90], [0, 0.0]); k[0] += 1; k[1] += e.device_time
print(f"# rows {N} dim {D} pooling {P} batch {B} lookups {B*P} unique {int(g.uniq_count[0])}")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{t / 5:10.1f} us/step {c / 5:5.1f}x  {n}")
if B * P > 16384:
    h = g._bwd_ws[:256].view(torch.int32).cpu().numpy().astype("int64") & 0xffffffff
    names = ["start", "keys", "pass0", "pass1", "pass2", "pass3", "segments", "fold", "long+scale"]
    prev = None
    for i, n in enumerate(names):
        v = int(h[16 + i])
        if v == 0: continue
        if prev is not None: print(f"   sort kernel phase {n:12s} {((v - prev) & 0xffffffff) / 1000.0:9.1f} us")
        prev = v
