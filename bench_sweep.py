#!/usr/bin/env python
"""BASELINE.json configs[4]: INT4 QAT EmbeddingBag fwd+bwd micro-benchmark sweep on ONE table, 1 B200,
against the CPU oracle (the reference's path restated, all host threads).

    python bench_sweep.py [--quick] [--cpu] [--out profiles/rNN_sweep.jsonl]

Per grid point (rows N, dim D, pooling P, batch B) it times, with CUDA events after warm-up:
  scan      dqrm_table_absmax_scale                          bytes N*D*4
  fwd       dqrm_embbag_fwd (scale given)                    bytes L(4D+8) + B(8+4D) + B*D (int8 codes)
  fwd_int4  dqrm_embbag_fwd_int4 on bit-packed tables        bytes L(D/2+8) + B(8+4D)
  bwd       dqrm_embbag_bwd + grad_pack + grad_merge_apply   bytes B*4D + L*8 + U*8D   (the data-parallel
            path at world 1: INT8 codes through the exchange slot)
  bwd_sgd   dqrm_embbag_bwd_sgd (sort + de-duplicate + SGD row update in place, the single-process path a5+a10)
                                                             bytes B*4D + L*8 + U*8D
and reports achieved GB/s on those ALGORITHMIC bytes (U = unique rows, counted on the device).
Tables are larger than the 126 MB L2 for N >= 4M (D=16) so no flush is needed there; for smaller tables a
256 MB buffer is written between iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def make_table(N, D):
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic, tables
    W = torch.empty((N, D), dtype=torch.float32, device="cuda")
    synthetic.table_weights_(W, 0, 99)
    return tables.EmbeddingTableGroup([W], embedding_bit=4)


def gpu_point(g, N, D, P, B, iters=5, flush=None):
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(5)
    idx = torch.randint(0, N, (B * P,), device=dev, generator=gen, dtype=torch.int64)
    off = (torch.arange(B, device=dev, dtype=torch.int64) * P).view(1, B)
    ib = [0, B * P]
    dout = torch.randn((1, B, D), device=dev, generator=gen) * 1e-6      # tiny gradients and learning rate: the update must not move the table maximum (the uniform init has a hard edge, so any visible update would change the scale and re-encode the INT4 shadow on every iteration)
    out = torch.empty((1, B, D), device=dev)
    L = B * P

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # every phase is captured in its own CUDA graph so that the timings are GPU time, not Python launch
    # overhead (the forward kernels run for ~5 us)
    def run_scan():
        g.scan_scales()

    def run_fwd():
        g.forward(idx, off, ib, B, out=out)

    def run_bwd():
        g.backward(dout, world=1); g.exchange(world=1, rank=0); g.merge_apply(1e-6)

    def run_bwd_sgd():
        g.backward_sgd(dout, 1e-6)

    def run_int4():
        g.forward_int4(idx, off, ib, B, out=out)

    def run_shadow():                                    # training forward on the packed-INT4 shadow rows (P == 1)
        g.forward(idx, off, ib, B, out=out)

    g.shadow = None
    g.scan_scales()
    g.pack_int4()
    phases = {"scan": run_scan, "fwd": run_fwd, "bwd": run_bwd, "bwd_sgd": run_bwd_sgd, "fwd_int4": run_int4}
    graphs = {}
    side = torch.cuda.Stream()

    def capture(name, fn):
        with torch.cuda.stream(side):
            for _ in range(2):
                fn()
        torch.cuda.synchronize()
        try:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                fn()
            graphs[name] = gr.replay
        except Exception:                                 # (a launch that cannot be captured: time it eagerly)
            torch.cuda.synchronize()
            graphs[name] = fn
    for name, fn in phases.items():
        capture(name, fn)
    if P == 1 and D % 16 == 0:
        g.enable_shadow()
        phases["fwd_shadow"] = run_shadow
        capture("fwd_shadow", run_shadow)
    torch.cuda.synchronize()
    t = {k: [] for k in phases}
    for it in range(iters + 2):
        for name in phases:
            if flush is not None:
                flush.add_(1.0)
            saved = g.shadow
            if name != "fwd_shadow":
                g.shadow = None                           # (eager fallbacks must take the fp32 path)
            a, b = ev(), ev()
            a.record(); graphs[name](); b.record()
            torch.cuda.synchronize()
            g.shadow = saved
            if it >= 2:
                t[name].append(a.elapsed_time(b))
    g.check_status()
    U = int(g.uniq_count[0].item())
    ms = {k: float(np.median(v)) for k, v in t.items()}
    by = {"scan": N * D * 4, "fwd": L * (4 * D + 8) + B * (8 + 4 * D) + B * D, "bwd": B * 4 * D + L * 8 + U * 8 * D}
    by["bwd_sgd"] = by["bwd"]
    by["fwd_int4"] = L * (D // 2 + 8) + B * (8 + 4 * D)
    if "fwd_shadow" in ms:
        by["fwd_shadow"] = L * (D // 2 + 8) + B * (8 + 4 * D) + B * D
    res = {"rows": N, "dim": D, "pooling": P, "batch": B, "lookups": L, "unique_rows": U, "ms": ms, "bytes": by,
           "GBps": {k: by[k] / (ms[k] * 1e-3) / 1e9 for k in ms},
           "fwd_bwd_GBps_with_scan": sum(by[k] for k in ("scan", "fwd", "bwd")) / (sum(ms[k] for k in ("scan", "fwd", "bwd")) * 1e-3) / 1e9,
           "fwd_bwd_GBps_gather_only": (by["fwd"] + by["bwd"]) / ((ms["fwd"] + ms["bwd"]) * 1e-3) / 1e9,
           "fwd_bwd_sgd_GBps_gather_only": (by["fwd"] + by["bwd_sgd"]) / ((ms["fwd"] + ms["bwd_sgd"]) * 1e-3) / 1e9}
    g.shadow = g._shadow_buf = None
    return res


def cpu_point(N, D, P, B):
    """Reference path on the host (oracle torch layer): scan + EmbeddingBag fwd + quant + bwd + coalesce +
    8-bit quantise + SGD row update, one iteration after one warm-up."""
    from oracle import dqrm_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(5)
    W = (torch.rand((N, D), generator=g) * 2 - 1) * float(np.sqrt(1 / N))
    E = O.OracleEmbeddingBag(N, D, 4, weight=W)
    idx = torch.randint(0, N, (B * P,), generator=g)
    off = torch.arange(B) * P
    dout = torch.randn((B, D), generator=g)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        y = E(idx, off)
        E.embedding_bag.weight.grad = None
        y.backward(dout)
        co = E.embedding_bag.weight.grad.coalesce()
        s = O.table_scale_torch(co.values(), 8).view(-1)
        q = O.quantize_torch(co.values(), 8, s)
        with torch.no_grad():
            E.embedding_bag.weight.data[co.indices()[0]] += -0.1 * (q * s.item())
        best = time.perf_counter() - t0
    return {"ms_total": best * 1e3, "cores": torch.get_num_threads()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle on the points with rows <= 4M")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    a = ap.parse_args()
    if a.quick:
        grid = [(1_000_000, 16, 1, 65536), (10_000_000, 16, 1, 65536), (10_000_000, 64, 16, 8192), (40_000_000, 128, 4, 8192)]
    else:      # the full BASELINE configs[4] grid: 4 x 3 x 4 x 3 = 144 points
        grid = [(N, D, P, B) for N in (1_000_000, 4_000_000, 10_000_000, 40_000_000) for D in (16, 64, 128)
                for P in (1, 4, 16, 64) for B in (1024, 8192, 65536)]
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")          # 256 MB > L2
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    cur, g = None, None
    with open(a.out, "w") as f:
        for (N, D, P, B) in grid:
            if cur != (N, D):
                g = None
                torch.cuda.empty_cache()
                g = make_table(N, D)
                cur = (N, D)
            r = gpu_point(g, N, D, P, B, flush=flush if N * D * 4 < 512 * 1024 * 1024 else None)
            if a.cpu and N <= 1_000_000 and B <= 8192 and P in (1, 16):      # a bounded CPU sample (seconds each)
                r["cpu"] = cpu_point(N, D, P, B)
                r["speedup_vs_cpu"] = r["cpu"]["ms_total"] / sum(r["ms"][k] for k in ("scan", "fwd", "bwd"))
            f.write(json.dumps(r) + "\n")
            f.flush()
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
