"""CPU oracle for the DQRM data-parallel hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the reference
(YangZhou08/Deep_Quantized_Recommendation_Model_DQRM) for the one path this
repository accelerates.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under
``deep_quantized_recommendation_model_dqrm_b200/`` imports it, and the product
path raises when its CUDA library is missing instead of falling back to this.

Parity pinning: the reference has no tests and no golden vectors (SURVEY.md
§4, §8c).  The oracle is therefore pinned by *executing the reference itself*
in the build container: ``oracle/make_golden.py`` imports the reference
modules from ``/root/reference`` (with the ``.cuda()`` shim that
``quant_utils.py:336`` needs on a CUDA-less host), runs them on seeded
inputs, and commits inputs + outputs under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function below against those
vectors.  The arithmetic that lives in the reference's third-party dependency
(PyTorch: ``nn.EmbeddingBag`` fwd/bwd, ``Tensor.coalesce``, dense += sparse,
``torch.round/clamp``; version unpinned by the reference, ``requirements.txt:5``)
is pinned to torch 2.11.0 CPU behaviour, the version in this image.

Each function cites the reference ``file:line`` it follows.  Abbreviations:
  qu    = quantization_supp/quant_utils.py
  qm    = quantization_supp/quant_modules_not_quantize_grad.py
  sgd   = sgd_quantized_gradients_parallel_comm.py
  drv   = dlrm_s_pytorch_comm_grad.py

Two layers are provided:
  * "spec" functions (numpy, explicit evaluation order) -- the bit-level
    statement of what the CUDA kernels must reproduce;
  * "torch" functions -- the same op sequence the reference issues through
    torch on CPU (used for whole-step checks and as the timed CPU baseline).
``tests/test_oracle_golden.py`` checks both layers against each other and
against the golden vectors.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

F32 = np.float32

# --------------------------------------------------------------------------
# (a16) batch shard                                              drv:993-997
# --------------------------------------------------------------------------

def get_my_slice(n: int, my_size: int, my_rank: int) -> slice:
    """Contiguous balanced shard of ``n`` items (drv:993-997; same as
    extend_distributed.py:47-51)."""
    k, m = divmod(n, my_size)
    return slice(my_rank * k + min(my_rank, m), (my_rank + 1) * k + min(my_rank + 1, m), 1)


# --------------------------------------------------------------------------
# (a1) per-table symmetric scale                                 qu:141-194
# --------------------------------------------------------------------------

def qrange(bits: int) -> int:
    """n = 2^(bits-1) - 1 (qu:189, qu:326)."""
    return 2 ** (bits - 1) - 1


def scale_from_absmax(absmax, bits: int) -> np.float32:
    """s = clamp(absmax, min=1e-8) / n, IEEE fp32 (qu:191-192)."""
    m = np.maximum(F32(absmax), F32(1e-8))
    return F32(m / F32(qrange(bits)))


def table_scale_spec(W: np.ndarray, bits: int) -> np.float32:
    """max(|min W|, |max W|) over the whole table, then clamp / n (qu:177-192)."""
    W = np.asarray(W, dtype=F32)
    w_min = W.min()
    w_max = W.max()
    return scale_from_absmax(max(abs(w_min), abs(w_max)), bits)


def table_scale_torch(W: torch.Tensor, bits: int) -> torch.Tensor:
    """Same op sequence as the reference issues (qu:177-192): two column
    reductions followed by two tiny ones, a Python ``max`` of two 0-dim
    tensors, clamp, divide.  Returns a 0-dim fp32 tensor."""
    with torch.no_grad():
        w_min, _ = torch.min(torch.min(W, dim=0).values, dim=0)
        w_max, _ = torch.max(torch.max(W, dim=0).values, dim=0)
        n = qrange(bits)
        scale = max(w_min.abs(), w_max.abs())
        scale = torch.clamp(scale, min=1e-8) / n
    return scale


# --------------------------------------------------------------------------
# (a4) symmetric fake-quantisation                      qu:75-101, 316-363
# --------------------------------------------------------------------------

def inv_scale(scale) -> np.ndarray:
    """``1. / scale`` evaluated in fp32 (qu:101)."""
    return (F32(1.0) / np.asarray(scale, dtype=F32)).astype(F32)


def quantize_spec(x: np.ndarray, bits: int, scale) -> np.ndarray:
    """clamp(round((1/s) * x + 0), -n-1, n) as integer-valued fp32.

    ``torch.round`` is round-half-to-even (= np.rint); the reciprocal is
    formed first and then multiplied (qu:101), the clamp comes after the
    round (qu:343).  ``scale`` is a scalar or broadcasts over rows
    (``scale.view(-1, 1)``, qu:90-93)."""
    x = np.asarray(x, dtype=F32)
    n = qrange(bits)
    inv = inv_scale(scale)
    if inv.ndim == 1 and x.ndim == 2 and inv.shape[0] != 1:
        inv = inv.reshape(-1, 1)
    q = np.rint((inv * x).astype(F32) + F32(0.0)).astype(F32)
    return np.clip(q, F32(-n - 1), F32(n)).astype(F32)


def quantize_torch(x: torch.Tensor, bits: int, scale: torch.Tensor) -> torch.Tensor:
    """SymmetricQuantFunction.forward without the autograd wrapper (qu:322-346)."""
    n = qrange(bits)
    zero_point = torch.tensor(0.0)
    if x.dim() == 2:
        if scale.dim() != 1 or scale.shape[0] != 1:
            scale = scale.view(-1, 1)
        zero_point = zero_point.view(-1, 1)
    else:
        scale = scale.view(-1)
        zero_point = zero_point.view(-1)
    q = torch.round(1.0 / scale * x + zero_point)
    return torch.clamp(q, -n - 1, n)


class SymmetricQuantSTE(torch.autograd.Function):
    """SymmetricQuantFunction with its straight-through backward
    ``grad / scale`` and no clipping mask (qu:348-363)."""

    @staticmethod
    def forward(ctx, x, k, scale):
        ctx.scale = scale
        return quantize_torch(x, k, scale)

    @staticmethod
    def backward(ctx, grad_output):
        scale = ctx.scale
        if grad_output.dim() == 2:
            scale = scale.view(-1, 1)
        else:
            scale = scale.view(-1)
        return grad_output / scale, None, None


# --------------------------------------------------------------------------
# (a3) QuantEmbeddingBagTwo.forward                               qm:317-395
# --------------------------------------------------------------------------

def bag_bounds(offsets: np.ndarray, L: int):
    """Bag b covers lookups [offsets[b], offsets[b+1]) with the last bag ending
    at L (``nn.EmbeddingBag`` with include_last_offset=False, qm:288,367)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    ends = np.concatenate([offsets[1:], np.array([L], dtype=np.int64)])
    return offsets, ends


def pool_sum_spec(W: np.ndarray, idx: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """Sum pooling as a left fold over each bag in index order -- the order
    ATen's EmbeddingBag uses on CPU and CUDA (qm:367)."""
    W = np.asarray(W, dtype=F32)
    idx = np.asarray(idx, dtype=np.int64)
    starts, ends = bag_bounds(offsets, idx.shape[0])
    lens = ends - starts
    B, D = starts.shape[0], W.shape[1]
    acc = np.zeros((B, D), dtype=F32)
    for p in range(int(lens.max()) if B else 0):
        live = np.nonzero(lens > p)[0]
        rows = W[idx[starts[live] + p]]
        if p == 0:
            acc[live] = rows
        else:
            acc[live] = (acc[live] + rows).astype(F32)
    return acc


def embbag_forward_spec(W, idx, offsets, bits: int, scale=None, full_precision=False):
    """(scale, codes, out): scale scan (a1), fp32 pooling, quantise the POOLED
    vector, dequantise (qm:337,367,378,393).  ``full_precision`` returns the
    pooled fp32 vector (qm:395)."""
    if scale is None:
        scale = table_scale_spec(W, bits)
    pooled = pool_sum_spec(W, idx, offsets)
    if full_precision:
        return scale, None, pooled
    codes = quantize_spec(pooled, bits, scale)
    out = (codes * F32(scale)).astype(F32)
    return F32(scale), codes, out


def embbag_backward_spec(dout: np.ndarray, idx: np.ndarray, offsets: np.ndarray, scale,
                         full_precision=False):
    """Uncoalesced sparse row gradient of (a3): ``dy = (g * s) / s`` (autograd
    of qm:393 then qu:363), one value row per lookup (ATen sparse EmbeddingBag
    backward, triggered at drv:1938).  Returns (rows[L], values[L, D])."""
    dout = np.asarray(dout, dtype=F32)
    idx = np.asarray(idx, dtype=np.int64)
    starts, ends = bag_bounds(offsets, idx.shape[0])
    bag_of = np.repeat(np.arange(starts.shape[0]), ends - starts)
    if full_precision:
        dy = dout
    else:
        s = F32(scale)
        dy = ((dout * s).astype(F32) / s).astype(F32)
    return idx.copy(), dy[bag_of]


# --------------------------------------------------------------------------
# packed INT4 export / serving forward
# --------------------------------------------------------------------------

def pack_int4_spec(W: np.ndarray, scale) -> np.ndarray:
    """[N, D] fp32 -> [N, D/2] bytes: code = clamp(rint((1/s) w), -8, 7) (qu:101,343); element d in byte d/2,
    low nibble for even d."""
    q = quantize_spec(W, 4, scale).astype(np.int8)
    n = (q & 0x0F).astype(np.uint8)
    return (n[:, 0::2] | (n[:, 1::2] << 4)).astype(np.uint8)


def unpack_int4_spec(packed: np.ndarray) -> np.ndarray:
    lo = (packed & 0x0F).astype(np.int8)
    hi = (packed >> 4).astype(np.int8)
    lo = np.where(lo > 7, lo - 16, lo)
    hi = np.where(hi > 7, hi - 16, hi)
    out = np.empty((packed.shape[0], packed.shape[1] * 2), dtype=np.int32)
    out[:, 0::2], out[:, 1::2] = lo, hi
    return out


def embbag_forward_int4_spec(packed: np.ndarray, idx, offsets, scale) -> np.ndarray:
    """out[b] = s * sum_l code(row_l): exact integer pooling of the stored codes, one fp32 multiply."""
    codes = unpack_int4_spec(packed)
    idx = np.asarray(idx, dtype=np.int64)
    starts, ends = bag_bounds(offsets, idx.shape[0])
    out = np.zeros((starts.shape[0], codes.shape[1]), dtype=np.int64)
    for b, (a, e) in enumerate(zip(starts, ends)):
        if e > a:
            out[b] = codes[idx[a:e]].sum(axis=0)
    return (out.astype(F32) * F32(scale)).astype(F32)


# --------------------------------------------------------------------------
# (a7 step 1) coalesce                                              sgd:859
# --------------------------------------------------------------------------

FOLD_BLOCK = 64      # include/dqrm_b200.h DQRM_FOLD_BLOCK


def coalesce_spec(rows: np.ndarray, values: np.ndarray, order: np.ndarray | None = None, block: int | None = FOLD_BLOCK):
    """``Tensor.coalesce``: sorted-ascending unique rows; duplicates of a row
    summed as a left fold.  Returns (uniq_rows[U] int64, sums[U, D] fp32).

    Fold order: torch 2.11's CPU coalesce folds duplicates in whatever order its
    (unstable) ``sort`` leaves them -- re-folding by ``torch.sort``'s permutation
    reproduces ``coalesce()`` bit-for-bit (tests/test_oracle_golden.py), and the
    CUDA coalesce uses yet another order -- so the duplicate order is NOT a
    reference contract.  The spec fixes ORIGINAL-OCCURRENCE order (stable sort),
    which is what the CUDA kernel implements; ``order`` lets a test pass
    torch's permutation to reproduce the reference's bits.  Row sets are exact
    either way; sums agree to fp32 rounding (north_star: 1e-5 relative).

    ``block``: a row with more than ``block`` duplicates is folded in blocks of
    ``block`` consecutive occurrences (each a left fold), then the block sums are
    folded left to right -- the shape the CUDA kernel uses so that thousands of
    duplicates of one row (3-row tables at batch 8192) do not form one serial
    chain.  Identical to the plain left fold for rows of <= ``block`` duplicates;
    ``block=None`` is the plain left fold for any length."""
    rows = np.asarray(rows, dtype=np.int64)
    values = np.asarray(values, dtype=F32)
    if order is None:
        order = np.argsort(rows, kind="stable")
    srows = rows[order]
    head = np.ones(srows.shape[0], dtype=bool)
    head[1:] = srows[1:] != srows[:-1]
    seg_start = np.nonzero(head)[0]
    seg_end = np.concatenate([seg_start[1:], [srows.shape[0]]])
    seg_len = seg_end - seg_start
    uniq = srows[seg_start]
    sums = values[order[seg_start]].copy()
    short = np.ones(seg_len.shape[0], dtype=bool) if block is None else seg_len <= block
    for r in range(1, int(seg_len[short].max()) if short.any() else 0):
        live = np.nonzero(short & (seg_len > r))[0]
        sums[live] = (sums[live] + values[order[seg_start[live] + r]]).astype(F32)
    for j in np.nonzero(~short)[0]:                                 # long rows: blocked fold
        occ = values[order[seg_start[j]:seg_end[j]]]
        total = None
        for b0 in range(0, occ.shape[0], block):
            part = occ[b0].copy()
            for v in occ[b0 + 1:b0 + block]:
                part = (part + v).astype(F32)
            total = part if total is None else (total + part).astype(F32)
        sums[j] = total
    return uniq, sums


# --------------------------------------------------------------------------
# (a7 steps 2-5) quantised sparse exchange, (a9) update    sgd:850-890, 601-628
# --------------------------------------------------------------------------

def grad_scale_spec(sums: np.ndarray, bits: int) -> np.float32:
    """Per-rank gradient scale: (a1) applied to the coalesced values (sgd:861)."""
    return table_scale_spec(sums, bits)


def mean_scale_spec(scales, world: int) -> np.float32:
    """``all_reduce(SUM)`` then ``mul_(1. / N)`` (sgd:865-866).  Summed in rank
    order 0..N-1 (Gloo's order is backend-defined; exact for N=2)."""
    acc = F32(scales[0])
    for s in scales[1:]:
        acc = F32(acc + F32(s))
    return F32(acc * F32(1.0 / world))


def exchange_emb_grad_spec(per_rank, bits: int, num_rows: int, s_bar=None):
    """Reference quantize_emb_grad(parallel=True) over N ranks (sgd:850-890).

    per_rank: list of (uniq_rows, sums) from coalesce_spec, one per rank.
    Returns dict(s_bar, codes=[per-rank codes], union_rows, qbar) where
    ``qbar = (sum_r q_r) * (1/N)`` on the union of rows (sgd:878,885)."""
    world = len(per_rank)
    s_local = [grad_scale_spec(s, bits) for _, s in per_rank]
    if s_bar is None:          # (tests may pass the reference's own mean: Gloo's SUM order at N > 2 is backend-internal)
        s_bar = mean_scale_spec(s_local, world)
    codes = [quantize_spec(s, bits, s_bar) for _, s in per_rank]
    all_rows = np.concatenate([r for r, _ in per_rank])
    all_codes = np.concatenate(codes, axis=0)
    union, qsum = coalesce_spec(all_rows, all_codes)   # integer-valued: any order is exact
    qbar = (qsum * F32(1.0 / world)).astype(F32)
    assert union.size == 0 or (union.min() >= 0 and union.max() < num_rows)
    return dict(s_local=s_local, s_bar=s_bar, codes=codes, union_rows=union, qbar=qbar)


def exchange_emb_grad_unquantized_spec(per_rank):
    """emb_grad_quantized=False across N ranks (sgd:319-329): coalesce, sparse all-reduce(SUM) -- coinciding
    rows summed in rank order --, ``mul_(1./N)``.  Returns (union_rows, gmean)."""
    world = len(per_rank)
    union = np.unique(np.concatenate([r for r, _ in per_rank]))
    D = per_rank[0][1].shape[1]
    acc = np.zeros((union.shape[0], D), dtype=F32)
    seen = np.zeros(union.shape[0], dtype=bool)
    for rows, sums in per_rank:
        pos = np.searchsorted(union, rows)
        first = ~seen[pos]
        acc[pos[first]] = sums[first]
        acc[pos[~first]] = (acc[pos[~first]] + sums[~first]).astype(F32)
        seen[pos] = True
    return union, (acc * F32(1.0 / world)).astype(F32)


def weight_update_emb_spec(W: np.ndarray, union_rows, qbar, s_bar, lr: float) -> None:
    """W[row] += (-lr) * (qbar * s_bar), each product rounded to fp32, in that
    association (sgd:618,622).  In place."""
    upd = (qbar * F32(s_bar)).astype(F32)
    upd = (F32(-lr) * upd).astype(F32)
    W[union_rows] = (W[union_rows] + upd).astype(F32)


def weight_update_emb_unquantized_spec(W, union_rows, gmean, lr: float) -> None:
    """``W.add_(-lr * grad)`` for emb_grad_quantized=False (sgd:626)."""
    W[union_rows] = (W[union_rows] + (F32(-lr) * gmean).astype(F32)).astype(F32)


def sgd_sparse_spec(W: np.ndarray, rows, values, lr: float) -> None:
    """(a10) ``torch.optim.SGD.step`` on an UNCOALESCED sparse grad:
    ``W.add_(grad, alpha=-lr)`` applies duplicates one at a time in storage
    order on CPU.  In place."""
    a = F32(-lr)
    for r, v in zip(np.asarray(rows), np.asarray(values, dtype=F32)):
        W[r] = (W[r] + (v * a).astype(F32)).astype(F32)


def rwsadagrad_rows_spec(W: np.ndarray, momentum: np.ndarray, rows, sums, lr: float, eps: float = 1e-10) -> None:
    """(f-2) row-wise sparse Adagrad on COALESCED row gradients, optim/rwsadagrad.py:97-113:
    ``m[row] += mean_d(g^2)``; ``std = sqrt(m[row]) + eps``; ``W[row] += (-clr) * (g / std)``.  In place.
    (The mean over the row is a float32 sum / D; torch's reduction order inside a row is not a contract, so the
    golden check allows fp32 rounding.)"""
    rows = np.asarray(rows)
    g = np.asarray(sums, dtype=F32)
    D = F32(g.shape[1])
    sq = (g * g).astype(F32)
    acc = np.zeros(g.shape[0], dtype=F32)
    for d in range(g.shape[1]):
        acc = (acc + sq[:, d]).astype(F32)
    momentum[rows] = (momentum[rows] + (acc / D).astype(F32)).astype(F32)
    std = (np.sqrt(momentum[rows]).astype(F32) + F32(eps)).astype(F32)
    W[rows] = (W[rows] + (F32(-lr) * (g / std[:, None]).astype(F32)).astype(F32)).astype(F32)


def topk_rows_spec(uniq_rows, sums, k: int):
    """(a8) north-star extension, no DQRM reference semantics (parity
    unpinned).  Score = ||g_row||^2 / D as in the only top-k in the tree
    (training_imagenet_speedup.py:138,149); keep the k largest, ties broken by
    lower row id; result returned in ascending row order."""
    sums = np.asarray(sums, dtype=F32)
    D = sums.shape[1]
    sq = np.zeros(sums.shape[0], dtype=F32)
    for d in range(D):                      # left fold over the row, fp32
        sq = (sq + (sums[:, d] * sums[:, d]).astype(F32)).astype(F32)
    score = (sq / F32(D)).astype(F32)
    if k >= sums.shape[0]:
        keep = np.arange(sums.shape[0])
    else:
        order = np.lexsort((np.asarray(uniq_rows), -score.astype(np.float64)))
        keep = np.sort(order[:k])
    return np.asarray(uniq_rows)[keep], sums[keep], score


# --------------------------------------------------------------------------
# (a14) dot interaction                                          drv:701-806
# --------------------------------------------------------------------------

def tril_pairs(ni: int, itself: bool = False):
    """Index lists of drv:719-722."""
    offset = 1 if itself else 0
    li = [i for i in range(ni) for j in range(i + offset)]
    lj = [j for i in range(ni) for j in range(i + offset)]
    return li, lj


def interact_features_torch(x: torch.Tensor, ly, itself: bool = False) -> torch.Tensor:
    """cat -> bmm(T, T^T) -> strict lower triangle -> cat (drv:706-725)."""
    (batch_size, d) = x.shape
    T = torch.cat([x] + list(ly), dim=1).view((batch_size, -1, d))
    Z = torch.bmm(T, torch.transpose(T, 1, 2))
    _, ni, nj = Z.shape
    li, lj = tril_pairs(ni, itself)
    Zflat = Z[:, torch.tensor(li), torch.tensor(lj)]
    return torch.cat([x] + [Zflat], dim=1)


# --------------------------------------------------------------------------
# (a15) QuantLinear.forward                                      qm:105-211
# --------------------------------------------------------------------------

def linear_scale_spec(W: np.ndarray, bits: int) -> np.ndarray:
    """Per-output-channel scale: max(|min_row|, |max_row|) clamp / n
    (qm:125-126, qu:213-215)."""
    W = np.asarray(W, dtype=F32)
    m = np.maximum(np.abs(W.min(axis=1)), np.abs(W.max(axis=1))).astype(F32)
    return (np.maximum(m, F32(1e-8)) / F32(qrange(bits))).astype(F32)


def linear_fakequant_spec(W, b, bits: int):
    """(W_int, b_int, s_row): the bias is quantised to ``bits`` with the
    WEIGHT's per-channel scale (qm:143-154; bias_bit == weight_bit, drv:318-319)."""
    s = linear_scale_spec(W, bits)
    W_int = quantize_spec(W, bits, s)
    n = qrange(bits)
    inv = inv_scale(s)
    b_int = np.clip(np.rint((inv * np.asarray(b, dtype=F32)).astype(F32)), -n - 1, n).astype(F32)
    return W_int, b_int, s


def quant_linear_forward_torch(x, weight, bias, bits: int, full_precision=False):
    """QuantLinear.forward with per_channel=True, quantize_activation=False
    (qm:105-211).  Autograd-capable (STE).  Returns (y, None)."""
    if full_precision:
        return F.linear(x, weight, bias), None
    w = weight.data.detach()
    w_min, _ = torch.min(w, dim=1)
    w_max, _ = torch.max(w, dim=1)
    n = qrange(bits)
    s, _ = torch.max(torch.stack([w_min.abs(), w_max.abs()], dim=1), dim=1)
    s = torch.clamp(s, min=1e-8) / n
    w_int = SymmetricQuantSTE.apply(weight, bits, s)
    b_int = SymmetricQuantSTE.apply(bias, bits, s.view(1, -1))
    return F.linear(x, w_int, b_int) * s.view(1, -1), None


# --------------------------------------------------------------------------
# (a11) MLP gradient quantisation + update             sgd:892-961, 642-663
# --------------------------------------------------------------------------

def linear_grad_scale_spec(G: np.ndarray, bits: int = 8) -> np.ndarray:
    """Per-row scale of a weight gradient (sgd:905-910 -> qu:213-215)."""
    return linear_scale_spec(G, bits)


def bias_grad_scale_spec(g: np.ndarray, bits: int = 8) -> np.float32:
    """Scalar scale of a bias gradient (sgd:945-947 -> qu:217-218)."""
    g = np.asarray(g, dtype=F32)
    return scale_from_absmax(max(abs(g.min()), abs(g.max())), bits)


def exchange_dense_grad_spec(per_rank_grads, per_rank_scales, bits: int = 8, s_bar=None):
    """scale all-reduce-mean, quantise with the mean scale, code all-reduce,
    ``* 1/N`` (sgd:912-924, 948-956).  Returns (s_bar, qbar)."""
    world = len(per_rank_grads)
    if s_bar is None:
        acc = np.asarray(per_rank_scales[0], dtype=F32).copy()
        for s in per_rank_scales[1:]:
            acc = (acc + np.asarray(s, dtype=F32)).astype(F32)
        s_bar = (acc * F32(1.0 / world)).astype(F32)
    s_bar = np.asarray(s_bar, dtype=F32)
    qsum = None
    for g in per_rank_grads:
        g = np.asarray(g, dtype=F32)
        if g.ndim == 2:
            q = quantize_spec(g, bits, s_bar)
        else:
            n = qrange(bits)
            q = np.clip(np.rint((inv_scale(s_bar) * g).astype(F32)), -n - 1, n).astype(F32)
        qsum = q if qsum is None else (qsum + q).astype(F32)
    return s_bar, (qsum * F32(1.0 / world)).astype(F32)


def exchange_dense_grad_ec_spec(per_rank_grads, per_rank_ec, bits: int = 8):
    """quantize_linear_grad / quantize_bias_grad with err_compensation=True (sgd:899-900,926-927,938-939,958-959):
    ``w_r = grad_r + ec_r``; scale from w_r; mean scale; q_r = Q(w_r); qbar = (sum_r q_r)/N;
    ``ec_r' = w_r - qbar * s_bar`` (the GLOBAL averaged update is subtracted from the LOCAL compensated gradient).
    Returns (s_bar, qbar, [ec_r'])."""
    ws = [(np.asarray(g, dtype=F32) + np.asarray(e, dtype=F32)).astype(F32) for g, e in zip(per_rank_grads, per_rank_ec)]
    if ws[0].ndim == 2:
        scales = [linear_grad_scale_spec(w, bits) for w in ws]
    else:
        scales = [bias_grad_scale_spec(w, bits) for w in ws]
    s_bar, qbar = exchange_dense_grad_spec(ws, scales, bits)
    sb = np.asarray(s_bar, dtype=F32).reshape(-1, 1) if ws[0].ndim == 2 else F32(s_bar)
    new_ec = [(w - (qbar * sb).astype(F32)).astype(F32) for w in ws]
    return s_bar, qbar, new_ec


def weight_update_linear_spec(W, b, qbar_w, s_w, qbar_b, s_b, lr: float) -> None:
    """``W.add_(-lr * grad * s.view(-1,1))``; ``b.add_(-lr * grad * s)``
    (sgd:642-643): ((-lr) * q) * s, left to right.  In place."""
    a = F32(-lr)
    W += ((a * qbar_w).astype(F32) * np.asarray(s_w, dtype=F32).reshape(-1, 1)).astype(F32)
    b += ((a * qbar_b).astype(F32) * F32(s_b)).astype(F32)


# --------------------------------------------------------------------------
# Whole-model restatement (torch CPU): DLRM_Net flow with QAT modules and the
# quantised-gradient DP "optimizer".                    drv:278-966, sgd:257-685
# --------------------------------------------------------------------------

class OracleEmbeddingBag(torch.nn.Module):
    """QuantEmbeddingBagTwo (qm:220-398) restated on torch CPU ops."""

    def __init__(self, num_embeddings, embedding_dim, embedding_bit=4, weight=None, embedding_id=None):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.embedding_bit = embedding_bit
        self.embedding_id = embedding_id
        self.full_precision_flag = False
        self.embedding_bag = torch.nn.EmbeddingBag(num_embeddings, embedding_dim, mode="sum", sparse=True)
        if weight is None:
            W = np.random.uniform(low=-np.sqrt(1 / num_embeddings), high=np.sqrt(1 / num_embeddings),
                                  size=(num_embeddings, embedding_dim)).astype(np.float32)   # qm:273-275
            weight = torch.tensor(W)
        self.embedding_bag.weight.data = weight.clone().requires_grad_(True)
        self.eb_scaling_factor = torch.zeros(128, 1)     # qm:258 (replaced by a 0-dim tensor on first use)
        self.emb_scaling_factor = torch.zeros(1)         # qm:279
        self.output_integer = None

    def forward(self, input, offsets=None, per_sample_weights=None, full_precision_flag=False, test_mode=False):
        fp = full_precision_flag or self.full_precision_flag
        if (not fp and not test_mode) or self.eb_scaling_factor.shape == (128, 1):   # qm:331
            self.eb_scaling_factor = table_scale_torch(self.embedding_bag.weight.data, self.embedding_bit)
        y = self.embedding_bag(input, offsets, per_sample_weights=None)               # qm:367
        if fp:
            self.output_integer = y
            return y
        self.output_integer = SymmetricQuantSTE.apply(y, self.embedding_bit, self.eb_scaling_factor)
        return self.output_integer * self.eb_scaling_factor                           # qm:393


class OracleQuantLinear(torch.nn.Module):
    """QuantLinear (qm:20-211), per_channel=True path."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, weight_bit=4, full_precision_flag=False):
        super().__init__()
        self.weight = torch.nn.Parameter(weight.clone())
        self.bias = torch.nn.Parameter(bias.clone())
        self.weight_bit = weight_bit
        self.full_precision_flag = full_precision_flag
        self.weight_scaling_factor = torch.zeros(weight.shape[0])
        self.bias_scaling_factor = torch.zeros(weight.shape[0])

    def forward(self, x):
        return quant_linear_forward_torch(x, self.weight, self.bias, self.weight_bit, self.full_precision_flag)[0]


def make_mlp_params(ln, rng: np.random.RandomState):
    """MLP init of drv:288-296: W ~ N(0, sqrt(2/(m+n))), b ~ N(0, sqrt(1/m))."""
    params = []
    for i in range(len(ln) - 1):
        n, m = int(ln[i]), int(ln[i + 1])
        W = rng.normal(0.0, np.sqrt(2 / (m + n)), size=(m, n)).astype(np.float32)
        b = rng.normal(0.0, np.sqrt(1 / m), size=m).astype(np.float32)
        params.append((W, b))
    return params


class OracleDLRM(torch.nn.Module):
    """DLRM_Net (drv:278-966) restricted to the hot configuration:
    arch_interaction_op="dot", quantization_flag, quantize_activation=False
    (``--linear_channel``), sigmoid on the last top layer, BCE-mean loss."""

    def __init__(self, ln_emb, m_spa, bot_params, top_params, embedding_bit=4, weight_bit=4,
                 quantize_mlp=True, emb_weights=None):
        super().__init__()
        self.emb_l = torch.nn.ModuleList([
            OracleEmbeddingBag(int(n), m_spa, embedding_bit, embedding_id=i,
                               weight=None if emb_weights is None else emb_weights[i])
            for i, n in enumerate(ln_emb)])
        mk = lambda ps: torch.nn.ModuleList([
            OracleQuantLinear(torch.tensor(W), torch.tensor(b), weight_bit, not quantize_mlp) for W, b in ps])
        self.bot_l = mk(bot_params)
        self.top_l = mk(top_params)

    def apply_mlp(self, x, layers, sigmoid_last):
        for i, layer in enumerate(layers):
            x = layer(x)
            x = torch.sigmoid(x) if (sigmoid_last and i == len(layers) - 1) else torch.relu(x)
        return x

    def apply_emb(self, lS_o, lS_i, test_mode=False):           # drv:614-679
        return [E(lS_i[k], lS_o[k], test_mode=test_mode) for k, E in enumerate(self.emb_l)]

    def forward(self, dense_x, lS_o, lS_i, test_mode=False):    # drv:855-859
        x = self.apply_mlp(dense_x, self.bot_l, False)
        ly = self.apply_emb(lS_o, lS_i, test_mode)
        z = interact_features_torch(x, ly)
        return self.apply_mlp(z, self.top_l, True)


def clear_gradients_torch(model) -> None:
    """sgd:714-734."""
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.grad is not None:
                p.grad.requires_grad_(False)
                p.grad.zero_()


def _allreduce_mean_scale(local_scales, world):
    acc = local_scales[0].clone()
    for s in local_scales[1:]:
        acc = acc + s
    return acc * (1.0 / world)


def grad_update_torch(models, emb_grad_quantized=True, num_bits=8, mlp_layer_quantized=True):
    """grad_update_parallel_comm over an in-process list of replicas (one per
    rank) -- the collectives of sgd:257-446 are evaluated as explicit sums in
    rank order.  Leaves on every replica: ``emb.grad_rows/grad_q`` (union rows,
    averaged codes), ``emb.emb_scaling_factor``, and for each linear layer
    ``weight_q/bias_q`` + ``weight_scaling_factor/bias_scaling_factor``."""
    world = len(models)
    with torch.no_grad():
        for t in range(len(models[0].emb_l)):
            co = [m.emb_l[t].embedding_bag.weight.grad.coalesce() for m in models]          # sgd:859
            if emb_grad_quantized:
                s_loc = [table_scale_torch(c.values(), num_bits) for c in co]               # sgd:861
                s_bar = _allreduce_mean_scale(s_loc, world).view(-1)                        # sgd:865-867
                qs = [torch.sparse_coo_tensor(c.indices(), quantize_torch(c.values(), num_bits, s_bar),
                                              size=c.size()) for c in co]                   # sgd:869
            else:
                s_bar, qs = None, co
            tot = qs[0]
            for q in qs[1:]:
                tot = tot + q
            tot = tot.coalesce()                                                            # sgd:878
            vals = tot.values() * (1.0 / world)                                             # sgd:885
            for m in models:
                e = m.emb_l[t]
                e.grad_rows, e.grad_q = tot.indices()[0].clone(), vals.clone()
                if s_bar is not None:
                    e.emb_scaling_factor = s_bar.clone()
        for group in ("bot_l", "top_l"):
            for li in range(len(getattr(models[0], group))):
                layers = [getattr(m, group)[li] for m in models]
                gw = [l.weight.grad for l in layers]
                gb = [l.bias.grad for l in layers]
                if mlp_layer_quantized:
                    sw = []
                    for g in gw:                                                             # sgd:905-910
                        w_min, _ = torch.min(g, dim=1)
                        w_max, _ = torch.max(g, dim=1)
                        s, _ = torch.max(torch.stack([w_min.abs(), w_max.abs()], dim=1), dim=1)
                        sw.append(torch.clamp(s, min=1e-8) / qrange(8))
                    sw_bar = _allreduce_mean_scale(sw, world)
                    qw = sum(quantize_torch(g, 8, sw_bar) for g in gw) * (1.0 / world)
                    sb = []
                    for g in gb:                                                             # sgd:945-947
                        s = max(torch.min(g, dim=0)[0].abs(), torch.max(g, dim=0)[0].abs())
                        sb.append(torch.clamp(s, min=1e-8) / qrange(8))
                    sb_bar = _allreduce_mean_scale(sb, world)
                    qb = sum(quantize_torch(g, 8, sb_bar) for g in gb) * (1.0 / world)
                else:
                    sw_bar = sb_bar = None
                    qw = sum(gw) * (1.0 / world)
                    qb = sum(gb) * (1.0 / world)
                for l in layers:
                    l.weight_q, l.bias_q = qw.clone(), qb.clone()
                    l.weight_scaling_factor, l.bias_scaling_factor = sw_bar, sb_bar


def weight_update_torch(model, lr, emb_grad_quantized=True, mlp_layer_quantized=True) -> None:
    """weight_update_parallel_comm (sgd:601-685) on one replica."""
    with torch.no_grad():
        for e in model.emb_l:
            W = e.embedding_bag.weight.data
            if emb_grad_quantized:
                upd = e.grad_q * e.emb_scaling_factor.item()                                # sgd:618
                W[e.grad_rows] += -lr * upd                                                 # sgd:622
            else:
                W[e.grad_rows] += -lr * e.grad_q                                            # sgd:626
        for group in (model.bot_l, model.top_l):
            for l in group:
                if mlp_layer_quantized:
                    l.weight.data.add_(-lr * l.weight_q * l.weight_scaling_factor.view(-1, 1))   # sgd:642
                    l.bias.data.add_(-lr * l.bias_q * l.bias_scaling_factor)                      # sgd:643
                else:
                    l.weight.data.add_(-lr * l.weight_q)
                    l.bias.data.add_(-lr * l.bias_q)


def train_step_torch(models, batches, lr, emb_grad_quantized=True, num_bits=8, mlp_layer_quantized=True):
    """One iteration of the hot loop (drv:1909-1957) over in-process replicas.
    ``batches[r] = (X, lS_o, lS_i, T)`` is rank r's shard.  Returns the list of
    per-rank losses (python floats)."""
    losses = []
    for m, (X, lS_o, lS_i, T) in zip(models, batches):
        Z = m(X, lS_o, lS_i)
        E = F.binary_cross_entropy(Z, T)                      # loss_fn_wrap, bce mean (drv:192-211)
        clear_gradients_torch(m)
        E.backward()
        losses.append(float(E.detach()))
    grad_update_torch(models, emb_grad_quantized, num_bits, mlp_layer_quantized)
    for m in models:
        weight_update_torch(m, lr, emb_grad_quantized, mlp_layer_quantized)
    return losses
