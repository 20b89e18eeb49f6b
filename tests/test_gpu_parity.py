"""GPU parity: the CUDA path (through the C ABI, via the reference-shaped Python surface) against the
oracle on identical seeded inputs, and against the golden vectors produced by the reference.

Bars (north_star): INT4/INT8 codes, scales, de-duplicated row sets and updated-row sets bit-exact;
pooled fp32 outputs and updated weights within 1e-5 relative (they are in fact bit-exact against the
oracle's explicit-order spec, which is asserted where the order is a contract)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import dqrm_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _mods():
    from deep_quantized_recommendation_model_dqrm_b200 import _lib, synthetic, tables
    from deep_quantized_recommendation_model_dqrm_b200.quantization_supp import quant_modules as qm
    from deep_quantized_recommendation_model_dqrm_b200.quantization_supp import quant_utils as qu
    return _lib, synthetic, tables, qm, qu


def bits_equal(a, b):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float32))
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def cpu(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------ (a1)
@pytest.mark.parametrize("rows,dim", [(3, 16), (1000, 16), (4097, 16), (100003, 16), (257, 64), (70001, 128), (5, 20)])
def test_table_scale_bit_exact(rows, dim):
    _lib, synthetic, tables, qm, qu = _mods()
    rng = np.random.RandomState(rows)
    W = synthetic.table_weights_numpy(rows, dim, rng)
    W[rng.randint(rows), rng.randint(dim)] *= -3.0           # a negative extreme decides the scale
    Wd = torch.tensor(W, device="cuda")
    for bits in (4, 8):
        s = qu.symmetric_linear_quantization_param_two(bits, Wd, None, None, None)
        assert s.dim() == 0 and s.is_cuda
        assert bits_equal(cpu(s), O.table_scale_spec(W, bits))


def test_table_scale_many_tables_one_launch_and_sharded():
    _lib, synthetic, tables, qm, qu = _mods()
    rows = [1460, 583, 101227, 22608, 305, 24, 12517, 633, 3, 93145, 5683, 83593, 3194, 27]
    rng = np.random.RandomState(1)
    Ws = [synthetic.table_weights_numpy(n, 16, rng) for n in rows]
    g = tables.EmbeddingTableGroup([torch.tensor(w, device="cuda") for w in Ws], embedding_bit=4)
    for _ in range(2):                                        # second launch checks the workspace reset
        g.scan_scales()
        want = np.array([O.table_scale_spec(w, 4) for w in Ws], dtype=np.float32)
        assert bits_equal(cpu(g.scale), want)
        assert bits_equal(cpu(g.inv_scale), (np.float32(1.0) / want).astype(np.float32))
    full = cpu(g.absmax).copy()
    for world in (2, 3, 8):                                   # row-sharded scan: max over shards == full scan
        acc = np.zeros_like(full)
        for r in range(world):
            rc = g.lib.dqrm_table_absmax_scale(g.T, g._wptrs(), g._rows_arr, g.dim, 4, r, world, g.absmax.data_ptr(),
                                               None, None, g._scan_ws.data_ptr(), _lib.stream_ptr())
            assert rc == 0
            acc = np.maximum(acc, cpu(g.absmax))
        assert bits_equal(acc, full)


# ------------------------------------------------------------------------------------ (a3), (a4), (a5)
EMB_CASES = ["emb_multihot_d16", "emb_multihot_d64_b8", "emb_onehot_zipf_d16", "emb_onehot_tiny_d16", "emb_ragged_d16"]


@pytest.mark.parametrize("name", EMB_CASES)
def test_emb_module_forward_backward_vs_golden_and_oracle(name):
    _lib, synthetic, tables, qm, qu = _mods()
    g = load_golden(name)
    bits = int(g["bits"])
    E = qm.QuantEmbeddingBagTwo(int(g["rows"]), int(g["dim"]), bits, embedding_id=0,
                                _weight=torch.tensor(g["W"], device="cuda"))
    idx, off = torch.tensor(g["idx"], device="cuda"), torch.tensor(g["off"], device="cuda")
    out = E(idx, off)
    # reference-produced vectors: scale + codes bit-exact, output 1e-5
    assert bits_equal(cpu(E.eb_scaling_factor), g["scale"])
    assert np.array_equal(cpu(E.output_integer), g["codes"])
    np.testing.assert_allclose(cpu(out), g["out"], rtol=RTOL, atol=0)
    # oracle spec: everything bit-exact
    s, codes, want = O.embbag_forward_spec(g["W"], g["idx"], g["off"], bits)
    assert bits_equal(cpu(out), want)
    out.backward(torch.tensor(g["dout"], device="cuda"))
    grp = E._group
    grp.check_status()
    sp = grp.sparse_grad(0)
    rows = cpu(sp._indices()[0])
    assert np.array_equal(rows, g["co_rows"])                                  # de-duplicated row set
    np.testing.assert_allclose(cpu(sp._values()), g["co_vals"], rtol=RTOL, atol=1e-5 * np.abs(g["dout"]).max())
    r0, v0 = O.embbag_backward_spec(g["dout"], g["idx"], g["off"], g["scale"])
    urows, sums = O.coalesce_spec(r0, v0)
    assert np.array_equal(rows, urows) and bits_equal(cpu(sp._values()), sums)   # original-order left fold
    assert bits_equal(cpu(grp.grad_scale_local[0]), O.grad_scale_spec(sums, 8))
    # full-precision branch (qm:395)
    pooled = E(idx, off, full_precision_flag=True)
    assert bits_equal(cpu(pooled), O.pool_sum_spec(g["W"], g["idx"], g["off"]))
    np.testing.assert_allclose(cpu(pooled), g["pooled"], rtol=RTOL, atol=1e-9)


def _c1_group(B=128, P=10, seed=7, dim=16, rows=None, bits=4):
    _lib, synthetic, tables, qm, qu = _mods()
    rows = rows or [10000] * 8
    rng = np.random.RandomState(seed)
    Ws = [synthetic.table_weights_numpy(n, dim, rng) for n in rows]
    X, lS_o, lS_i, T = synthetic.random_batch(rows, B, P, seed=seed + 1)
    g = tables.EmbeddingTableGroup([torch.tensor(w, device="cuda") for w in Ws], embedding_bit=bits)
    idx, off, idx_begin, bags = tables.EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cuda")
    return g, Ws, lS_i, lS_o, (idx, off, idx_begin, bags), rng


@pytest.mark.parametrize("dim,P", [(16, 10), (64, 4), (128, 3), (32, 1)])
def test_group_fwd_bwd_exchange_update_world1(dim, P):
    """Config C1 shape: 8 tables x 10k rows, multi-hot bags; the whole (a1..a9) chain for one rank."""
    g, Ws, lS_i, lS_o, (idx, off, idx_begin, bags), rng = _c1_group(P=P, dim=dim)
    g.scan_scales()
    out = g.forward(idx, off, idx_begin, bags)
    dout = rng.randn(g.T, bags, dim).astype(np.float32)
    g.backward(torch.tensor(dout, device="cuda"), world=1)
    g.keep_debug = True
    g.exchange(world=1, rank=0)
    g.merge_apply(0.1)
    g.check_status()
    cnt, srows, scodes = g.slot_views(0)
    for t in range(g.T):
        i_t, o_t = lS_i[t].numpy(), lS_o[t].numpy()
        s, codes, want = O.embbag_forward_spec(Ws[t], i_t, o_t, 4)
        assert bits_equal(cpu(g.scale[t]), s) and bits_equal(cpu(out[t]), want)
        assert np.array_equal(cpu(g.codes[t]).astype(np.float32), codes)
        r0, v0 = O.embbag_backward_spec(dout[t], i_t, o_t, s)
        urows, sums = O.coalesce_spec(r0, v0)
        ex = O.exchange_emb_grad_spec([(urows, sums)], 8, Ws[t].shape[0])
        U = int(cnt[t])
        assert U == len(urows) and np.array_equal(cpu(srows[t, :U]), urows)
        assert bits_equal(cpu(g.grad_scale_mean[t]), ex["s_bar"])
        assert np.array_equal(cpu(scodes[t, :U]).astype(np.float32), ex["codes"][0])          # INT8 codes
        Wn = Ws[t].copy()
        O.weight_update_emb_spec(Wn, ex["union_rows"], ex["qbar"], ex["s_bar"], 0.1)
        assert bits_equal(cpu(g.weights[t]), Wn)                                             # updated table
        nu = int(g.updated_count[t])
        order = np.argsort(cpu(g.updated_rows[t, :nu]))
        assert np.array_equal(cpu(g.updated_rows[t, :nu])[order], ex["union_rows"])           # updated-row set
        assert np.array_equal(cpu(g.qbar[t, :nu])[order], ex["qbar"])      # values (the oracle keeps rint's -0.0)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchange_emulated_ranks(world):
    """N ranks' slots merged on one GPU (each 'rank' is a replica group over its own copy of the tables)."""
    _lib, synthetic, tables, qm, qu = _mods()
    rows = [50, 3, 1000, 20000]
    dim, B = 16, 64
    rng = np.random.RandomState(world)
    Ws = [synthetic.table_weights_numpy(n, dim, rng) for n in rows]
    groups, per_rank = [], []
    for r in range(world):
        g = tables.EmbeddingTableGroup([torch.tensor(w, device="cuda") for w in Ws], embedding_bit=4)
        X, lS_o, lS_i, T = synthetic.criteo_batch(rows, B, seed=50 + r, zipf=1.2 if r % 2 else None)
        idx, off, idx_begin, bags = tables.EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cuda")
        g.scan_scales()
        g.forward(idx, off, idx_begin, bags)
        dout = (rng.randn(g.T, bags, dim) * 0.01).astype(np.float32)
        g.backward(torch.tensor(dout, device="cuda"), world=world)
        g.keep_debug = True
        groups.append(g)
        per_rank.append((lS_i.numpy(), lS_o.numpy(), dout))
    for r, g in enumerate(groups):
        g.stage_scale(r)
    allsc = torch.stack([g.grad_scale_local for g in groups])
    for g in groups:
        g.gathered_scales.copy_(allsc)
    for r, g in enumerate(groups):
        g.pack(r)
    sb = groups[0].slot_bytes
    for g in groups:
        for r2, g2 in enumerate(groups):
            if g2 is not g:
                g.gathered[r2 * sb:(r2 + 1) * sb].copy_(g2.gathered[r2 * sb:(r2 + 1) * sb])
    for g in groups:
        g.merge_apply(0.1)
        g.check_status()
    for t in range(len(rows)):
        s = O.table_scale_spec(Ws[t], 4)
        pr = []
        for (li, lo, dout) in per_rank:
            r0, v0 = O.embbag_backward_spec(dout[t], li[t], lo[t], s)
            pr.append(O.coalesce_spec(r0, v0))
        ex = O.exchange_emb_grad_spec(pr, 8, rows[t])
        Wn = Ws[t].copy()
        O.weight_update_emb_spec(Wn, ex["union_rows"], ex["qbar"], ex["s_bar"], 0.1)
        for r, g in enumerate(groups):
            assert bits_equal(cpu(g.grad_scale_mean[t]), ex["s_bar"])
            cnt, srows, scodes = g.slot_views(r)
            U = int(cnt[t])
            assert np.array_equal(cpu(srows[t, :U]), pr[r][0])
            assert np.array_equal(cpu(scodes[t, :U]).astype(np.float32), ex["codes"][r])
            nu = int(g.updated_count[t])
            order = np.argsort(cpu(g.updated_rows[t, :nu]))
            assert np.array_equal(cpu(g.updated_rows[t, :nu])[order], ex["union_rows"])
            assert np.array_equal(cpu(g.qbar[t, :nu])[order], ex["qbar"])      # values (the oracle keeps rint's -0.0)
            assert bits_equal(cpu(g.weights[t]), Wn)                    # replicas bit-identical to the oracle


def test_large_table_path_matches_oracle():
    """> DQRM_BWD_CTA_MAX_LOOKUPS lookups on one table: multi-block radix-sort path."""
    _lib, synthetic, tables, qm, qu = _mods()
    rows, dim, B, P = 5000, 16, 4096, 8
    rng = np.random.RandomState(3)
    W = synthetic.table_weights_numpy(rows, dim, rng)
    idx, off = synthetic.random_bags(rows, B, P, rng, fixed=True)
    assert idx.numel() > _lib.BWD_CTA_MAX_LOOKUPS
    g = tables.EmbeddingTableGroup([torch.tensor(W, device="cuda")], embedding_bit=4)
    i2, o2, ib, bags = tables.EmbeddingTableGroup.pack_inputs([idx], [off], "cuda")
    g.scan_scales()
    out = g.forward(i2, o2, ib, bags)
    s, codes, want = O.embbag_forward_spec(W, idx.numpy(), off.numpy(), 4)
    assert bits_equal(cpu(out[0]), want)
    dout = rng.randn(1, bags, dim).astype(np.float32)
    g.backward(torch.tensor(dout, device="cuda"), world=1)
    g.check_status()
    sp = g.sparse_grad(0)
    r0, v0 = O.embbag_backward_spec(dout[0], idx.numpy(), off.numpy(), s)
    urows, sums = O.coalesce_spec(r0, v0)
    assert np.array_equal(cpu(sp._indices()[0]), urows)
    assert bits_equal(cpu(sp._values()), sums)
    assert bits_equal(cpu(g.grad_scale_local[0]), O.grad_scale_spec(sums, 8))


@pytest.mark.parametrize("case", ["zipf_onehot", "three_rows", "ragged_multihot", "dim64"])
def test_large_path_one_kernel_sort(case):
    """> DQRM_BWD_CTA_MAX_LOOKUPS lookups on one table: the persistent radix-sort kernel (csrc/embbag_bwd_large.cu)
    against the oracle -- ascending unique rows, blocked left fold of heavy duplicates, gradient scale: bit-exact."""
    _lib, synthetic, tables, qm, qu = _mods()
    rng = np.random.RandomState(11)
    dim = 64 if case == "dim64" else 16
    if case == "zipf_onehot":
        rows, B = 300, 40000
        idx = torch.from_numpy(np.minimum(rng.zipf(1.1, size=B) - 1, rows - 1).astype(np.int64))
        off = torch.arange(B, dtype=torch.int64)
    elif case == "three_rows":
        rows, B = 3, 50000
        idx = torch.from_numpy(rng.randint(0, 3, size=B).astype(np.int64))
        off = torch.arange(B, dtype=torch.int64)
    elif case == "ragged_multihot":
        rows, B = 100000, 9000
        idx, off = synthetic.random_bags(rows, B, 7, rng)
        off[1000:1100] = off[1000]                       # a run of empty bags
    else:
        rows, B = 1 << 20, 20000
        idx, off = synthetic.random_bags(rows, B, 2, rng, fixed=True)
    assert idx.numel() > _lib.BWD_CTA_MAX_LOOKUPS
    W = synthetic.table_weights_numpy(rows, dim, rng)
    g = tables.EmbeddingTableGroup([torch.tensor(W, device="cuda")], embedding_bit=4)
    i2, o2, ib, bags = tables.EmbeddingTableGroup.pack_inputs([idx], [off], "cuda")
    g.scan_scales()
    g.forward(i2, o2, ib, bags)
    dout = rng.randn(1, bags, dim).astype(np.float32)
    for rep in range(2):                                 # the second call reuses the workspace (barrier state reset)
        g.backward(torch.tensor(dout, device="cuda"), world=1)
        g.check_status()
        sp = g.sparse_grad(0)
        r0, v0 = O.embbag_backward_spec(dout[0], idx.numpy(), off.numpy(), O.table_scale_spec(W, 4))
        urows, sums = O.coalesce_spec(r0, v0)
        assert np.array_equal(cpu(sp._indices()[0]), urows)
        assert bits_equal(cpu(sp._values()), sums)
        assert bits_equal(cpu(g.grad_scale_local[0]), O.grad_scale_spec(sums, 8))


@pytest.mark.parametrize("path,lookups", [("cta", None), ("sort", None), ("auto", None), ("cta", 6000), ("auto", 6000),
                                          ("cta_cluster4", 6000), ("cta_cluster8", None), ("cta_bitonic", 6000)])
def test_max_cta_lookups_boundary(path, lookups, monkeypatch):
    """Exactly DQRM_BWD_CTA_MAX_LOOKUPS lookups (largest single-CTA sort) with heavy duplication, on the single-CTA
    path, on the whole-chip sort kernel, and on whichever the cost model picks (one table of >= 4k lookups: the sort
    kernel): the same bits either way."""
    _lib, synthetic, tables, qm, qu = _mods()
    if path.startswith("cta_cluster"):         # the single-CTA sort with a thread-block cluster folding (DSMEM reads)
        monkeypatch.setenv("DQRM_BWD_CLUSTER", path[-1])
        path = "cta"
    if path == "cta_bitonic":                  # 6000 lookups sort by shared-memory radix passes by default
        monkeypatch.setenv("DQRM_BWD_CTA_SORT", "bitonic")
        path = "cta"
    if path != "auto":
        monkeypatch.setenv("DQRM_BWD_PATH", path)
    rows, dim, B = 300, 16, lookups or _lib.BWD_CTA_MAX_LOOKUPS
    rng = np.random.RandomState(4)
    W = synthetic.table_weights_numpy(rows, dim, rng)
    idx = torch.from_numpy(np.minimum(rng.zipf(1.1, size=B) - 1, rows - 1).astype(np.int64))
    off = torch.arange(B, dtype=torch.int64)
    g = tables.EmbeddingTableGroup([torch.tensor(W, device="cuda")], embedding_bit=4)
    i2, o2, ib, bags = tables.EmbeddingTableGroup.pack_inputs([idx], [off], "cuda")
    g.scan_scales()
    g.forward(i2, o2, ib, bags)
    dout = rng.randn(1, bags, dim).astype(np.float32)
    g.backward(torch.tensor(dout, device="cuda"), world=1)
    g.check_status()
    sp = g.sparse_grad(0)
    r0, v0 = O.embbag_backward_spec(dout[0], idx.numpy(), off.numpy(), O.table_scale_spec(W, 4))
    urows, sums = O.coalesce_spec(r0, v0)
    assert np.array_equal(cpu(sp._indices()[0]), urows) and bits_equal(cpu(sp._values()), sums)


def test_out_of_range_index_is_flagged():
    _lib, synthetic, tables, qm, qu = _mods()
    E = qm.QuantEmbeddingBagTwo(10, 16, 4, _weight=torch.zeros(10, 16, device="cuda"))
    E(torch.tensor([1, 10, 2], device="cuda"), torch.tensor([0, 1, 2], device="cuda"))
    with pytest.raises(IndexError):
        E._group.check_status()
    E._group.check_status()        # flag is cleared after being reported


# ------------------------------------------------------------------------------------------ (a8)
def test_topk_rows():
    g, Ws, lS_i, lS_o, (idx, off, idx_begin, bags), rng = _c1_group(B=256, P=1, seed=11)
    g.scan_scales()
    g.forward(idx, off, idx_begin, bags)
    dout = rng.randn(g.T, bags, 16).astype(np.float32)
    g.backward(torch.tensor(dout, device="cuda"), world=1)
    before = [(cpu(g.uniq_rows[t, :int(g.uniq_count[t])]).copy(), cpu(g.grad_sums[t, :int(g.uniq_count[t])]).copy())
              for t in range(g.T)]
    k = 50
    g.topk(k)
    for t in range(g.T):
        rows, sums, _ = O.topk_rows_spec(before[t][0], before[t][1], k)
        U = int(g.uniq_count[t])
        assert U == min(k, len(before[t][0]))
        assert np.array_equal(cpu(g.uniq_rows[t, :U]), rows)
        assert bits_equal(cpu(g.grad_sums[t, :U]), sums)
        assert bits_equal(cpu(g.grad_scale_local[t]), O.grad_scale_spec(sums, 8))
    g.topk(10 ** 6)                                    # k >= count: identity
    assert int(g.uniq_count[0]) == min(k, len(before[0][0]))


# ----------------------------------------------------------------------------------------- (a10)
def test_sgd_rows_and_rwsadagrad():
    g, Ws, lS_i, lS_o, (idx, off, idx_begin, bags), rng = _c1_group(B=64, P=3, seed=13)
    g.scan_scales()
    g.forward(idx, off, idx_begin, bags)
    dout = rng.randn(g.T, bags, 16).astype(np.float32)
    g.backward(torch.tensor(dout, device="cuda"), world=1)
    sums = [(cpu(g.uniq_rows[t, :int(g.uniq_count[t])]).astype(np.int64), cpu(g.grad_sums[t, :int(g.uniq_count[t])]).copy())
            for t in range(g.T)]
    g.sgd_apply(0.1)
    for t in range(g.T):
        Wn = Ws[t].copy()
        O.weight_update_emb_unquantized_spec(Wn, sums[t][0], sums[t][1], 0.1)
        assert bits_equal(cpu(g.weights[t]), Wn)
    # row-wise sparse Adagrad (optim/rwsadagrad.py:97-113) against the same formula in torch
    mom = [torch.zeros(w.shape[0], device="cuda") for w in g.weights]
    W0 = [cpu(w).copy() for w in g.weights]
    g.sgd_apply(0.05, momentum=mom, eps=1e-10)
    for t in range(g.T):
        r, v = sums[t]
        m = np.zeros(W0[t].shape[0], dtype=np.float32)
        m[r] += (v ** 2).mean(axis=1)
        Wn = W0[t].copy()
        Wn[r] += -0.05 * (v / (np.sqrt(m[r]) + 1e-10)[:, None])
        np.testing.assert_allclose(cpu(mom[t]), m, rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(cpu(g.weights[t]), Wn, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("case", ["zipf_onehot", "three_rows", "ragged_multihot", "dim64", "dim128", "small_cta", "dim8",
                                  "dim4_multihot", "dim256", "dim512"])
@pytest.mark.parametrize("adagrad", [False, True])
def test_bwd_sgd_fused(case, adagrad):
    """dqrm_embbag_bwd_sgd (north-star kernel 3: sort + de-duplicate + row update in ONE kernel on the radix-sort
    path): updated-row lists equal to the oracle's coalesce, tables bit-identical to backward() + sgd_apply() and to
    the oracle's W.add_(-lr * coalesced grad) (sgd_quantized_gradients_parallel_comm.py:626); RW-Adagrad state and
    tables against optim/rwsadagrad.py:97-113 as restated by the oracle (fp32 rounding of the row mean allowed)."""
    _lib, synthetic, tables, qm, qu = _mods()
    rng = np.random.RandomState(17)
    dim = {"dim64": 64, "dim128": 128, "dim8": 8, "dim4_multihot": 4, "dim256": 256, "dim512": 512}.get(case, 16)
    if case == "zipf_onehot":
        rows, B = 300, 40000
        idx = torch.from_numpy(np.minimum(rng.zipf(1.1, size=B) - 1, rows - 1).astype(np.int64))
        off = torch.arange(B, dtype=torch.int64)
    elif case == "three_rows":
        rows, B = 3, 50000
        idx = torch.from_numpy(rng.randint(0, 3, size=B).astype(np.int64))
        off = torch.arange(B, dtype=torch.int64)
    elif case in ("ragged_multihot", "dim4_multihot"):
        rows, B = 100000, 9000
        idx, off = synthetic.random_bags(rows, B, 7, rng)
        off[1000:1100] = off[1000]
    elif case in ("dim256", "dim512"):                   # wide rows: 2 / 4 float4 columns per lane, fewer rows in flight
        rows, B = 50000, 20000
        idx = torch.from_numpy(np.minimum(rng.zipf(1.05, size=B) - 1, rows - 1).astype(np.int64))
        off = torch.arange(B, dtype=torch.int64)
    elif case == "small_cta":
        rows, B = 5000, 300                              # few lookups: single-CTA de-duplication + row-update kernel
        idx, off = synthetic.random_bags(rows, B, 3, rng)
    else:
        rows, B = 1 << 20, 20000
        idx, off = synthetic.random_bags(rows, B, 3, rng, fixed=True)
    W = synthetic.table_weights_numpy(rows, dim, rng)
    lr = 0.05
    i2, o2, ib, bags = tables.EmbeddingTableGroup.pack_inputs([idx], [off], "cuda")
    dout = rng.randn(1, bags, dim).astype(np.float32)
    res = []
    for fused in (True, False):
        g = tables.EmbeddingTableGroup([torch.tensor(W, device="cuda")], embedding_bit=4)
        mom = [torch.zeros(rows, device="cuda")] if adagrad else None
        for rep in range(2):                             # two steps: the workspace and the state are reused
            g.scan_scales()
            g.forward(i2, o2, ib, bags)
            d = torch.tensor(dout * (rep + 1), device="cuda")
            if fused:
                g.backward_sgd(d, lr, momentum=mom, eps=1e-10)
            else:
                g.backward(d, world=1)
                g.sgd_apply(lr, momentum=mom, eps=1e-10)
            g.check_status()
        U = int(g.uniq_count[0])
        res.append((cpu(g.weights[0]).copy(), cpu(g.uniq_rows[0, :U]).copy(), cpu(mom[0]).copy() if adagrad else None))
    (Wf, rf, mf), (Wu, ru, mu) = res
    assert np.array_equal(rf, ru)
    assert bits_equal(Wf, Wu)
    if adagrad:
        assert bits_equal(mf, mu)
    # against the oracle
    Wo, mo = W.copy(), np.zeros(rows, dtype=np.float32)
    for rep in range(2):
        r0, v0 = O.embbag_backward_spec(dout[0] * np.float32(rep + 1), idx.numpy(), off.numpy(), O.table_scale_spec(Wo, 4))
        urows, sums = O.coalesce_spec(r0, v0)
        if adagrad:
            O.rwsadagrad_rows_spec(Wo, mo, urows, sums, lr, 1e-10)
        else:
            O.weight_update_emb_unquantized_spec(Wo, urows, sums, lr)
    assert np.array_equal(rf, urows)
    if adagrad:
        np.testing.assert_allclose(mf, mo, rtol=2e-6, atol=1e-12)
        np.testing.assert_allclose(Wf, Wo, rtol=1e-5, atol=1e-7)
    else:
        assert bits_equal(Wf, Wo)



@pytest.mark.parametrize("case", ["cta", "sort"])
def test_bwd_sgd_vs_uncoalesced_reference_update(case):
    """(a10) the reference's single-process update is torch.optim.SGD on the UNCOALESCED sparse gradient
    (dlrm_s_pytorch_single_gpu.py:1944-1946): duplicates are applied one at a time, W += -lr*g1; W += -lr*g2 (oracle
    sgd_sparse_spec), in an order that is undefined on CUDA.  dqrm_embbag_bwd_sgd sums a row's duplicates first (fixed
    order) and applies once: equal to the reference's result within fp32 rounding of the row -- 1e-5 relative, the
    north-star tolerance for updated weights -- and the set of changed rows is exactly unique(indices)."""
    _lib, synthetic, tables, qm, qu = _mods()
    rng = np.random.RandomState(23)
    if case == "cta":
        rows, B, P = 5000, 300, 3
    else:
        rows, B, P = 100000, 9000, 7
    idx, off = synthetic.random_bags(rows, B, P, rng)
    W = synthetic.table_weights_numpy(rows, 16, rng)
    g = tables.EmbeddingTableGroup([torch.tensor(W, device="cuda")], embedding_bit=4)
    i2, o2, ib, bags = tables.EmbeddingTableGroup.pack_inputs([idx], [off], "cuda")
    g.scan_scales()
    g.forward(i2, o2, ib, bags)
    dout = rng.randn(1, bags, 16).astype(np.float32)
    g.backward_sgd(torch.tensor(dout, device="cuda"), 0.1)
    g.check_status()
    r0, v0 = O.embbag_backward_spec(dout[0], idx.numpy(), off.numpy(), O.table_scale_spec(W, 4))
    Wref = W.copy()
    O.sgd_sparse_spec(Wref, r0, v0, 0.1)
    got = cpu(g.weights[0])
    np.testing.assert_allclose(got, Wref, rtol=1e-5, atol=1e-7)
    changed = np.nonzero((got != W).any(axis=1))[0]
    assert set(changed.tolist()) <= set(np.unique(idx.numpy()).tolist())
    U = int(g.uniq_count[0])
    assert np.array_equal(cpu(g.uniq_rows[0, :U]), np.unique(idx.numpy()))


def test_bwd_sgd_vs_reference_optimizer_golden():
    """dqrm_embbag_bwd_sgd against the REFERENCE's single-process path (tests/golden/sgd_single.npz: its
    QuantEmbeddingBagTwo + torch.optim.SGD on the uncoalesced sparse gradient, three steps, duplicate rows): forward
    scales bit-exact at every step, tables within 1e-5."""
    _lib, synthetic, tables, qm, qu = _mods()
    gd = load_golden("sgd_single")
    rows, dim, lr = int(gd["rows"]), int(gd["dim"]), float(gd["lr"])
    g = tables.EmbeddingTableGroup([torch.tensor(gd["W_init"], device="cuda")], embedding_bit=int(gd["bits"]))
    for s in range(int(gd["steps"])):
        idx, off = torch.tensor(gd[f"idx{s}"]), torch.tensor(gd[f"off{s}"])
        i2, o2, ib, bags = tables.EmbeddingTableGroup.pack_inputs([idx], [off], "cuda")
        g.scan_scales()
        assert cpu(g.scale[0]).tobytes() == gd[f"scale{s}"].astype(np.float32).tobytes()
        g.forward(i2, o2, ib, bags)
        g.backward_sgd(torch.tensor(gd[f"dout{s}"], device="cuda").view(1, bags, dim), lr)
        g.check_status()
        np.testing.assert_allclose(cpu(g.weights[0]), gd[f"W{s}"], rtol=1e-5, atol=1e-7)


def test_rwsadagrad_rows_vs_reference_golden():
    """dqrm_sgd_rows with momentum against 4 steps of the REFERENCE's optim/rwsadagrad.py:97-113
    (tests/golden/rwsadagrad_rows.npz: sparse gradients with duplicate rows, state and weights after every step)."""
    _lib, synthetic, tables, qm, qu = _mods()
    gd = load_golden("rwsadagrad_rows")
    rows, dim, B = int(gd["rows"]), int(gd["dim"]), 24
    g = tables.EmbeddingTableGroup([torch.tensor(gd["W_init"], device="cuda")], embedding_bit=4)
    mom = [torch.zeros(rows, device="cuda")]
    off = torch.arange(B, dtype=torch.int64, device="cuda").view(1, B)
    for step in range(int(gd["steps"])):
        idx = torch.tensor(gd[f"idx{step}"], device="cuda")
        g.scan_scales()
        g.forward(idx, off, [0, B], B, full_precision=True)
        g.backward(torch.tensor(gd[f"vals{step}"], device="cuda").view(1, B, dim), world=1)   # coalesce (dedup + fold)
        g.sgd_apply(float(gd["lr"]), momentum=mom, eps=float(gd["eps"]))
        g.check_status()
        np.testing.assert_allclose(cpu(mom[0]), gd[f"m{step}"], rtol=2e-6, atol=1e-12)
        np.testing.assert_allclose(cpu(g.weights[0]), gd[f"W{step}"], rtol=1e-5, atol=1e-7)
        # and the kernel against the oracle spec that the CPU suite pins on the same golden
    W, m = gd["W_init"].copy(), np.zeros(rows, dtype=np.float32)
    for step in range(int(gd["steps"])):
        r, sums = O.coalesce_spec(gd[f"idx{step}"], gd[f"vals{step}"])[:2]
        O.rwsadagrad_rows_spec(W, m, r, sums, float(gd["lr"]), float(gd["eps"]))
    np.testing.assert_allclose(cpu(mom[0]), m, rtol=2e-6, atol=1e-12)
    np.testing.assert_allclose(cpu(g.weights[0]), W, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------ (a14)
@pytest.mark.parametrize("name", ["interact_kaggle", "interact_tb", "interact_small"])
def test_interact_vs_golden(name):
    from deep_quantized_recommendation_model_dqrm_b200 import dlrm_s_pytorch_comm_grad as drv
    g = load_golden(name)
    x = torch.tensor(g["x"], device="cuda", requires_grad=True)
    ly = torch.tensor(g["ly"], device="cuda", requires_grad=True)          # [T, B, D]
    m = drv.DLRM_Net()
    m.arch_interaction_op, m.arch_interaction_itself, m.quantization_flag, m.quantize_activation = "dot", False, True, False
    R, none = m.interact_features(x, list(ly.unbind(0)))
    assert none is None
    np.testing.assert_allclose(cpu(R), g["R"], rtol=RTOL, atol=1e-5)
    R.backward(torch.tensor(g["dR"], device="cuda"))
    np.testing.assert_allclose(cpu(x.grad), g["dx"], rtol=RTOL, atol=2e-5)
    np.testing.assert_allclose(cpu(ly.grad), g["dly"], rtol=RTOL, atol=2e-5)


# ------------------------------------------------------------------------------------------ (a15)
@pytest.mark.parametrize("name", ["linear_13_64", "linear_367_32", "linear_64_1"])
def test_quant_linear_vs_golden(name):
    _lib, synthetic, tables, qm, qu = _mods()
    g = load_golden(name)
    LL = torch.nn.Linear(g["W"].shape[1], g["W"].shape[0])
    LL.weight.data, LL.bias.data = torch.tensor(g["W"]), torch.tensor(g["b"])
    Q = qm.QuantLinear(weight_bit=int(g["bits"]), bias_bit=int(g["bits"]), per_channel=True)
    Q.set_param(LL)
    Q = Q.cuda()
    x = torch.tensor(g["x"], device="cuda", requires_grad=True)
    y, none = Q(x)
    assert none is None
    assert bits_equal(cpu(Q.fc_scaling_factor), g["scale"])
    assert np.array_equal(cpu(Q.weight_integer), g["W_int"]) and np.array_equal(cpu(Q.bias_integer), g["b_int"])
    np.testing.assert_allclose(cpu(y), g["y"], rtol=RTOL, atol=1e-5)
    y.backward(torch.tensor(g["dy"], device="cuda"))
    np.testing.assert_allclose(cpu(x.grad), g["dx"], rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(cpu(Q.weight.grad), g["dW"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(cpu(Q.bias.grad), g["db"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["linear_13_64", "linear_367_32", "linear_64_1"])
def test_fused_quant_linear_vs_golden(name):
    """The kernels the training step actually runs (dqrm_mlp_fakequant_all + dqrm_linear_fwd / dqrm_linear_bwd through
    QuantLinear.forward_fused) against the reference's QuantLinear forward/backward (qm:105-211): scales and integer
    weights bit-exact; y, dx, dW, db within 1e-5 (relative to the tensor's largest magnitude for the gradients,
    whose entries are sums with cancellation)."""
    _lib, synthetic, tables, qm, qu = _mods()
    from deep_quantized_recommendation_model_dqrm_b200.dense import DenseArena
    g = load_golden(name)
    LL = torch.nn.Linear(g["W"].shape[1], g["W"].shape[0])
    LL.weight.data, LL.bias.data = torch.tensor(g["W"]), torch.tensor(g["b"])
    Q = qm.QuantLinear(weight_bit=int(g["bits"]), bias_bit=int(g["bits"]), per_channel=True)
    Q.set_param(LL)
    Q = Q.cuda()
    arena = DenseArena([Q], torch.device("cuda"))
    assert arena.fused_ok
    arena.fakequant_all()
    assert bits_equal(cpu(Q.fc_scaling_factor), g["scale"])
    assert np.array_equal(cpu(Q.weight_integer), g["W_int"]) and np.array_equal(cpu(Q.bias_integer), g["b_int"])
    x = torch.tensor(g["x"], device="cuda", requires_grad=True)
    arena.zero_grad()
    y = Q.forward_fused(x, 0)
    np.testing.assert_allclose(cpu(y), g["y"], rtol=RTOL, atol=1e-5 * np.abs(g["y"]).max())
    y.backward(torch.tensor(g["dy"], device="cuda"))
    arena.join()
    torch.cuda.synchronize()
    for got, want in ((x.grad, g["dx"]), (Q.weight.grad, g["dW"]), (Q.bias.grad, g["db"])):
        np.testing.assert_allclose(cpu(got), want, rtol=RTOL, atol=1e-5 * np.abs(want).max())


def test_symmetric_quant_function_api():
    _lib, synthetic, tables, qm, qu = _mods()
    rng = np.random.RandomState(0)
    x = rng.randn(37, 16).astype(np.float32)
    s = np.float32(0.05)
    xd = torch.tensor(x, device="cuda", requires_grad=True)
    sd = torch.tensor(s, device="cuda")
    q = qu.SymmetricQuantFunction.apply(xd, 4, sd)
    assert np.array_equal(cpu(q), O.quantize_spec(x, 4, s))
    q.sum().backward()
    assert bits_equal(cpu(xd.grad), (np.ones_like(x) / s).astype(np.float32))
    srow = np.abs(rng.randn(37)).astype(np.float32) + 0.01
    q = qu.SymmetricQuantFunction.apply(xd, 8, torch.tensor(srow, device="cuda"))
    assert np.array_equal(cpu(q), O.quantize_spec(x, 8, srow))
    with pytest.raises(ValueError):
        qu.SymmetricQuantFunction.apply(xd, 4, None)


@pytest.mark.parametrize("world", [2, 4])
def test_unquantized_exchange_emulated_ranks(world):
    """emb_grad_quantized=False across ranks: fp32 payload in the same slots, rank-ordered sums."""
    _lib, synthetic, tables, qm, qu = _mods()
    rows, dim, B = [40, 3000], 16, 48
    rng = np.random.RandomState(100 + world)
    Ws = [synthetic.table_weights_numpy(n, dim, rng) for n in rows]
    groups, per_rank = [], []
    for r in range(world):
        g = tables.EmbeddingTableGroup([torch.tensor(w, device="cuda") for w in Ws], embedding_bit=4)
        X, lS_o, lS_i, T = synthetic.criteo_batch(rows, B, seed=150 + r, zipf=1.2)
        idx, off, ib, bags = tables.EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cuda")
        g.scan_scales()
        g.forward(idx, off, ib, bags)
        dout = (rng.randn(g.T, bags, dim) * 0.01).astype(np.float32)
        g.backward(torch.tensor(dout, device="cuda"), world=world)
        g.set_grad_bit(32)
        g.keep_debug = True
        groups.append(g)
        per_rank.append((lS_i.numpy(), lS_o.numpy(), dout))
    for r, g in enumerate(groups):
        g.stage_scale(r)
        g.pack(r)
    sb = groups[0].slot_bytes
    for g in groups:
        for r2, g2 in enumerate(groups):
            if g2 is not g:
                g.gathered[r2 * sb:(r2 + 1) * sb].copy_(g2.gathered[r2 * sb:(r2 + 1) * sb])
    for g in groups:
        g.merge_apply(0.1)
        g.check_status()
    for t in range(len(rows)):
        s = O.table_scale_spec(Ws[t], 4)
        pr = []
        for (li, lo, dout) in per_rank:
            r0, v0 = O.embbag_backward_spec(dout[t], li[t], lo[t], s)
            pr.append(O.coalesce_spec(r0, v0))
        union, gmean = O.exchange_emb_grad_unquantized_spec(pr)
        Wn = Ws[t].copy()
        O.weight_update_emb_unquantized_spec(Wn, union, gmean, 0.1)
        for g in groups:
            nu = int(g.updated_count[t])
            order = np.argsort(cpu(g.updated_rows[t, :nu]))
            assert np.array_equal(cpu(g.updated_rows[t, :nu])[order], union)
            assert bits_equal(cpu(g.qbar[t, :nu])[order], gmean)
            assert bits_equal(cpu(g.weights[t]), Wn)


@pytest.mark.parametrize("with_ec", [False, True])
def test_dense_local_one_launch_equals_three(with_ec):
    """dqrm_dense_quant_apply_local (world 1: scale + quantise + update in one launch) against dqrm_dense_grad_scale ->
    dqrm_dense_grad_quant -> dqrm_dense_apply on the same data: every buffer bit-identical, with and without
    error compensation (sgd_quantized_gradients_parallel_comm.py:892-961, 630-663)."""
    _lib, synthetic, tables, qm, qu = _mods()
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(77)
    lens = [13, 512, 367, 1, 64, 200, 5, 129, 128, 1024]
    cb = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device="cuda")
    n, C = int(cb[-1]), len(lens)
    st = _lib.stream_ptr()
    lr_dev = torch.full((1,), 0.37, device="cuda")

    def fresh():
        gg = torch.Generator(device="cuda").manual_seed(78)
        param = torch.randn(n, device="cuda", generator=gg)
        grad = torch.randn(n, device="cuda", generator=gg) * 1e-3
        grad[cb[3]:cb[4]] = 0.0                                          # an all-zero channel: scale floor 1e-8
        ec = torch.randn(n, device="cuda", generator=gg) * 1e-5 if with_ec else None
        return param, grad, ec, torch.zeros(C, device="cuda"), torch.zeros(n, device="cuda"), torch.zeros(C, device="cuda")

    p1, g1, e1, sl1, co1, sm1 = fresh()
    _lib.check(lib.dqrm_dense_grad_scale(g1.data_ptr(), _lib.ptr(e1), cb.data_ptr(), C, 8, sl1.data_ptr(), st), "scale")
    _lib.check(lib.dqrm_dense_grad_quant(g1.data_ptr(), cb.data_ptr(), C, sl1.data_ptr(), 1.0, 8, co1.data_ptr(),
                                         sm1.data_ptr(), st), "quant")
    _lib.check(lib.dqrm_dense_apply(p1.data_ptr(), co1.data_ptr(), cb.data_ptr(), C, sm1.data_ptr(), 1.0, 0.0,
                                    lr_dev.data_ptr(), g1.data_ptr() if with_ec else None, _lib.ptr(e1), st), "apply")
    p2, g2, e2, sl2, co2, sm2 = fresh()
    _lib.check(lib.dqrm_dense_quant_apply_local(p2.data_ptr(), g2.data_ptr(), _lib.ptr(e2), cb.data_ptr(), C, 8,
                                                sl2.data_ptr(), co2.data_ptr(), sm2.data_ptr(), 0.0, lr_dev.data_ptr(), st),
               "fused")
    torch.cuda.synchronize()
    for a, b, what in ((p1, p2, "param"), (g1, g2, "grad"), (sl1, sl2, "scale_local"), (co1, co2, "codes"), (sm1, sm2, "scale_mean")):
        assert torch.equal(a, b), what
    if with_ec:
        assert torch.equal(e1, e2)
    assert float(co1.abs().max()) == 127.0
