"""CPU: the oracle (oracle/dqrm_oracle.py) against the golden vectors that
oracle/make_golden.py produced by executing the reference.  Codes, scales and
row sets bit-exact; floating outputs within 1e-5 relative (north_star)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import dqrm_oracle as O
from deep_quantized_recommendation_model_dqrm_b200 import synthetic

EMB_CASES = ["emb_multihot_d16", "emb_multihot_d64_b8", "emb_onehot_zipf_d16", "emb_onehot_tiny_d16",
             "emb_ragged_d16"]
RTOL = 1e-5


@pytest.mark.parametrize("name", EMB_CASES)
def test_emb_forward_spec(name):
    g = load_golden(name)
    bits = int(g["bits"])
    scale, codes, out = O.embbag_forward_spec(g["W"], g["idx"], g["off"], bits)
    assert scale.tobytes() == g["scale"].astype(np.float32).tobytes()
    assert np.array_equal(codes, g["codes"])
    np.testing.assert_allclose(out, g["out"], rtol=RTOL, atol=0)
    _, _, pooled = O.embbag_forward_spec(g["W"], g["idx"], g["off"], bits, full_precision=True)
    np.testing.assert_allclose(pooled, g["pooled"], rtol=RTOL, atol=1e-9)
    # torch-layer scale scan issues the reference's op sequence: same bits
    assert O.table_scale_torch(torch.from_numpy(g["W"]), bits).numpy().tobytes() == scale.tobytes()


@pytest.mark.parametrize("name", EMB_CASES)
def test_emb_backward_and_coalesce_spec(name):
    g = load_golden(name)
    rows, vals = O.embbag_backward_spec(g["dout"], g["idx"], g["off"], g["scale"])
    assert np.array_equal(rows, g["grad_rows"])
    assert np.array_equal(vals, g["grad_vals"])          # (g*s)/s reproduced bit-for-bit
    urows, sums = O.coalesce_spec(rows, vals)
    assert np.array_equal(urows, g["co_rows"])
    # duplicate-fold order is implementation-defined in torch (unstable sort): fp32-rounding agreement ...
    np.testing.assert_allclose(sums, g["co_vals"], rtol=RTOL, atol=1e-5 * np.abs(vals).max())
    # ... and bit-exact once the fold follows torch.sort's own permutation
    perm = torch.from_numpy(rows).sort(0)[1].numpy()
    urows2, sums2 = O.coalesce_spec(rows, vals, order=perm, block=None)
    assert np.array_equal(urows2, g["co_rows"]) and np.array_equal(sums2, g["co_vals"])


def test_coalesce_fold_order_heavy_duplicates():
    rng = np.random.RandomState(5)
    rows = rng.randint(0, 37, size=20000)
    vals = rng.randn(20000, 16).astype(np.float32)
    sp = torch.sparse_coo_tensor(torch.from_numpy(rows)[None], torch.from_numpy(vals), size=(37, 16)).coalesce()
    perm = torch.from_numpy(rows).sort(0)[1].numpy()
    urows, sums = O.coalesce_spec(rows, vals, order=perm, block=None)   # torch: one left fold in its sort's order
    assert np.array_equal(urows, sp.indices()[0].numpy())
    assert np.array_equal(sums, sp.values().numpy())
    _, sums_stable = O.coalesce_spec(rows, vals, block=None)
    np.testing.assert_allclose(sums_stable, sums, rtol=1e-4, atol=1e-3)
    # the spec's blocked fold (rows with > FOLD_BLOCK duplicates; here ~540 each): same rows, fp32-rounding agreement,
    # and identical to the plain fold when no row exceeds the block
    urows_b, sums_b = O.coalesce_spec(rows, vals)
    assert np.array_equal(urows_b, urows)
    np.testing.assert_allclose(sums_b, sums, rtol=1e-4, atol=1e-3)
    few = rng.randint(0, 5000, size=300)
    a = O.coalesce_spec(few, vals[:300])
    b = O.coalesce_spec(few, vals[:300], block=None)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("name", ["linear_13_64", "linear_367_32", "linear_64_1"])
def test_quant_linear(name):
    g = load_golden(name)
    bits = int(g["bits"])
    W_int, b_int, s = O.linear_fakequant_spec(g["W"], g["b"], bits)
    assert np.array_equal(s, g["scale"])
    assert np.array_equal(W_int, g["W_int"])
    assert np.array_equal(b_int, g["b_int"])
    x = torch.tensor(g["x"], requires_grad=True)
    W = torch.nn.Parameter(torch.tensor(g["W"]))
    b = torch.nn.Parameter(torch.tensor(g["b"]))
    y, _ = O.quant_linear_forward_torch(x, W, b, bits)
    y.backward(torch.tensor(g["dy"]))
    for got, want in ((y, "y"), (x.grad, "dx"), (W.grad, "dW"), (b.grad, "db")):
        np.testing.assert_allclose(got.detach().numpy(), g[want], rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("name", ["interact_kaggle", "interact_tb", "interact_small"])
def test_interact(name):
    g = load_golden(name)
    x = torch.tensor(g["x"], requires_grad=True)
    ly = [torch.tensor(t, requires_grad=True) for t in g["ly"]]
    R = O.interact_features_torch(x, ly)
    np.testing.assert_allclose(R.detach().numpy(), g["R"], rtol=RTOL, atol=1e-6)
    R.backward(torch.tensor(g["dR"]))
    np.testing.assert_allclose(x.grad.numpy(), g["dx"], rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(np.stack([t.grad.numpy() for t in ly]), g["dly"], rtol=RTOL, atol=1e-5)


def _shard(world, rank, multihot, step, rows):
    Bg = 16 * world
    sl = slice(rank * 16, (rank + 1) * 16)
    if multihot == "nodup":
        X, lS_o, lS_i, T = synthetic.criteo_batch_nodup(rows, world, 16, seed=400 + step)
        lS_i = lS_i[:, sl]
        return X[sl], lS_o[:, 0:lS_i.shape[1]], lS_i, T[sl]
    if multihot:
        X, lS_o, lS_i, T = synthetic.random_batch(rows, Bg, 4, seed=400 + step)
        li, lo = [], []
        for i, o in zip(lS_i, lS_o):
            ends = torch.cat([o[1:], torch.tensor([i.shape[0]])])
            a, b = int(o[sl][0]), int(ends[sl][-1])
            li.append(i[a:b])
            lo.append(o[sl] - a)
        return X[sl], lo, li, T[sl]
    X, lS_o, lS_i, T = synthetic.criteo_batch(rows, Bg, seed=400 + step, zipf=1.3)
    lS_i = lS_i[:, sl]
    return X[sl], lS_o[:, 0:lS_i.shape[1]], lS_i, T[sl]


def build_oracle_models(world, cfg, seed=300):
    models = []
    for _ in range(world):
        rng = np.random.RandomState(seed)
        emb = [torch.from_numpy(synthetic.table_weights_numpy(n, cfg["dim"], rng)) for n in cfg["rows"]]
        ln_top = synthetic.top_mlp_sizes(len(cfg["rows"]), cfg["dim"], cfg["ln_top_hidden"])
        bot = synthetic.mlp_params(cfg["ln_bot"], rng)
        top = synthetic.mlp_params(ln_top, rng)
        models.append(O.OracleDLRM(cfg["rows"], cfg["dim"], bot, top, emb_weights=emb))
    return models


C_SMALL = dict(rows=[50, 3, 1000, 200], dim=16, ln_bot=[13, 32, 16], ln_top_hidden=[32, 1])


C_NODUP = dict(rows=[1500, 40, 1000, 200], dim=16, ln_bot=[13, 32, 16], ln_top_hidden=[32, 1])


@pytest.mark.parametrize("name", ["dp1_onehot", "dp2_onehot", "dp2_multihot", "dp4_onehot", "dp2_nodup"])
def test_dp_train_steps(name):
    """Two iterations of the reference's custom-DP loop (N Gloo ranks) against
    the oracle's in-process replicas: union rows, averaged codes and scales
    bit-exact; losses and final weights within 1e-5."""
    g = load_golden(name)
    world, multihot = int(g["world"]), bool(g["multihot"])
    cfg = C_SMALL
    if "nodup" in g and bool(g["nodup"]):
        multihot, cfg = "nodup", C_NODUP
    models = build_oracle_models(world, cfg)
    for step in range(2):
        batches = [_shard(world, r, multihot, step, cfg["rows"]) for r in range(world)]
        scales_before = None
        losses = O.train_step_torch(models, batches, lr=0.1) if False else None
        # run the step in two halves so the exchanged artefacts can be inspected
        losses = []
        for m, (X, lS_o, lS_i, T) in zip(models, batches):
            Z = m(X, lS_o, lS_i)
            E = torch.nn.functional.binary_cross_entropy(Z, T)
            O.clear_gradients_torch(m)
            E.backward()
            losses.append(float(E.detach()))
        O.grad_update_torch(models)
        for r in range(world):
            assert abs(losses[r] - float(g[f"rank{r}_loss{step}"])) <= 1e-5 * abs(losses[r]) + 1e-7
        m0 = models[0]
        for k, e in enumerate(m0.emb_l):
            assert np.array_equal(e.grad_rows.numpy(), g[f"s{step}_emb{k}_rows"]), (step, k)
            assert e.eb_scaling_factor.numpy().tobytes() == g[f"s{step}_emb{k}_eb_scale"].tobytes()
            if world <= 2:   # scale mean is order-free for N<=2: bit-exact
                assert np.array_equal(e.emb_scaling_factor.numpy(), g[f"s{step}_emb{k}_sbar"]), (step, k)
                assert np.array_equal(e.grad_q.numpy(), g[f"s{step}_emb{k}_qbar"]), (step, k)
            else:
                np.testing.assert_allclose(e.emb_scaling_factor.numpy(), g[f"s{step}_emb{k}_sbar"], rtol=1e-6)
        for m in models:
            O.weight_update_torch(m, 0.1)
    m0 = models[0]
    for k, e in enumerate(m0.emb_l):
        np.testing.assert_allclose(e.embedding_bag.weight.data.numpy(), g[f"final_emb{k}"], rtol=RTOL, atol=1e-7)
    for grp, layers in (("bot", m0.bot_l), ("top", m0.top_l)):
        for i, l in enumerate(layers):
            np.testing.assert_allclose(l.weight.data.numpy(), g[f"final_{grp}{i}_W"], rtol=RTOL, atol=1e-7)
            np.testing.assert_allclose(l.bias.data.numpy(), g[f"final_{grp}{i}_b"], rtol=RTOL, atol=1e-7)


def test_spec_exchange_matches_torch_layer():
    """numpy spec of the exchange == torch-layer restatement (2 ranks)."""
    rng = np.random.RandomState(9)
    per_rank = []
    for r in range(2):
        rows = rng.randint(0, 40, size=64)
        vals = rng.randn(64, 16).astype(np.float32) * 0.01
        per_rank.append(O.coalesce_spec(rows, vals))
    ex = O.exchange_emb_grad_spec(per_rank, 8, 40)
    W = rng.randn(40, 16).astype(np.float32)
    W2 = W.copy()
    O.weight_update_emb_spec(W2, ex["union_rows"], ex["qbar"], ex["s_bar"], 0.1)
    # torch layer
    s_loc = [O.table_scale_torch(torch.from_numpy(s), 8) for _, s in per_rank]
    s_bar = ((s_loc[0] + s_loc[1]) * 0.5).view(-1)
    assert s_bar.numpy()[0].tobytes() == ex["s_bar"].tobytes()
    sp = None
    for rws, s in per_rank:
        q = O.quantize_torch(torch.from_numpy(s), 8, s_bar)
        t = torch.sparse_coo_tensor(torch.from_numpy(rws)[None], q, size=(40, 16))
        sp = t if sp is None else sp + t
    sp = sp.coalesce()
    assert np.array_equal(sp.indices()[0].numpy(), ex["union_rows"])
    assert np.array_equal((sp.values() * 0.5).numpy(), ex["qbar"])
    Wt = torch.from_numpy(W.copy())
    upd = (sp * 0.5) * s_bar.item()
    Wt.add_(-0.1 * upd)
    assert np.array_equal(Wt.numpy(), W2)


def test_get_my_slice():
    for n, w in ((128, 8), (10, 3), (7, 7), (5, 8)):
        got = []
        for r in range(w):
            got += list(range(n))[O.get_my_slice(n, w, r)]
        assert got == list(range(n))


def test_int4_pack_roundtrip_and_onehot_equivalence():
    rng = np.random.RandomState(3)
    W = synthetic.table_weights_numpy(257, 32, rng)
    s = O.table_scale_spec(W, 4)
    packed = O.pack_int4_spec(W, s)
    assert packed.shape == (257, 16) and packed.dtype == np.uint8
    assert np.array_equal(O.unpack_int4_spec(packed), O.quantize_spec(W, 4, s).astype(np.int32))
    idx = rng.randint(0, 257, size=40)
    off = np.arange(40)
    _, _, out = O.embbag_forward_spec(W, idx, off, 4)
    assert np.array_equal(O.embbag_forward_int4_spec(packed, idx, off, s), out)            # one-index bags


def test_rwsadagrad_rows_spec_vs_reference_optimizer():
    """oracle.rwsadagrad_rows_spec (what dqrm_sgd_rows with momentum must reproduce) against 4 steps of the
    reference's optim/rwsadagrad.py on sparse gradients with duplicate rows."""
    g = load_golden("rwsadagrad_rows")
    W = g["W_init"].copy()
    m = np.zeros(int(g["rows"]), dtype=np.float32)
    for step in range(int(g["steps"])):
        rows, sums = O.coalesce_spec(g[f"idx{step}"], g[f"vals{step}"])[:2]
        O.rwsadagrad_rows_spec(W, m, rows, sums, float(g["lr"]), float(g["eps"]))
        np.testing.assert_allclose(m, g[f"m{step}"], rtol=2e-6, atol=1e-12)
        np.testing.assert_allclose(W, g[f"W{step}"], rtol=1e-5, atol=1e-7)


def test_dp_unquantized_exchange_vs_reference():
    """emb_grad_quantized=False (sgd:319-329, 626): 2 Gloo ranks of the reference against the oracle replicas --
    union rows exact, averaged fp32 row gradients, losses and final weights within 1e-5."""
    g = load_golden("dp2_unquantized")
    world = int(g["world"])
    models = build_oracle_models(world, C_SMALL)
    for step in range(2):
        batches = [_shard(world, r, False, step, C_SMALL["rows"]) for r in range(world)]
        losses = []
        for m, (X, lS_o, lS_i, T) in zip(models, batches):
            Z = m(X, lS_o, lS_i)
            E = torch.nn.functional.binary_cross_entropy(Z, T)
            O.clear_gradients_torch(m)
            E.backward()
            losses.append(float(E.detach()))
        O.grad_update_torch(models, emb_grad_quantized=False)
        for r in range(world):
            assert abs(losses[r] - float(g[f"rank{r}_loss{step}"])) <= 1e-5 * abs(losses[r]) + 1e-7
        for k, e in enumerate(models[0].emb_l):
            assert np.array_equal(e.grad_rows.numpy(), g[f"s{step}_emb{k}_rows"]), (step, k)
            np.testing.assert_allclose(e.grad_q.numpy(), g[f"s{step}_emb{k}_qbar"], rtol=RTOL, atol=1e-9)
        for m in models:
            O.weight_update_torch(m, 0.1, emb_grad_quantized=False)
    m0 = models[0]
    for k, e in enumerate(m0.emb_l):
        np.testing.assert_allclose(e.embedding_bag.weight.data.numpy(), g[f"final_emb{k}"], rtol=RTOL, atol=1e-7)
    for grp, layers in (("bot", m0.bot_l), ("top", m0.top_l)):
        for i, l in enumerate(layers):
            np.testing.assert_allclose(l.weight.data.numpy(), g[f"final_{grp}{i}_W"], rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("name", ["xchg1", "xchg2", "xchg4", "xchg2_ec"])
def test_exchange_on_injected_gradients_bit_exact(name):
    """The exchange + update half of the iteration on INJECTED gradients (no forward/backward, so no GEMM or
    duplicate-fold rounding enters): the reference's grad_update_parallel_comm + weight_update_parallel_comm on
    1/2/4 Gloo ranks (oracle/make_golden.py xchg_worker) vs the oracle spec -- gradient scales, INT8 codes,
    merged row sets, MLP channel scales / codes, error-compensation residuals and every updated weight BIT-EXACT."""
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic
    g = load_golden(name)
    world, ec, steps = int(g["world"]), bool(g["ec"]), int(g["steps"])
    cfg = synthetic.XCHG
    emb_w, mlp_w = synthetic.xchg_weights(cfg)
    emb_w = [w.copy() for w in emb_w]
    mlp_w = [(W.copy(), b.copy()) for W, b in mlp_w]
    ec_w = [[np.zeros_like(W) for W, _ in mlp_w] for _ in range(world)]
    ec_b = [[np.zeros_like(b) for _, b in mlp_w] for _ in range(world)]
    for step in range(steps):
        inj = [synthetic.injected_grads(cfg, world, r, step) for r in range(world)]
        for k, n in enumerate(cfg["rows"]):
            per_rank = [O.coalesce_spec(*inj[r][0][k]) for r in range(world)]
            ex = O.exchange_emb_grad_spec(per_rank, 8, n)
            want_s = g[f"s{step}_emb{k}_sbar"].astype(np.float32)
            if world <= 2:
                assert np.float32(ex["s_bar"]).tobytes() == want_s.tobytes()
            else:       # Gloo's fp32 SUM order over > 2 ranks is backend-internal: <= 1 ulp, then pin the rest on ITS mean
                np.testing.assert_allclose(ex["s_bar"], want_s[0], rtol=2.5e-7)
                ex = O.exchange_emb_grad_spec(per_rank, 8, n, s_bar=want_s[0])
            assert np.array_equal(ex["union_rows"], g[f"s{step}_emb{k}_rows"])
            assert np.array_equal(ex["qbar"], g[f"s{step}_emb{k}_qbar"])         # integer codes / N (-0.0 == 0.0)
            O.weight_update_emb_spec(emb_w[k], ex["union_rows"], ex["qbar"], ex["s_bar"], 0.1)
        for i in range(len(cfg["layers"])):
            gws = [inj[r][1][i][0] for r in range(world)]
            gbs = [inj[r][1][i][1] for r in range(world)]
            if ec:
                s_w, q_w, new_w = O.exchange_dense_grad_ec_spec(gws, [ec_w[r][i] for r in range(world)])
                s_b, q_b, new_b = O.exchange_dense_grad_ec_spec(gbs, [ec_b[r][i] for r in range(world)])
                for r in range(world):
                    ec_w[r][i], ec_b[r][i] = new_w[r], new_b[r]
                    assert new_w[r].tobytes() == g[f"rank{r}_s{step}_lin{i}_ec_w"].tobytes()
                    assert new_b[r].tobytes() == g[f"rank{r}_s{step}_lin{i}_ec_b"].tobytes()
            else:
                s_w, q_w = O.exchange_dense_grad_spec(gws, [O.linear_grad_scale_spec(x) for x in gws])
                s_b, q_b = O.exchange_dense_grad_spec(gbs, [O.bias_grad_scale_spec(x) for x in gbs])
                if world > 2:
                    np.testing.assert_allclose(s_w, g[f"s{step}_lin{i}_s_w"], rtol=2.5e-7)
                    np.testing.assert_allclose(s_b, g[f"s{step}_lin{i}_s_b"][0], rtol=2.5e-7)
                    s_w, q_w = O.exchange_dense_grad_spec(gws, None, s_bar=g[f"s{step}_lin{i}_s_w"])
                    s_b, q_b = O.exchange_dense_grad_spec(gbs, None, s_bar=g[f"s{step}_lin{i}_s_b"][0])
            assert np.asarray(s_w, np.float32).tobytes() == g[f"s{step}_lin{i}_s_w"].tobytes()
            assert np.asarray(s_b, np.float32).reshape(-1).tobytes() == g[f"s{step}_lin{i}_s_b"].tobytes()
            assert np.array_equal(q_w, g[f"s{step}_lin{i}_qbar_w"])
            assert np.array_equal(q_b, g[f"s{step}_lin{i}_qbar_b"])
            O.weight_update_linear_spec(mlp_w[i][0], mlp_w[i][1], q_w, s_w, q_b, s_b, 0.1)
    for k in range(len(cfg["rows"])):
        assert emb_w[k].tobytes() == g[f"final_emb{k}"].tobytes()
    for i in range(len(cfg["layers"])):
        assert mlp_w[i][0].tobytes() == g[f"final_lin{i}_W"].tobytes()
        assert mlp_w[i][1].tobytes() == g[f"final_lin{i}_b"].tobytes()


def test_single_process_sgd_vs_reference_optimizer():
    """(a10) tests/golden/sgd_single.npz: the reference's QuantEmbeddingBagTwo stepped by torch.optim.SGD on its
    UNCOALESCED sparse gradient (dlrm_s_pytorch_single_gpu.py:1944-1946), three iterations with duplicate rows.  The
    oracle's per-duplicate update (sgd_sparse_spec, storage order) and the coalesced form the CUDA path uses (sum the
    duplicates in the fixed fold order, apply once) both reproduce the reference's tables within 1e-5 (+ 1e-7 absolute:
    elements that cancel to ~0);
    forward scales bit-exact at every step."""
    g = load_golden("sgd_single")
    lr = float(g["lr"])
    W_dup, W_co = g["W_init"].copy(), g["W_init"].copy()
    for s in range(int(g["steps"])):
        sc = O.table_scale_spec(W_co, 4)
        assert np.asarray(sc, dtype=np.float32).tobytes() == g[f"scale{s}"].astype(np.float32).tobytes()
        r0, v0 = O.embbag_backward_spec(g[f"dout{s}"], g[f"idx{s}"], g[f"off{s}"], O.table_scale_spec(W_dup, 4))
        O.sgd_sparse_spec(W_dup, r0, v0, lr)
        np.testing.assert_allclose(W_dup, g[f"W{s}"], rtol=1e-5, atol=1e-7)     # (torch adds a row's duplicates in its own order)
        r1, v1 = O.embbag_backward_spec(g[f"dout{s}"], g[f"idx{s}"], g[f"off{s}"], sc)
        ur, sums = O.coalesce_spec(r1, v1)
        O.weight_update_emb_unquantized_spec(W_co, ur, sums, lr)
        np.testing.assert_allclose(W_co, g[f"W{s}"], rtol=1e-5, atol=1e-7)
