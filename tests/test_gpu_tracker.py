"""GPU: the exact incremental scale tracker against the full rescan -- scales must be
bit-identical after every update, including updates that shrink the current maximum."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _two_groups(rows, dim, seed, policy="incremental"):
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic, tables
    rng = np.random.RandomState(seed)
    Ws = [synthetic.table_weights_numpy(n, dim, rng) for n in rows]
    a = tables.EmbeddingTableGroup([torch.tensor(w, device="cuda") for w in Ws], embedding_bit=4)
    b = tables.EmbeddingTableGroup([torch.tensor(w, device="cuda") for w in Ws], embedding_bit=4)
    b.scale_policy = policy
    return a, b, rng


@pytest.mark.parametrize("policy", ["incremental", "pipelined"])
@pytest.mark.parametrize("dim,quantized", [(16, True), (64, True), (16, False), (8, True)])
def test_tracker_bit_identical_to_full_scan(dim, quantized, policy):
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic, tables
    rows = [3, 70, 1000, 4097, 50000]
    full, inc, rng = _two_groups(rows, dim, 21, policy)
    B = 256
    for step in range(6):
        X, lS_o, lS_i, T = synthetic.criteo_batch(rows, B, seed=70 + step, zipf=1.3 if step % 2 else None)
        if step == 3:                                   # force the arg-max row of every table into the batch
            for k, w in enumerate(full.weights):
                lS_i[k, 0] = int(w.abs().max(dim=1)[0].argmax())
        idx, off, ib, bags = tables.EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cuda")
        dout = torch.tensor(rng.randn(len(rows), B, dim).astype(np.float32) * (5.0 if step >= 3 else 0.05), device="cuda")
        used = []
        for g in (full, inc):
            g.scan_scales()
            used.append(g.scale.clone())             # the scale this step's forward quantises with
            g.forward(idx, off, ib, bags)
            g.backward(dout, world=1)
            if quantized:
                g.exchange(world=1, rank=0)
                g.merge_apply(0.5)
            else:
                g.sgd_apply(0.5)
        assert torch.equal(used[0], used[1]), step
        for wa, wb in zip(full.weights, inc.weights):
            assert torch.equal(wa, wb)
    full.scan_scales(); inc.scan_scales()
    assert torch.equal(full.absmax, inc.absmax) and torch.equal(full.scale, inc.scale) and torch.equal(full.inv_scale, inc.inv_scale)
    # external mutation: a replaced table is detected, an in-place one needs invalidate_tracker()
    torch.cuda.synchronize()
    inc.weights[2].mul_(3.0); full.weights[2].mul_(3.0)
    inc.invalidate_tracker()
    inc.scale_valid = False          # pipelined: forces the bootstrap pass
    full.scan_scales(); inc.scan_scales()
    assert torch.equal(full.scale, inc.scale)


@pytest.mark.parametrize("dim,world", [(16, 1), (16, 3), (64, 8), (128, 2)])
def test_pipelined_scan_shards_cover_every_block(dim, world):
    """dqrm_blockmax_scan / _reduce on every shard of a W-way split: the MAX over the shards' absmax equals the
    full scan bit-for-bit, block maxima equal ATen's per-block abs().max(), and the fix-up honours the shard."""
    from deep_quantized_recommendation_model_dqrm_b200 import _lib, synthetic, tables
    rows = [1, 63, 64, 65, 1000, 4097, 200000]
    rng = np.random.RandomState(3)
    Ws = [torch.tensor(synthetic.table_weights_numpy(n, dim, rng), device="cuda") for n in rows]
    g = tables.EmbeddingTableGroup(Ws, embedding_bit=4)
    g.scan_scales()
    want_absmax = g.absmax.clone()
    lib, st = g.lib, _lib.stream_ptr()
    g._ensure_blockmax()
    g._bm_buf.fill_(-1.0)
    parts = []
    for r in range(world):
        _lib.check(lib.dqrm_blockmax_scan(g.T, g._wptrs(), g._rows_arr, dim, g.block_rows, g._bm_ptrs, r, world, st), "scan")
        am = torch.zeros(g.T, device="cuda")
        _lib.check(lib.dqrm_blockmax_reduce(g.T, g._rows_arr, g.block_rows, g._bm_ptrs, r, world, 4, am.data_ptr(), None,
                                            None, g._scan_ws.data_ptr(), st), "reduce")
        parts.append(am)
    assert torch.equal(torch.stack(parts).max(dim=0)[0], want_absmax)
    for k, (n, W) in enumerate(zip(rows, Ws)):
        nb = (n + 63) // 64
        pad = torch.zeros((nb * 64, dim), device="cuda")
        pad[:n] = W.abs()
        assert torch.equal(g._bm_views[k][:nb], pad.view(nb, -1).max(dim=1)[0]), k
    # unsharded reduce with scale output == the full scan's scale
    sc, inv, am = torch.zeros(g.T, device="cuda"), torch.zeros(g.T, device="cuda"), torch.zeros(g.T, device="cuda")
    _lib.check(lib.dqrm_blockmax_reduce(g.T, g._rows_arr, g.block_rows, g._bm_ptrs, 0, 1, 4, am.data_ptr(), sc.data_ptr(),
                                        inv.data_ptr(), g._scan_ws.data_ptr(), st), "reduce")
    assert torch.equal(sc, g.scale) and torch.equal(inv, g.inv_scale) and torch.equal(am, want_absmax)


def test_pipelined_graph_step_matches_serial_rescan():
    """GraphedTrainStep with the overlapped rescan (two graphs + side stream) against the serial rescan:
    losses, scales and every table bit-identical over several iterations."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import build_cuda_model
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic
    from deep_quantized_recommendation_model_dqrm_b200.graph_step import GraphedTrainStep
    cfg = dict(rows=[50, 3, 100000, 2000, 70000], dim=16, ln_bot=[13, 32, 16], ln_top_hidden=[32, 1])
    B = 64
    batches = [synthetic.criteo_batch(cfg["rows"], B, seed=900 + i, zipf=1.2 if i % 2 else None) for i in range(6)]
    results = {}
    for policy, graph in (("full", True), ("pipelined", True), ("pipelined", False)):
        m = build_cuda_model(cfg, seed=11)
        m._ensure_group().scale_policy = policy
        step = GraphedTrainStep(m, *batches[0], lr=0.5, world_size=1, rank=0, grad_bits=8, warmup=1, use_graph=graph)
        losses = []
        with torch.cuda.stream(step.stream):
            for b in batches:
                step.load(*b)
                step.run()
                losses.append(step.loss.clone())
        torch.cuda.synchronize()
        m.emb_group.check_status()
        if policy == "full":                 # pipelined already holds the post-update scale of the next forward
            m.emb_group.scan_scales()
        results[(policy, graph)] = (torch.stack(losses).cpu(), m.emb_group.scale.clone().cpu(),
                                    [w.detach().clone().cpu() for w in m.emb_group.weights])
    ref = results[("full", True)]
    for key in (("pipelined", True), ("pipelined", False)):
        got = results[key]
        assert torch.equal(ref[0], got[0]), key
        assert torch.equal(ref[1], got[1]), key
        for a, b in zip(ref[2], got[2]):
            assert torch.equal(a, b), key


def test_periodic_scale_update_policy():
    """--scale-update-period P: the scale is refreshed on the first forward and then every P+1 forwards
    (the reference's commented-out bookkeeping, qm:303-315, 354-363)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import C_SMALL, build_cuda_model, emulated_dp_step, shard
    from deep_quantized_recommendation_model_dqrm_b200 import _lib
    m = build_cuda_model(C_SMALL)
    m.scale_update_period = 2
    scans = []
    for step in range(7):
        n0 = _lib.launch_counts["dqrm_table_absmax_scale"]
        emulated_dp_step([m], [shard(1, 0, False, step, C_SMALL["rows"])], lr=0.1, keep_debug=False)
        scans.append(_lib.launch_counts["dqrm_table_absmax_scale"] - n0)
    assert scans == [1, 0, 0, 1, 0, 0, 1]


def test_rwsadagrad_optimizer_class():
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic, tables
    from deep_quantized_recommendation_model_dqrm_b200.optim.rwsadagrad import RWSAdagrad
    rows, dim, B = [500, 7], 16, 64
    rng = np.random.RandomState(8)
    Ws = [synthetic.table_weights_numpy(n, dim, rng) for n in rows]
    g = tables.EmbeddingTableGroup([torch.nn.Parameter(torch.tensor(w, device="cuda")) for w in Ws], embedding_bit=4)
    opt = RWSAdagrad(g.weights, lr=0.05, eps=1e-10)
    mom = opt.attach_table_group(g)
    ref_m = [np.zeros(n, dtype=np.float32) for n in rows]
    ref_W = [w.copy() for w in Ws]
    for step in range(3):
        X, lS_o, lS_i, T = synthetic.criteo_batch(rows, B, seed=30 + step)
        idx, off, ib, bags = tables.EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cuda")
        g.scan_scales(); g.forward(idx, off, ib, bags)
        dout = torch.tensor(rng.randn(2, B, dim).astype(np.float32), device="cuda")
        g.backward(dout, world=1)
        for t in range(2):
            U = int(g.uniq_count[t])
            r = g.uniq_rows[t, :U].cpu().numpy().astype(np.int64)
            v = g.grad_sums[t, :U].cpu().numpy()
            ref_m[t][r] += (v ** 2).mean(axis=1)
            ref_W[t][r] += -0.05 * (v / (np.sqrt(ref_m[t][r]) + 1e-10)[:, None])
        opt.step()
    for t in range(2):
        np.testing.assert_allclose(mom[t].cpu().numpy(), ref_m[t], rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(g.weights[t].detach().cpu().numpy(), ref_W[t], rtol=1e-5, atol=1e-7)


def test_fused_update_through_autograd_and_optimizer():
    """group.enable_fused_update / RWSAdagrad.attach_table_group(fused=True): the autograd backward of the table group
    applies the row update itself (dqrm_embbag_bwd_sgd); tables, row-wise state and the optimizer's schedule come out
    bit-identical to backward + RWSAdagrad.step (sgd_apply), and plain SGD to weight_update_parallel_comm's
    un-quantised branch (sgd_quantized_gradients_parallel_comm.py:626)."""
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic, tables
    from deep_quantized_recommendation_model_dqrm_b200.optim.rwsadagrad import RWSAdagrad
    from deep_quantized_recommendation_model_dqrm_b200.quantization_supp.quant_modules import EmbBagGroupFunction
    rows, dim, B = [500, 7, 30000], 16, 6000           # the 30000-row table takes the one-kernel sort path
    rng = np.random.RandomState(8)
    Ws = [synthetic.table_weights_numpy(n, dim, rng) for n in rows]
    douts = [torch.tensor(rng.randn(3, B, dim).astype(np.float32), device="cuda") for _ in range(3)]
    for adagrad in (True, False):
        res = []
        for fused in (True, False):
            g = tables.EmbeddingTableGroup([torch.nn.Parameter(torch.tensor(w, device="cuda")) for w in Ws], embedding_bit=4)
            g.dp_world, g.dp_rank, g.process_group, g.materialize_grads, g.modules = 1, 0, None, False, None
            if adagrad:
                opt = RWSAdagrad(g.weights, lr=0.05, lr_decay=0.1, eps=1e-10)
                mom = opt.attach_table_group(g, fused=fused)
            elif fused:
                g.enable_fused_update(0.05)
            for step in range(3):
                X, lS_o, lS_i, T = synthetic.criteo_batch(rows, B, seed=30 + step)
                idx, off, ib, bags = tables.EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cuda")
                g.scan_scales()
                out = EmbBagGroupFunction.apply(g, idx, off, ib, bags, False, *g.weights)
                if not adagrad and fused:
                    g.fused_update["lr"] = 0.05 / (step + 1)
                out.backward(douts[step])
                assert g.applied_fused == fused
                if adagrad:
                    opt.step()
                elif fused:
                    g.applied_fused = False
                else:
                    g.sgd_apply(0.05 / (step + 1))
                g.check_status()
            res.append(([w.detach().cpu().numpy().copy() for w in g.weights],
                        [m.cpu().numpy().copy() for m in mom] if adagrad else []))
        for a, b in zip(res[0][0] + res[0][1], res[1][0] + res[1][1]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
