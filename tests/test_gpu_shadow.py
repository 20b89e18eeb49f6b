"""GPU: packed-INT4 shadow rows in the training forward (csrc/shadow.cu, north-star kernel 2) -- the forward that reads
4-bit codes for one-index bags must return the SAME bits as the fp32-row forward (reference semantics: pool, then
quantise, quant_modules_not_quantize_grad.py:367,378,393), through scale changes, row updates and multi-hot bags."""
import numpy as np
import pytest
import torch

from helpers import C_SMALL, build_cuda_model
from deep_quantized_recommendation_model_dqrm_b200 import synthetic
from deep_quantized_recommendation_model_dqrm_b200 import dlrm_s_pytorch_comm_grad as drv
from deep_quantized_recommendation_model_dqrm_b200.graph_step import GraphedTrainStep
from deep_quantized_recommendation_model_dqrm_b200.tables import EmbeddingTableGroup
from oracle import dqrm_oracle as O

pytestmark = pytest.mark.gpu


def cpu(t):
    return t.detach().cpu().numpy()


def _group(rows, dim, seed, shadow):
    rng = np.random.RandomState(seed)
    g = EmbeddingTableGroup([torch.tensor(synthetic.table_weights_numpy(n, dim, rng), device="cuda") for n in rows], embedding_bit=4)
    if shadow:
        g.enable_shadow()
    return g


@pytest.mark.parametrize("dim", [16, 64])
def test_shadow_forward_bits_equal_fp32_forward(dim):
    rows = [3, 40, 1460, 5000, 100000]          # staged in shared memory (<= 32 KiB packed) and not
    a, b = _group(rows, dim, 9, False), _group(rows, dim, 9, True)
    for seed, zipf in ((1, None), (2, 1.2)):
        X, lS_o, lS_i, T = synthetic.criteo_batch(rows, 300, seed=seed, zipf=zipf)
        idx, off, ib, bags = EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cuda")
        for g in (a, b):
            g.scan_scales()
        oa, ob = a.forward(idx, off, ib, bags), b.forward(idx, off, ib, bags)
        assert torch.equal(oa, ob) and torch.equal(a.codes, b.codes)
        for g in (a, b):
            g.check_status()
    # the shadow holds exactly the codes of the oracle's packing, for every table
    for t, n in enumerate(rows):
        want = O.pack_int4_spec(cpu(b.weights[t]), cpu(b.scale[t]))
        assert np.array_equal(cpu(b.shadow[t]), want)
    # multi-hot bags (incl. empty ones) fall back to the fp32 rows bag by bag
    X, lS_o, lS_i, T = synthetic.random_batch(rows, 64, 5, seed=3)
    lS_o[1][10:14] = lS_o[1][10]
    idx, off, ib, bags = EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cuda")
    oa, ob = a.forward(idx, off, ib, bags), b.forward(idx, off, ib, bags)
    assert torch.equal(oa, ob) and torch.equal(a.codes, b.codes)


def test_shadow_survives_updates_and_scale_changes():
    """Several training iterations with and without the shadow: losses, tables, MLPs bit-identical; after every step
    the shadow of every table whose scale is unchanged equals a fresh packing of the updated fp32 rows."""
    ma, mb = build_cuda_model(C_SMALL, seed=21), build_cuda_model(C_SMALL, seed=21)
    gb = mb._ensure_group()
    gb.enable_shadow()
    changed = 0
    prev = None
    for step in range(8):
        X, lS_o, lS_i, T = synthetic.criteo_batch(C_SMALL["rows"], 64, seed=700 + step, zipf=1.2 if step % 2 else None)
        la = drv.train_iteration(ma, X, lS_o, lS_i, T, lr=0.3)
        lb = drv.train_iteration(mb, X, lS_o, lS_i, T, lr=0.3)
        assert torch.equal(la, lb)
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.equal(pa, pb)
        gb.check_status()
        sc = cpu(gb.scale).copy()
        if prev is not None:
            changed += int((sc != prev).sum())
        prev = sc
        # rows updated by this step were re-encoded with this step's scale
        for t in range(gb.T):
            assert cpu(gb.shadow_scale[t]).tobytes() == cpu(gb.scale[t]).tobytes()
            want = O.pack_int4_spec(cpu(gb.weights[t]), cpu(gb.scale[t]))
            assert np.array_equal(cpu(gb.shadow[t]), want), (step, t)
    assert changed > 0, "the test must see at least one scale change (full re-encode path)"


def test_shadow_in_captured_graph_step():
    cfg = C_SMALL
    ma, mb = build_cuda_model(cfg, seed=22), build_cuda_model(cfg, seed=22)
    mb._ensure_group().enable_shadow()
    b0 = [t.cuda() for t in synthetic.criteo_batch(cfg["rows"], 32, seed=800)]
    sa = GraphedTrainStep(ma, *b0, lr=0.2, warmup=1, use_graph=True)
    sb = GraphedTrainStep(mb, *b0, lr=0.2, warmup=1, use_graph=True)
    for step in range(5):
        bt = [t.cuda() for t in synthetic.criteo_batch(cfg["rows"], 32, seed=801 + step, zipf=1.3)]
        for s in (sa, sb):
            with torch.cuda.stream(s.stream):
                s.load(*bt)
                s.run()
        torch.cuda.synchronize()
        assert torch.equal(sa.loss, sb.loss)
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.equal(pa, pb)
    mb.emb_group.check_status()
