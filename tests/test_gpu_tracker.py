"""GPU: the exact incremental scale tracker against the full rescan -- scales must be
bit-identical after every update, including updates that shrink the current maximum."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _two_groups(rows, dim, seed):
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic, tables
    rng = np.random.RandomState(seed)
    Ws = [synthetic.table_weights_numpy(n, dim, rng) for n in rows]
    a = tables.EmbeddingTableGroup([torch.tensor(w, device="cuda") for w in Ws], embedding_bit=4)
    b = tables.EmbeddingTableGroup([torch.tensor(w, device="cuda") for w in Ws], embedding_bit=4)
    b.scale_policy = "incremental"
    return a, b, rng


@pytest.mark.parametrize("dim,quantized", [(16, True), (64, True), (16, False)])
def test_tracker_bit_identical_to_full_scan(dim, quantized):
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic, tables
    rows = [3, 70, 1000, 4097, 50000]
    full, inc, rng = _two_groups(rows, dim, 21)
    B = 256
    for step in range(6):
        X, lS_o, lS_i, T = synthetic.criteo_batch(rows, B, seed=70 + step, zipf=1.3 if step % 2 else None)
        if step == 3:                                   # force the arg-max row of every table into the batch
            for k, w in enumerate(full.weights):
                lS_i[k, 0] = int(w.abs().max(dim=1)[0].argmax())
        idx, off, ib, bags = tables.EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cuda")
        dout = torch.tensor(rng.randn(len(rows), B, dim).astype(np.float32) * (5.0 if step >= 3 else 0.05), device="cuda")
        for g in (full, inc):
            g.scan_scales()
            g.forward(idx, off, ib, bags)
            g.backward(dout, world=1)
            if quantized:
                g.exchange(world=1, rank=0)
                g.merge_apply(0.5)
            else:
                g.sgd_apply(0.5)
        assert torch.equal(full.scale, inc.scale), step
        for wa, wb in zip(full.weights, inc.weights):
            assert torch.equal(wa, wb)
    full.scan_scales(); inc.scan_scales()
    assert torch.equal(full.absmax, inc.absmax) and torch.equal(full.scale, inc.scale) and torch.equal(full.inv_scale, inc.inv_scale)
    # external mutation: a replaced table is detected, an in-place one needs invalidate_tracker()
    inc.weights[2].mul_(3.0); full.weights[2].mul_(3.0)
    inc.invalidate_tracker()
    full.scan_scales(); inc.scan_scales()
    assert torch.equal(full.scale, inc.scale)
