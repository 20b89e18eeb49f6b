"""GPU: bit-packed INT4 tables (export format) and the forward that reads them, against the oracle and against
the fp32 QAT forward (bit-identical for one-index bags)."""
import numpy as np
import pytest
import torch

from oracle import dqrm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,P,B", [(16, 1, 256), (64, 1, 128), (128, 1, 64), (16, 6, 64), (64, 40, 48), (128, 16, 32)])
def test_pack_and_forward_int4(dim, P, B):
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic, tables
    rows = [3, 1000, 20011]
    rng = np.random.RandomState(dim + P)
    Ws = [synthetic.table_weights_numpy(n, dim, rng) for n in rows]
    g = tables.EmbeddingTableGroup([torch.tensor(w, device="cuda") for w in Ws], embedding_bit=4)
    lS_i, lS_o = [], []
    for n in rows:
        if P == 1:
            i, o = torch.from_numpy(rng.randint(0, n, size=B).astype(np.int64)), torch.arange(B, dtype=torch.int64)
        else:
            i, o = synthetic.random_bags(n, B, P, rng, fixed=(P >= 16))
        lS_i.append(i); lS_o.append(o)
    idx, off, ib, bags = tables.EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cuda")
    g.scan_scales()
    packed, scale = g.pack_int4()
    out4 = g.forward_int4(idx, off, ib, bags)
    g.check_status()
    for t, n in enumerate(rows):
        s = O.table_scale_spec(Ws[t], 4)
        want_packed = O.pack_int4_spec(Ws[t], s)
        assert packed[t].shape == (n, dim // 2)
        assert np.array_equal(packed[t].cpu().numpy(), want_packed)                      # INT4 codes, bit-exact
        want = O.embbag_forward_int4_spec(want_packed, lS_i[t].numpy(), lS_o[t].numpy(), s)
        assert out4[t].cpu().numpy().tobytes() == want.tobytes()
    if P == 1:                                                                           # == the QAT forward
        out = g.forward(idx, off, ib, bags)
        assert torch.equal(out, out4)
