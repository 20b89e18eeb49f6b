"""GPU: the exchange + update half of the iteration on INJECTED gradients against vectors produced by the
reference's own grad_update_parallel_comm / weight_update_parallel_comm on 1/2/4 Gloo ranks
(oracle/make_golden.py xchg_worker).  No forward/backward runs on either side, so neither GEMM rounding nor the
duplicate-fold order of coalesce() enters: gradient scales (sgd:861-866), INT8 codes (:869), merged row sets and
averaged codes (:878,885), per-channel MLP scales / codes (:905-925, 945-957), error-compensation residuals
(:899-900,926-927,938-939,958-959) and every updated weight (:618-622, 642-643) must be BIT-EXACT."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import emulated_exchange
from deep_quantized_recommendation_model_dqrm_b200 import synthetic
from deep_quantized_recommendation_model_dqrm_b200.dense import DenseArena
from deep_quantized_recommendation_model_dqrm_b200.quantization_supp.quant_modules import QuantLinear
from deep_quantized_recommendation_model_dqrm_b200.tables import EmbeddingTableGroup

pytestmark = pytest.mark.gpu


def cpu(t):
    return t.detach().cpu().numpy()


def build_rank(cfg, emb_w, mlp_w, ec):
    grp = EmbeddingTableGroup([torch.tensor(w, device="cuda") for w in emb_w], embedding_bit=4, grad_bit=8)
    grp.keep_debug, grp.fixed_capacity = True, cfg["per_rank"]
    layers = []
    for (n_in, n_out), (W, b) in zip(cfg["layers"], mlp_w):
        LL = torch.nn.Linear(n_in, n_out)
        LL.weight.data, LL.bias.data = torch.tensor(W), torch.tensor(b)
        Q = QuantLinear(weight_bit=4, bias_bit=4, per_channel=True)
        Q.set_param(LL)
        layers.append(Q.cuda())
    arena = DenseArena(layers, torch.device("cuda"))
    arena.error_compensation = ec
    return grp, arena


@pytest.mark.parametrize("name", ["xchg1", "xchg2", "xchg4", "xchg2_ec"])
def test_exchange_on_injected_gradients_vs_reference(name):
    g = load_golden(name)
    world, ec, steps = int(g["world"]), bool(g["ec"]), int(g["steps"])
    cfg = synthetic.XCHG
    T, U, D = len(cfg["rows"]), cfg["per_rank"], cfg["dim"]
    emb_w, mlp_w = synthetic.xchg_weights(cfg)
    ranks = [build_rank(cfg, emb_w, mlp_w, ec) for _ in range(world)]
    groups, arenas = [r[0] for r in ranks], [r[1] for r in ranks]
    exact = world <= 2          # Gloo's fp32 SUM order of the scale all-reduce over > 2 ranks is backend-internal
    off = torch.arange(U, dtype=torch.int64, device="cuda").repeat(T, 1).contiguous()
    for step in range(steps):
        for r, (grp, arena) in enumerate(ranks):
            emb_g, mlp_g = synthetic.injected_grads(cfg, world, r, step)
            idx = torch.from_numpy(np.concatenate([i for i, _ in emb_g])).cuda()
            dout = torch.from_numpy(np.stack([v for _, v in emb_g])).cuda()
            grp.scan_scales()
            grp.forward(idx, off, [k * U for k in range(T + 1)], U)
            grp.backward(dout, world=world, ste_done=True)          # dy = injected values exactly
            arena.zero_grad()
            for l, (gw, gb) in zip(arena.layers, mlp_g):
                l.weight.grad.copy_(torch.from_numpy(gw))
                l.bias.grad.copy_(torch.from_numpy(gb))
        emulated_exchange(groups, arenas)
        for grp, arena in ranks:
            grp.merge_apply(0.1)
            arena.apply(0.1, world=world, quantized=True)
            grp.check_status()
        grp, arena = ranks[0]
        for k in range(T):
            want_s = g[f"s{step}_emb{k}_sbar"].astype(np.float32)
            got_s = cpu(grp.grad_scale_mean[k:k + 1])
            nu = int(grp.updated_count[k])
            rows = cpu(grp.updated_rows[k, :nu])
            order = np.argsort(rows)
            assert np.array_equal(rows[order], g[f"s{step}_emb{k}_rows"])                    # merged row set
            qbar = cpu(grp.qbar[k, :nu])[order]
            if exact:
                assert got_s.tobytes() == want_s.tobytes()                                    # s_bar bit-exact
                assert np.array_equal(qbar, g[f"s{step}_emb{k}_qbar"])                        # (sum of INT8 codes) / N
            else:
                np.testing.assert_allclose(got_s, want_s, rtol=2.5e-7)
                assert np.abs(qbar - g[f"s{step}_emb{k}_qbar"]).max() <= 1.0 / world + 1e-7
        c = 0
        for i, l in enumerate(arena.layers):
            o, n_in = l.weight.shape
            s_w, s_b = cpu(arena.scale_mean[c:c + o]), cpu(arena.scale_mean[c + o:c + o + 1])
            c += o + 1
            qw = cpu(l.weight.grad)            # not overwritten: compare the summed codes instead
            base = l.weight.data.data_ptr() - arena.flat.data_ptr()
            qsum_w = cpu(arena.codes[base // 4: base // 4 + o * n_in]).reshape(o, n_in) * np.float32(1.0 / world)
            bb = l.bias.data.data_ptr() - arena.flat.data_ptr()
            qsum_b = cpu(arena.codes[bb // 4: bb // 4 + o]) * np.float32(1.0 / world)
            if exact:
                assert s_w.tobytes() == g[f"s{step}_lin{i}_s_w"].tobytes()
                assert s_b.tobytes() == g[f"s{step}_lin{i}_s_b"].tobytes()
                assert np.array_equal(qsum_w, g[f"s{step}_lin{i}_qbar_w"])
                assert np.array_equal(qsum_b, g[f"s{step}_lin{i}_qbar_b"])
            else:
                np.testing.assert_allclose(s_w, g[f"s{step}_lin{i}_s_w"], rtol=2.5e-7)
                assert np.abs(qsum_w - g[f"s{step}_lin{i}_qbar_w"]).max() <= 1.0 / world + 1e-7
            if ec:
                for r, (_, ar) in enumerate(ranks):
                    lr_ = ar.layers[i]
                    assert cpu(lr_.error_compensation_weight).tobytes() == g[f"rank{r}_s{step}_lin{i}_ec_w"].tobytes()
                    assert cpu(lr_.error_compensation_bias).tobytes() == g[f"rank{r}_s{step}_lin{i}_ec_b"].tobytes()
    grp, arena = ranks[0]
    for (g2, a2) in ranks[1:]:                                                               # replicas identical
        for w0, w in zip(grp.weights, g2.weights):
            assert torch.equal(w0, w)
        assert torch.equal(arena.flat, a2.flat)
    for k in range(T):
        if exact:
            assert cpu(grp.weights[k]).tobytes() == g[f"final_emb{k}"].tobytes()
        else:
            np.testing.assert_allclose(cpu(grp.weights[k]), g[f"final_emb{k}"], rtol=1e-5, atol=1e-7)
    for i, l in enumerate(arena.layers):
        if exact:
            assert cpu(l.weight).tobytes() == g[f"final_lin{i}_W"].tobytes()
            assert cpu(l.bias).tobytes() == g[f"final_lin{i}_b"].tobytes()
        else:
            np.testing.assert_allclose(cpu(l.weight), g[f"final_lin{i}_W"], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(cpu(l.bias), g[f"final_lin{i}_b"], rtol=1e-5, atol=1e-7)
