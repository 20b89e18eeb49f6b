"""GPU: ONE full training iteration at the BASELINE.json config sizes through the object bench.py times
(GraphedTrainStep: scan + CUDA-graph replay) against the oracle's restatement of the reference iteration
(dlrm_s_pytorch_comm_grad.py:1909-1957) on the SAME tables -- the device arena is copied to the host.
  configs[1]  Kaggle shape, 26 tables / 33.76 M rows / 2.16 GB, dim 16, batch 128
  configs[3]  Terabyte shape (dim 64, bot 13-512-256-64, top 512-512-256-1), batch 8192 -- the batch > 2048 MLP
              branch -- with the row counts capped so the CPU oracle finishes in seconds
Bars: loss <= 1e-5 relative, every table scale bit-exact, updated-row sets equal, updated rows and MLP
parameters <= 1e-5."""
import numpy as np
import pytest
import torch

from deep_quantized_recommendation_model_dqrm_b200 import synthetic
from deep_quantized_recommendation_model_dqrm_b200 import dlrm_s_pytorch_comm_grad as drv
from deep_quantized_recommendation_model_dqrm_b200.graph_step import GraphedTrainStep
from oracle import dqrm_oracle as O

pytestmark = pytest.mark.gpu


def cpu(t):
    return t.detach().cpu().numpy()


def _mlp_params(layers):
    return [(cpu(l.weight).copy(), cpu(l.bias).copy()) for l in layers if isinstance(l, drv.QuantLinear)]


def _close_or_one_code(got, want, step, what):
    """<= 1e-5 everywhere, except that a gradient within fp32 rounding of a quantisation boundary may land on the
    neighbouring INT8 code (fp32 sums of thousands of terms in another order than the CPU BLAS): such entries
    differ by exactly one code step = lr * s_bar, and there must be few of them."""
    diff = np.abs(got - want)
    bad = diff > 1e-5 * np.abs(want) + 2e-6
    if bad.any():
        assert bad.mean() <= 0.02, f"{what}: {bad.mean():.3%} of the entries off"
        assert diff[bad].max() <= 1.01 * step + 2e-6, f"{what}: off by {diff[bad].max():.3e} > one code step {step:.3e}"


def _one_step_vs_oracle(cfg, B, zipf=None, use_graph=True):
    rows, dim = cfg["rows"], cfg["dim"]
    ln_top = synthetic.top_mlp_sizes(len(rows), dim, cfg["ln_top_hidden"])
    np.random.seed(123)
    m = drv.DLRM_Net(dim, np.array(rows), np.array(cfg["ln_bot"]), np.array(ln_top), arch_interaction_op="dot",
                     sigmoid_bot=-1, sigmoid_top=len(ln_top) - 2, loss_function="bce", quantization_flag=True,
                     embedding_bit=4, weight_bit=4, quantize_act_and_lin=True, mlp_channelwise=True,
                     quantize_activation=False, device="cuda", table_seed=1234)
    g = m._ensure_group()
    g.keep_debug = True                                   # the captured merge also emits the updated-row lists
    b0 = [t.cuda() for t in synthetic.criteo_batch(rows, B, seed=41, zipf=zipf)]
    step = GraphedTrainStep(m, *b0, lr=0.1, warmup=1, use_graph=use_graph)     # warm-up iterations update the model
    torch.cuda.synchronize()
    # snapshot AFTER the warm-up: this is the state the checked iteration starts from
    om = O.OracleDLRM(rows, dim, _mlp_params(m.bot_l), _mlp_params(m.top_l),
                      emb_weights=[e.embedding_bag.weight.detach().cpu() for e in m.emb_l])
    batch = synthetic.criteo_batch(rows, B, seed=42, zipf=zipf)
    with torch.cuda.stream(step.stream):
        step.load(*[t.cuda() for t in batch])
        loss = step.run()
    torch.cuda.synchronize()
    g.check_status()
    want = O.train_step_torch([om], [batch], lr=0.1)[0]
    got = float(loss)
    assert abs(got - want) <= 1e-5 * abs(want) + 1e-6, (got, want)
    for k in range(g.T):
        assert cpu(g.scale[k]).tobytes() == om.emb_l[k].eb_scaling_factor.numpy().tobytes(), k     # (a1) bit-exact
        nu = int(g.updated_count[k])
        upd = np.sort(cpu(g.updated_rows[k, :nu]).astype(np.int64))
        assert np.array_equal(upd, om.emb_l[k].grad_rows.numpy()), k                                # updated-row set
        assert np.array_equal(upd, np.unique(batch[2][k].numpy()))
        idx = torch.from_numpy(upd)
        Wg = m.emb_l[k].embedding_bag.weight.detach()[idx.cuda()].cpu().numpy()
        Wo = om.emb_l[k].embedding_bag.weight.data[idx].numpy()
        _close_or_one_code(Wg, Wo, 0.1 * float(g.grad_scale_mean[k]), f"table {k}")
    for mine, theirs in ((m.bot_l, om.bot_l), (m.top_l, om.top_l)):
        mine = [l for l in mine if isinstance(l, drv.QuantLinear)]
        for l, ol in zip(mine, theirs):
            _close_or_one_code(cpu(l.weight), ol.weight.data.numpy(), 0.1 * float(l.weight_scaling_factor.max()), "MLP weight")
            _close_or_one_code(cpu(l.bias), ol.bias.data.numpy(), 0.1 * float(l.bias_scaling_factor), "MLP bias")
    del step, m, g, om
    torch.cuda.empty_cache()


def test_one_step_kaggle_config_full_size_vs_oracle():
    """BASELINE configs[1] at full size (the bench headline workload)."""
    _one_step_vs_oracle(synthetic.KAGGLE, 128)


def test_one_step_kaggle_full_size_batch_2048_vs_oracle():
    """BASELINE configs[2] per-step shape on one rank: Kaggle tables, batch 2048 (duplicates on the small tables)."""
    _one_step_vs_oracle(synthetic.KAGGLE, 2048)


def test_one_step_terabyte_shape_batch_8192_vs_oracle():
    """BASELINE configs[3] shape -- dim 64, Terabyte MLPs, batch 8192 (the batch > 2048 MLP branch; thousands of
    duplicate lookups on the tiny tables) -- with table rows capped at 200k."""
    cfg = dict(synthetic.TERABYTE)
    cfg["rows"] = [min(n, 200_000) for n in cfg["rows"]]
    _one_step_vs_oracle(cfg, 8192)
