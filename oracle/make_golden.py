"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE in this container.

TEST INFRASTRUCTURE ONLY.  Run here (the build container has the reference at
/root/reference; the GPU box does not):

    python oracle/make_golden.py

The reference has no tests or golden vectors of its own, so the
vectors written here -- outputs of the reference's own modules on seeded
inputs -- are what pins the oracle (oracle/dqrm_oracle.py) and, through it, the
CUDA path.  Inputs that are cheap to store are stored; large inputs are
regenerated from their seed through ``synthetic.py`` (numpy RandomState streams
are stable across versions).

Reference entry points exercised (file:line):
  QuantEmbeddingBagTwo.forward / backward      quantization_supp/quant_modules_not_quantize_grad.py:317-395
  symmetric_linear_quantization_param_two      quantization_supp/quant_utils.py:141-194
  SymmetricQuantFunction                       quantization_supp/quant_utils.py:316-363
  QuantLinear.forward / backward               quantization_supp/quant_modules_not_quantize_grad.py:105-211
  clear_gradients, grad_update_parallel_comm, weight_update_parallel_comm, weight_syncc
                                               sgd_quantized_gradients_parallel_comm.py:714,257,601,963
  DLRM_Net.forward / interact_features         dlrm_s_pytorch_comm_grad.py:809,701
  RWSAdagrad.step (sparse branch)              optim/rwsadagrad.py:97-113
The only modification is the ``.cuda()`` no-op shim (quant_utils.py:336 hard-codes
``.cuda()``).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

sys.path.insert(0, ROOT)
from deep_quantized_recommendation_model_dqrm_b200 import synthetic  # noqa: E402


def import_reference():
    torch.Tensor.cuda = lambda self, *a, **k: self          # the shim
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import quantization_supp.quant_modules_not_quantize_grad as qm
    import quantization_supp.quant_utils as qu
    import sgd_quantized_gradients_parallel_comm as sgd
    return qm, qu, sgd


def _np(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------- embeddings
def emb_case(qm, name, rows, dim, bits, idx, off, seed):
    rng = np.random.RandomState(seed)
    W = synthetic.table_weights_numpy(rows, dim, rng)
    E = qm.QuantEmbeddingBagTwo(rows, dim, bits, embedding_id=0)
    E.embedding_bag.weight.data = torch.tensor(W, requires_grad=True)
    out = E(idx, off)
    dout = torch.from_numpy(rng.randn(off.shape[0], dim).astype(np.float32))
    out.backward(dout)
    g = E.embedding_bag.weight.grad
    gc = g.coalesce()
    # full-precision branch (qm:395)
    E2 = qm.QuantEmbeddingBagTwo(rows, dim, bits, embedding_id=0)
    E2.embedding_bag.weight.data = torch.tensor(W, requires_grad=True)
    pooled = E2(idx, off, full_precision_flag=True)
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        rows=rows, dim=dim, bits=bits, seed=seed, W=W, idx=_np(idx), off=_np(off), dout=_np(dout),
        scale=_np(E.eb_scaling_factor), codes=_np(E.output_integer), out=_np(out), pooled=_np(pooled),
        grad_rows=_np(g._indices()[0]), grad_vals=_np(g._values()),
        co_rows=_np(gc.indices()[0]), co_vals=_np(gc.values()))


def gen_embeddings(qm):
    rng = np.random.RandomState(11)
    idx, off = synthetic.random_bags(1000, 32, 10, rng)
    emb_case(qm, "emb_multihot_d16", 1000, 16, 4, idx, off, seed=101)
    idx, off = synthetic.random_bags(300, 24, 6, rng)
    emb_case(qm, "emb_multihot_d64_b8", 300, 64, 8, idx, off, seed=102)
    # Criteo-shaped: one index per bag, heavy duplication (zipf)
    B = 128
    idx = torch.from_numpy(np.minimum(rng.zipf(1.2, size=B) - 1, 4999).astype(np.int64))
    emb_case(qm, "emb_onehot_zipf_d16", 5000, 16, 4, idx, torch.arange(B, dtype=torch.int64), seed=103)
    # tiny table (3 rows, as Kaggle table 8), every bag hits one of 3 rows
    idx = torch.from_numpy(rng.randint(0, 3, size=64).astype(np.int64))
    emb_case(qm, "emb_onehot_tiny_d16", 3, 16, 4, idx, torch.arange(64, dtype=torch.int64), seed=104)
    # ragged: empty bags in the middle and at the end
    idx = torch.tensor([5, 1, 1, 7, 7, 7, 2], dtype=torch.int64)
    off = torch.tensor([0, 0, 2, 3, 3, 6, 7, 7], dtype=torch.int64)
    emb_case(qm, "emb_ragged_d16", 9, 16, 4, idx, off, seed=105)


# ---------------------------------------------------------------- single-process SGD on the sparse gradient (a10)
def gen_sgd_single(qm):
    """The reference's single-process update (dlrm_s_pytorch_single_gpu.py:1944-1946): the reference's own
    QuantEmbeddingBagTwo (forward, STE backward, UNCOALESCED sparse gradient) stepped by torch.optim.SGD for three
    iterations on multi-hot bags with duplicate rows; tables after every step -> tests/golden/sgd_single.npz."""
    rng = np.random.RandomState(31)
    rows, dim, bits, B, lr = 400, 16, 4, 96, 0.1
    W = synthetic.table_weights_numpy(rows, dim, rng)
    E = qm.QuantEmbeddingBagTwo(rows, dim, bits, embedding_id=0)
    E.embedding_bag.weight.data = torch.tensor(W)
    opt = torch.optim.SGD(E.parameters(), lr=lr)
    rec = dict(rows=rows, dim=dim, bits=bits, lr=lr, steps=3, W_init=W)
    for step in range(3):
        idx, off = synthetic.random_bags(rows, B, 6, rng)
        dout = rng.randn(B, dim).astype(np.float32)
        opt.zero_grad()
        out = E(idx, off)
        out.backward(torch.from_numpy(dout))
        assert E.embedding_bag.weight.grad.is_sparse and not E.embedding_bag.weight.grad.is_coalesced()
        opt.step()
        rec[f"idx{step}"], rec[f"off{step}"], rec[f"dout{step}"] = _np(idx), _np(off), dout
        rec[f"scale{step}"] = _np(E.eb_scaling_factor).copy()
        rec[f"W{step}"] = _np(E.embedding_bag.weight.data).copy()
    np.savez_compressed(os.path.join(GOLD, "sgd_single.npz"), **rec)


# ---------------------------------------------------------------- QuantLinear
def gen_linear(qm):
    rng = np.random.RandomState(21)
    for name, (n_in, n_out, B, bits) in {"linear_13_64": (13, 64, 16, 4), "linear_367_32": (367, 32, 8, 4),
                                         "linear_64_1": (64, 1, 8, 4)}.items():
        (W, b), = synthetic.mlp_params([n_in, n_out], rng)
        x = rng.randn(B, n_in).astype(np.float32)
        LL = torch.nn.Linear(n_in, n_out)
        LL.weight.data = torch.tensor(W)
        LL.bias.data = torch.tensor(b)
        Q = qm.QuantLinear(weight_bit=bits, bias_bit=bits, full_precision_flag=False, per_channel=True,
                           quantize_activation=False)
        Q.set_param(LL)
        xt = torch.tensor(x, requires_grad=True)
        y, _ = Q(xt)
        dy = torch.from_numpy(rng.randn(B, n_out).astype(np.float32))
        y.backward(dy)
        np.savez_compressed(
            os.path.join(GOLD, name + ".npz"), W=W, b=b, x=x, dy=_np(dy), bits=bits,
            scale=_np(Q.fc_scaling_factor), W_int=_np(Q.weight_integer), b_int=_np(Q.bias_integer).reshape(-1),
            y=_np(y), dx=_np(xt.grad), dW=_np(Q.weight.grad), db=_np(Q.bias.grad))


# ---------------------------------------------------------------- full model / DP step
C_SMALL = dict(rows=[50, 3, 1000, 200], dim=16, ln_bot=[13, 32, 16], ln_top_hidden=[32, 1])
# duplicate-free shards (tables >> per-rank batch, indices drawn without replacement inside a rank, overlapping
# between ranks): coalesce() has nothing to fold, so the reference's gradient scales and INT8 codes are
# order-independent and must be reproduced BIT-EXACTLY (sgd_quantized_gradients_parallel_comm.py:861-869)
C_NODUP = dict(rows=[1500, 40, 1000, 200], dim=16, ln_bot=[13, 32, 16], ln_top_hidden=[32, 1])


def build_reference_model(cfg, seed, embedding_bit=4, weight_bit=4):
    """Reference DLRM_Net with weights overwritten from synthetic.py streams."""
    import dlrm_s_pytorch_comm_grad as drv
    drv.full_precision_flag = False                       # train(): = args.pretrain_and_quantize (drv:1425-1426)
    rows, dim = cfg["rows"], cfg["dim"]
    ln_bot = np.array(cfg["ln_bot"])
    ln_top = np.array(synthetic.top_mlp_sizes(len(rows), dim, cfg["ln_top_hidden"]))
    m = drv.DLRM_Net(dim, np.array(rows), ln_bot, ln_top, arch_interaction_op="dot", sigmoid_bot=-1,
                     sigmoid_top=ln_top.size - 2, ndevices=-1, loss_function="bce", quantization_flag=True,
                     embedding_bit=embedding_bit, weight_bit=weight_bit, quantize_act_and_lin=True,
                     mlp_channelwise=True, quantize_activation=False)
    rng = np.random.RandomState(seed)
    for k, n in enumerate(rows):
        m.emb_l[k].embedding_bag.weight.data = torch.tensor(synthetic.table_weights_numpy(n, dim, rng),
                                                            requires_grad=True)
    for layers, ln in ((m.bot_l, ln_bot), (m.top_l, ln_top)):
        qls = [l for l in layers if hasattr(l, "weight_bit")]
        for l, (W, b) in zip(qls, synthetic.mlp_params(ln, rng)):
            l.weight.data = torch.tensor(W)
            l.bias.data = torch.tensor(b)
    return m


def model_state(m):
    d = {}
    for k, E in enumerate(m.emb_l):
        d[f"emb{k}"] = _np(E.embedding_bag.weight.data).copy()
    for g, layers in (("bot", m.bot_l), ("top", m.top_l)):
        for i, l in enumerate([l for l in layers if hasattr(l, "weight_bit")]):
            d[f"{g}{i}_W"] = _np(l.weight.data).copy()
            d[f"{g}{i}_b"] = _np(l.bias.data).copy()
    return d


def dp_worker(rank, world, port, steps, multihot, out_path, emb_q=True):
    """One rank of the reference's custom DP loop (drv:1909-1957) on CPU/Gloo."""
    import torch.distributed as dist
    qm, qu, sgd = import_reference()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nodup = multihot == "nodup"
    multihot = multihot is True
    cfg = C_NODUP if nodup else C_SMALL
    m = build_reference_model(cfg, seed=300)
    lr = 0.1
    rec = {}
    for step in range(steps):
        Bg = 16 * world
        if nodup:
            X, lS_o, lS_i, T = synthetic.criteo_batch_nodup(cfg["rows"], world, 16, seed=400 + step)
        elif multihot:
            X, lS_o, lS_i, T = synthetic.random_batch(cfg["rows"], Bg, 4, seed=400 + step)
        else:
            X, lS_o, lS_i, T = synthetic.criteo_batch(cfg["rows"], Bg, seed=400 + step, zipf=1.3)
        sl = slice(rank * 16, (rank + 1) * 16)
        if multihot:   # shard the bags of each table
            li, lo = [], []
            for i, o in zip(lS_i, lS_o):
                ends = torch.cat([o[1:], torch.tensor([i.shape[0]])])
                a, b = int(o[sl][0]), int(ends[sl][-1])
                li.append(i[a:b])
                lo.append(o[sl] - a)
            lS_i, lS_o = li, lo
        else:
            lS_i = lS_i[:, sl]
            lS_o = lS_o[:, 0:lS_i.shape[1]]
        Z = m(X[sl], lS_o, lS_i)
        E = torch.nn.BCELoss(reduction="mean")(Z, T[sl])
        sgd.clear_gradients(m)
        E.backward()
        sgd.grad_update_parallel_comm(m, world, emb_grad_quantized=emb_q, num_bits=8, ranking_range=False,
                                      rank_for_debug=rank, iteration_count=step)
        rec[f"loss{step}"] = float(E.detach())
        for k, Etab in enumerate(m.emb_l):
            g = Etab.embedding_bag.weight.grad.coalesce()
            rec[f"s{step}_emb{k}_rows"] = _np(g.indices()[0]).copy()
            rec[f"s{step}_emb{k}_qbar"] = _np(g.values()).copy()     # emb_q=False: the averaged fp32 row gradients
            rec[f"s{step}_emb{k}_sbar"] = _np(Etab.emb_scaling_factor).copy()
            rec[f"s{step}_emb{k}_eb_scale"] = _np(Etab.eb_scaling_factor).copy()
        sgd.weight_update_parallel_comm(m, lr, emb_grad_quantized=emb_q, update_embedding=True, num_gpus=world,
                                        rank_for_debug=rank)
    for k, v in model_state(m).items():
        rec["final_" + k] = v
    np.savez_compressed(out_path.format(rank=rank), **rec)
    dist.destroy_process_group()


def gen_dp(world, multihot, name, port, emb_q=True):
    import torch.multiprocessing as mp
    tmp = os.path.join(GOLD, name + "_rank{rank}.npz")
    mp.spawn(dp_worker, args=(world, port, 2, multihot, tmp, emb_q), nprocs=world, join=True)
    # ranks must agree on everything except the per-rank loss; keep rank 0 + all losses
    recs = [dict(np.load(tmp.format(rank=r))) for r in range(world)]
    for r in range(1, world):
        for k in recs[0]:
            if not k.startswith("loss"):
                assert np.array_equal(recs[0][k], recs[r][k]), (name, r, k)
    out = dict(recs[0])
    for r in range(world):
        for k in recs[r]:
            if k.startswith("loss"):
                out[f"rank{r}_{k}"] = recs[r][k]
        os.remove(tmp.format(rank=r))
    out["world"] = world
    out["multihot"] = multihot is True
    out["nodup"] = multihot == "nodup"
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)


def xchg_worker(rank, world, port, steps, ec, out_path):
    """The exchange + update half of the reference's iteration on injected gradients (no forward/backward, so
    no GEMM rounding enters): sgd...parallel_comm.py grad_update_parallel_comm :257 + weight_update_parallel_comm
    :601 on CPU/Gloo.  ec=True: the MLP tensors go through quantize_linear_grad / quantize_bias_grad with
    err_compensation=True (:899-900,926-927,938-939,958-959) in the way grad_update_parallel_comm calls them."""
    import torch.distributed as dist
    qm, qu, sgd = import_reference()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cfg = synthetic.XCHG
    emb_w, mlp_w = synthetic.xchg_weights(cfg)

    class Model(torch.nn.Module):
        pass
    m = Model()
    m.emb_l = torch.nn.ModuleList()
    for n, W in zip(cfg["rows"], emb_w):
        E = qm.QuantEmbeddingBagTwo(n, cfg["dim"], 4, embedding_id=len(m.emb_l))
        E.embedding_bag.weight.data = torch.tensor(W)
        m.emb_l.append(E)
    m.bot_l = torch.nn.ModuleList()
    for (n_in, n_out), (W, b) in zip(cfg["layers"], mlp_w):
        LL = torch.nn.Linear(n_in, n_out)
        LL.weight.data, LL.bias.data = torch.tensor(W), torch.tensor(b)
        Q = qm.QuantLinear(weight_bit=4, bias_bit=4, per_channel=True)
        Q.set_param(LL)
        m.bot_l.append(Q)
    m.top_l = torch.nn.ModuleList()
    lr, rec = 0.1, {}
    for step in range(steps):
        emb_g, mlp_g = synthetic.injected_grads(cfg, world, rank, step)
        for E, (idx, vals), n in zip(m.emb_l, emb_g, cfg["rows"]):
            E.embedding_bag.weight.grad = torch.sparse_coo_tensor(torch.from_numpy(idx)[None], torch.from_numpy(vals),
                                                                  size=(n, cfg["dim"]))
        for Q, (gw, gb) in zip(m.bot_l, mlp_g):
            Q.weight.grad, Q.bias.grad = torch.from_numpy(gw.copy()), torch.from_numpy(gb.copy())
        if not ec:
            sgd.grad_update_parallel_comm(m, world, emb_grad_quantized=True, num_bits=8, ranking_range=False,
                                          rank_for_debug=rank, iteration_count=step)
        else:
            mlp_saved = [(Q.weight.grad.clone(), Q.bias.grad.clone()) for Q in m.bot_l]
            bot, m.bot_l = m.bot_l, torch.nn.ModuleList()           # embeddings through the stock entry point
            sgd.grad_update_parallel_comm(m, world, emb_grad_quantized=True, num_bits=8, ranking_range=False,
                                          rank_for_debug=rank, iteration_count=step)
            m.bot_l = bot
            with torch.no_grad():
                for Q in m.bot_l:                                   # the body of :341-356 with err_compensation=True
                    upd, sc = sgd.quantize_linear_grad(Q, num_bits=8, parallel=True, num_gpus=world, err_compensation=True)
                    Q.weight_scaling_factor = sc
                    Q.weight.grad.zero_(); Q.weight.grad.add_(upd)
                    upd, sc = sgd.quantize_bias_grad(Q, num_bits=8, parallel=True, num_gpus=world, err_compensation=True)
                    Q.bias_scaling_factor = sc
                    Q.bias.grad.zero_(); Q.bias.grad.add_(upd)
        for k, E in enumerate(m.emb_l):
            g = E.embedding_bag.weight.grad.coalesce()
            rec[f"s{step}_emb{k}_rows"] = _np(g.indices()[0]).copy()
            rec[f"s{step}_emb{k}_qbar"] = _np(g.values()).copy()
            rec[f"s{step}_emb{k}_sbar"] = _np(E.emb_scaling_factor).copy()
        for i, Q in enumerate(m.bot_l):
            rec[f"s{step}_lin{i}_qbar_w"] = _np(Q.weight.grad).copy()
            rec[f"s{step}_lin{i}_qbar_b"] = _np(Q.bias.grad).copy()
            rec[f"s{step}_lin{i}_s_w"] = _np(Q.weight_scaling_factor).copy()
            rec[f"s{step}_lin{i}_s_b"] = _np(Q.bias_scaling_factor).reshape(-1).copy()
            if ec:
                rec[f"s{step}_lin{i}_ec_w"] = _np(Q.error_compensation_weight).copy()
                rec[f"s{step}_lin{i}_ec_b"] = _np(Q.error_compensation_bias).copy()
        sgd.weight_update_parallel_comm(m, lr, emb_grad_quantized=True, update_embedding=True, num_gpus=world,
                                        rank_for_debug=rank)
    for k, E in enumerate(m.emb_l):
        rec[f"final_emb{k}"] = _np(E.embedding_bag.weight.data).copy()
    for i, Q in enumerate(m.bot_l):
        rec[f"final_lin{i}_W"], rec[f"final_lin{i}_b"] = _np(Q.weight.data).copy(), _np(Q.bias.data).copy()
    np.savez_compressed(out_path.format(rank=rank), **rec)
    dist.destroy_process_group()


def gen_xchg(world, name, port, ec=False, steps=2):
    import torch.multiprocessing as mp
    tmp = os.path.join(GOLD, name + "_rank{rank}.npz")
    mp.spawn(xchg_worker, args=(world, port, steps, ec, tmp), nprocs=world, join=True)
    recs = [dict(np.load(tmp.format(rank=r))) for r in range(world)]
    for r in range(1, world):
        for k in recs[0]:
            if "_ec_" in k:
                continue          # the residual is per rank (local grad - global update): keep every rank's
            assert np.array_equal(recs[0][k], recs[r][k]), (name, r, k)
    out = dict(recs[0])
    for r in range(world):
        for k in recs[r]:
            if "_ec_" in k:
                out[f"rank{r}_{k}"] = recs[r][k]
        os.remove(tmp.format(rank=r))
    out = {k: v for k, v in out.items() if "_ec_" not in k or k.startswith("rank")}
    out["world"], out["ec"], out["steps"] = world, ec, steps
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)


def gen_interact():
    import dlrm_s_pytorch_comm_grad as drv

    class _Self:
        modify_feature_interaction = False
        arch_interaction_op = "dot"
        arch_interaction_itself = False
        quantization_flag = True
        quantize_activation = False
    rng = np.random.RandomState(31)
    for name, (B, F, D) in {"interact_kaggle": (8, 26, 16), "interact_tb": (4, 26, 64), "interact_small": (5, 3, 16)}.items():
        x = torch.tensor(rng.randn(B, D).astype(np.float32), requires_grad=True)
        ly = [torch.tensor(rng.randn(B, D).astype(np.float32), requires_grad=True) for _ in range(F)]
        R, _ = drv.DLRM_Net.interact_features(_Self(), x, ly)
        dR = torch.from_numpy(rng.randn(*R.shape).astype(np.float32))
        R.backward(dR)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), x=_np(x), ly=np.stack([_np(t) for t in ly]),
                            R=_np(R), dR=_np(dR), dx=_np(x.grad), dly=np.stack([_np(t.grad) for t in ly]))


def gen_rwsadagrad():
    """optim/rwsadagrad.py (the reference's CPU-only row-wise sparse Adagrad) on sparse gradients with
    duplicate rows: 4 steps, state and weights after every step."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from optim.rwsadagrad import RWSAdagrad
    rng = np.random.RandomState(41)
    rows, dim, lr, eps = 60, 16, 0.05, 1e-10
    W0 = synthetic.table_weights_numpy(rows, dim, rng)
    W = torch.nn.Parameter(torch.tensor(W0))
    opt = RWSAdagrad([W], lr=lr, eps=eps)
    rec = dict(W_init=W0, rows=rows, dim=dim, lr=lr, eps=eps, steps=4)
    for step in range(4):
        idx = rng.randint(0, rows, size=24).astype(np.int64)
        vals = (rng.randn(24, dim) * 10.0 ** rng.uniform(-3, 0)).astype(np.float32)
        W.grad = torch.sparse_coo_tensor(torch.from_numpy(idx)[None], torch.from_numpy(vals), size=(rows, dim))
        opt.step()
        rec[f"idx{step}"], rec[f"vals{step}"] = idx, vals
        rec[f"W{step}"] = _np(W.data).copy()
        rec[f"m{step}"] = _np(opt.state[W]["momentum"]).copy()
    np.savez_compressed(os.path.join(GOLD, "rwsadagrad_rows.npz"), **rec)


def main():
    os.makedirs(GOLD, exist_ok=True)
    if "--only-rwsadagrad" in sys.argv:
        gen_rwsadagrad()
        return
    if "--only-nodup" in sys.argv:
        import_reference()
        gen_dp(2, "nodup", "dp2_nodup", 29616)
        return
    if "--only-xchg" in sys.argv:
        import_reference()
        gen_xchg(1, "xchg1", 29620)
        gen_xchg(2, "xchg2", 29618)
        gen_xchg(4, "xchg4", 29619)
        gen_xchg(2, "xchg2_ec", 29621, ec=True, steps=3)
        return
    if "--only-sgd-single" in sys.argv:
        qm, qu, sgd = import_reference()
        gen_sgd_single(qm)
        return
    if "--only-unquantized" in sys.argv:
        import_reference()
        gen_dp(2, False, "dp2_unquantized", 29615, emb_q=False)
        return
    qm, qu, sgd = import_reference()
    gen_embeddings(qm)
    gen_linear(qm)
    gen_interact()
    gen_dp(1, False, "dp1_onehot", 29611)
    gen_dp(2, False, "dp2_onehot", 29612)
    gen_dp(2, True, "dp2_multihot", 29613)
    gen_dp(4, False, "dp4_onehot", 29614)
    gen_rwsadagrad()
    gen_sgd_single(qm)
    gen_dp(2, False, "dp2_unquantized", 29615, emb_q=False)     # emb_grad_quantized=False (sgd:319-329, 626)
    gen_dp(2, "nodup", "dp2_nodup", 29616)
    gen_xchg(1, "xchg1", 29620)
    gen_xchg(2, "xchg2", 29618)
    gen_xchg(4, "xchg4", 29619)
    gen_xchg(2, "xchg2_ec", 29621, ec=True, steps=3)
    tot = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print("golden files:", sorted(os.listdir(GOLD)), "total bytes:", tot)


if __name__ == "__main__":
    main()
