"""Shared builders for the GPU parity tests: identical synthetic weights into the CUDA model and the
oracle, and an in-process emulation of N data-parallel ranks on one GPU (collectives replaced by
explicit copies between the replicas' buffers, so no process waits on another -- see
B200_PROFILING.md on why ranks must not be emulated as concurrent processes on one GPU)."""
import numpy as np
import torch

from deep_quantized_recommendation_model_dqrm_b200 import synthetic
from deep_quantized_recommendation_model_dqrm_b200 import dlrm_s_pytorch_comm_grad as drv
from deep_quantized_recommendation_model_dqrm_b200 import sgd_quantized_gradients_parallel_comm as sgd
from oracle import dqrm_oracle as O

C_SMALL = dict(rows=[50, 3, 1000, 200], dim=16, ln_bot=[13, 32, 16], ln_top_hidden=[32, 1])
C_NODUP = dict(rows=[1500, 40, 1000, 200], dim=16, ln_bot=[13, 32, 16], ln_top_hidden=[32, 1])   # oracle/make_golden.py
C1 = synthetic.RANDOM_SMALL


def weights_for(cfg, seed):
    rng = np.random.RandomState(seed)
    emb = [synthetic.table_weights_numpy(n, cfg["dim"], rng) for n in cfg["rows"]]
    ln_top = synthetic.top_mlp_sizes(len(cfg["rows"]), cfg["dim"], cfg["ln_top_hidden"])
    bot = synthetic.mlp_params(cfg["ln_bot"], rng)
    top = synthetic.mlp_params(ln_top, rng)
    return emb, bot, top, ln_top


def build_cuda_model(cfg, seed=300, device="cuda", embedding_bit=4, weight_bit=4):
    emb, bot, top, ln_top = weights_for(cfg, seed)
    m = drv.DLRM_Net(cfg["dim"], np.array(cfg["rows"]), np.array(cfg["ln_bot"]), np.array(ln_top),
                     arch_interaction_op="dot", sigmoid_bot=-1, sigmoid_top=len(ln_top) - 2, loss_function="bce",
                     quantization_flag=True, embedding_bit=embedding_bit, weight_bit=weight_bit,
                     quantize_act_and_lin=True, mlp_channelwise=True, quantize_activation=False)
    for k, W in enumerate(emb):
        m.emb_l[k].embedding_bag.weight.data = torch.tensor(W)
    for layers, ps in ((m.bot_l, bot), (m.top_l, top)):
        qls = [l for l in layers if isinstance(l, drv.QuantLinear)]
        for l, (W, b) in zip(qls, ps):
            l.weight.data = torch.tensor(W)
            l.bias.data = torch.tensor(b)
    return m.to(device)


def build_oracle_model(cfg, seed=300):
    emb, bot, top, _ = weights_for(cfg, seed)
    return O.OracleDLRM(cfg["rows"], cfg["dim"], bot, top, emb_weights=[torch.from_numpy(w) for w in emb])


def shard(world, rank, multihot, step, rows, per_rank=16):
    """The batch shard of tests/golden dp*: same generator calls as oracle/make_golden.py."""
    Bg = per_rank * world
    sl = slice(rank * per_rank, (rank + 1) * per_rank)
    if multihot == "nodup":
        X, lS_o, lS_i, T = synthetic.criteo_batch_nodup(rows, world, per_rank, seed=400 + step)
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


def emulated_dp_step(models, batches, lr, device="cuda", keep_debug=True):
    """One iteration of the custom-DP loop over `len(models)` replicas living on ONE GPU."""
    world = len(models)
    losses = []
    cap = 1                       # common slot capacity: the largest per-table lookup count of any rank
    for (_, _, lS_i, _) in batches:
        cap = max(cap, max(int(t.shape[0]) for t in lS_i) if not torch.is_tensor(lS_i) else int(lS_i.shape[1]))
    for r, (m, (X, lS_o, lS_i, T)) in enumerate(zip(models, batches)):
        g = m._ensure_group()
        g.dp_world, g.dp_rank = world, r
        g.fixed_capacity = cap
        g.keep_debug = keep_debug
        Z = drv.dlrm_wrap(m, X, lS_o, lS_i, True, device)
        E = drv.loss_fn_wrap(Z, T, True, device)
        sgd.clear_gradients(m)
        E.backward()
        losses.append(E.detach())
    groups = [m.emb_group for m in models]
    arenas = [sgd._dense_arena(m) for m in models]
    emulated_exchange(groups, arenas)
    for m in models:
        sgd.weight_update_parallel_comm(m, lr, emb_grad_quantized=True, update_embedding=True, num_gpus=world)
    return [float(l) for l in losses]


def emulated_exchange(groups, arenas):
    """The collectives of grad_update_parallel_comm between W replicas on one GPU, as explicit copies."""
    world = len(groups)
    # collective 1: all-gather of the per-table local scales
    for r, g in enumerate(groups):
        g.stage_scale(r)
    all_scales = torch.stack([g.grad_scale_local for g in groups])
    for g in groups:
        g.gathered_scales.copy_(all_scales)
    for r, g in enumerate(groups):
        g.pack(r)
    # collective 2: all-gather of the packed slots
    sb = groups[0].slot_bytes
    for g in groups:
        for r2, g2 in enumerate(groups):
            if g2 is not g:
                g.gathered[r2 * sb:(r2 + 1) * sb].copy_(g2.gathered[r2 * sb:(r2 + 1) * sb])
    # MLP: SUM all-reduce of scales, then of codes (rank order)
    for a in arenas:
        a.local_scale(8)
    ssum = arenas[0].scale_local.clone()
    for a in arenas[1:]:
        ssum = ssum + a.scale_local
    for a in arenas:
        a.scale_local.copy_(ssum)
        a.quantize(world, 8)
    csum = arenas[0].codes.clone()
    for a in arenas[1:]:
        csum = csum + a.codes
    for a in arenas:
        a.codes.copy_(csum)
