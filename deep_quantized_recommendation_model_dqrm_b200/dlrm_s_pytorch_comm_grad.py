"""Hot-path subset of the reference driver dlrm_s_pytorch_comm_grad.py: the DLRM_Net model flow
(create_mlp / create_emb / apply_mlp / apply_emb / interact_features / forward), its DQRM command-line
flags, the batch shard (get_my_slice) and the body of the custom-DP training iteration
(dlrm_s_pytorch_comm_grad.py:1909-1957) -- with every hot operator routed to libdqrm_b200.

Out of scope: dataset loading, evaluation, checkpoints, TensorBoard,
QR/MD embedding tricks, activation quantisation, the "cat" interaction, multi-device model
parallelism.  Those options raise instead of silently diverging.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np
import torch
import torch.nn as nn

from . import _lib, extend_distributed as ext_dist
from .quantization_supp.quant_modules import QuantAct, QuantEmbeddingBagTwo, QuantLinear  # noqa: F401
from .quantization_supp.quant_modules_not_quantize_grad import QuantEmbeddingBagTwo, QuantLinear  # noqa: F401,F811
from .quantization_supp.quant_modules import EmbBagGroupFunction, _new_group
from .sgd_quantized_gradients_parallel_comm import _emb_groups
from .sgd_quantized_gradients_parallel_comm import (clear_gradients, grad_update_parallel_comm,  # noqa: F401
                                                    weight_syncc, weight_update_parallel_comm)
from .tables import EmbeddingTableGroup

# module-level switches of the reference (dlrm_s_pytorch_comm_grad.py:147-159); train() sets
# full_precision_flag = args.pretrain_and_quantize (:1425-1426)
full_precision_flag = False
change_bitw = False
change_bitw2 = 4
change_lin_full_quantize = False


def get_my_slice(n, my_size, my_rank):
    """dlrm_s_pytorch_comm_grad.py:993-997."""
    k, m = divmod(n, my_size)
    return slice(my_rank * k + min(my_rank, m), (my_rank + 1) * k + min(my_rank + 1, m), 1)


class EmbOutputs(list):
    """``ly``: list of per-table [B, D] tensors (the reference's return type) that also carries the
    [T, B, D] tensor they are views of, so interact_features can skip the concatenation."""
    stacked = None
    group = None


class _InteractFunction(torch.autograd.Function):
    """Fused dot interaction (a14), forward and backward in one kernel each.  When `ly` is the output of the
    fused QAT EmbeddingBag (`group` given) the backward also applies that op's straight-through estimator
    (g*s)/s in its epilogue and tells the group, so the de-duplicating backward skips it."""

    @staticmethod
    def forward(ctx, x, ly, itself, group):
        lib = _lib.load()
        T, B, D = ly.shape
        nf = T + 1
        npairs = nf * (nf - 1) // 2 + (nf if itself else 0)
        x = x.contiguous()
        R = torch.empty((B, D + npairs), dtype=torch.float32, device=x.device)
        rc = lib.dqrm_interact_fwd(x.data_ptr(), ly.data_ptr(), ly.stride(0), ly.stride(1), B, T, D, int(itself),
                                   R.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "dqrm_interact_fwd")
        ctx.save_for_backward(x, ly)
        ctx.itself = itself
        ctx.group = group
        ctx.ste = group is not None and group.last is not None and not group.last[5]    # quantised forward
        return R

    @staticmethod
    def backward(ctx, dR):
        lib = _lib.load()
        x, ly = ctx.saved_tensors
        T, B, D = ly.shape
        dR = dR.contiguous()
        dx = torch.empty_like(x)
        dly = torch.empty((T, B, D), dtype=torch.float32, device=x.device)
        g = ctx.group
        rc = lib.dqrm_interact_bwd(x.data_ptr(), ly.data_ptr(), ly.stride(0), ly.stride(1), dR.data_ptr(), B, T, D,
                                   int(ctx.itself), dx.data_ptr(), dly.data_ptr(), dly.stride(0), dly.stride(1),
                                   g.scale.data_ptr() if ctx.ste else None, _lib.stream_ptr())
        _lib.check(rc, "dqrm_interact_bwd")
        if ctx.ste:
            g.ste_done_for = dly.data_ptr()      # the gradient tensor that already carries (g*s)/s
        return dx, dly, None, None


class DLRM_Net(nn.Module):
    """Reference DLRM_Net (dlrm_s_pytorch_comm_grad.py:278-966), hot configuration.

    Extensions (keyword-only, default = reference behaviour):
      device        build the embedding tables directly on this CUDA device inside ONE arena
                    (``table_arena``), initialised on the device with the reference's uniform law;
                    the reference draws them with numpy on the host (2.16 GB at Kaggle shape).
      table_seed    seed for that device-side init (per table: seed + table id).
    """

    def create_mlp(self, ln, sigmoid_layer, quant_linear_layer=False, channelwise_lin=False, quantize_activation=False):
        layers = nn.ModuleList()
        for i in range(0, ln.size - 1):
            n, m = ln[i], ln[i + 1]
            LL = nn.Linear(int(n), int(m), bias=True)
            mean = 0.0
            std_dev = np.sqrt(2 / (m + n))
            W = np.random.normal(mean, std_dev, size=(m, n)).astype(np.float32)
            std_dev = np.sqrt(1 / m)
            bt = np.random.normal(mean, std_dev, size=m).astype(np.float32)
            LL.weight.data = torch.tensor(W, requires_grad=True)
            LL.bias.data = torch.tensor(bt, requires_grad=True)
            if self.quantization_flag and quant_linear_layer:
                QuantLnr = QuantLinear(weight_bit=self.weight_bit, bias_bit=self.weight_bit,
                                       full_precision_flag=not self.quantize_act_and_lin,
                                       per_channel=channelwise_lin, quantize_activation=quantize_activation)
                QuantLnr.set_param(LL)
                layers.append(QuantLnr)
            else:
                layers.append(LL)
            layers.append(nn.Sigmoid() if i == sigmoid_layer else nn.ReLU())
        return layers

    def create_emb(self, m, ln, weighted_pooling=None):
        emb_l = nn.ModuleList()
        v_W_l = []
        if not self.quantization_flag:
            raise NotImplementedError("the un-quantised FP32 DLRM baseline is not part of the DQRM hot path")
        arena = None
        if self._build_device is not None:
            total = int(sum(int(n) for n in ln))
            arena = torch.empty((total, m), dtype=torch.float32, device=self._build_device)
            self.table_arena = arena
        row0 = 0
        for i in range(0, ln.size):
            n = int(ln[i])
            w = None
            if arena is not None:
                from . import synthetic
                w = synthetic.table_weights_(arena[row0:row0 + n], i, self._table_seed)
                row0 += n
            EE = QuantEmbeddingBagTwo(n, m, self.embedding_bit, embedding_id=i, _weight=w)
            v_W_l.append(None if weighted_pooling is None else torch.ones(n, dtype=torch.float32))
            emb_l.append(EE)
        return emb_l, v_W_l

    def __init__(self, m_spa=None, ln_emb=None, ln_bot=None, ln_top=None, arch_interaction_op=None,
                 arch_interaction_itself=False, sigmoid_bot=-1, sigmoid_top=-1, sync_dense_params=True,
                 loss_threshold=0.0, ndevices=-1, qr_flag=False, qr_operation="mult", qr_collisions=0,
                 qr_threshold=200, md_flag=False, md_threshold=200, weighted_pooling=None, loss_function="bce",
                 quantization_flag=False, embedding_bit=32, modify_feature_interaction=False, weight_bit=8,
                 quantize_act_and_lin=False, mlp_channelwise=False, quantize_activation=False, deviceid=None,
                 *, device=None, table_seed=1234):
        super().__init__()
        self.emb_group = None
        self.table_arena = None
        self._build_device = torch.device(device) if device is not None else None
        self._table_seed = table_seed
        if (m_spa is None) or (ln_emb is None) or (ln_bot is None) or (ln_top is None) or (arch_interaction_op is None):
            return
        if qr_flag or md_flag:
            raise NotImplementedError("QR / mixed-dimension embeddings are orthogonal tricks outside the hot path")
        if weighted_pooling is not None:
            raise NotImplementedError("weighted pooling is ignored by QuantEmbeddingBagTwo (qm:364-367)")
        if modify_feature_interaction:
            raise NotImplementedError("--modify_feature_interaction is unused by the canonical scripts")
        ln_emb, ln_bot, ln_top = np.asarray(ln_emb), np.asarray(ln_bot), np.asarray(ln_top)
        self.ndevices = ndevices
        self.output_d = 0
        self.arch_interaction_op = arch_interaction_op
        self.arch_interaction_itself = arch_interaction_itself
        self.sync_dense_params = sync_dense_params
        self.loss_threshold = loss_threshold
        self.loss_function = loss_function
        self.quantization_flag = quantization_flag
        self.embedding_bit = embedding_bit
        self.modify_feature_interaction = modify_feature_interaction
        self.weight_bit = weight_bit
        self.quantize_act_and_lin = quantize_act_and_lin and quantization_flag
        self.change_lin_from_full_to_quantized = False
        self.channelwise_lin = mlp_channelwise
        self.quantize_activation = quantize_activation
        self.deviceid = deviceid
        self.weighted_pooling = weighted_pooling
        self.qr_flag, self.md_flag = qr_flag, md_flag
        if self.quantization_flag:
            if self.weight_bit is not None:
                ab = self.weight_bit if self.weight_bit >= 8 else 8
                self.quant_input = QuantAct(activation_bit=ab, act_range_momentum=-1)
                self.quant_feature_outputs = QuantAct(fixed_point_quantization=True, activation_bit=ab, act_range_momentum=-1)
            self.register_buffer("feature_xmin", torch.zeros(1))
            self.register_buffer("feature_xmax", torch.zeros(1))
            self.register_buffer("features_scaling_factor", torch.zeros(1))
        self.emb_l, self.v_W_l = self.create_emb(m_spa, ln_emb, weighted_pooling)
        self.bot_l = self.create_mlp(ln_bot, sigmoid_bot, quant_linear_layer=True, channelwise_lin=self.channelwise_lin,
                                     quantize_activation=self.quantize_activation)
        self.top_l = self.create_mlp(ln_top, sigmoid_top, quant_linear_layer=True, channelwise_lin=self.channelwise_lin,
                                     quantize_activation=self.quantize_activation)
        self.quantize_emb = False
        self.emb_l_q = []
        self.quantize_bits = 32
        if self.loss_function == "mse":
            self.loss_fn = torch.nn.MSELoss(reduction="mean")
        elif self.loss_function == "bce":
            self.loss_fn = torch.nn.BCELoss(reduction="mean")
        else:
            sys.exit("ERROR: --loss-function=" + self.loss_function + " is not supported")
        if self._build_device is not None:
            self.to(self._build_device)

    # ------------------------------------------------------------------
    def _ensure_group(self):
        ws = [e.embedding_bag.weight for e in self.emb_l]
        g = self.emb_group
        if g is None or any(a is not b for a, b in zip(g.weights, ws)) or g.device != ws[0].device \
                or g.embedding_bit != self.embedding_bit:
            g = _new_group(ws, self.embedding_bit)
            g.modules = list(self.emb_l)
            for t, e in enumerate(self.emb_l):
                e._group, e._group_index = g, t
            self.emb_group = g
        if ext_dist.my_size > 1:
            g.dp_world, g.dp_rank = ext_dist.my_size, ext_dist.my_rank
        return g

    fuse_mlp = True               # fused QuantLinear+activation kernels over the dense arena (a15)
    fuse_mlp_max_batch = 1 << 30  # no limit: from 256 rows the fused layers contract on the tensor cores
                                  # (csrc/mlp_tc.cu), so no batch size falls back to cuBLAS any more (round 1:
                                  # batch > 2048 did, because the FFMA weight-gradient kernel outgrew its cluster split)

    def _fused_mlp_arena(self):
        """The dense arena if the fused MLP path applies (all layers quantised per channel on CUDA)."""
        if not self.fuse_mlp or not self.quantize_act_and_lin or not self.channelwise_lin:
            return None
        from .sgd_quantized_gradients_parallel_comm import _dense_arena
        arena = _dense_arena(self)
        return arena if (arena.fused_ok and arena.flat.is_cuda) else None

    def apply_mlp(self, x, layers, prev_act_scaling_factor=None):
        fused = self._mlp_arena is not None
        i, n = 0, len(layers)
        while i < n:
            layer = layers[i]
            if isinstance(layer, QuantLinear):
                if fused and not layer.full_precision_flag and prev_act_scaling_factor is None:
                    nxt = layers[i + 1] if i + 1 < n else None
                    act = 1 if isinstance(nxt, nn.ReLU) else (2 if isinstance(nxt, nn.Sigmoid) else 0)
                    x = layer.forward_fused(x, act)
                    i += 2 if act else 1
                    continue
                x, prev_act_scaling_factor = layer(x, prev_act_scaling_factor)
            else:
                x = layer(x)
            i += 1
        return x

    _mlp_arena = None

    def apply_emb(self, lS_o, lS_i, emb_l, v_W_l, test_mode=False):
        """All tables in two launches (scale scan + fused forward) instead of a Python loop over
        tables with ~16 launches and a host sync each (dlrm_s_pytorch_comm_grad.py:614-679)."""
        if emb_l is not self.emb_l:
            return [E(lS_i[k], lS_o[k], full_precision_flag=full_precision_flag, test_mode=test_mode)
                    for k, E in enumerate(emb_l)]
        g = self._ensure_group()
        fp = full_precision_flag or any(e.full_precision_flag for e in self.emb_l)
        idx, off, idx_begin, bags = EmbeddingTableGroup.pack_inputs(lS_i, lS_o, g.device)
        if ((not fp and not test_mode) or not g.scale_valid) and not self.external_scan:
            # periodic update (qm:303-315,354-363; the paper's period-200 numbers): after the first scan, rescan
            # only when `scale_update_period` forwards have passed; 0 = every forward (the shipped behaviour)
            if not g.scale_valid or self._since_scan >= self.scale_update_period:
                g.scan_scales(shard_rank=g.dp_rank if self.shard_scan else 0,
                              shard_world=g.dp_world if self.shard_scan else 1)
                self._since_scan = 0
            else:
                self._since_scan += 1
        if not self._scale_views_bound:
            for t, e in enumerate(self.emb_l):
                e.eb_scaling_factor = g.scale[t]
            self._scale_views_bound = True
        out = EmbBagGroupFunction.apply(g, idx, off, idx_begin, bags, fp, *[e.embedding_bag.weight for e in self.emb_l])
        ly = EmbOutputs(out.unbind(0))
        ly.stacked, ly.group = out, g
        return ly

    scale_update_period = 0       # 0: rescan every forward (shipped reference); P: rescan every P+1 forwards
    _since_scan = 0
    shard_scan = False            # multi-GPU: scan 1/world of every table per rank + MAX all-reduce
    external_scan = False         # the caller launches group.scan_scales() itself before every forward
    _scale_views_bound = False

    def interact_features(self, x, ly):
        if self.arch_interaction_op != "dot":
            sys.exit("ERROR: --arch-interaction-op=" + self.arch_interaction_op + " is not supported")
        stacked = getattr(ly, "stacked", None)
        group = getattr(ly, "group", None)
        if stacked is None:
            stacked, group = torch.stack(list(ly), dim=0), None
        R = _InteractFunction.apply(x, stacked, bool(self.arch_interaction_itself), group)
        if not self.quantization_flag:
            return R
        if self.quantize_activation:
            raise NotImplementedError("activation quantisation is outside the hot path")
        return R, None

    def forward_bottom(self, dense_x):
        """The part of the forward that does not depend on the embedding scales: MLP weight fake-quantisation and the
        bottom MLP (dlrm_s_pytorch_comm_grad.py:855).  graph_step runs it on its own stream beside the table scan and
        hands the result to forward(x_bottom=...)."""
        if not self.quantization_flag or self.quantize_activation:
            raise NotImplementedError("only the --quantization_flag --linear_channel flow is built "
                                      "(dlrm_s_pytorch_comm_grad.py:855-859)")
        self._mlp_arena = self._fused_mlp_arena() if dense_x.shape[0] <= self.fuse_mlp_max_batch else None
        if self._mlp_arena is not None:
            self._mlp_arena.fakequant_all()          # all 7 layers' weights + biases, one launch
        return self.apply_mlp(dense_x, self.bot_l, prev_act_scaling_factor=None)

    def forward(self, dense_x, lS_o, lS_i, test_mode=False, x_bottom=None):
        x = self.forward_bottom(dense_x) if x_bottom is None else x_bottom
        ly = self.apply_emb(lS_o, lS_i, self.emb_l, self.v_W_l, test_mode=test_mode)
        z, feature_scaling_factor = self.interact_features(x, ly)
        p = self.apply_mlp(z, self.top_l, prev_act_scaling_factor=feature_scaling_factor)
        if 0.0 < self.loss_threshold < 1.0:
            return torch.clamp(p, min=self.loss_threshold, max=(1.0 - self.loss_threshold))
        return p


# ----------------------------------------------------------------------------
def dlrm_wrap(dlrm, X, lS_o, lS_i, use_gpu, device, ndevices=1, test_mode=False):
    """dlrm_s_pytorch_comm_grad.py:170-189: H2D copies of the batch, then forward."""
    if use_gpu:
        lS_i = [S_i.to(device, non_blocking=True) for S_i in lS_i] if isinstance(lS_i, list) else lS_i.to(device, non_blocking=True)
        lS_o = [S_o.to(device, non_blocking=True) for S_o in lS_o] if isinstance(lS_o, list) else lS_o.to(device, non_blocking=True)
    return dlrm(X.to(device, non_blocking=True), lS_o, lS_i, test_mode=test_mode)


def loss_fn_wrap(Z, T, use_gpu, device, args=None):
    """dlrm_s_pytorch_comm_grad.py:192-211 (bce / mse)."""
    fn = torch.nn.MSELoss(reduction="mean") if (args is not None and args.loss_function == "mse") else torch.nn.BCELoss(reduction="mean")
    return fn(Z, T.to(device, non_blocking=True))


def train_iteration(dlrm, X, lS_o, lS_i, T, lr, world_size=1, rank=0, device=None,
                    quantize_embedding_bag_gradient=True, embedding_bag_gradient_bit_num=8,
                    mlp_layer_quantized=True, args=None):
    """Body of the reference hot loop for one (global) batch (dlrm_s_pytorch_comm_grad.py:1909-1957):
    shard -> forward -> loss -> clear_gradients -> backward -> grad_update_parallel_comm ->
    weight_update_parallel_comm.  Returns the loss tensor (on the device; no sync)."""
    device = device if device is not None else next(dlrm.parameters()).device
    mbs = T.shape[0]
    if world_size > 1:
        sl = get_my_slice(mbs, world_size, rank)
        X, T = X[sl], T[sl]
        if torch.is_tensor(lS_i):
            lS_i = lS_i[:, sl]
            lS_o = lS_o[:, 0:lS_i.shape[1]]
        else:
            raise NotImplementedError("the reference shards only stacked Criteo index tensors (:1910-1911)")
    Z = dlrm_wrap(dlrm, X, lS_o, lS_i, True, device)
    E = loss_fn_wrap(Z, T, True, device, args)
    clear_gradients(dlrm)
    for g in _emb_groups(dlrm):
        if g.fused_update is not None:               # single process, un-quantised: the backward applies the rows itself
            g.fused_update["lr"] = float(lr)
    E.backward()
    grad_update_parallel_comm(dlrm, world_size, emb_grad_quantized=quantize_embedding_bag_gradient,
                              num_bits=embedding_bag_gradient_bit_num, ranking_range=False, rank_for_debug=rank,
                              mlp_layer_quantized=mlp_layer_quantized)
    weight_update_parallel_comm(dlrm, lr, emb_grad_quantized=quantize_embedding_bag_gradient, update_embedding=True,
                                num_gpus=world_size, rank_for_debug=rank, mlp_layer_quantized=mlp_layer_quantized)
    return E.detach()


def make_parser():
    """The reference's flat argparse (dlrm_s_pytorch_comm_grad.py:999-1137), hot-path subset with the
    same flag names and defaults."""
    p = argparse.ArgumentParser(description="DQRM hot path on B200")
    p.add_argument("--arch-sparse-feature-size", type=int, default=2)
    p.add_argument("--arch-embedding-size", type=str, default="4-3-2")
    p.add_argument("--arch-mlp-bot", type=str, default="4-3-2")
    p.add_argument("--arch-mlp-top", type=str, default="4-2-1")
    p.add_argument("--arch-interaction-op", type=str, choices=["dot", "cat"], default="dot")
    p.add_argument("--arch-interaction-itself", action="store_true", default=False)
    p.add_argument("--quantization_flag", action="store_true", default=False)
    p.add_argument("--embedding_bit", type=int, default=None)
    p.add_argument("--weight_bit", type=int, default=None)
    p.add_argument("--loss-function", type=str, default="mse")
    p.add_argument("--loss-threshold", type=float, default=0.0)
    p.add_argument("--round-targets", type=bool, default=False)
    p.add_argument("--data-size", type=int, default=1)
    p.add_argument("--num-batches", type=int, default=0)
    p.add_argument("--data-generation", type=str, default="random")
    p.add_argument("--num-indices-per-lookup", type=int, default=10)
    p.add_argument("--num-indices-per-lookup-fixed", type=bool, default=False)
    p.add_argument("--max-ind-range", type=int, default=-1)
    p.add_argument("--rand-data-dist", type=str, default="uniform")
    p.add_argument("--rand-data-min", type=float, default=0)
    p.add_argument("--rand-data-max", type=float, default=1)
    p.add_argument("--rand-data-mu", type=float, default=-1)
    p.add_argument("--rand-data-sigma", type=float, default=1)
    p.add_argument("--data-trace-file", type=str, default="./input/dist_emb_j.log")
    p.add_argument("--data-trace-enable-padding", type=bool, default=False)
    p.add_argument("--data-set", type=str, default="kaggle")
    p.add_argument("--raw-data-file", type=str, default="")
    p.add_argument("--processed-data-file", type=str, default="")
    p.add_argument("--data-randomize", type=str, default="total")
    p.add_argument("--data-sub-sample-rate", type=float, default=0.0)
    p.add_argument("--memory-map", action="store_true", default=False)
    p.add_argument("--dataset-multiprocessing", action="store_true", default=False)
    p.add_argument("--num-workers", type=int, default=0)
    p.add_argument("--test-mini-batch-size", type=int, default=-1)
    p.add_argument("--test-num-workers", type=int, default=-1)
    p.add_argument("--mini-batch-size", type=int, default=1)
    p.add_argument("--nepochs", type=int, default=1)
    p.add_argument("--learning-rate", type=float, default=0.01)
    p.add_argument("--numpy-rand-seed", type=int, default=123)
    p.add_argument("--lr-num-warmup-steps", type=int, default=0)
    p.add_argument("--lr-decay-start-step", type=int, default=0)
    p.add_argument("--lr-num-decay-steps", type=int, default=0)
    p.add_argument("--optimizer", type=str, default="sgd")
    p.add_argument("--use-gpu", action="store_true", default=False)
    p.add_argument("--print-freq", type=int, default=1)
    p.add_argument("--print-time", action="store_true", default=False)
    p.add_argument("--pretrain_and_quantize", action="store_true", default=False)
    p.add_argument("--modify_feature_interaction", action="store_true", default=False)
    p.add_argument("--linear_shift_down_bit_width", action="store_true", default=False)
    p.add_argument("--documenting_table_weight", action="store_true", default=False)
    p.add_argument("--pretrain_and_quantize_lin", action="store_true", default=False)
    p.add_argument("--quantize_activation", action="store_true", default=False)
    p.add_argument("--linear_channel", action="store_true", default=False)
    p.add_argument("--quantize_act_and_lin", action="store_true", default=False)
    p.add_argument("--quantize_embedding_bag_gradient", action="store_true", default=False)
    p.add_argument("--embedding_bag_gradient_bit_num", type=int, default=16)
    p.add_argument("--scale-update-period", type=int, default=0,
                   help="extension: rescan the tables for their scale every P+1 iterations (0 = every iteration)")
    p.add_argument("--scale-policy", type=str, default="full", choices=["full", "incremental"],
                   help="extension: full rescan (reference) or the exact incremental block-max tracker")
    p.add_argument("-n", "--nodes", default=1, type=int, metavar="N")
    p.add_argument("-g", "--gpus", default=1, type=int)
    p.add_argument("-nr", "--nr", default=0, type=int)
    return p


def parse_args(argv=None):
    args = make_parser().parse_args(argv)
    if args.linear_channel:                      # dlrm_s_pytorch_comm_grad.py:1155-1156
        args.quantize_activation = False
    args.world_size = args.gpus * args.nodes     # :1158
    if args.test_mini_batch_size < 0:            # :1370-1373
        args.test_mini_batch_size = args.mini_batch_size
    if args.test_num_workers < 0:
        args.test_num_workers = args.num_workers
    return args


class LRPolicyScheduler:
    """The learning rate the reference feeds into weight_update_parallel_comm: ``lr_scheduler.get_lr()[-1]`` of
    its LRPolicyScheduler (dlrm_s_pytorch_comm_grad.py:221-255, used at :1738, :1957-1960) -- linear warm-up,
    constant, quadratic decay, then frozen.  Stand-alone (no optimizer object: this path has none); ``step_count``
    follows torch's _LRScheduler convention that the reference relies on: 1 after construction, +1 per step()."""

    def __init__(self, base_lr, num_warmup_steps, decay_start_step, num_decay_steps):
        if decay_start_step < num_warmup_steps:
            sys.exit("Learning rate warmup must finish before the decay starts")
        self.base_lrs = [float(base_lr)]
        self.num_warmup_steps, self.decay_start_step = num_warmup_steps, decay_start_step
        self.num_decay_steps, self.decay_end_step = num_decay_steps, decay_start_step + num_decay_steps
        self._step_count = 0
        self.last_lr = list(self.base_lrs)
        self.step()                                   # _LRScheduler.__init__ performs the initial step

    def get_lr(self):
        n = self._step_count
        if n < self.num_warmup_steps:
            scale = 1.0 - (self.num_warmup_steps - n) / self.num_warmup_steps
            self.last_lr = [b * scale for b in self.base_lrs]
        elif self.decay_start_step <= n < self.decay_end_step:
            scale = ((self.num_decay_steps - (n - self.decay_start_step)) / self.num_decay_steps) ** 2
            self.last_lr = [max(0.0000001, b * scale) for b in self.base_lrs]
        elif self.num_decay_steps <= 0:
            return list(self.base_lrs)
        return list(self.last_lr)

    def step(self):
        self._step_count += 1
        self.get_lr()                                 # the base class evaluates get_lr() inside step()


def train(args, rank=0, world_size=1, device=None, log=print):
    """The training loop of the reference's train() (dlrm_s_pytorch_comm_grad.py:1400-1995), hot-path subset:
    build the loaders (random or pre-processed Criteo), the model, and run the custom-DP iteration for
    --nepochs / --num-batches, printing the loss every --print-freq iterations.  Evaluation, checkpoints,
    TensorBoard and LR schedules are outside the hot path.  Returns the list of per-iteration losses."""
    from . import dlrm_data_pytorch as dp
    if not torch.cuda.is_available():
        raise _lib.DqrmLibraryError("train(): no CUDA device -- the product path has no CPU fallback")
    device = torch.device(device) if device is not None else torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(device)
    np.random.seed(args.numpy_rand_seed)
    torch.manual_seed(args.numpy_rand_seed)
    ln_bot = np.fromstring(args.arch_mlp_bot, dtype=int, sep="-")
    if args.data_generation == "dataset":
        train_data, train_ld, _, _ = dp.make_criteo_data_and_loaders(args)
        ln_emb = np.array(train_data.counts)
        if args.max_ind_range > 0:               # :1476-1478
            ln_emb = np.array([min(int(c), args.max_ind_range) for c in ln_emb])
        m_den = train_data.m_den
        ln_bot[0] = m_den
    else:
        ln_emb = np.fromstring(args.arch_embedding_size, dtype=int, sep="-")
        m_den = ln_bot[0]
        train_data, train_ld, _, _ = dp.make_random_data_and_loader(args, ln_emb, m_den)
    m_spa = args.arch_sparse_feature_size
    num_int = ln_emb.size + 1
    n_pairs = num_int * (num_int + 1) // 2 if args.arch_interaction_itself else num_int * (num_int - 1) // 2
    ln_top = np.fromstring(str(m_spa + n_pairs) + "-" + args.arch_mlp_top, dtype=int, sep="-")   # :1500-1519
    if m_spa != ln_bot[-1]:
        sys.exit("ERROR: arch-sparse-feature-size " + str(m_spa) + " does not match last dim of bottom mlp "
                 + str(ln_bot[-1]))
    global full_precision_flag
    full_precision_flag = args.pretrain_and_quantize
    dlrm = DLRM_Net(m_spa, ln_emb, ln_bot, ln_top, arch_interaction_op=args.arch_interaction_op,
                    arch_interaction_itself=args.arch_interaction_itself, sigmoid_bot=-1, sigmoid_top=ln_top.size - 2,
                    loss_threshold=args.loss_threshold, loss_function=args.loss_function,
                    quantization_flag=args.quantization_flag, embedding_bit=args.embedding_bit,
                    weight_bit=args.weight_bit, quantize_act_and_lin=args.quantize_act_and_lin,
                    mlp_channelwise=args.linear_channel, quantize_activation=args.quantize_activation,
                    device=device, table_seed=args.numpy_rand_seed)
    dlrm._ensure_group().scale_policy = args.scale_policy
    dlrm.scale_update_period = args.scale_update_period
    dlrm.shard_scan = world_size > 1
    lr_scheduler = LRPolicyScheduler(args.learning_rate, args.lr_num_warmup_steps, args.lr_decay_start_step,
                                     args.lr_num_decay_steps)
    losses, it = [], 0
    for epoch in range(args.nepochs):
        for X, lS_o, lS_i, T in train_ld:
            if world_size > 1 and T.shape[0] % world_size != 0:       # ragged last batch is skipped (:1900-1905)
                log("Warning: Skiping the batch %d with size %d" % (it, T.shape[0]))
                continue
            E = train_iteration(dlrm, X, lS_o, lS_i, T, lr_scheduler.get_lr()[-1], world_size=world_size, rank=rank,
                                device=device, quantize_embedding_bag_gradient=args.quantize_embedding_bag_gradient,
                                embedding_bag_gradient_bit_num=args.embedding_bag_gradient_bit_num, args=args)   # :1940-1957
            lr_scheduler.step()                                        # :1960
            it += 1
            if args.print_freq > 0 and it % args.print_freq == 0:
                losses.append(float(E))                                # the reference syncs here too (:1928)
                _poll_status(dlrm)                                     # the host is synchronised anyway: poll device errors
                log("Finished training it {}/{} of epoch {}, loss {:.6f}".format(it, len(train_ld), epoch, losses[-1]))
            if args.num_batches > 0 and it >= args.num_batches * (epoch + 1):
                break
    _poll_status(dlrm)
    return losses


def _poll_status(dlrm):
    """Device-side error words of the embedding group and the MLP arena (index range, capacity, exchange timeout)."""
    dlrm._ensure_group().check_status()
    arena = getattr(dlrm, "_dense_arena", None)
    if arena is not None:
        arena.check_status()


def main(argv=None):
    args = parse_args(argv)
    if args.world_size > 1:
        raise SystemExit("multi-GPU runs are launched with torchrun through bench.py / GraphedTrainStep; "
                         "this entry point is the single-process loop")
    train(args)


if __name__ == "__main__":
    main()
