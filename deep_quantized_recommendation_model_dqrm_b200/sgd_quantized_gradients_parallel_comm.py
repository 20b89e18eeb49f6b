"""Drop-in for sgd_quantized_gradients_parallel_comm.py -- the reference's custom data-parallel
"optimizer" (SURVEY.md section 8 a6, a7, a9, a11, a12).  Same function names and arguments; the model
is duck-typed through ``.emb_l / .bot_l / .top_l`` exactly as in the reference
(sgd_quantized_gradients_parallel_comm.py:277-278, 337-339, 374-376).

Per step the reference issues, for 26 tables and 14 MLP tensors, 2x26 + 28 Gloo collectives with
host staging and ~53 host syncs.  Here:
  embeddings  one all-gather of 26 scales + one all-gather of the packed int8 slots (NCCL), then one
              merge/update kernel;  MLP  one all-reduce of the per-channel scales + one of the codes.
Nothing synchronises with the host.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from .dense import DenseArena
from .quantization_supp.quant_modules import QuantEmbeddingBagTwo, QuantLinear
from .quantization_supp.quant_utils import *  # noqa: F401,F403  (star-import kept from the reference, :19)

__all__ = ["clear_gradients", "grad_update_parallel_comm", "weight_update_parallel_comm", "weight_syncc",
           "quantized_gradients_update", "quantize_emb_grad", "quantize_linear_grad", "quantize_bias_grad",
           "grad_precision_and_scale"]


def _rank_world(number_of_gpus):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _emb_groups(model):
    """Table groups holding this step's embedding gradients."""
    if model.emb_l is None:
        raise Warning("Cannot find the list of embedding tables")
    g = getattr(model, "emb_group", None)
    if g is not None and g.last is not None:
        return [g]
    groups = []
    for e in model.emb_l:
        if not isinstance(e, QuantEmbeddingBagTwo) or e._group is None:
            raise _lib.DqrmLibraryError("grad_update_parallel_comm: embedding table has no recorded forward/backward")
        if e._group not in groups:
            groups.append(e._group)
    return groups


def _quant_layers(model):
    if model.bot_l is None:
        raise Warning("Cannot find the list of bottom linear layers")
    if model.top_l is None:
        raise Warning("Cannot find the list of top linear layers")
    return [l for l in list(model.bot_l) + list(model.top_l) if isinstance(l, QuantLinear)]


def _dense_arena(model) -> DenseArena:
    a = getattr(model, "_dense_arena", None)
    layers = _quant_layers(model)
    if a is None or a.layers != layers or not a.intact():
        a = DenseArena(layers, layers[0].weight.device)
        object.__setattr__(model, "_dense_arena", a)
    return a


def clear_gradients(model):
    """Zero every gradient (sgd_quantized_gradients_parallel_comm.py:714-734).  Embedding row
    gradients live in the table group's step buffers and are overwritten by the next backward."""
    with torch.no_grad():
        arena = getattr(model, "_dense_arena", None)
        if arena is not None and arena.intact():
            arena.zero_grad()
            seen = {id(p) for p in arena.params}
        else:
            seen = set()
        for _, param in model.named_parameters():
            if id(param) in seen or param.grad is None:
                continue
            if param.grad.grad_fn is not None:
                param.grad.detach_()
            else:
                param.grad.requires_grad_(False)
            param.grad.zero_()


def grad_precision_and_scale(model, number_of_gpus, rank_for_debug, output_flag=False):
    """Per-table gradient bit width from the gradient range (sgd_quantized_gradients_parallel_comm.py:158-255):
    an experiment (``ranking_range=True``) whose only call sites are commented out in the reference drivers
    (dlrm_s_pytorch_comm_grad.py:1946-1951); it also samples the ranking with numpy on rank 0 for exactly 26
    tables. Out of scope for the hot path: fails loudly instead of silently doing something else."""
    raise NotImplementedError("grad_precision_and_scale / ranking_range: commented out in the reference drivers "
                              "(dlrm_s_pytorch_comm_grad.py:1946-1951); not part of the data-parallel hot path")


def grad_update_parallel_comm(model, number_of_gpus, emb_grad_quantized=True, num_bits=16, ranking_range=False,
                              rank_for_debug=None, iteration_count=None, mlp_layer_quantized=True):
    """Quantise and exchange the gradients (sgd_quantized_gradients_parallel_comm.py:257-446)."""
    if ranking_range:
        raise NotImplementedError("ranking_range (mixed precision by range) is an experiment whose call sites are "
                                  "commented out in the reference (dlrm_s_pytorch_comm_grad.py:1946-1951)")
    rank, world = _rank_world(number_of_gpus)
    if world != number_of_gpus:
        raise ValueError(f"number_of_gpus={number_of_gpus} but the process group has {world} ranks")
    with torch.no_grad():
        groups = _emb_groups(model)
        arena = _dense_arena(model)
        if groups and arena.status is not groups[0].status:
            arena.status = groups[0].status          # ONE status word: a timeout on any site stops every update
        # a group whose backward (and exchange) already runs on its side stream (graph_step: backward_async) is joined
        # AFTER the dense exchange has been issued: the two do not depend on each other, and inside a captured graph
        # the position of the join is the dependency
        running = [g for g in groups if emb_grad_quantized and g.grad_bit == num_bits and
                   (g.exchange_started or g._bwd_forked is not None)]
        for g in groups:
            if g in running:
                continue
            if emb_grad_quantized:
                started = g.finish_exchange()        # (a side-stream run with another code width: join, redo)
                if g.applied_eagerly:
                    raise RuntimeError("embedding update already applied with a different gradient bit width")
                if g.grad_bit != num_bits:
                    g.set_grad_bit(num_bits)
                    started = False
                if not started:
                    g.exchange(world=world, rank=rank)
                if g.modules is not None:
                    for t, e in enumerate(g.modules):
                        e.emb_scaling_factor = g.grad_scale_mean[t:t + 1]
            elif world > 1:
                # emb_grad_quantized=False (sgd:319-329): same slots, fp32 payload, summed in rank order
                g.finish_exchange()
                if g.grad_bit != 32:
                    g.set_grad_bit(32)
                g.exchange(world=world, rank=rank)
            else:
                g.finish_exchange()
        arena.quantize_exchange(world=world, bits=8, quantized=mlp_layer_quantized)
        for g in running:
            if not g.finish_exchange():
                g.exchange(world=world, rank=rank)
            if g.modules is not None:
                for t, e in enumerate(g.modules):
                    e.emb_scaling_factor = g.grad_scale_mean[t:t + 1]


def weight_update_parallel_comm(model, lr, emb_grad_quantized=True, update_embedding=True, num_gpus=1,
                                rank_for_debug=None, ranking_range=False, use_ec=False, mlp_layer_quantized=True):
    """SGD update from the exchanged gradients (sgd_quantized_gradients_parallel_comm.py:601-685)."""
    if ranking_range or use_ec:
        raise NotImplementedError("ranking_range / error compensation are not called by the reference drivers")
    with torch.no_grad():
        for g in _emb_groups(model):
            if g.applied_fused and (emb_grad_quantized or num_gpus > 1 or not update_embedding):
                raise RuntimeError("the embedding update already ran inside the backward (fused_update is the "
                                   "single-process un-quantised path)")
            if g.applied_eagerly and not (update_embedding and emb_grad_quantized):
                raise RuntimeError("the embedding update already ran behind the exchange (eager_apply)")
        if update_embedding:
            for g in _emb_groups(model):
                if g.applied_fused:                  # done inside the de-duplicating backward (dqrm_embbag_bwd_sgd)
                    g.applied_fused = False
                elif g.applied_eagerly:              # done on the side stream, joined by grad_update_parallel_comm
                    g.applied_eagerly = False
                elif emb_grad_quantized or num_gpus > 1:
                    g.merge_apply(lr)
                else:
                    g.sgd_apply(lr, inv_world=1.0)
        _dense_arena(model).apply(lr, world=num_gpus, quantized=mlp_layer_quantized)


def weight_syncc(dlrm, num_gpus):
    """All-reduce-average every parameter (sgd_quantized_gradients_parallel_comm.py:963-970).  Our updates
    are deterministic and identical on every rank, so replicas never drift; kept for API parity and as
    the way to make differently-initialised replicas agree before training (comm_grad.py:1848)."""
    with torch.no_grad():
        if not (dist.is_available() and dist.is_initialized()) or num_gpus == 1:
            return
        done = set()
        arena = getattr(dlrm, "_dense_arena", None)
        if arena is not None and arena.intact():
            dist.all_reduce(arena.flat)
            arena.flat.mul_(1.0 / num_gpus)
            done = {id(p) for p in arena.params}
        table_arena = getattr(dlrm, "table_arena", None)
        if table_arena is not None:
            dist.all_reduce(table_arena)
            table_arena.mul_(1.0 / num_gpus)
            done |= {id(e.embedding_bag.weight) for e in dlrm.emb_l}
        # the tables changed behind the scale tracker's back: block maxima, cached scales and a rescan in flight
        # are stale (incremental / pipelined policies must stay bit-identical to a full rescan)
        _invalidate_scale_state(dlrm)
        for _, param in dlrm.named_parameters():
            if id(param) in done:
                continue
            dist.all_reduce(param.data)
            param.data.mul_(1.0 / num_gpus)
        _invalidate_scale_state(dlrm)


def _invalidate_scale_state(model):
    """Call after mutating embedding tables outside merge_apply / sgd_apply (weight_syncc, checkpoint load)."""
    groups = []
    g = getattr(model, "emb_group", None)
    if g is not None:
        groups.append(g)
    for e in getattr(model, "emb_l", []) or []:
        for cand in (getattr(e, "_group", None), getattr(e, "_solo", None)):
            if cand is not None and cand not in groups:
                groups.append(cand)
    for g in groups:
        g.invalidate_tracker()
        g.scale_valid = False
        g.pipe_pending = False


def quantized_gradients_update(model, arg, lr, num_gpus):
    """Un-quantised all-reduce + SGD for every parameter (sgd_quantized_gradients_parallel_comm.py:687-712)."""
    with torch.no_grad():
        for _, param in model.named_parameters():
            if param.grad is None:
                continue
            update = param.grad
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(update)
            param.add_(update / num_gpus * (-lr[-1]))


# ---- per-tensor entry points kept for API parity (operate on one table / layer) -----------------
def quantize_emb_grad(embedding_table, embedding_table_grad, num_bits, parallel, num_gpus=None, scale=None,
                      use_ec=False, table_id=None):
    """Single-table quantize_emb_grad (sgd_quantized_gradients_parallel_comm.py:850-890) on a sparse COO
    gradient; returns (sparse fp32 codes averaged over ranks, scale[1]).  The fused multi-table path
    used by grad_update_parallel_comm never builds these sparse tensors."""
    if use_ec:
        raise NotImplementedError("error compensation is a broken stub in the reference (:821 uses undefined `scale`)")
    with torch.no_grad():
        g = embedding_table_grad.coalesce()
        if scale is None:
            scale = symmetric_linear_quantization_param_two(num_bits, g.values(), None, None, None)  # noqa: F405
        if parallel:
            dist.all_reduce(scale)
            scale.mul_(1.0 / num_gpus)
        scale = scale.view(-1)
        q = SymmetricQuantFunction.apply(g.values(), num_bits, scale)  # noqa: F405
        out = torch.sparse_coo_tensor(g.indices(), q, size=g.size(), device=g.device)
        if parallel:
            # NCCL has no sparse all-reduce: gather (rows, codes) and let coalesce() merge
            world = dist.get_world_size()
            n = torch.tensor([q.shape[0]], device=q.device)
            ns = [torch.zeros_like(n) for _ in range(world)]
            dist.all_gather(ns, n)
            cap = int(max(int(x) for x in ns))
            rows = torch.zeros(cap, dtype=torch.int64, device=q.device)
            vals = torch.zeros((cap, q.shape[1]), dtype=q.dtype, device=q.device)
            rows[:q.shape[0]] = g.indices()[0]
            vals[:q.shape[0]] = q
            rl = [torch.zeros_like(rows) for _ in range(world)]
            vl = [torch.zeros_like(vals) for _ in range(world)]
            dist.all_gather(rl, rows)
            dist.all_gather(vl, vals)
            rows = torch.cat([r[:int(k)] for r, k in zip(rl, ns)])
            vals = torch.cat([v[:int(k)] for v, k in zip(vl, ns)])
            out = torch.sparse_coo_tensor(rows[None], vals, size=g.size(), device=g.device).coalesce()
            out = out * (1.0 / num_gpus)
        return out, scale


def quantize_linear_grad(layer, num_bits, parallel, num_gpus=None, per_channel=True, scale=None, err_compensation=False):
    """Per-layer quantize_linear_grad (sgd_quantized_gradients_parallel_comm.py:892-929)."""
    if err_compensation or not per_channel:
        raise NotImplementedError("only per_channel=True, err_compensation=False is used (sgd...:341)")
    with torch.no_grad():
        w = layer.weight.grad
        if scale is None:
            scale = symmetric_linear_quantization_params(num_bits, w.min(dim=1)[0], w.max(dim=1)[0], True)  # noqa: F405
        if parallel:
            dist.all_reduce(scale)
            scale.mul_(1.0 / num_gpus)
        q = SymmetricQuantFunction.apply(w, num_bits, scale)  # noqa: F405
        if parallel:
            dist.all_reduce(q)
            q.mul_(1.0 / num_gpus)
        return q, scale


def quantize_bias_grad(layer, num_bits, parallel, num_gpus=None, scale=None, err_compensation=False):
    """Per-layer quantize_bias_grad (sgd_quantized_gradients_parallel_comm.py:931-961)."""
    if err_compensation:
        raise NotImplementedError("err_compensation=False on every call site (sgd...:350)")
    with torch.no_grad():
        b = layer.bias.grad
        if scale is None:
            scale = symmetric_linear_quantization_params(num_bits, b.min(dim=0)[0], b.max(dim=0)[0])  # noqa: F405
        if parallel:
            dist.all_reduce(scale)
            scale.mul_(1.0 / num_gpus)
        q = SymmetricQuantFunction.apply(b, num_bits, scale.view(1))  # noqa: F405
        if parallel:
            dist.all_reduce(q)
            q.mul_(1.0 / num_gpus)
        return q, scale
