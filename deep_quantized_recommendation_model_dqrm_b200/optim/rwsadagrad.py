"""Row-wise sparse Adagrad for the embedding tables -- drop-in for optim/rwsadagrad.py (SURVEY.md section 8
f-2).  The reference implementation is CPU-only (the driver exits on GPUs, dlrm_s_pytorch_comm_grad.py:
1724-1725); here the row update runs in dqrm_sgd_rows on the de-duplicated row gradients that the fused
backward leaves in the table group:

    m[row] += mean_d(g[row]^2);   W[row] -= clr * g[row] / (sqrt(m[row]) + eps)       (rwsadagrad.py:97-113)

Dense parameters follow the reference's dense branch (plain Adagrad, rwsadagrad.py:115-118).
"""
from __future__ import annotations

import torch
from torch.optim import Optimizer


class RWSAdagrad(Optimizer):
    def __init__(self, params, lr=1e-2, lr_decay=0.0, weight_decay=0.0, initial_accumulator_value=0.0, eps=1e-10):
        if lr < 0 or lr_decay < 0 or weight_decay < 0 or initial_accumulator_value < 0 or eps < 0:
            raise ValueError("Invalid RWSAdagrad hyper-parameter")
        defaults = dict(lr=lr, lr_decay=lr_decay, eps=eps, weight_decay=weight_decay,
                        initial_accumulator_value=initial_accumulator_value)
        super().__init__(params, defaults)
        self._groups = []          # (EmbeddingTableGroup, momentum tensors, step counter)

    def attach_table_group(self, group, fused=False):
        """Register a fused table group; its tables get row-wise state and are updated by step().  With `fused` the
        row update runs inside the group's de-duplicating backward (dqrm_embbag_bwd_sgd: one kernel sorts,
        de-duplicates and applies m[row] / W[row]); step() then only advances the learning-rate schedule."""
        init = self.defaults["initial_accumulator_value"]
        mom = [torch.full((n,), init, dtype=torch.float32, device=group.device) for n in group.rows]
        self._groups.append([group, mom, 0])
        if fused:
            group.enable_fused_update(self.defaults["lr"], momentum=mom, eps=self.defaults["eps"])
        return mom

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        d = self.defaults
        for rec in self._groups:
            group, mom, _ = rec
            rec[2] += 1
            clr = d["lr"] / (1.0 + (rec[2] - 1.0) * d["lr_decay"])
            if group.applied_fused:                  # this step's update ran inside the backward; set the next step's rate
                group.applied_fused = False
                group.fused_update["lr"] = d["lr"] / (1.0 + rec[2] * d["lr_decay"])
                continue
            group.sgd_apply(clr, inv_world=1.0, momentum=mom, eps=d["eps"])
        table_params = {id(w) for rec in self._groups for w in rec[0].weights}
        for pg in self.param_groups:
            for p in pg["params"]:
                if p.grad is None or id(p) in table_params:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("sparse gradients are handled through attach_table_group() (fused backward)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["sum"] = torch.full_like(p.data, pg["initial_accumulator_value"])
                st["step"] += 1
                grad = p.grad
                if pg["weight_decay"] != 0:
                    grad = grad.add(p.data, alpha=pg["weight_decay"])
                clr = pg["lr"] / (1.0 + (st["step"] - 1.0) * pg["lr_decay"])
                st["sum"].addcmul_(grad, grad, value=1.0)
                p.data.addcdiv_(grad, st["sum"].sqrt().add_(pg["eps"]), value=-clr)
        return loss
