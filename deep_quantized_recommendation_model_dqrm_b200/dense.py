"""Flat arena for the dense (MLP) parameters and their 8-bit per-channel quantised gradient
exchange.  All QuantLinear weights and biases of a model live in ONE fp32
buffer (parameters are views), their gradients in a second one, so that scale / quantise / apply are
one kernel launch each and the exchange is two all-reduces in total (the reference: 28 Gloo
all-reduces per step, sgd_quantized_gradients_parallel_comm.py:341-394)."""
from __future__ import annotations

import torch

from . import _lib, p2p as _p2p


class DenseArena:
    def __init__(self, layers, device):
        """layers: QuantLinear modules in reference order (bot_l then top_l)."""
        self.lib = _lib.load()
        self.layers = list(layers)
        self.device = device
        params = []
        for l in self.layers:
            params.append(l.weight)
            if l.bias is not None:
                params.append(l.bias)
        total = sum(p.numel() for p in params)
        self.flat = torch.empty(total, dtype=torch.float32, device=device)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=device)
        chan = [0]
        off = 0
        for p in params:
            n = p.numel()
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view(p.shape)
            if p.grad is not None:                # adopt a gradient produced before the arena existed
                self.flat_grad[off:off + n].copy_(p.grad.reshape(-1))
            p.grad = self.flat_grad[off:off + n].view(p.shape)
            if p.dim() == 2:                      # weight: one channel per output row  (sgd...:905-906)
                for r in range(p.shape[0]):
                    chan.append(off + (r + 1) * p.shape[1])
            else:                                 # bias: one scalar scale for the vector (sgd...:945-947)
                chan.append(off + n)
            off += n
        self.params = params
        self.total = total
        self.num_chan = len(chan) - 1
        self.chan_begin = torch.tensor(chan, dtype=torch.int64, device=device)
        # fake-quantised copies (same layout as `flat`) + per-output-row scales, refreshed by fakequant_all()
        self.flat_int = torch.zeros(total, dtype=torch.float32, device=device)
        nrows = sum(l.weight.shape[0] for l in self.layers)
        self.fc_scale = torch.zeros(nrows, dtype=torch.float32, device=device)
        off, r0 = 0, 0
        Wp, bp, Wi, bi, sp, outs, ins = [], [], [], [], [], [], []
        for l in self.layers:
            o, i = l.weight.shape
            l._w_int = self.flat_int[off:off + o * i].view(o, i)
            Wp.append(l.weight.data.data_ptr()); Wi.append(l._w_int.data_ptr())
            off += o * i
            if l.bias is not None:
                l._b_int = self.flat_int[off:off + o]
                bp.append(l.bias.data.data_ptr()); bi.append(l._b_int.data_ptr())
                off += o
            else:
                l._b_int = None
                bp.append(None); bi.append(None)
            l._fc_scale = self.fc_scale[r0:r0 + o]
            sp.append(l._fc_scale.data_ptr())
            r0 += o
            outs.append(o); ins.append(i)
        import ctypes as C
        n = len(self.layers)
        mk = lambda vals: (C.c_void_p * n)(*vals)
        self._fq_args = (mk(Wp), mk(bp), (C.c_int32 * n)(*outs), (C.c_int32 * n)(*ins), mk(Wi), mk(bi), mk(sp))
        self.fused_ok = n <= 16 and len({l.weight_bit for l in self.layers}) == 1 and \
            all((l.bias is None) or (l.quantize_bias and l.bias_bit == l.weight_bit) for l in self.layers)
        # weight-gradient GEMMs run on a side stream, off the critical dx chain (None = same stream)
        import os
        self.side_stream = torch.cuda.Stream(device=device, priority=-1) \
            if (self.flat.is_cuda and not os.environ.get("DQRM_NO_SIDE_STREAM")) else None
        # the 7 weight-gradient GEMMs are independent of each other (each needs only its layer's dout): two more
        # streams, used round-robin, let one start the moment its dout exists instead of queueing behind the previous
        # layer's (on ONE side stream they had become the step's critical path: 55 us back to back at batch 128)
        n_side = max(1, min(4, int(os.environ.get("DQRM_DW_STREAMS", "3"))))
        self.side_streams = ([self.side_stream] + [torch.cuda.Stream(device=device, priority=-1) for _ in range(n_side - 1)]) \
            if self.side_stream is not None else []
        self._side_rr = 0
        self.keepalive = []
        for l in self.layers:
            l._arena = self
        self.scale_local = torch.zeros(self.num_chan, dtype=torch.float32, device=device)
        self.scale_mean = torch.zeros(self.num_chan, dtype=torch.float32, device=device)
        self.codes = torch.zeros(total, dtype=torch.float32, device=device)
        self.p2p = None                    # PeerArena of the NVLink exchange (world > 1, DQRM_EXCHANGE=p2p)
        self.slot_world = 0                # world size the exchange slots were built for (0: all-reduce form)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        # error compensation of the quantised MLP gradients (quantize_linear_grad / quantize_bias_grad with
        # err_compensation=True, sgd...parallel_comm.py:899-900,926-927,938-939,958-959): the residual
        # (grad + ec) - qbar * s_bar of every element, carried to the next step.  Off by default like every
        # call site of the reference (:341,350); allocated on first use.
        self.error_compensation = False
        self.ec = None
        self.lr_dev = None                 # device fp32 [1]: when set, apply() reads the learning rate from it
        self.fuse_local = False            # world 1: quantize_exchange() + apply() as ONE launch (graph_step sets it)
        self._local_pending = None
        # world > 1 over NVLink: scale -> exchange -> quantise -> exchange -> update as ONE launch issued by apply()
        # (csrc/dense_xchg.cu; graph_step sets it -- between the two calls scale_mean is not up to date yet)
        self.fuse_xchg = False             # (lr_dev set: issued by quantize_exchange() already, beside the embedding exchange)
        self._xchg_pending = None
        self._xchg_plan = None
        self.xchg_early = False            # ... and the top MLP's share already from inside the backward (after_dw)
        self.early_from_layer = 0          # index of the first top-MLP layer in `layers` (graph_step sets both)
        self.xchg_stream = None
        self._xchg_early_done = False
        self.lazy_zero = False             # zero_grad() only marks the layers clean (see zero_grad)
        self._bind_scale_views()

    def _bind_scale_views(self):
        """weight_scaling_factor / bias_scaling_factor of every layer = views of scale_mean (sgd...:343,352)."""
        c = 0
        for l in self.layers:
            o = l.weight.shape[0]
            l.weight_scaling_factor = self.scale_mean[c:c + o]
            c += o
            if l.bias is not None:
                l.bias_scaling_factor = self.scale_mean[c]
                c += 1

    def intact(self):
        """True while every parameter (and its grad) still aliases the arena."""
        base, gbase = self.flat.data_ptr(), self.flat_grad.data_ptr()
        off = 0
        for p in self.params:
            if p.data.data_ptr() != base + 4 * off or p.grad is None or p.grad.data_ptr() != gbase + 4 * off:
                return False
            off += p.numel()
        return True

    def fakequant_all(self):
        """Fake-quantise every layer's weight and bias in ONE launch (qm:125-154 for all layers); binds the
        results to the modules' ``weight_integer / bias_integer / fc_scaling_factor`` attributes."""
        Wp, bp, outs, ins, Wi, bi, sp = self._fq_args
        rc = self.lib.dqrm_mlp_fakequant_all(len(self.layers), Wp, bp, outs, ins, int(self.layers[0].weight_bit),
                                             Wi, bi, sp, _lib.stream_ptr())
        _lib.check(rc, "dqrm_mlp_fakequant_all")
        if not getattr(self, "_int_views_bound", False):
            for l in self.layers:
                l.weight_integer, l.bias_integer, l.fc_scaling_factor = l._w_int, l._b_int, l._fc_scale
            self._int_views_bound = True

    def check_status(self):
        """Device-side errors of the NVLink exchange (D2H sync).  A timeout is fatal and sticky: the bit is left set,
        so the apply kernels stay no-ops and every later poll raises again."""
        s = int(self.status.item())
        if s & _lib.STATUS_P2P_TIMEOUT:
            raise _p2p.ExchangeTimeout("a peer never signalled an NVLink exchange site within DQRM_P2P_TIMEOUT_S; the "
                                       "MLP update was NOT applied and replicas can no longer be trusted -- abort the job")
        if s:
            self.status.zero_()
            raise RuntimeError(f"dqrm dense exchange status {s}")

    def join(self):
        """Make the current stream wait for the side-stream weight-gradient kernels of this step."""
        if self.side_stream is not None and self.keepalive:
            cur = torch.cuda.current_stream()
            for st in self.side_streams:
                cur.wait_stream(st)
            self.keepalive.clear()
            self._side_rr = 0

    def next_side_stream(self):
        """Stream for the next weight-gradient GEMM (None: run it in line)."""
        if self.side_stream is None:
            return None
        st = self.side_streams[self._side_rr % len(self.side_streams)]
        self._side_rr += 1
        return st

    def zero_grad(self):
        """clear_gradients() for the MLP arena.  The fused backward OVERWRITES a clean layer's gradient (accumulate = 0),
        so when every layer's fused backward is known to run before the gradients are read (graph_step sets
        `lazy_zero`), marking the layers clean is all there is to do: no fill kernel in the step."""
        self.join()
        if not self.lazy_zero:
            self.flat_grad.zero_()
        for l in self.layers:
            l._grad_dirty = False

    def xchg_partition(self, num_ctas=None, split_chan=None):
        """Contiguous runs of channels for the CTAs of the one-kernel exchange, balanced by elements + a per-channel
        cost; the same on every rank (pure function of the layer shapes).  With `split_chan` no run crosses that
        channel, so the CTAs before / from `split_cta` can be launched on their own (late / early bucket).
        -> dict(cta_chan, cta_word, elems, chans, split_cta)"""
        import bisect
        cb = self.chan_begin.cpu().tolist()
        cost = [cb[c + 1] - cb[c] + 24 for c in range(self.num_chan)]
        pre = [0]
        for x in cost:
            pre.append(pre[-1] + x)
        tot = pre[-1]
        if num_ctas is None:
            num_ctas = max(1, min(148, tot // 2048))
        num_ctas = max(1, min(int(num_ctas), self.num_chan))

        def cut(c_lo, c_hi, k):                                   # k runs over channels [c_lo, c_hi)
            cuts = [c_lo]
            for b in range(1, k):                                 # cut b at the channel boundary nearest b/k of the cost
                target = pre[c_lo] + (pre[c_hi] - pre[c_lo]) * b / k
                c = bisect.bisect_left(pre, target)
                if c > 0 and target - pre[c - 1] < pre[min(c, self.num_chan)] - target:
                    c -= 1
                cuts.append(min(max(c, cuts[-1] + 1), c_hi - (k - b)))    # every CTA owns at least one channel
            return cuts

        if split_chan is None or not (0 < split_chan < self.num_chan) or num_ctas < 2:
            cuts, split_cta = cut(0, self.num_chan, num_ctas) + [self.num_chan], 0
        else:
            k_lo = min(max(1, round(num_ctas * pre[split_chan] / tot)), num_ctas - 1, split_chan)
            k_hi = min(num_ctas - k_lo, self.num_chan - split_chan)
            cuts, split_cta = cut(0, split_chan, k_lo) + cut(split_chan, self.num_chan, k_hi) + [self.num_chan], k_lo
        n = len(cuts) - 1
        elems = max(cb[cuts[b + 1]] - cb[cuts[b]] for b in range(n))
        chans = max(cuts[b + 1] - cuts[b] for b in range(n))
        words = [0]                                               # seven int8 codes per exchanged word, runs padded to a word
        for b in range(n):
            words.append(words[-1] + (cb[cuts[b + 1]] - cb[cuts[b]] + 6) // 7)
        return dict(cta_chan=cuts, cta_word=words, elems=elems, chans=chans, split_cta=split_cta)

    def _ensure_slots(self, world):
        """Per-rank slots of the two MLP exchange sites (channel scales fp32, codes int8): views of this rank's peer
        arena (NVLink form, collective on first use) or plain device buffers gathered by NCCL.  Same layout, same
        consumers, so the two transports give bit-identical updates."""
        if self.slot_world == world:
            return
        import torch.distributed as dist
        self.release()
        live = dist.is_available() and dist.is_initialized() and dist.get_world_size() == world
        a = None
        if live and _p2p.backend() == "p2p":
            try:
                part = self.xchg_partition(split_chan=self._early_split_chan())
                a = _p2p.PeerArena({"mlp_scale": self.num_chan * 4, "mlp_codes": self.total,
                                    "mlp_xscale": self.num_chan * 8, "mlp_xcodes": part["cta_word"][-1] * 8},
                                   world, dist.get_rank(), self.device)
            except _p2p.P2PUnavailable as e:         # raised on every rank together
                _p2p.fall_back_to_nccl(e)
        if a is not None:
            self.p2p = a
            self._scale_slots = a.slots("mlp_scale", torch.float32)
            self._code_slots = a.slots("mlp_codes", torch.int8)
            rank = a.rank
            self.bind_xchg(a, part)
        else:
            pad = lambda n: (n + 15) // 16 * 16
            self._scale_slots = torch.zeros((world, pad(self.num_chan * 4) // 4), dtype=torch.float32, device=self.device)
            self._code_slots = torch.zeros((world, pad(self.total)), dtype=torch.int8, device=self.device)
            rank = dist.get_rank() if live else 0
        self.slot_world, self.slot_rank = world, rank
        self.scale_local = self._scale_slots[rank, :self.num_chan]
        self._codes_mine = self._code_slots[rank, :self.total]

    def bind_xchg(self, arena, part):
        """Plan of the one-kernel exchange on `arena` (sites mlp_xscale / mlp_xcodes): the per-CTA channel runs
        (xchg_partition) and the device-side sequence numbers (zero like the fresh arena's words)."""
        G = len(part["cta_chan"]) - 1
        self._xchg_plan = dict(arena=arena, num_ctas=G, elems=int(part["elems"]), chans=int(part["chans"]),
                               cta_chan=torch.tensor(part["cta_chan"], dtype=torch.int32, device=self.device),
                               cta_word=torch.tensor(part["cta_word"], dtype=torch.int32, device=self.device),
                               seq=torch.zeros(G, dtype=torch.int32, device=self.device),
                               split_cta=int(part.get("split_cta", 0)))

    def _early_split_chan(self):
        """First channel of layer `early_from_layer` (the top MLP): the early bucket of the exchange (0: no split)."""
        k = self.early_from_layer
        if not k or k >= len(self.layers):
            return None
        return sum(l.weight.shape[0] + (1 if l.bias is not None else 0) for l in self.layers[:k])

    def exchange_apply_fused(self, lr, bits=8, part="all"):
        """quantize_exchange() + apply() of a multi-rank step in ONE launch (csrc/dense_xchg.cu); bit-identical to
        local_scale -> all-gather -> dqrm_dense_grad_quant_gathered -> all-gather -> dqrm_dense_apply_gathered.
        part = "early" / "late": only the CTAs from / before the plan's split (channels are independent)."""
        pl = self._xchg_plan
        a = pl["arena"]
        b0, b1 = {"all": (0, pl["num_ctas"]), "late": (0, pl["split_cta"]), "early": (pl["split_cta"], pl["num_ctas"])}[part]
        if b1 <= b0:
            return
        sc, co = a.sites["mlp_xscale"], a.sites["mlp_xcodes"]
        rc = self.lib.dqrm_dense_exchange_apply(a.ptrs, a.world, a.rank, sc["data_off"], sc["stride"], co["data_off"],
                                                co["stride"], self.flat.data_ptr(), self.flat_grad.data_ptr(),
                                                self._ec_ptr(), self.chan_begin.data_ptr(),
                                                pl["cta_chan"].data_ptr() + 4 * b0, pl["cta_word"].data_ptr() + 4 * b0,
                                                b1 - b0, pl["elems"], pl["chans"], int(bits), self.scale_mean.data_ptr(),
                                                pl["seq"].data_ptr() + 4 * b0, float(lr), _lib.ptr(self.lr_dev),
                                                self.status.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "dqrm_dense_exchange_apply")

    def after_dw(self, module):
        """Called by the fused backward right after a layer's weight-gradient GEMM was issued.  When the LAST top-MLP
        layer (in backward order) is through, the top MLP's share of the dense exchange + update (4/5 of the bytes at
        Kaggle shape) starts on its own stream, beside the interaction / bottom-MLP backward, instead of behind the last
        bottom-layer GEMM; quantize_exchange() then only issues the bottom MLP's share.  Safe: the backward reads the
        fake-quantised copies (flat_int), never the parameters this updates."""
        if not (self.xchg_early and self.fuse_xchg and self.lr_dev is not None and self._xchg_plan is not None
                and self._xchg_plan["split_cta"] > 0 and module is self.layers[self.early_from_layer]):
            return
        if self.xchg_stream is None:
            self.xchg_stream = torch.cuda.Stream(device=self.device, priority=-1)
        for st in self.side_streams[:min(self._side_rr, len(self.side_streams))]:     # (only streams used this step: an
            self.xchg_stream.wait_stream(st)                                          # idle one is outside a graph capture)
        with torch.cuda.stream(self.xchg_stream):
            self.exchange_apply_fused(0.0, 8, part="early")
        self._xchg_early_done = True

    def release(self):
        """Close the MLP exchange arena (barrier first: no rank may still be storing into it).  Collective."""
        a, self.p2p = self.p2p, None
        self._xchg_plan = None
        if a is None:
            return
        import torch.distributed as dist
        live = dist.is_available() and dist.is_initialized()
        torch.cuda.synchronize()
        if live:
            dist.barrier()
        self._scale_slots = self._code_slots = self._codes_mine = None
        self.scale_local = torch.zeros(self.num_chan, dtype=torch.float32, device=self.device)
        self.slot_world = 0
        a.close(dist.barrier if live else None)

    def _allgather(self, site, slots, process_group):
        if self.p2p is not None:
            self.p2p.allgather(site, self.status)
        else:
            import torch.distributed as dist
            dist.all_gather_into_tensor(slots.view(-1), slots[self.slot_rank], group=process_group)

    def quantize_exchange(self, world=1, process_group=None, bits=8, quantized=True):
        """quantize_linear_grad / quantize_bias_grad for every tensor at once
        (sgd_quantized_gradients_parallel_comm.py:892-961): local scales -> SUM over ranks ->
        quantise with the mean scale -> SUM of the codes over ranks.  Both sums are taken in rank order by the
        consumer kernels from all-gathered slots (deterministic, identical on every rank); the codes travel as
        int8.  Transport: one-kernel NVLink all-gathers (default) or NCCL all-gathers (DQRM_EXCHANGE=nccl)."""
        if world > 1:
            import torch.distributed as dist
        self.join()
        self.exchanged_gathered = False
        if not quantized:
            self.codes.copy_(self.flat_grad)
            if world > 1:
                dist.all_reduce(self.codes, group=process_group)
            return
        live = world > 1 and dist.is_available() and dist.is_initialized() and dist.get_world_size() == world
        if live and bits <= 8:
            self._ensure_slots(world)
            if self.fuse_xchg and self.p2p is not None and self._xchg_plan is not None:
                # the whole exchange + update in ONE launch.  With the learning rate in device memory it is issued HERE
                # (grad_update_parallel_comm: before the embedding streams are joined, so it runs beside the embedding
                # exchange) and apply() has nothing left to do; otherwise apply() issues it with its lr argument.
                if self.lr_dev is not None:
                    early, self._xchg_early_done = self._xchg_early_done and bits == 8, False
                    self.exchange_apply_fused(0.0, bits, part="late" if early else "all")
                    if early:                    # (inside a captured graph the position of this join is the dependency)
                        torch.cuda.current_stream().wait_stream(self.xchg_stream)
                    self._xchg_pending = "done"
                else:
                    self._xchg_pending = bits
                return
            self.local_scale(bits)                                  # -> this rank's slot of the scale site
            self._allgather("mlp_scale", self._scale_slots, process_group)
            rc = self.lib.dqrm_dense_grad_quant_gathered(self.flat_grad.data_ptr(), self.chan_begin.data_ptr(),
                                                         self.num_chan, self._scale_slots.data_ptr(),
                                                         self._scale_slots.stride(0), world, bits,
                                                         self._codes_mine.data_ptr(), self.scale_mean.data_ptr(),
                                                         _lib.stream_ptr())
            _lib.check(rc, "dqrm_dense_grad_quant_gathered")
            self._allgather("mlp_codes", self._code_slots, process_group)
            self.exchanged_gathered = True
            return
        # single rank, emulated ranks (tests copy between replicas) or > 8-bit codes: the all-reduce form
        if self.slot_world != 0:
            self.scale_local = torch.zeros(self.num_chan, dtype=torch.float32, device=self.device)
            self.slot_world = 0
        if world == 1 and self.fuse_local:
            self._local_pending = bits       # scale + quantise + update in ONE launch, issued by apply()
            return
        self.local_scale(bits)
        if world > 1 and live:
            dist.all_reduce(self.scale_local, group=process_group)
        self.quantize(world, bits)
        if world > 1 and live:
            dist.all_reduce(self.codes, group=process_group)

    def _ec_ptr(self):
        if not self.error_compensation:
            return None
        if self.ec is None:
            self.ec = torch.zeros(self.total, dtype=torch.float32, device=self.device)
            off = 0
            for l in self.layers:                         # the reference's per-layer buffers (qm:87,95) as views
                o, i = l.weight.shape
                l.error_compensation_weight = self.ec[off:off + o * i].view(o, i)
                off += o * i
                if l.bias is not None:
                    l.error_compensation_bias = self.ec[off:off + o]
                    off += o
        return self.ec.data_ptr()

    def local_scale(self, bits=8):
        """Per-channel local scales; with error compensation flat_grad becomes grad + ec in place first."""
        self.join()
        _lib.check(self.lib.dqrm_dense_grad_scale(self.flat_grad.data_ptr(), self._ec_ptr(), self.chan_begin.data_ptr(),
                                                  self.num_chan, bits, self.scale_local.data_ptr(), _lib.stream_ptr()),
                   "dqrm_dense_grad_scale")

    def quantize(self, world=1, bits=8):
        """scale_local must hold the SUM over ranks of the local scales."""
        _lib.check(self.lib.dqrm_dense_grad_quant(self.flat_grad.data_ptr(), self.chan_begin.data_ptr(), self.num_chan,
                                                  self.scale_local.data_ptr(), float(1.0 / world), bits,
                                                  self.codes.data_ptr(), self.scale_mean.data_ptr(), _lib.stream_ptr()),
                   "dqrm_dense_grad_quant")

    def apply(self, lr, world=1, quantized=True):
        """MLP half of weight_update_parallel_comm (sgd_quantized_gradients_parallel_comm.py:630-663)."""
        st = _lib.stream_ptr()
        ec = self._ec_ptr() if quantized else None
        if self._xchg_pending is not None:
            bits, self._xchg_pending = self._xchg_pending, None
            assert quantized and world == self.slot_world
            if bits != "done":
                self.exchange_apply_fused(lr, bits)
            return
        if self._local_pending is not None:
            bits, self._local_pending = self._local_pending, None
            assert quantized and world == 1
            rc = self.lib.dqrm_dense_quant_apply_local(self.flat.data_ptr(), self.flat_grad.data_ptr(), ec,
                                                       self.chan_begin.data_ptr(), self.num_chan, bits,
                                                       self.scale_local.data_ptr(), self.codes.data_ptr(),
                                                       self.scale_mean.data_ptr(), float(lr), _lib.ptr(self.lr_dev), st)
            _lib.check(rc, "dqrm_dense_quant_apply_local")
            return
        comp = self.flat_grad.data_ptr() if ec is not None else None        # = grad + ec since local_scale()
        if quantized and getattr(self, "exchanged_gathered", False):
            rc = self.lib.dqrm_dense_apply_gathered(self.flat.data_ptr(), self._code_slots.data_ptr(),
                                                    self._code_slots.stride(0), world, self.chan_begin.data_ptr(),
                                                    self.num_chan, self.scale_mean.data_ptr(), float(lr),
                                                    _lib.ptr(self.lr_dev), comp, ec, self.status.data_ptr(), st)
            _lib.check(rc, "dqrm_dense_apply_gathered")
            return
        _lib.check(self.lib.dqrm_dense_apply(self.flat.data_ptr(), self.codes.data_ptr(), self.chan_begin.data_ptr(),
                                             self.num_chan, self.scale_mean.data_ptr() if quantized else None,
                                             float(1.0 / world), float(lr), _lib.ptr(self.lr_dev), comp, ec, st),
                   "dqrm_dense_apply")
