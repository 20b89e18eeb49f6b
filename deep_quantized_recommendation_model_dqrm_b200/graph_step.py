"""The custom-DP training iteration (dlrm_s_pytorch_comm_grad.py:1909-1957) captured as CUDA graphs.

The reference step is launch- and sync-bound (~1.6k ATen launches, >= 53 host syncs, 80 Gloo
collectives at Kaggle shape).  Every kernel of this implementation is sync-free and its shapes are
static (fixed-capacity exchange slots), so one iteration is captured once and replayed: static input
buffers are refilled (H2D or D2D), the table scan is launched (kept outside the graph so bench.py can
bracket it with CUDA events inside the timed region), and the graph runs forward, loss, backward,
de-duplication, the two embedding all-gathers, the MLP all-reduces and all updates.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from . import dlrm_s_pytorch_comm_grad as drv


def _packed_layout(X, lS_o, lS_i, T):
    """Byte layout of one batch in a single staging buffer: X | lS_o | lS_i | T, each 16-byte aligned."""
    lay, off = [], 0
    for t in (X, lS_o, lS_i, T):
        n = t.numel() * t.element_size()
        lay.append((off, n, t.dtype, tuple(t.shape)))
        off += (n + 15) // 16 * 16
    return lay, off


def self_pipelined(dlrm):
    return dlrm._ensure_group().scale_policy == "pipelined"


class GraphedTrainStep:
    def __init__(self, dlrm, X, lS_o, lS_i, T, lr, world_size=1, rank=0, grad_bits=8, warmup=3, use_graph=True,
                 mlp_layer_quantized=True):
        """X, lS_o, lS_i, T: an example LOCAL batch (this rank's shard) fixing the static shapes;
        lS_i / lS_o must be stacked [T, B] tensors (Criteo shape).

        With ``group.scale_policy == "pipelined"`` the table rescan runs on the group's low-priority side stream
        concurrently with the step (tables.py); the step itself must then run on ``self.stream`` (high priority)
        so that its short kernels are scheduled ahead of the scan's CTAs:
        ``with torch.cuda.stream(step.stream): step.load(...); step.run()``."""
        self.dlrm, self.lr, self.world, self.rank = dlrm, float(lr), world_size, rank
        self.grad_bits = grad_bits
        self.mlp_layer_quantized = mlp_layer_quantized
        dev = next(dlrm.parameters()).device
        self.device = dev
        # the static inputs are views of ONE staging buffer, so a host batch packed the same way (pack_host) arrives
        # with a single H2D copy (load_packed); load() still accepts the four tensors separately
        self._layout, nbytes = _packed_layout(X, lS_o, lS_i, T)
        self._stage = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        views = [self._stage[o:o + n].view(dt).view(shape) for (o, n, dt, shape) in self._layout]
        self.X, self.lS_o, self.lS_i, self.T = views
        for v, src in zip(views, (X, lS_o, lS_i, T)):
            v.copy_(src)
        self.loss = torch.zeros((), device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev, priority=-1)
        self._copy_done, self._copy_pending = torch.cuda.Event(), False
        # the learning rate lives in device memory: the update kernels read it at run time, so set_lr() takes effect on
        # the next replay of the captured graph (LRPolicyScheduler warm-up / decay without re-capture)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.fused_loss = dlrm.loss_function == "bce" and not (0.0 < dlrm.loss_threshold < 1.0)
        self.group = dlrm._ensure_group()
        self.group.dp_world, self.group.dp_rank = world_size, rank
        self.group.lr_dev = self.lr_dev
        from .sgd_quantized_gradients_parallel_comm import _dense_arena as _arena_of
        _arena_of(dlrm).lr_dev = self.lr_dev
        # one rank: nothing to exchange between the MLP gradient quantisation and the update -> one launch, same bits
        _arena_of(dlrm).fuse_local = world_size == 1 and os.environ.get("DQRM_FUSE_LOCAL_DENSE", "1") != "0"
        # N ranks over NVLink: the five launches of the dense exchange (scale, all-gather, quantise, all-gather, update)
        # as ONE kernel whose CTAs synchronise pairwise with their peers (csrc/dense_xchg.cu), same bits
        _arena_of(dlrm).fuse_xchg = world_size > 1 and os.environ.get("DQRM_FUSE_DENSE_XCHG", "1") != "0"
        # ... optionally (DQRM_DENSE_XCHG_EARLY=1) the top MLP's share of it is issued from inside the backward, as soon
        # as the top MLP's weight gradients exist (DenseArena.after_dw).  Measured at two GPUs and OFF by default: the
        # launch is latency-, not byte-bound, so the bottom MLP's share left for the tail takes as long as the whole
        # (18 us), while the early launch slows the bottom-MLP backward it runs beside (step 0.293 -> 0.341 ms)
        from .quantization_supp.quant_modules import QuantLinear as _QL
        _arena_of(dlrm).early_from_layer = sum(1 for l in dlrm.bot_l if isinstance(l, _QL))
        _arena_of(dlrm).xchg_early = (world_size > 1 and mlp_layer_quantized and not self_pipelined(dlrm) and
                                      os.environ.get("DQRM_DENSE_XCHG_EARLY", "0") == "1")
        # every step runs the fused backward of all 7 layers, which overwrites clean gradients: no zero-fill launch
        _arena_of(dlrm).lazy_zero = X.shape[0] <= dlrm.fuse_mlp_max_batch and dlrm._fused_mlp_arena() is not None
        self.pipelined = self.group.scale_policy == "pipelined"
        # row-sharded scan: the absmax exchange + scale are issued right before the embedding forward (inside the
        # graph, after the bottom MLP) instead of right behind the scan -- its round trip is off the critical path
        self.group.defer_scan_reduce = world_size > 1 and self.group.scale_policy == "full"
        # better still (graph replays): the exchange runs on its own stream BESIDE the bottom MLP, which is replayed
        # from its own small graph between the scan launch and the join (pre_mode "after_scan", below)
        self.group.side_scan_reduce = (self.group.defer_scan_reduce and use_graph and dlrm.shard_scan and
                                       os.environ.get("DQRM_SIDE_SCAN_REDUCE", "1") != "0")
        # multi-rank: launch the embedding exchange from inside the backward (side stream), overlapping the two
        # all-gathers + pack with the bottom-MLP backward; grad_update_parallel_comm then only joins
        self.group.eager_exchange = (world_size > 1 and self.group.grad_bit == grad_bits and not self.pipelined and
                                     os.environ.get("DQRM_EAGER_EXCHANGE", "1") != "0")
        # ... and the de-duplicating backward kernel itself: the bottom-MLP backward does not depend on it
        self.group.side_backward = not self.pipelined and os.environ.get("DQRM_SIDE_BACKWARD", "1") != "0"
        # ... and, behind the exchange on that stream, the row update (lr comes from lr_dev)
        self.group.eager_apply = (self.group.side_backward and self.group.grad_bit == grad_bits and
                                  (world_size == 1 or self.group.eager_exchange) and
                                  os.environ.get("DQRM_EAGER_APPLY", "1") != "0")
        self.stream = torch.cuda.Stream(device=dev, priority=-1)
        if self.pipelined:
            # measured on B200: a graph with forked branches is not co-scheduled with the side-stream pass (its
            # kernels wait for the whole scan); a linear graph is.  Keep the weight-gradient GEMMs in line.
            from .sgd_quantized_gradients_parallel_comm import _dense_arena
            _dense_arena(dlrm).side_stream = None
        dlrm.external_scan = True
        self.graph = self.graph_b = self.graph_pre = None
        self.scan_in_graph = False
        self._xb = None
        # The bottom MLP (and the MLP weight fake-quantisation) does not depend on the table scales, so they CAN be
        # captured into their own linear graph and replayed on a second stream beside the scan kernel
        # (DQRM_OVERLAP_BOTTOM=1, or =force regardless of the scan size; serial-slice FFMA forward: the bits of the
        # cluster split-K kernel without a cluster launch).  Measured on B200 and OFF by default: with
        # the scan's default L1/shared split the GEMMs cannot become resident before the pass drains (an SM's split only
        # changes while it is empty: step 0.440 -> 0.453 ms), and with a shared-heavy split the pass drops to 5.5 TB/s
        # while every dependent load of the co-running kernels sees the loaded HBM latency (fake-quant 4 -> 88 us, one
        # 512x256 layer 24 -> 229 us: step 0.497 ms).
        from . import _lib as _l
        scan_rows = sum(int(w.shape[0]) for w in self.group.weights)
        scan_bytes = scan_rows * self.group.dim * 4 // (world_size if (world_size > 1 and dlrm.shard_scan) else 1)
        mode = os.environ.get("DQRM_OVERLAP_BOTTOM", "0")          # 0: never (default), 1: long scans, force: always (tests)
        self.overlap_bottom = (use_graph and self.group.scale_policy == "full" and mode != "0" and
                               (scan_bytes >= 400_000_000 or mode == "force") and
                               X.shape[0] <= dlrm.fuse_mlp_max_batch and dlrm._fused_mlp_arena() is not None)
        self.pre_after_scan = (self.group.side_scan_reduce and not self.overlap_bottom and
                               X.shape[0] <= dlrm.fuse_mlp_max_batch and dlrm._fused_mlp_arena() is not None)
        if not self.pre_after_scan:
            self.group.side_scan_reduce = False
        if self.overlap_bottom:
            self.pre_stream = torch.cuda.Stream(device=dev, priority=-1)
            self._pre_done = torch.cuda.Event()
            tc_min = int(os.environ.get("DQRM_MLP_TC_MIN_BATCH", "256"))
            if _l.linear_path == _l.LINEAR_FFMA or (_l.linear_path == _l.LINEAR_AUTO and X.shape[0] < tc_min):
                for layer in dlrm.bot_l:
                    if hasattr(layer, "forward_fused"):
                        layer.fwd_path = _l.LINEAR_FFMA_SERIAL
        torch.cuda.synchronize()
        with torch.cuda.stream(self.stream):
            self.group.pipe_external_join = False
            self.scan()
            for _ in range(max(warmup, 1)):           # eager warm-up: creates arenas / step buffers / cuBLAS handles
                self.scan()
                self._body_a()
                self._body_b()
            torch.cuda.synchronize()
            if use_graph:
                self.scan()
                self.graph = torch.cuda.CUDAGraph()
                if self.pipelined:
                    # two graphs: everything that does not need the scan (A), then update + fix-up + reduce (B);
                    # run() orders B after the side-stream pass with an event wait between the two replays
                    self.group.pipe_external_join = True
                    with torch.cuda.graph(self.graph, stream=self.stream):
                        self._body_a()
                    torch.cuda.current_stream().wait_event(self.group.pipe_event)
                    self.graph_b = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph_b, pool=self.graph.pool(), stream=self.stream):
                        self._body_b()
                elif self.overlap_bottom or self.pre_after_scan:
                    self.graph_pre = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph_pre, stream=self.pre_stream if self.overlap_bottom else self.stream):
                        self._xb = self.dlrm.forward_bottom(self.X)
                    self.group.finish_scan()              # (pre_after_scan: join the eager absmax exchange)
                    with torch.cuda.graph(self.graph, pool=self.graph_pre.pool(), stream=self.stream):
                        self._body_a()                    # forward(x_bottom=self._xb): the rest of the step
                        self._body_b()
                    self._xb = None
                else:
                    # DQRM_SCAN_IN_GRAPH=1 (experiment): the scan kernel as the first node of the step graph, so no graph
                    # launch sits between it and the forward (bench.py then cannot bracket it with events)
                    self.scan_in_graph = world_size == 1 and os.environ.get("DQRM_SCAN_IN_GRAPH", "0") == "1"
                    with torch.cuda.graph(self.graph, stream=self.stream):
                        if self.scan_in_graph:
                            self.scan()
                        self._body_a()
                        self._body_b()
                torch.cuda.synchronize()

    def scan(self, events=None):
        """(a1) one launch for all tables; with world > 1 each rank scans 1/world of the rows."""
        g = self.group
        sharded = self.world > 1 and self.dlrm.shard_scan
        g.scan_scales(shard_rank=self.rank if sharded else 0, shard_world=self.world if sharded else 1,
                      events=events)

    def _body_a(self):
        d = self.dlrm
        Z = d(self.X, self.lS_o, self.lS_i, x_bottom=self._xb)
        if self.fused_loss:
            # BCELoss(mean) + the first backward step in ONE launch (ATen: 8 small kernels), loss written in place
            Zd = Z.detach()
            dZ = torch.empty_like(Zd)
            _lib.check(_lib.load().dqrm_bce_loss_grad(Zd.data_ptr(), self.T.data_ptr(), Zd.numel(), self.loss.data_ptr(),
                                                      dZ.data_ptr(), _lib.stream_ptr()), "dqrm_bce_loss_grad")
            drv.clear_gradients(d)
            Z.backward(dZ)
        else:
            E = torch.nn.functional.binary_cross_entropy(Z, self.T)
            drv.clear_gradients(d)
            E.backward()
            self.loss.copy_(E.detach())
        drv.grad_update_parallel_comm(d, self.world, emb_grad_quantized=True, num_bits=self.grad_bits,
                                      ranking_range=False, rank_for_debug=self.rank,
                                      mlp_layer_quantized=self.mlp_layer_quantized)

    def _body_b(self):
        drv.weight_update_parallel_comm(self.dlrm, self.lr, emb_grad_quantized=True, update_embedding=True,
                                        num_gpus=self.world, rank_for_debug=self.rank,
                                        mlp_layer_quantized=self.mlp_layer_quantized)

    def _body(self):
        self._body_a()
        if self.pipelined and self.group.pipe_external_join:
            torch.cuda.current_stream().wait_event(self.group.pipe_event)
        self._body_b()

    def load(self, X, lS_o, lS_i, T):
        """Refill the static inputs (host pinned or device tensors); asynchronous."""
        self.X.copy_(X, non_blocking=True)
        self.lS_o.copy_(lS_o, non_blocking=True)
        self.lS_i.copy_(lS_i, non_blocking=True)
        self.T.copy_(T, non_blocking=True)

    def pack_host(self, X, lS_o, lS_i, T, pin=True):
        """Collate one host batch into the staging layout (one pinned uint8 tensor) for load_packed()."""
        buf = torch.zeros(self._stage.numel(), dtype=torch.uint8)
        if pin:
            buf = buf.pin_memory()
        for (o, n, dt, shape), t in zip(self._layout, (X, lS_o, lS_i, T)):
            assert tuple(t.shape) == shape and t.dtype == dt, "batch does not match the captured static shapes"
            buf[o:o + n].view(dt).view(shape).copy_(t)
        return buf

    def load_packed(self, packed):
        """Refill all static inputs with ONE (H2D or D2D) copy of a pack_host()-shaped buffer; asynchronous.
        The copy is issued on a copy stream, ordered after everything queued on the current stream (the previous step
        still reads the inputs) -- the table scan that run() launches next does not need the batch, so the transfer
        (a PCIe round trip from pinned host memory) hides behind it; run() joins before the first consumer."""
        cur = torch.cuda.current_stream()
        self.copy_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream):
            self._stage.copy_(packed, non_blocking=True)
            self._copy_done.record()
        self._copy_pending = True

    def _join_copy(self):
        if self._copy_pending:
            torch.cuda.current_stream().wait_event(self._copy_done)
            self._copy_pending = False

    def replay(self):
        """Everything after the scan launch, on the current stream."""
        if self.graph is None:
            self._body()
            return
        self.graph.replay()
        if self.graph_b is not None:
            torch.cuda.current_stream().wait_event(self.group.pipe_event)
            self.graph_b.replay()

    def run(self, events=None):
        """scan + (graph replay | eager body); returns the device loss tensor (no sync)."""
        if self.graph_pre is not None and self.pre_after_scan:
            # row-sharded scan: scan kernel -> [absmax exchange + scale on scan_stream] beside [fake-quant + bottom MLP
            # replayed here] -> join -> the rest of the step
            self.scan(events)
            self._join_copy()
            self.graph_pre.replay()
            self.group.finish_scan()
        elif self.graph_pre is not None:
            self._join_copy()
            # fake-quant + bottom MLP on their own stream, beside the scan: ordered after everything already queued on
            # the current stream (the previous step's update, this step's input copy), joined before the replay
            cur = torch.cuda.current_stream()
            self.pre_stream.wait_stream(cur)
            with torch.cuda.stream(self.pre_stream):
                self.graph_pre.replay()
                self._pre_done.record()
            self.scan(events)
            cur.wait_event(self._pre_done)
        elif not self.scan_in_graph:
            self.scan(events)
        elif events is not None:                                    # (no kernel in between: reads as 0)
            events[0].record(); events[1].record()
        self._join_copy()
        self.replay()
        return self.loss

    def set_lr(self, lr):
        """New learning rate for the following steps (asynchronous; ordered on the current stream before the next
        run()).  Works for graph replays and the eager body alike."""
        self.lr = float(lr)
        self._lr_host[0] = self.lr
        self.lr_dev.copy_(self._lr_host, non_blocking=True)

    def input_bytes(self):
        return int(self._stage.numel())
