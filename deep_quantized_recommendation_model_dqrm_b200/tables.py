"""Host-side state of a group of QAT embedding tables on one GPU.

`EmbeddingTableGroup` owns what the reference scatters over 26 module
instances and their autograd/optimizer side effects: the per-table scales
(``eb_scaling_factor``), the de-duplicated row gradients of the current step,
the gradient scales (``emb_scaling_factor``) and the exchange slots.  Every
method is a thin call into libdqrm_b200 on the current CUDA stream; nothing
here computes on the host or synchronises the device, so a whole train step is
CUDA-graph capturable.

Data layout in HBM:
  weights      T tables, fp32 [rows_k, D] row-major (optionally views of one arena)
  scale/inv    fp32 [T]                         (a1)
  out          fp32 [T, B, D]                   (a3)   codes int8/int16 [T, B, D]
  uniq_rows    int32 [T, cap]; uniq_count int32 [T]; grad_sums fp32 [T, cap, D]     (a5/a7-1)
  slot         count[T] | rows[T][cap] | codes[T][cap][D]   (a7-4), gathered = world slots
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, p2p as _p2p


def _dist_world():
    import torch.distributed as dist
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


class EmbeddingTableGroup:
    def __init__(self, weights, embedding_bit=4, grad_bit=8):
        lib = _lib.load()
        assert len(weights) >= 1
        w0 = weights[0]
        if not w0.is_cuda:
            raise _lib.DqrmLibraryError("EmbeddingTableGroup needs CUDA tensors: there is no CPU path")
        self.lib = lib
        self.weights = list(weights)                 # Parameters or tensors; .data re-read at every call
        self.T = len(weights)
        self.dim = int(w0.shape[1])
        self.rows = [int(w.shape[0]) for w in weights]
        self.device = w0.device
        self.embedding_bit = int(embedding_bit)
        self.grad_bit = int(grad_bit)
        self._rows_arr = _lib.i64_array(self.rows)
        dev = self.device
        self.absmax = torch.zeros(self.T, dtype=torch.float32, device=dev)
        self.scale = torch.zeros(self.T, dtype=torch.float32, device=dev)        # eb_scaling_factor per table
        self.inv_scale = torch.zeros(self.T, dtype=torch.float32, device=dev)
        self.scale_valid = False
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._scan_ws = torch.zeros(int(lib.dqrm_scan_workspace_bytes(self.T)), dtype=torch.uint8, device=dev)
        # step state
        self.capacity = 0
        self.world = 1
        self.bags = 0
        self.uniq_rows = self.uniq_count = self.grad_sums = None
        self.grad_scale_local = torch.zeros(self.T, dtype=torch.float32, device=dev)
        self.grad_scale_mean = torch.zeros(self.T, dtype=torch.float32, device=dev)  # emb_scaling_factor per table
        self.gathered_scales = None
        self.slot = self.gathered = None
        self.updated_rows = self.updated_count = self.qbar = None
        self._bwd_ws = None
        self.last = None          # (indices, offsets, idx_begin, idx_begin_arr, bags, full_precision)
        # (a1) scale policy: "full" = rescan every table on every call (reference semantics, HBM-bound);
        # "incremental" = exact block-max tracker (bit-identical scales, O(touched rows) per step);
        # "pipelined" = full rescan of every table on every step, but overlapped with the step on a side stream
        #               (block maxima + fix-up of the updated blocks after the row update; bit-identical scales)
        self.scale_policy = "full"
        self.pipe_stream = None            # side stream of the pipelined rescan (created on first use)
        self.pipe_event = None             # recorded after the block-max pass
        self.pipe_pending = False          # a block-max pass was issued since the last table update
        self.pipe_shard = (0, 1)
        self.pipe_group = None             # process group of the MAX all-reduce when the pass is row-sharded
        self.pipe_external_join = False    # the caller orders update-after-scan itself (graph replay)
        self.block_rows = 64
        self._bm_buf = None
        self._bm_valid = False
        self._bm_wptrs = None
        self.p2p = None             # PeerArena of the NVLink exchange (world > 1, DQRM_EXCHANGE=p2p)
        self.eager_exchange = False  # start the embedding exchange from inside the backward, on xchg_stream
        self.side_backward = False   # run the de-duplicating backward itself on xchg_stream (graph_step sets it)
        self._bwd_forked = None
        self.eager_apply = False     # ... and merge_apply() right behind the exchange, on the same stream
        self.applied_eagerly = False
        self.fused_update = None     # dict(lr, momentum, eps): single-process row update INSIDE the de-duplicating backward
        self.applied_fused = False   # ... which then already ran when the optimizer / weight_update_parallel_comm is called
        self.xchg_stream = None
        self.exchange_started = False
        self.dp_world, self.dp_rank = 1, 0
        self.fixed_capacity = None  # rows per table in the exchange slots (default: this step's largest table)
        self.keep_debug = False   # also emit updated_rows / qbar in merge (parity tests, .grad materialisation)
        # packed-INT4 shadow rows for the training forward (csrc/shadow.cu): None = off (fp32 rows only)
        self.shadow = None
        self.shadow_scale = self._shadow_flags = self._shadow_ptrs = self._shadow_buf = None
        self.lr_dev = None                 # device fp32 [1]: when set, the update kernels read the learning rate from it
        self.defer_scan_reduce = False     # sharded scan: leave the MAX over ranks to finish_scan() (called by forward)
        self._scan_reduce_pending = None
        self.side_scan_reduce = False      # ... or run it on scan_stream right behind the scan kernel (graph_step)
        self.scan_stream = None

    # ---- helpers --------------------------------------------------------
    def _wptrs(self):
        ws = []
        for w in self.weights:
            d = w.data
            if d.dtype != torch.float32 or not d.is_contiguous() or d.device != self.device:
                raise _lib.DqrmLibraryError("embedding weights must be contiguous fp32 on the group's device")
            ws.append(d)
        return _lib.ptr_array(ws)

    def check_status(self):
        """Device-side data errors (D2H sync). Raises on out-of-range indices etc."""
        s = int(self.status.item())
        if s & _lib.STATUS_P2P_TIMEOUT:
            # fatal and sticky: the bit stays set, merge_apply stays a no-op, every later poll raises again
            raise _p2p.ExchangeTimeout("a peer never signalled an NVLink exchange site within DQRM_P2P_TIMEOUT_S; the "
                                       "update was NOT applied and replicas can no longer be trusted -- abort the job")
        if s:
            self.status.zero_()
            names = [n for b, n in ((1, "index out of range"), (2, "offsets not monotone"), (4, "capacity exceeded"),
                                      (8, "a peer never signalled an NVLink exchange site (timeout)")) if s & b]
            raise IndexError("dqrm kernel status: " + ", ".join(names))

    @staticmethod
    def pack_inputs(lS_i, lS_o, device):
        """Reference batch formats -> (indices[L_total], offsets[T,B], idx_begin list, bags).
        lS_i/lS_o are a [T,B] int64 tensor pair (Criteo) or lists of 1-D tensors
        (random data); dlrm_data_pytorch.py:328-345, 1099-1157."""
        if torch.is_tensor(lS_i):
            T, B = lS_i.shape
            idx = lS_i.to(device=device, dtype=torch.int64).contiguous().view(-1)
            idx_begin = [k * B for k in range(T + 1)]
        else:
            lens = [int(t.shape[0]) for t in lS_i]
            idx = torch.cat([t.to(device=device, dtype=torch.int64).view(-1) for t in lS_i])
            idx_begin = [0]
            for n in lens:
                idx_begin.append(idx_begin[-1] + n)
        if torch.is_tensor(lS_o):
            off = lS_o.to(device=device, dtype=torch.int64).contiguous()
        else:
            off = torch.stack([t.to(device=device, dtype=torch.int64) for t in lS_o]).contiguous()
        return idx, off, idx_begin, int(off.shape[1])

    # ---- (a1) incremental tracker ------------------------------------------
    def _ensure_blockmax(self):
        if self._bm_buf is not None:
            return
        ent = [int(self.lib.dqrm_blockmax_entries(n, self.block_rows)) for n in self.rows]
        pad = [(e + 3) // 4 * 4 for e in ent]                       # keep every table's segment 16-byte aligned
        self._bm_buf = torch.zeros(max(sum(pad), 4), dtype=torch.float32, device=self.device)
        views, off = [], 0
        for p_ in pad:
            views.append(self._bm_buf[off:off + p_])
            off += p_
        self._bm_views = views
        self._bm_ptrs = _lib.ptr_array(views)
        self._bm_entries = _lib.i64_array(ent)

    def invalidate_tracker(self):
        """Call after mutating a table outside merge_apply / sgd_apply (e.g. loading a checkpoint)."""
        self._bm_valid = False
        self.invalidate_shadow()

    def _tracker_scan(self, events=None):
        lib, st = self.lib, _lib.stream_ptr()
        self._ensure_blockmax()
        wp = self._wptrs()
        cur = tuple(wp)
        if self._bm_wptrs != cur:                                  # a table's storage was replaced
            self._bm_valid, self._bm_wptrs = False, cur
        if not self._bm_valid:
            _lib.check(lib.dqrm_blockmax_build(self.T, wp, self._rows_arr, self.dim, self.block_rows, self._bm_ptrs, st),
                       "dqrm_blockmax_build")
            self._bm_valid = True
        if events is not None:
            events[0].record()
        rc = lib.dqrm_table_absmax_scale(self.T, self._bm_ptrs, self._bm_entries, 1, self.embedding_bit, 0, 1,
                                         self.absmax.data_ptr(), self.scale.data_ptr(), self.inv_scale.data_ptr(),
                                         self._scan_ws.data_ptr(), st)
        if events is not None:
            events[1].record()
        _lib.check(rc, "dqrm_table_absmax_scale(blockmax)")
        self.scale_valid = True

    def _tracker_update(self, from_slots):
        if self.scale_policy == "pipelined":
            return self._pipe_finish(from_slots)
        if self.scale_policy != "incremental" or not self._bm_valid:
            return
        st = _lib.stream_ptr()
        if from_slots:
            rc = self.lib.dqrm_blockmax_update(self.T, self._wptrs(), self._rows_arr, self.dim, self.block_rows,
                                               self._bm_ptrs, self.gathered.data_ptr(), self.world, self.capacity,
                                               self.grad_bit, None, None, st)
        else:
            rc = self.lib.dqrm_blockmax_update(self.T, self._wptrs(), self._rows_arr, self.dim, self.block_rows,
                                               self._bm_ptrs, None, 0, self.capacity, self.grad_bit,
                                               self.uniq_rows.data_ptr(), self.uniq_count.data_ptr(), st)
        _lib.check(rc, "dqrm_blockmax_update")

    # ---- (a1) pipelined full rescan ----------------------------------------
    def _pipe_reduce(self):
        """Block maxima of this rank's shard -> absmax -> (MAX all-reduce) -> scale, on the current stream."""
        lib, st = self.lib, _lib.stream_ptr()
        r, w = self.pipe_shard
        sharded = w > 1
        rc = lib.dqrm_blockmax_reduce(self.T, self._rows_arr, self.block_rows, self._bm_ptrs, r, w, self.embedding_bit,
                                      self._absmax_out(sharded).data_ptr(), None if sharded else self.scale.data_ptr(),
                                      None if sharded else self.inv_scale.data_ptr(), self._scan_ws.data_ptr(), st)
        _lib.check(rc, "dqrm_blockmax_reduce")
        if sharded:
            self._allreduce_absmax_to_scale(self.pipe_group)
        self.scale_valid = True

    def _pipe_pass(self, events=None):
        """The HBM-bound part: every byte of this rank's shard once -> block maxima (current stream)."""
        r, w = self.pipe_shard
        if events is not None:
            events[0].record()
        rc = self.lib.dqrm_blockmax_scan(self.T, self._wptrs(), self._rows_arr, self.dim, self.block_rows,
                                         self._bm_ptrs, r, w, _lib.stream_ptr())
        if events is not None:
            events[1].record()
        _lib.check(rc, "dqrm_blockmax_scan")

    def _pipelined_scan(self, shard_rank, shard_world, process_group, events):
        """Called where the reference rescans (before a forward).  The scale of THIS forward was produced by the
        previous update's fix-up (or is bootstrapped here); the pass issued now, on the side stream, feeds the
        scale of the NEXT forward and overlaps with this step."""
        self._ensure_blockmax()
        if (shard_rank, shard_world) != self.pipe_shard:
            self.pipe_shard, self.scale_valid = (shard_rank, shard_world), False
        self.pipe_group = process_group
        cur = torch.cuda.current_stream()
        if not self.scale_valid:                       # bootstrap: serial pass + reduce
            self._pipe_pass()
            self._pipe_reduce()
        if self.pipe_stream is None:
            self.pipe_stream = torch.cuda.Stream(device=self.device)      # default (lowest) priority
            self.pipe_event = torch.cuda.Event()
        side = self.pipe_stream
        side.wait_stream(cur)                          # the tables are final as of the current stream
        with torch.cuda.stream(side):
            self._pipe_pass(events)
            self.pipe_event.record(side)
        self.pipe_pending = True

    def _pipe_finish(self, from_slots):
        """After a row update: recompute the touched blocks, reduce -> scale of the next forward."""
        if not self.pipe_pending:
            self.scale_valid = False                   # tables changed with no pass in flight: next scan bootstraps
            return
        if not self.pipe_external_join:
            torch.cuda.current_stream().wait_event(self.pipe_event)
        st = _lib.stream_ptr()
        r, w = self.pipe_shard
        if from_slots:
            rc = self.lib.dqrm_blockmax_update_shard(self.T, self._wptrs(), self._rows_arr, self.dim, self.block_rows,
                                                     self._bm_ptrs, self.gathered.data_ptr(), self.world,
                                                     self.capacity, self.grad_bit, None, None, r, w, st)
        else:
            rc = self.lib.dqrm_blockmax_update_shard(self.T, self._wptrs(), self._rows_arr, self.dim, self.block_rows,
                                                     self._bm_ptrs, None, 0, self.capacity, self.grad_bit,
                                                     self.uniq_rows.data_ptr(), self.uniq_count.data_ptr(), r, w, st)
        _lib.check(rc, "dqrm_blockmax_update_shard")
        self._pipe_reduce()
        self.pipe_pending = False

    # ---- (a1) -----------------------------------------------------------
    def scan_scales(self, shard_rank=0, shard_world=1, process_group=None, events=None):
        """Recompute every table's scale from a full max-abs pass (one launch).
        With shard_world > 1 each rank scans 1/world of the rows and the maxima
        are combined with a MAX all-reduce (replicas are bit-identical)."""
        if self.scale_policy == "incremental":
            return self._tracker_scan(events)
        if self.scale_policy == "pipelined":
            return self._pipelined_scan(shard_rank, shard_world, process_group, events)
        lib, st = self.lib, _lib.stream_ptr()
        sharded = shard_world > 1
        if events is not None:            # (start, end) CUDA events bracketing exactly the scan kernel
            events[0].record()
        rc = lib.dqrm_table_absmax_scale(self.T, self._wptrs(), self._rows_arr, self.dim, self.embedding_bit,
                                         shard_rank, shard_world, self._absmax_out(sharded).data_ptr(),
                                         None if sharded else self.scale.data_ptr(),
                                         None if sharded else self.inv_scale.data_ptr(),
                                         self._scan_ws.data_ptr(), st)
        if events is not None:
            events[1].record()
        _lib.check(rc, "dqrm_table_absmax_scale")
        if sharded:
            if self.side_scan_reduce:
                # the absmax exchange + scale on their own stream, right behind the scan kernel: they run beside the
                # bottom MLP and cost the step nothing; finish_scan() (before the embedding forward) joins
                cur = torch.cuda.current_stream()
                if self.scan_stream is None:
                    self.scan_stream = torch.cuda.Stream(device=self.device, priority=-1)
                self.scan_stream.wait_stream(cur)
                with torch.cuda.stream(self.scan_stream):
                    self._allreduce_absmax_to_scale(process_group)
                self._scan_reduce_pending = "side"
            elif self.defer_scan_reduce:
                # the cross-rank MAX + scale only has to be done before the embedding forward: the bottom MLP runs in
                # between (dlrm_s_pytorch_comm_grad.py:855-857), which hides this exchange's round trip and rank skew
                self._scan_reduce_pending = (process_group,)
            else:
                self._allreduce_absmax_to_scale(process_group)
        self.scale_valid = True

    def finish_scan(self):
        """Complete a row-sharded scan whose cross-rank reduction was deferred (defer_scan_reduce)."""
        if self._scan_reduce_pending == "side":
            self._scan_reduce_pending = None
            torch.cuda.current_stream().wait_stream(self.scan_stream)
        elif self._scan_reduce_pending is not None:
            (pg,), self._scan_reduce_pending = self._scan_reduce_pending, None
            self._allreduce_absmax_to_scale(pg)

    # ---- packed-INT4 shadow rows in the training forward (north-star kernel 2) ------------------------------
    def enable_shadow(self):
        """Keep a bit-packed INT4 copy of every table (rows_k x D/2 bytes, 1/8 of the fp32 arena) that the forward
        reads instead of the fp32 rows whenever that gives the same bits: bags of one index (all Criteo lookups)
        of a table whose scale is bit-identical to the one its shadow was encoded with.  fp32 rows stay the master
        copy; stale tables are re-encoded after the scan, updated rows after the update (csrc/shadow.cu)."""
        if self.embedding_bit != 4 or self.dim % 16:
            raise ValueError("the INT4 shadow needs embedding_bit == 4 and dim % 16 == 0")
        half, offs, off = self.dim // 2, [], 0
        for n in self.rows:
            offs.append(off)
            off += (n * half + 15) // 16 * 16                     # every table 16-byte aligned (128-bit staging loads)
        self._shadow_buf = torch.zeros(off + 16, dtype=torch.uint8, device=self.device)
        self.shadow = [self._shadow_buf[o:o + n * half].view(n, half) for o, n in zip(offs, self.rows)]
        self._shadow_ptrs = _lib.ptr_array(self.shadow)
        self.shadow_scale = torch.zeros(self.T, dtype=torch.float32, device=self.device)   # 0 != any scale: all stale
        self._shadow_flags = torch.zeros(self.T, dtype=torch.int32, device=self.device)

    def invalidate_shadow(self):
        """Tables were changed outside merge_apply / sgd_apply: every shadow is stale."""
        if self.shadow is not None:
            self.shadow_scale.zero_()

    def _shadow_refresh(self):
        rc = self.lib.dqrm_shadow_refresh(self.T, self._wptrs(), self._rows_arr, self.dim, self.scale.data_ptr(),
                                          self.inv_scale.data_ptr(), self._shadow_ptrs, self.shadow_scale.data_ptr(),
                                          self._shadow_flags.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "dqrm_shadow_refresh")

    def _shadow_update_rows(self, from_slots):
        if self.shadow is None:
            return
        if from_slots:
            rc = self.lib.dqrm_shadow_update_rows(self.T, self._wptrs(), self._rows_arr, self.dim, self.inv_scale.data_ptr(),
                                                  self._shadow_ptrs, self.gathered.data_ptr(), self.world, self.capacity,
                                                  self.grad_bit, None, None, _lib.stream_ptr())
        else:
            rc = self.lib.dqrm_shadow_update_rows(self.T, self._wptrs(), self._rows_arr, self.dim, self.inv_scale.data_ptr(),
                                                  self._shadow_ptrs, None, 0, self.capacity, self.grad_bit,
                                                  self.uniq_rows.data_ptr(), self.uniq_count.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "dqrm_shadow_update_rows")

    # ---- (a3) -----------------------------------------------------------
    def forward(self, indices, offsets, idx_begin, bags, full_precision=False, want_codes=True, out=None):
        self.finish_scan()
        lib, st = self.lib, _lib.stream_ptr()
        dev = self.device
        if out is None:
            out = torch.empty((self.T, bags, self.dim), dtype=torch.float32, device=dev)
        codes = None
        if want_codes and not full_precision:
            codes = torch.empty((self.T, bags, self.dim), dtype=torch.int8 if self.embedding_bit <= 8 else torch.int16,
                                device=dev)
        ib = _lib.i64_array(idx_begin)
        if self.shadow is not None and not full_precision:
            self._shadow_refresh()                   # re-encode the tables whose scale changed since their shadow was made
            rc = lib.dqrm_embbag_fwd_shadow(self.T, self._wptrs(), self._shadow_ptrs, self._rows_arr, self.dim,
                                            indices.data_ptr(), offsets.data_ptr(), ib, bags, self.scale.data_ptr(),
                                            self.inv_scale.data_ptr(), self.shadow_scale.data_ptr(), out.data_ptr(),
                                            out.stride(0), out.stride(1), _lib.ptr(codes), self.status.data_ptr(), st)
            _lib.check(rc, "dqrm_embbag_fwd_shadow")
            self.last = (indices, offsets, idx_begin, ib, bags, full_precision)
            self.codes = codes
            return out
        rc = lib.dqrm_embbag_fwd(self.T, self._wptrs(), self._rows_arr, self.dim, indices.data_ptr(),
                                 offsets.data_ptr(), ib, bags,
                                 None if full_precision else self.scale.data_ptr(),
                                 None if full_precision else self.inv_scale.data_ptr(), self.embedding_bit,
                                 out.data_ptr(), out.stride(0), out.stride(1), _lib.ptr(codes),
                                 self.status.data_ptr(), st)
        _lib.check(rc, "dqrm_embbag_fwd")
        self.last = (indices, offsets, idx_begin, ib, bags, full_precision)
        self.codes = codes
        return out

    # ---- packed INT4 export / serving forward --------------------------------
    def pack_int4(self):
        """Quantise every table with its current scale into bit-packed INT4 (two codes per byte).
        Returns (packed uint8 views [rows_k, D/2], scale [T] clone) -- the 8x smaller checkpoint / serving
        format (paper Table 3: 2.161 GB -> 0.270 GB at Kaggle shape)."""
        if self.embedding_bit != 4:
            raise ValueError("pack_int4 needs embedding_bit == 4")
        if not self.scale_valid:
            self.scan_scales()
        half = self.dim // 2
        total = sum(self.rows) * half
        buf = torch.empty(max(total, 8), dtype=torch.uint8, device=self.device)
        views, off = [], 0
        for n in self.rows:
            views.append(buf[off:off + n * half].view(n, half))
            off += n * half
        rc = self.lib.dqrm_table_pack_int4(self.T, self._wptrs(), self._rows_arr, self.dim, self.inv_scale.data_ptr(),
                                           _lib.ptr_array(views), _lib.stream_ptr())
        _lib.check(rc, "dqrm_table_pack_int4")
        self.packed, self.packed_buf, self.packed_scale = views, buf, self.scale.clone()
        return views, self.packed_scale

    def save_int4(self, path):
        """pack_int4() (if not done yet) and write the tables + scales as one file (int4_checkpoint.py)."""
        from . import int4_checkpoint
        if getattr(self, "packed", None) is None:
            self.pack_int4()
        return int4_checkpoint.save(path, self.packed, self.packed_scale, self.dim)

    def load_int4(self, path):
        """Read an INT4 table file onto this group's device for forward_int4(); shapes must match the group."""
        from . import int4_checkpoint
        packed, scale, rows, dim = int4_checkpoint.load(path, device=self.device)
        if rows != self.rows or dim != self.dim:
            raise ValueError(f"{path}: tables {rows} x {dim} do not match this group ({self.rows} x {self.dim})")
        self.packed, self.packed_scale = packed, scale
        return packed, scale

    def forward_int4(self, indices, offsets, idx_begin, bags, packed=None, scale=None, out=None):
        """Gather + dequantise + sum-pool straight from the packed INT4 tables."""
        packed = packed if packed is not None else self.packed
        scale = scale if scale is not None else self.packed_scale
        if out is None:
            out = torch.empty((self.T, bags, self.dim), dtype=torch.float32, device=self.device)
        rc = self.lib.dqrm_embbag_fwd_int4(self.T, _lib.ptr_array(packed), self._rows_arr, self.dim, indices.data_ptr(),
                                           offsets.data_ptr(), _lib.i64_array(idx_begin), bags, scale.data_ptr(),
                                           out.data_ptr(), out.stride(0), out.stride(1), self.status.data_ptr(),
                                           _lib.stream_ptr())
        _lib.check(rc, "dqrm_embbag_fwd_int4")
        return out

    # ---- (a4 bwd, a5, a7-1/2) ---------------------------------------------
    def _ensure_step_buffers(self, capacity, world):
        if self.capacity == capacity and self.world == world and self.uniq_rows is not None:
            return
        dev, T, D = self.device, self.T, self.dim
        self.capacity, self.world = capacity, world
        self.uniq_rows = torch.zeros((T, capacity), dtype=torch.int32, device=dev)
        self.uniq_count = torch.zeros(T, dtype=torch.int32, device=dev)
        self.grad_sums = torch.zeros((T, capacity, D), dtype=torch.float32, device=dev)
        self._alloc_exchange_buffers()
        self.slot = None   # view of this rank's slice of `gathered`, set by exchange()
        self.updated_rows = self.updated_count = self.qbar = None
        wsb = int(self.lib.dqrm_bwd_workspace_bytes(T, capacity, D))
        self._bwd_ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
        self._bwd_ws_bytes = wsb

    def _alloc_exchange_buffers(self):
        """gathered_scales [world, T(+pad)], gathered [world * slot_bytes] (+ the local scale slot): plain device
        buffers, or -- with world > 1 real ranks and DQRM_EXCHANGE=p2p -- views of this rank's peer arena, which
        the other ranks write into directly over NVLink (collective: all ranks get here in the same step)."""
        dev, T, D, world = self.device, self.T, self.dim, self.world
        sb = int(self.lib.dqrm_slot_bytes(T, self.capacity, D, self.grad_bit))
        self.slot_bytes = sb
        self.release()                       # an arena of another geometry: unmap + free it first (collective)
        a = None
        if world > 1 and _p2p.backend() == "p2p" and _dist_world() == world:
            try:
                a = _p2p.PeerArena({"emb_scale": T * 4, "emb_slot": sb, "absmax": T * 4}, world, self.dp_rank, dev)
            except _p2p.P2PUnavailable as e:         # raised on every rank together
                _p2p.fall_back_to_nccl(e)
        if a is not None:
            self.p2p = a
            self.gathered_scales = a.slots("emb_scale", torch.float32)
            self.grad_scale_local = a.my_slot("emb_scale", torch.float32, T)
            self.gathered = a.slots("emb_slot").view(-1)              # slot stride == slot_bytes (16-byte multiple)
            assert a.stride("emb_slot") == sb
            self._absmax_slots = a.slots("absmax", torch.float32)
            self._absmax_mine = a.my_slot("absmax", torch.float32, T)
        else:
            if self.grad_scale_local.numel() != T or self.grad_scale_local.data_ptr() % 16:     # was an arena view
                self.grad_scale_local = torch.zeros(T, dtype=torch.float32, device=dev)
            self.gathered_scales = torch.zeros((world, T), dtype=torch.float32, device=dev)
            self.gathered = torch.zeros(world * sb, dtype=torch.uint8, device=dev)

    def release(self):
        """Close this group's peer arena (after a barrier, so no rank still stores into it) and drop the views of
        it.  Collective when an arena exists: every rank gets here at the same point (geometry change, teardown)."""
        a, self.p2p = self.p2p, None
        if a is None:
            return
        import torch.distributed as dist
        live = dist.is_available() and dist.is_initialized()
        torch.cuda.synchronize()
        if live:
            dist.barrier()
        self.gathered_scales = self.gathered = self.slot = None
        self._absmax_slots = self._absmax_mine = None
        self.grad_scale_local = torch.zeros(self.T, dtype=torch.float32, device=self.device)
        a.close(dist.barrier if live else None)

    def _agree_capacity(self, cap):
        """All ranks must carve identical exchange slots: with real ranks and no caller-fixed capacity, take the
        MAX of the first step's per-table lookup counts over the ranks once (host sync, outside any capture) and
        keep it; a later step that needs more raises instead of silently re-creating arenas of another geometry."""
        import torch.distributed as dist
        if self.fixed_capacity is not None or not (dist.is_available() and dist.is_initialized()) \
                or dist.get_world_size() == 1 or self.dp_world != dist.get_world_size():
            return cap
        t = torch.tensor([cap], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        self.fixed_capacity = int(t.item())
        return self.fixed_capacity

    def _allreduce_absmax_to_scale(self, process_group):
        """Row-sharded scan: per-shard maxima in the absmax out-buffer -> MAX over ranks -> scale, 1/scale."""
        lib, st = self.lib, _lib.stream_ptr()
        if self.p2p is not None:
            self.p2p.allgather("absmax", self.status)
            rc = lib.dqrm_scale_from_absmax_gathered(self.T, self._absmax_slots.data_ptr(), self._absmax_slots.stride(0),
                                                     self.world, self.embedding_bit, self.absmax.data_ptr(),
                                                     self.scale.data_ptr(), self.inv_scale.data_ptr(), st)
            _lib.check(rc, "dqrm_scale_from_absmax_gathered")
            return
        import torch.distributed as dist
        dist.all_reduce(self.absmax, op=dist.ReduceOp.MAX, group=process_group)
        _lib.check(lib.dqrm_scale_from_absmax(self.T, self.absmax.data_ptr(), self.embedding_bit,
                                              self.scale.data_ptr(), self.inv_scale.data_ptr(), st), "dqrm_scale_from_absmax")

    def _absmax_out(self, sharded):
        """Where a sharded scan writes its per-shard maxima: this rank's slot of the absmax site, else self.absmax."""
        return self._absmax_mine if (sharded and self.p2p is not None) else self.absmax

    def backward(self, dout, world=1, last=None, ste_done=False):
        """De-duplicated row gradients of a forward (default: the last one) from dOut [T, B, D]-strided."""
        lib, st = self.lib, _lib.stream_ptr()
        indices, offsets, idx_begin, ib, bags, full_precision = last if last is not None else self.last
        cap = max(idx_begin[k + 1] - idx_begin[k] for k in range(self.T))
        cap = max(cap, 1)
        if world > 1 and self.fixed_capacity is None:
            cap = self._agree_capacity(cap)
        if self.fixed_capacity is not None:
            # all ranks must agree on the slot capacity; with one index per bag (Criteo) it is simply the
            # local batch, with ragged multi-hot bags the caller fixes a common upper bound
            if cap > self.fixed_capacity:
                raise _lib.DqrmLibraryError(f"{cap} lookups on one table exceed fixed_capacity={self.fixed_capacity}")
            cap = self.fixed_capacity
        self._ensure_step_buffers(cap, world)
        rc = lib.dqrm_embbag_bwd(self.T, self._rows_arr, self.dim, indices.data_ptr(), offsets.data_ptr(), ib, bags,
                                 dout.data_ptr(), dout.stride(0), dout.stride(1),
                                 None if (full_precision or ste_done) else self.scale.data_ptr(),
                                 self.capacity, self.uniq_rows.data_ptr(), self.uniq_count.data_ptr(),
                                 self.grad_sums.data_ptr(), min(self.grad_bit, 16), self.grad_scale_local.data_ptr(),
                                 self.status.data_ptr(), self._bwd_ws.data_ptr(), self._bwd_ws_bytes, st)
        _lib.check(rc, "dqrm_embbag_bwd")

    def backward_sgd(self, dout, lr, inv_world=1.0, momentum=None, eps=1e-10, last=None, ste_done=False):
        """(a5 + a10 in one call) de-duplicate the row gradients of a forward and apply the single-process row update
        in place -- SGD, or RW-Adagrad with `momentum` -- without the sums travelling through memory on the
        radix-sort path.  Same table bits as backward() followed by sgd_apply()."""
        lib, st = self.lib, _lib.stream_ptr()
        if getattr(self, "dp_world", 1) > 1:
            raise _lib.DqrmLibraryError("backward_sgd is the single-process path (the data-parallel step exchanges the "
                                        "de-duplicated gradients first: backward + exchange + merge_apply)")
        indices, offsets, idx_begin, ib, bags, full_precision = last if last is not None else self.last
        cap = max(max(idx_begin[k + 1] - idx_begin[k] for k in range(self.T)), 1)
        if self.fixed_capacity is not None:
            if cap > self.fixed_capacity:
                raise _lib.DqrmLibraryError(f"{cap} lookups on one table exceed fixed_capacity={self.fixed_capacity}")
            cap = self.fixed_capacity
        self._ensure_step_buffers(cap, 1)
        mom = _lib.ptr_array(momentum) if momentum is not None else None
        rc = lib.dqrm_embbag_bwd_sgd(self.T, self._wptrs(), self._rows_arr, self.dim, indices.data_ptr(),
                                     offsets.data_ptr(), ib, bags, dout.data_ptr(), dout.stride(0), dout.stride(1),
                                     None if (full_precision or ste_done) else self.scale.data_ptr(),
                                     self.capacity, self.uniq_rows.data_ptr(), self.uniq_count.data_ptr(),
                                     self.grad_sums.data_ptr(), float(lr), _lib.ptr(self.lr_dev), float(inv_world), mom,
                                     float(eps), self.status.data_ptr(), self._bwd_ws.data_ptr(), self._bwd_ws_bytes, st)
        _lib.check(rc, "dqrm_embbag_bwd_sgd")
        self._shadow_update_rows(from_slots=False)
        self._tracker_update(from_slots=False)

    def enable_fused_update(self, lr, momentum=None, eps=1e-10):
        """Single process, un-quantised row gradients (torch.optim.SGD on the sparse gradient, or RW-Adagrad with
        `momentum`): let the autograd backward of the group apply the row update itself (dqrm_embbag_bwd_sgd) --
        weight_update_parallel_comm / RWSAdagrad.step then find it done.  `lr` may be changed between steps
        (fused_update["lr"]); disable with group.fused_update = None."""
        self.fused_update = {"lr": float(lr), "momentum": momentum, "eps": float(eps)}

    def set_grad_bit(self, bits):
        """Change the gradient code width after a backward: re-size the slots, refresh the local scale."""
        self.grad_bit = int(bits)
        if self.uniq_rows is None:
            return
        self._alloc_exchange_buffers()
        if self.grad_bit != 32:
            _lib.check(self.lib.dqrm_grad_absmax_scale(self.T, self.dim, self.grad_sums.data_ptr(),
                                                       self.uniq_count.data_ptr(), self.capacity, self.grad_bit,
                                                       self.grad_scale_local.data_ptr(), _lib.stream_ptr()),
                       "dqrm_grad_absmax_scale")

    def topk(self, k):
        """(a8) keep the k highest-energy rows per table, then refresh the local scale."""
        st = _lib.stream_ptr()
        _lib.check(self.lib.dqrm_grad_topk(self.T, self.dim, self.grad_sums.data_ptr(), self.uniq_rows.data_ptr(),
                                           self.uniq_count.data_ptr(), self.capacity, int(k), st), "dqrm_grad_topk")
        _lib.check(self.lib.dqrm_grad_absmax_scale(self.T, self.dim, self.grad_sums.data_ptr(),
                                                   self.uniq_count.data_ptr(), self.capacity, self.grad_bit,
                                                   self.grad_scale_local.data_ptr(), st), "dqrm_grad_absmax_scale")

    # ---- (a7-3..5) --------------------------------------------------------
    def exchange(self, world=1, rank=0, process_group=None):
        """Scale all-gather -> pack -> slot all-gather (the only two collectives of the
        embedding exchange, replacing 52 Gloo calls).  Composed of the three phases below so that
        tests can emulate several ranks on one GPU by copying between replicas' buffers."""
        assert world == self.world, "backward() and exchange() must agree on world size"
        if self.p2p is not None:
            # one-kernel all-gathers over NVLink peer memory (csrc/p2p.cu): the backward already wrote the local
            # scales into this rank's slot, pack() writes the codes into its slot of the gathered buffer
            self.p2p.allgather("emb_scale", self.status)
            self.pack(rank)
            self.p2p.allgather("emb_slot", self.status)
            return
        self.stage_scale(rank)
        if world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.gathered_scales.view(-1), self.grad_scale_local, group=process_group)
        self.pack(rank)
        if world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.gathered, self.slot, group=process_group)

    def start_exchange(self):
        """Run exchange() on a side stream right after the de-duplicating backward, so the two all-gathers and
        the pack overlap with the rest of the backward pass (the bottom MLP).  finish_exchange() joins."""
        main = torch.cuda.current_stream()
        if self.xchg_stream is None:
            self.xchg_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.xchg_stream.wait_stream(main)
        with torch.cuda.stream(self.xchg_stream):
            self.exchange(world=self.world, rank=self.dp_rank)
        self.exchange_started = True

    def backward_async(self, dout, last=None, ste_done=False):
        """backward() -- and the exchange when it may start eagerly (always at world 1, where it is just the pack) --
        on the side stream, forked from the current one.  `dout` is kept alive until finish_exchange() joins: inside
        a graph capture its block must not be handed to an allocation of the parallel branch."""
        main = torch.cuda.current_stream()
        if self.xchg_stream is None:
            self.xchg_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.xchg_stream.wait_stream(main)
        with torch.cuda.stream(self.xchg_stream):
            self.backward(dout, world=self.dp_world, last=last, ste_done=ste_done)
            if self.dp_world == 1 or self.eager_exchange:
                self.exchange(world=self.world, rank=self.dp_rank)
                self.exchange_started = True
                if self.eager_apply and self.lr_dev is not None:
                    # ... and the row update itself (learning rate read from lr_dev): nothing else in the step touches
                    # the tables any more, so it need not wait for the dense exchange on the main stream.
                    # weight_update_parallel_comm() sees applied_eagerly and skips this group.
                    self.merge_apply(0.0)
                    self.applied_eagerly = True
        self._bwd_forked = dout

    def finish_exchange(self):
        """Join the side stream (backward_async / start_exchange) into the current one.  True if an exchange was
        started there and is now complete."""
        if not self.exchange_started and self._bwd_forked is None:
            return False
        torch.cuda.current_stream().wait_stream(self.xchg_stream)
        started, self.exchange_started, self._bwd_forked = self.exchange_started, False, None
        return started

    def stage_scale(self, rank=0):
        """Phase 1: this rank's 26 local gradient scales into row `rank` of gathered_scales."""
        if self.p2p is None:
            self.gathered_scales[rank, :self.T].copy_(self.grad_scale_local)

    def pack(self, rank=0):
        """Phase 2 (after the scale all-gather): quantise with the rank-mean scale into slot `rank`."""
        sb = self.slot_bytes
        self.slot = self.gathered[rank * sb:(rank + 1) * sb]
        rc = self.lib.dqrm_grad_pack(self.T, self.dim, self.grad_sums.data_ptr(), self.uniq_rows.data_ptr(),
                                     self.uniq_count.data_ptr(), self.capacity, self.gathered_scales.data_ptr(),
                                     self.gathered_scales.stride(0), self.world, self.grad_bit, self.slot.data_ptr(), self.grad_scale_mean.data_ptr(),
                                     _lib.stream_ptr())
        _lib.check(rc, "dqrm_grad_pack")

    def merge_apply(self, lr):
        """(a7-5, a9): W[row] += (-lr) * ((sum_r q_r * 1/N) * s_bar) on the union of rows."""
        lib, st = self.lib, _lib.stream_ptr()
        if self.keep_debug and self.updated_rows is None:
            n = self.world * self.capacity
            self.updated_rows = torch.zeros((self.T, n), dtype=torch.int32, device=self.device)
            self.updated_count = torch.zeros(self.T, dtype=torch.int32, device=self.device)
            self.qbar = torch.zeros((self.T, n, self.dim), dtype=torch.float32, device=self.device)
        dbg = self.keep_debug
        rc = lib.dqrm_grad_merge_apply(self.T, self._wptrs(), self._rows_arr, self.dim, self.gathered.data_ptr(),
                                       self.world, self.capacity, self.grad_bit, self.grad_scale_mean.data_ptr(),
                                       float(lr), _lib.ptr(self.lr_dev), _lib.ptr(self.updated_rows) if dbg else None,
                                       _lib.ptr(self.updated_count) if dbg else None,
                                       _lib.ptr(self.qbar) if dbg else None, self.status.data_ptr(), st)
        _lib.check(rc, "dqrm_grad_merge_apply")
        self._shadow_update_rows(from_slots=True)
        self._tracker_update(from_slots=True)

    def sgd_apply(self, lr, inv_world=1.0, momentum=None, eps=1e-10):
        """(a10) un-quantised row update from the local de-duplicated sums (optionally RW-Adagrad)."""
        st = _lib.stream_ptr()
        mom = _lib.ptr_array(momentum) if momentum is not None else None
        rc = self.lib.dqrm_sgd_rows(self.T, self._wptrs(), self._rows_arr, self.dim, self.uniq_rows.data_ptr(),
                                    self.uniq_count.data_ptr(), self.grad_sums.data_ptr(), self.capacity, float(lr),
                                    _lib.ptr(self.lr_dev), float(inv_world), mom, float(eps), st)
        _lib.check(rc, "dqrm_sgd_rows")
        self._shadow_update_rows(from_slots=False)
        self._tracker_update(from_slots=False)

    # ---- views for tests / API compatibility (these synchronise) ----------
    def slot_views(self, rank=0):
        """(count[T], rows[T,cap], codes[T,cap,D]) views of one gathered slot."""
        ro, co = C.c_size_t(), C.c_size_t()
        self.lib.dqrm_slot_layout(self.T, self.capacity, self.dim, self.grad_bit, C.byref(ro), C.byref(co))
        sb = self.slot_bytes
        s = self.gathered[rank * sb:(rank + 1) * sb]
        cnt = s[:self.T * 4].view(torch.int32)
        rows = s[ro.value:ro.value + self.T * self.capacity * 4].view(torch.int32).view(self.T, self.capacity)
        cb, dt = (4, torch.float32) if self.grad_bit == 32 else ((1, torch.int8) if self.grad_bit <= 8 else (2, torch.int16))
        codes = s[co.value:co.value + self.T * self.capacity * self.dim * cb]
        codes = codes.view(dt).view(self.T, self.capacity, self.dim)
        return cnt, rows, codes

    def sparse_grad(self, t):
        """Coalesced sparse COO gradient of table t (what the reference leaves in
        ``embedding_bag.weight.grad`` after ``.coalesce()``, sgd...parallel_comm.py:859)."""
        u = int(self.uniq_count[t].item())
        return torch.sparse_coo_tensor(self.uniq_rows[t, :u].long()[None], self.grad_sums[t, :u].clone(),
                                       size=(self.rows[t], self.dim))
