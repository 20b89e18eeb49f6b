"""Peer arenas for the one-kernel NVLink exchanges (csrc/p2p.cu, include/dqrm_b200.h "exchange over NVLink peer
memory").  One cudaMalloc'd arena per rank, mapped by every rank of the box through CUDA IPC; the 64-byte handles
travel through torch.distributed (plumbing).  Sites are carved at identical offsets on every rank.

``PeerArena.local_group`` builds W arenas inside ONE process (plain pointers, no IPC) so that the kernels' protocol
can be tested on a single GPU with the W "ranks" on W streams.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib


def backend() -> str:
    """'p2p' (default) or 'nccl' -- how the step's exchanges travel when world > 1 (env DQRM_EXCHANGE)."""
    b = os.environ.get("DQRM_EXCHANGE", "p2p").lower()
    if b not in ("p2p", "nccl"):
        raise ValueError(f"DQRM_EXCHANGE={b!r}: expected 'p2p' or 'nccl'")
    return b


class ExchangeTimeout(RuntimeError):
    """DQRM_STATUS_P2P_TIMEOUT was raised by an exchange kernel: fatal (see csrc/p2p.cu)."""


class P2PUnavailable(RuntimeError):
    """Raised on EVERY rank when any rank could not allocate / export / map a peer arena."""


def fall_back_to_nccl(err):
    """The NVLink transport is unavailable on this box: say so loudly and use the NCCL all-gathers (same slots,
    same consumer kernels, bit-identical results).  Called on every rank after a P2PUnavailable."""
    import sys
    os.environ["DQRM_EXCHANGE"] = "nccl"
    print(f"[dqrm-b200] WARNING: NVLink peer-memory exchange unavailable ({err}); using NCCL all-gathers instead",
          file=sys.stderr, flush=True)


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so torch can view memory this library allocated."""

    def __init__(self, ptr, nbytes, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self._owner = owner


class PeerArena:
    def __init__(self, sites, world, rank, device, process_group=None, _local=None):
        """sites: ordered {name: slot_bytes}.  Collective: every rank of `process_group` must construct the same
        arena at the same point of its program."""
        lib = _lib.load()
        self.lib, self.world, self.rank, self.device = lib, int(world), int(rank), torch.device(device)
        self.sites, off = {}, 0
        for name, slot_bytes in sites.items():
            fo, do, st = C.c_size_t(), C.c_size_t(), C.c_size_t()
            _lib.check(lib.dqrm_p2p_site_layout(self.world, int(slot_bytes), C.byref(fo), C.byref(do), C.byref(st)),
                       "dqrm_p2p_site_layout")
            self.sites[name] = dict(off=off, slot_bytes=int(slot_bytes), data_off=off + do.value, stride=st.value)
            off += (int(lib.dqrm_p2p_site_bytes(self.world, int(slot_bytes))) + 255) // 256 * 256
        self.bytes = max(off, 256)
        self._opened = []
        if _local is not None:                       # single-process emulation: bases handed in by local_group
            self.base = _local[self.rank]
            bases = list(_local)
        else:
            import torch.distributed as dist
            with torch.cuda.device(self.device):
                # every step below is attempted on every rank and the outcomes are agreed on with a MIN all-reduce,
                # so a failure on one rank raises P2PUnavailable on all of them (no rank is left in a collective)
                ptr, handle, err = C.c_void_p(), (C.c_ubyte * 64)(), ""
                try:
                    if os.environ.get("DQRM_P2P_FAIL_RANK") == str(self.rank):      # failure injection (tests)
                        raise _lib.DqrmLibraryError("injected failure")
                    _lib.check(lib.dqrm_p2p_alloc(self.bytes, C.byref(ptr), handle), "dqrm_p2p_alloc")
                except _lib.DqrmLibraryError as e:
                    err = str(e)
                self.base = ptr.value
                mine = torch.tensor(list(handle) + [0 if err else 1], dtype=torch.uint8, device=self.device)
                allh = torch.empty(self.world * 65, dtype=torch.uint8, device=self.device)
                dist.all_gather_into_tensor(allh, mine, group=process_group)
                allh = allh.cpu().view(self.world, 65)
                ok = bool(allh[:, 64].min().item())
                bases = []
                for r in range(self.world):
                    if r == self.rank or not ok:
                        bases.append(self.base)
                        continue
                    h = (C.c_ubyte * 64)(*allh[r, :64].tolist())
                    p = C.c_void_p()
                    rc = lib.dqrm_p2p_open(h, C.byref(p))
                    if rc != 0:
                        err, ok = f"dqrm_p2p_open(rank {r}): {_lib.last_error()}", False
                        bases.append(None)
                        continue
                    bases.append(p.value)
                    self._opened.append(p.value)
                flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=process_group)
                torch.cuda.synchronize()
                if int(flag.item()) == 0:
                    for q in self._opened:
                        lib.dqrm_p2p_close(q)
                    if self.base:
                        lib.dqrm_p2p_free(self.base)
                    raise P2PUnavailable(err or "a peer rank could not set up its arena")
                dist.barrier(group=process_group)    # nobody signals into an arena that is not mapped and zeroed yet
        self.bases = bases
        self.ptrs = (C.c_void_p * self.world)(*bases)
        self._raw = torch.as_tensor(_RawCuda(self.base, self.bytes, self), device=self.device)

    @classmethod
    def local_group(cls, sites, world, device="cuda"):
        """W arenas in this process (tests): returns [PeerArena for rank 0..W-1]."""
        lib = _lib.load()
        total = sum((int(lib.dqrm_p2p_site_bytes(world, int(b))) + 255) // 256 * 256 for b in sites.values())
        bufs = [torch.zeros(max(total, 256), dtype=torch.uint8, device=device) for _ in range(world)]
        arenas = [cls(sites, world, r, device, _local=[b.data_ptr() for b in bufs]) for r in range(world)]
        for a in arenas:
            a._keep = bufs                           # the arenas alias these tensors
        return arenas

    def close(self, barrier=None):
        """Unmap the peers' arenas and free the local one (collective in spirit: call on every rank, after a
        barrier, when no exchange is in flight).  `barrier` (a callable) runs between the unmap and the free, so no
        exporter frees memory a peer still has mapped.  Views handed out by slots()/my_slot() must not be used
        afterwards."""
        for q in self._opened:
            self.lib.dqrm_p2p_close(q)
        self._opened = []
        if barrier is not None:
            barrier()
        if getattr(self, "_keep", None) is None and self.base:
            self.lib.dqrm_p2p_free(self.base)
        self.base = None

    # ---- views of the LOCAL arena ---------------------------------------------------------------------------
    def slots(self, name, dtype=torch.uint8):
        """[world, stride/itemsize] view of a site's slots (row r = rank r's contribution after allgather)."""
        s = self.sites[name]
        flat = self._raw[s["data_off"]:s["data_off"] + self.world * s["stride"]]
        return flat.view(dtype).view(self.world, -1)

    def my_slot(self, name, dtype=torch.uint8, numel=None):
        v = self.slots(name, dtype)[self.rank]
        return v if numel is None else v[:numel]

    def stride(self, name, itemsize=1):
        return self.sites[name]["stride"] // itemsize

    def allgather(self, name, status):
        s = self.sites[name]
        rc = self.lib.dqrm_p2p_allgather(self.ptrs, self.world, self.rank, s["off"], s["slot_bytes"],
                                         status.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, f"dqrm_p2p_allgather({name})")
