"""Drop-in for extend_distributed.py, thin by design: in the reference's DP-quantised drivers only the
module globals, ``get_my_slice`` and the process-group bring-up matter;
the butterfly ``alltoall`` belongs to the hybrid model-parallel drivers and is out of scope.

B200 mapping: one process per GPU, ``torch.distributed`` with the NCCL backend over NVLink 5 /
NVSwitch (the reference hard-codes Gloo because it needs sparse all-reduce,
dlrm_s_pytorch_comm_grad.py:1413-1418; the packed-slot all-gather removes that need).  ``gloo`` is
accepted for CPU-side tests of the host logic.
"""
from __future__ import annotations

import builtins
import os

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: F401  (re-export, extend_distributed.py:14)

my_rank = -1
my_size = -1
my_local_rank = -1
my_local_size = -1
alltoall_supported = False


def env2int(env_list, default=-1):
    for e in env_list:
        val = int(os.environ.get(e, -1))
        if val >= 0:
            return val
    return default


def get_my_slice(n):
    """extend_distributed.py:47-51."""
    k, m = divmod(n, my_size)
    return slice(my_rank * k + min(my_rank, m), (my_rank + 1) * k + min(my_rank + 1, m), 1)


def get_split_lengths(n):
    """extend_distributed.py:54-62."""
    k, m = divmod(n, my_size)
    if m == 0:
        splits = None
        my_len = k
    else:
        splits = [(k + 1) if i < m else k for i in range(my_size)]
        my_len = splits[my_rank]
    return (my_len, splits)


def init_distributed(rank=-1, local_rank=-1, size=-1, use_gpu=False, backend=""):
    """Bring up the process group (extend_distributed.py:65-194).  Rank / size come from the arguments or
    the torchrun / MPI environment; backend defaults to nccl on GPUs."""
    global my_rank, my_size, my_local_rank, my_local_size
    if rank == -1:
        rank = env2int(["PMI_RANK", "OMPI_COMM_WORLD_RANK", "MV2_COMM_WORLD_RANK", "RANK"], 0)
    if size == -1:
        size = env2int(["PMI_SIZE", "OMPI_COMM_WORLD_SIZE", "MV2_COMM_WORLD_SIZE", "WORLD_SIZE"], 1)
    if local_rank == -1:
        local_rank = env2int(["MPI_LOCALRANKID", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "LOCAL_RANK"], 0)
    if not backend:
        backend = "nccl" if use_gpu else "gloo"
    if backend not in ("nccl", "gloo"):
        raise ValueError(f"backend {backend!r}: only nccl (GPUs) and gloo (host-logic tests) are supported")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if use_gpu:
        torch.cuda.set_device(local_rank)
    if size > 1 and not dist.is_initialized():
        kw = {}
        if use_gpu and backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, rank=rank, world_size=size, **kw)
    my_rank, my_size, my_local_rank = rank, size, local_rank
    my_local_size = env2int(["MPI_LOCALNRANKS", "OMPI_COMM_WORLD_LOCAL_SIZE", "MV2_COMM_WORLD_LOCAL_SIZE", "LOCAL_WORLD_SIZE"], 1)
    print("Running on %d ranks using %s backend" % (my_size, backend))


def alltoall(inputs, per_rank_table_splits):
    raise NotImplementedError("alltoall belongs to the hybrid model-parallel drivers (extend_distributed.py:545-582), "
                              "which are outside the data-parallel hot path")


def all_gather(input, lengths, dim=0):
    """extend_distributed.py:585-588 (forward only; equal lengths)."""
    if my_size <= 1:
        return input
    if lengths and len(set(lengths)) != 1:
        raise NotImplementedError("ragged all_gather is only used by the hybrid drivers")
    out = [torch.empty_like(input) for _ in range(my_size)]
    dist.all_gather(out, input)
    return torch.cat(out, dim=dim)


def barrier():
    if my_size > 1:
        dist.barrier()


orig_print = builtins.print


def rank0_print(*args, **kwargs):
    """Rank-0-only print (the reference installs this over builtins.print as an import side effect,
    extend_distributed.py:597-605; here it is opt-in via install_rank0_print())."""
    if my_rank <= 0 or kwargs.pop("print_all", False):
        orig_print(*args, **kwargs)


def install_rank0_print():
    builtins.print = rank0_print


def print_all(*args, **kwargs):
    orig_print(*args, **kwargs)
