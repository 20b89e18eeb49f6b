"""On-disk format of the bit-packed INT4 embedding tables ("packed INT4 checkpoint /
inference export"; the paper's Table 3 model size, 2.16 GB -> 0.27 GB at Kaggle shape).

The reference has no such format: its QAT checkpoints keep fp32 tables (torch.save of the state dict,
dlrm_s_pytorch_comm_grad.py:1964-1976) and its PTQ path re-packs with ATen's per-ROW scale/bias 4-bit format
(dlrm_s_pytorch.py:428-440).  DQRM's quantiser is per-TABLE symmetric, so a table is just its codes plus one fp32
scale; this file stores exactly what ``EmbeddingTableGroup.pack_int4()`` produces and what
``dqrm_embbag_fwd_int4`` consumes (element d of a row in byte d/2, low nibble for even d):

    offset 0    magic  b"DQRMINT4"
           8    u32 version (1) | u32 num_tables T | u32 dim D | u32 bits (4)
          24    u64 rows[T]
    24+8T       f32 scale[T]
                zero padding to a multiple of 256
                table 0 codes [rows_0, D/2] bytes | pad to 256 | table 1 ... (each table 256-byte aligned)

Everything is little-endian.  ``load`` memory-maps the file, so a serving process can hand the mapped tables to the
GPU without an intermediate copy in host RAM.
"""
from __future__ import annotations

import struct

import numpy as np
import torch

MAGIC = b"DQRMINT4"
VERSION = 1
_ALIGN = 256


def _pad(n):
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


def layout(rows, dim):
    """(header_bytes, [table offsets], total_bytes) of a file holding tables with these row counts."""
    T = len(rows)
    off = _pad(24 + 8 * T + 4 * T)
    offs = []
    for n in rows:
        offs.append(off)
        off += _pad(int(n) * (dim // 2))
    return offs[0] if offs else off, offs, off


def save(path, packed, scale, dim):
    """packed: list of uint8 [rows_k, dim/2] tensors or arrays (device or host); scale: fp32 [T]."""
    if dim % 2:
        raise ValueError("dim must be even (two codes per byte)")
    tabs = [t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t) for t in packed]
    sc = (scale.detach().cpu().numpy() if torch.is_tensor(scale) else np.asarray(scale)).astype("<f4").reshape(-1)
    rows = [int(t.shape[0]) for t in tabs]
    if len(sc) != len(tabs):
        raise ValueError(f"{len(tabs)} tables but {len(sc)} scales")
    for k, t in enumerate(tabs):
        if t.dtype != np.uint8 or t.ndim != 2 or t.shape[1] != dim // 2:
            raise ValueError(f"table {k}: expected uint8 [rows, {dim // 2}], got {t.dtype} {t.shape}")
    _, offs, total = layout(rows, dim)
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<4I", VERSION, len(tabs), dim, 4))
        f.write(np.asarray(rows, dtype="<u8").tobytes())
        f.write(sc.tobytes())
        for off, t in zip(offs, tabs):
            f.write(b"\0" * (off - f.tell()))
            f.write(np.ascontiguousarray(t).tobytes())
        f.write(b"\0" * (total - f.tell()))
    return total


def load(path, device=None):
    """-> (packed: list of uint8 [rows_k, dim/2] tensors, scale fp32 [T], rows list, dim).  With device=None the
    tensors are views of a read-only memory map; otherwise they are copied to `device`."""
    with open(path, "rb") as f:
        head = f.read(24)
        if len(head) < 24 or head[:8] != MAGIC:
            raise ValueError(f"{path}: not a DQRM INT4 table file")
        version, T, dim, bits = struct.unpack("<4I", head[8:24])
        if version != VERSION or bits != 4 or dim % 2:
            raise ValueError(f"{path}: unsupported version {version} / bits {bits} / dim {dim}")
        rows = np.frombuffer(f.read(8 * T), dtype="<u8").astype(np.int64).tolist()
        scale = np.frombuffer(f.read(4 * T), dtype="<f4").copy()
    _, offs, total = layout(rows, dim)
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    if mm.shape[0] != total:
        raise ValueError(f"{path}: {mm.shape[0]} bytes, expected {total} for the table sizes in its header")
    packed = []
    for off, n in zip(offs, rows):
        a = mm[off:off + n * (dim // 2)].reshape(n, dim // 2)
        # np.array(a) copies out of the map (needed before a device copy from pageable memory anyway); the mapped
        # view itself is read-only, which torch only warns about -- callers must not write to it
        t = torch.from_numpy(np.array(a)) if device is not None else torch.from_numpy(np.asarray(a))
        packed.append(t.to(device) if device is not None else t)
    sc = torch.from_numpy(scale)
    return packed, (sc.to(device) if device is not None else sc), rows, dim
