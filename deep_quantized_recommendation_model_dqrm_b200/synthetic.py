"""Seeded synthetic inputs with the shapes and distributions of the reference's
data path.  There is no dataset in the build/bench
environment, so every test and benchmark draws its inputs from here.

Reference semantics mirrored:
  * table init  U(-sqrt(1/N), +sqrt(1/N))  quantization_supp/quant_modules_not_quantize_grad.py:273-275
  * MLP init    W ~ N(0, sqrt(2/(m+n))), b ~ N(0, sqrt(1/m))   dlrm_s_pytorch_comm_grad.py:288-296
  * Criteo batch (X[B,13] f32, lS_o[T,B] i64 = arange(B), lS_i[T,B] i64, T[B,1] f32)
                dlrm_data_pytorch.py:328-345
  * random multi-hot batch (P ~ round(max(1, r*min(N, P_max))), unique sorted indices per bag)
                dlrm_data_pytorch.py:1099-1157
"""
from __future__ import annotations

import numpy as np
import torch

# python_profiling_script/finding_kaggle_compression_ratio.py:2 (also bash_scripts/Kaggle/emb_bit_4.txt:16-41)
KAGGLE_ROWS = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27,
               14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]
# python_profiling_script/finding_kaggle_compression_ratio.py:5 (Terabyte, --max-ind-range=10000000)
TERABYTE_ROWS_10M = [9980200, 26095, 17224, 7383, 20152, 3, 7112, 1435, 62, 9756762, 1332128, 314263, 10,
                     2208, 11168, 122, 4, 971, 14, 9994101, 7267918, 9946670, 415284, 12422, 102, 36]
# public MLPerf-DLRM counts at --max-ind-range=40000000 (not in the reference tree)
TERABYTE_ROWS_40M = [39884406, 39043, 17289, 7420, 20263, 3, 7120, 1543, 63, 38532951, 2953546, 403346, 10,
                     2208, 11938, 155, 4, 976, 14, 39979771, 25641295, 39664984, 585935, 12972, 108, 36]

KAGGLE = dict(rows=KAGGLE_ROWS, dim=16, ln_bot=[13, 512, 256, 64, 16], ln_top_hidden=[512, 256, 1])
TERABYTE = dict(rows=TERABYTE_ROWS_40M, dim=64, ln_bot=[13, 512, 256, 64], ln_top_hidden=[512, 512, 256, 1])
RANDOM_SMALL = dict(rows=[10000] * 8, dim=16, ln_bot=[13, 512, 256, 64, 16], ln_top_hidden=[512, 256, 1])


def top_mlp_sizes(num_tables: int, dim: int, hidden) -> list:
    """ln_top = [num_int] + hidden, num_int = (num_fea choose 2) + dim for the
    dot interaction without self-interaction (dlrm_s_pytorch_comm_grad.py:1530-1548)."""
    num_fea = num_tables + 1
    return [num_fea * (num_fea - 1) // 2 + dim] + list(hidden)


def table_weights_numpy(rows: int, dim: int, rng: np.random.RandomState) -> np.ndarray:
    b = np.sqrt(1 / rows)
    return rng.uniform(low=-b, high=b, size=(rows, dim)).astype(np.float32)


def table_weights_(out: torch.Tensor, table_id: int, seed: int = 1234) -> torch.Tensor:
    """Fill ``out`` ([rows, dim] fp32, any device) in place with the reference's
    uniform init, generated on ``out``'s device so 10M-row tables never touch
    the host.  Deterministic per (seed, table_id, device type)."""
    rows = out.shape[0]
    g = torch.Generator(device=out.device)
    g.manual_seed(seed + table_id)
    b = float(np.sqrt(1 / rows))
    out.uniform_(-b, b, generator=g)
    return out


def mlp_params(ln, rng: np.random.RandomState):
    params = []
    for i in range(len(ln) - 1):
        n, m = int(ln[i]), int(ln[i + 1])
        W = rng.normal(0.0, np.sqrt(2 / (m + n)), size=(m, n)).astype(np.float32)
        b = rng.normal(0.0, np.sqrt(1 / m), size=m).astype(np.float32)
        params.append((W, b))
    return params


def criteo_batch(rows, batch: int, seed: int, zipf: float | None = None, dense: int = 13):
    """One Criteo-shaped batch: every bag has exactly one index."""
    rng = np.random.RandomState(seed)
    X = np.log(1.0 + rng.randint(0, 100, size=(batch, dense))).astype(np.float32)
    idx = np.empty((len(rows), batch), dtype=np.int64)
    for k, n in enumerate(rows):
        if zipf is None:
            idx[k] = rng.randint(0, n, size=batch)
        else:
            idx[k] = np.minimum(rng.zipf(zipf, size=batch) - 1, n - 1)
    off = np.tile(np.arange(batch, dtype=np.int64), (len(rows), 1))
    T = np.round(rng.rand(batch, 1).astype(np.float32)).astype(np.float32)
    return torch.from_numpy(X), torch.from_numpy(off), torch.from_numpy(idx), torch.from_numpy(T)


def criteo_batch_nodup(rows, world: int, per_rank: int, seed: int, dense: int = 13):
    """Criteo-shaped global batch of world*per_rank samples whose per-rank shards (contiguous, get_my_slice)
    hold no duplicate index inside a table, while ranks may share rows: the case in which the reference's
    coalesce() folds nothing, so its gradient scales and INT8 codes do not depend on a fold order."""
    rng = np.random.RandomState(seed)
    batch = world * per_rank
    X = np.log(1.0 + rng.randint(0, 100, size=(batch, dense))).astype(np.float32)
    idx = np.empty((len(rows), batch), dtype=np.int64)
    for k, n in enumerate(rows):
        for r in range(world):
            idx[k, r * per_rank:(r + 1) * per_rank] = rng.choice(n, size=per_rank, replace=False)
    off = np.tile(np.arange(batch, dtype=np.int64), (len(rows), 1))
    T = np.round(rng.rand(batch, 1).astype(np.float32)).astype(np.float32)
    return torch.from_numpy(X), torch.from_numpy(off), torch.from_numpy(idx), torch.from_numpy(T)


# isolated exchange cases (tests/golden/xchg*.npz): gradients are INJECTED after the backward, so everything
# downstream (scales, INT8 codes, merged rows, updated weights) is order-independent and bit-exact by contract
XCHG = dict(rows=[600, 40, 300, 100], dim=16, per_rank=16, layers=[(13, 32), (32, 16)])


def injected_grads(cfg, world: int, rank: int, step: int, seed: int = 700):
    """Per-rank gradients for one exchange step: per table (unique unsorted rows [U], values [U, D]) and per
    linear layer (dW [out, in], db [out]); magnitudes differ per rank / channel so the local scales differ."""
    rng = np.random.RandomState(seed + 1000 * step + 10 * rank + world)
    U, D = cfg["per_rank"], cfg["dim"]
    emb = []
    for n in cfg["rows"]:
        idx = rng.choice(n, size=U, replace=False).astype(np.int64)
        vals = (rng.randn(U, D) * 10.0 ** rng.uniform(-4, -1)).astype(np.float32)
        emb.append((idx, vals))
    mlp = []
    for li, (n_in, n_out) in enumerate(cfg["layers"]):
        gw = (rng.randn(n_out, n_in) * 10.0 ** rng.uniform(-4, -1, size=(n_out, 1))).astype(np.float32)
        gb = (rng.randn(n_out) * 10.0 ** rng.uniform(-4, -1)).astype(np.float32)
        if li == 0:
            gw[0, :] = 0.0                       # an all-zero channel: scale = 1e-8 / 127 (quant_utils.py:213-214)
        mlp.append((gw, gb))
    return emb, mlp


def xchg_weights(cfg, seed: int = 701):
    rng = np.random.RandomState(seed)
    emb = [table_weights_numpy(n, cfg["dim"], rng) for n in cfg["rows"]]
    mlp = [mlp_params([a, b], rng)[0] for a, b in cfg["layers"]]
    return emb, mlp


def random_bags(rows: int, batch: int, p_max: int, rng: np.random.RandomState, fixed: bool = False):
    """Indices/offsets of one table, RandomDataset style (unique sorted per bag)."""
    offsets, indices, offset = [], [], 0
    for _ in range(batch):
        if fixed:
            size = p_max
        else:
            size = int(np.round(max(1.0, rng.random_sample() * min(rows, p_max))))
        grp = np.unique(np.round(rng.random_sample(size) * (rows - 1)).astype(np.int64))
        offsets.append(offset)
        indices.append(grp)
        offset += grp.size
    idx = np.concatenate(indices) if indices else np.zeros(0, dtype=np.int64)
    return torch.from_numpy(idx), torch.tensor(offsets, dtype=torch.int64)


def random_batch(rows, batch: int, p_max: int, seed: int, fixed: bool = False, dense: int = 13):
    """One random-data batch: lists of per-table 1-D index / offset tensors."""
    rng = np.random.RandomState(seed)
    X = torch.from_numpy(rng.rand(batch, dense).astype(np.float32))
    lS_i, lS_o = [], []
    for n in rows:
        i, o = random_bags(int(n), batch, p_max, rng, fixed)
        lS_i.append(i)
        lS_o.append(o)
    T = torch.from_numpy(np.round(rng.rand(batch, 1).astype(np.float32)).astype(np.float32))
    return X, lS_o, lS_i, T
