"""Binary Criteo (Terabyte / MLPerf) batch reader -- the input format of BASELINE.json configs[3].

Reference: data_loader_terabyte.py  CriteoBinDataset :197-241, _transform_features :68-87, numpy_to_binary :243-280.
Record = 40 little-endian int32: label | 13 dense counts | 26 categorical ids; one dataset item = one whole batch
``(X [B,13] = log(dense + 1), lS_o [26,B] = arange(B) per table, lS_i [26,B] = ids (% max_ind_range), T [B,1])``.

The file is memory-mapped (the reference seeks + reads), and ``read_packed`` writes a batch straight into the pinned
single-copy staging layout of ``graph_step.GraphedTrainStep`` (one H2D per step).  Pinned against the reference's own
reader and writer by tests/test_data_formats.py (goldens: oracle/make_golden_data.py).
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
from torch.utils.data import Dataset

TAR_FEA, DEN_FEA, SPA_FEA = 1, 13, 26
TOT_FEA = TAR_FEA + DEN_FEA + SPA_FEA


def _transform_features(x_int_batch, x_cat_batch, y_batch, max_ind_range, flag_input_torch_tensor=False):
    """:68-87.  Accepts numpy arrays or torch tensors (int32 views of the records)."""
    as_t = (lambda a: a.clone().detach()) if flag_input_torch_tensor else (lambda a: torch.as_tensor(np.asarray(a)))
    if max_ind_range > 0:
        x_cat_batch = x_cat_batch % max_ind_range
    X = torch.log(as_t(x_int_batch).type(torch.float) + 1)
    cat = as_t(x_cat_batch).type(torch.long)
    T = as_t(y_batch).type(torch.float32).view(-1, 1)
    B, F = cat.shape
    lS_o = torch.arange(B).reshape(1, -1).repeat(F, 1)
    return X, lS_o, cat.t(), T


class CriteoBinDataset(Dataset):
    """Batches of a ``*_data.bin`` file written by numpy_to_binary; the last batch may be short."""

    def __init__(self, data_file, counts_file, batch_size=1, max_ind_range=-1, bytes_per_feature=4):
        if bytes_per_feature != 4:
            raise ValueError("records are int32 (numpy_to_binary): bytes_per_feature must be 4")
        self.tar_fea, self.den_fea, self.spa_fea = TAR_FEA, DEN_FEA, SPA_FEA
        self.tad_fea, self.tot_fea = TAR_FEA + DEN_FEA, TOT_FEA
        self.batch_size, self.max_ind_range = batch_size, max_ind_range
        self.bytes_per_entry = bytes_per_feature * TOT_FEA * batch_size
        size = os.path.getsize(data_file)
        if size % (4 * TOT_FEA):
            raise ValueError(f"{data_file}: {size} bytes is not a whole number of {4 * TOT_FEA}-byte records")
        self.num_samples = size // (4 * TOT_FEA)
        self.num_entries = math.ceil(size / self.bytes_per_entry)
        self.records = np.memmap(data_file, dtype=np.int32, mode="r", shape=(self.num_samples, TOT_FEA))
        with np.load(counts_file) as data:
            self.counts = data["counts"]
        self.m_den = DEN_FEA

    def __len__(self):
        return self.num_entries

    def _rows(self, idx):
        if idx < 0 or idx >= self.num_entries:
            raise IndexError(idx)
        return torch.from_numpy(np.array(self.records[idx * self.batch_size:(idx + 1) * self.batch_size]))

    def __getitem__(self, idx):
        t = self._rows(idx)
        return _transform_features(t[:, 1:14], t[:, 14:], t[:, 0], self.max_ind_range, flag_input_torch_tensor=True)

    def read_packed(self, idx, step, out=None):
        """Batch `idx` in the packed staging layout of `step` (a GraphedTrainStep whose static shapes match
        batch_size): returns a pinned uint8 buffer for ``step.load_packed``."""
        X, lS_o, lS_i, T = self[idx]
        if out is None:
            return step.pack_host(X, lS_o, lS_i.contiguous(), T)
        for (o, n, dt, shape), t in zip(step._layout, (X, lS_o, lS_i, T)):
            if tuple(t.shape) != shape:
                raise ValueError(f"batch {idx} has shape {tuple(t.shape)}, the step was captured for {shape}")
            out[o:o + n].view(dt).view(shape).copy_(t)
        return out


def numpy_to_binary(input_files, output_file_path, split="train"):
    """Pre-processed per-day npz files (y, X_int, X_cat) -> one int32 record file (:243-280).  'train' concatenates
    all inputs; 'test' / 'val' take the first / second half (midpoint = ceil(n/2)) of the single input."""
    def records(path):
        with np.load(path) as d:
            return np.concatenate([d["y"].reshape(-1, 1), d["X_int"], d["X_cat"]], axis=1).astype(np.int32)
    with open(output_file_path, "wb") as f:
        if split == "train":
            for path in input_files:
                f.write(records(path).tobytes())
            return
        if len(input_files) != 1:
            raise ValueError("test/val are cut from exactly one input file")
        rec = records(input_files[0])
        mid = int(np.ceil(rec.shape[0] / 2.0))
        if split == "test":
            rec = rec[:mid]
        elif split == "val":
            rec = rec[mid:]
        else:
            raise ValueError("Unknown split value: " + str(split))
        f.write(rec.tobytes())
