"""Input side of the hot path ("next"): the batch formats the reference's training loop
consumes, produced with the reference's own semantics so that a run can be switched over batch for batch.

Reference: dlrm_data_pytorch.py
  RandomDataset / generate_dist_input_batch / generate_random_output_batch   :773-865, 1036-1044, 1092-1157
  collate_wrapper_random_offset, make_random_data_and_loader                 :868-960
  CriteoDataset (pre-processed, in-memory branch), __getitem__                :44-300
  collate_wrapper_criteo_offset, make_criteo_data_and_loaders (plain branch)  :328-345, 423-577

What is built: the random-data generator with the reference's exact numpy call sequence (same seed -> the same
batches, bit for bit -- tests/test_data_formats.py pins this against batches produced by the reference), the Criteo
dataset reader for ALREADY PRE-PROCESSED ``*_processed.npz`` files with the reference's train/val/test split and
randomisation rules, both collate functions, and ``PackedCriteoCollate``, which collates straight into the pinned
single-copy staging layout of ``graph_step.GraphedTrainStep`` (one H2D per step).
Out of scope (raises): the raw-text pre-processing pipeline (data_utils.getCriteoAdData), memory-mapped per-day
files, the MLPerf binary loader, synthetic trace-driven generation.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

ra = np.random          # the reference draws from numpy's GLOBAL legacy generator (dlrm_data_pytorch.py:29)


# ---------------------------------------------------------------------------------------------------------------
# random data (BASELINE.json configs[0])
# ---------------------------------------------------------------------------------------------------------------
def _bag_size(table_rows, num_indices_per_lookup, fixed):
    """Lookups of one bag before de-duplication (:1108-1117): fixed, or round(max(1, r * min(rows, P)))."""
    if fixed:
        return np.int64(num_indices_per_lookup)
    r = ra.random(1)
    return np.int64(np.round(max([1.0], r * min(table_rows, num_indices_per_lookup))))


def generate_dist_input_batch(m_den, ln_emb, n, num_indices_per_lookup, num_indices_per_lookup_fixed,
                              rand_data_dist="uniform", rand_data_min=1, rand_data_max=1, rand_data_mu=-1,
                              rand_data_sigma=1):
    """One batch of dense features and per-table (offsets, indices) lists (:1092-1157).  The order and sizes of
    the draws from numpy's global generator are the reference's, so equal seeds give equal batches."""
    X = torch.tensor(ra.rand(n, m_den).astype(np.float32))
    all_offsets, all_indices = [], []
    for rows in ln_emb:
        offsets, flat, cursor = np.empty(n, dtype=np.int64), [], 0
        for b in range(n):
            k = _bag_size(rows, num_indices_per_lookup, num_indices_per_lookup_fixed)
            if rand_data_dist == "gaussian":
                if rand_data_mu == -1:
                    rand_data_mu = (rand_data_max + rand_data_min) / 2.0
                draw = np.clip(ra.normal(rand_data_mu, rand_data_sigma, k), rand_data_min, rand_data_max)
                bag = np.unique(draw).astype(np.int64)
            elif rand_data_dist == "uniform":
                bag = np.unique(np.round(ra.random(k) * (rows - 1)).astype(np.int64))
            else:
                raise ValueError(f"{rand_data_dist} distribution is not supported. please select uniform or gaussian")
            offsets[b] = cursor
            flat.append(bag)
            cursor += bag.size
        all_offsets.append(torch.from_numpy(offsets))
        all_indices.append(torch.from_numpy(np.concatenate(flat) if flat else np.zeros(0, dtype=np.int64)))
    return X, all_offsets, all_indices


def generate_random_output_batch(n, num_targets, round_targets=False):
    """Click probabilities (:1036-1044)."""
    P = ra.rand(n, num_targets).astype(np.float32)
    if round_targets:
        P = np.round(P).astype(np.float32)
    return torch.tensor(P)


class RandomDataset(Dataset):
    """One item = one whole mini-batch (X, lS_o, lS_i, T), generated on access (:773-865)."""

    def __init__(self, m_den, ln_emb, data_size, num_batches, mini_batch_size, num_indices_per_lookup,
                 num_indices_per_lookup_fixed, num_targets=1, round_targets=False, data_generation="random",
                 trace_file="", enable_padding=False, reset_seed_on_access=False, rand_data_dist="uniform",
                 rand_data_min=1, rand_data_max=1, rand_data_mu=-1, rand_data_sigma=1, rand_seed=0):
        if data_generation != "random":
            raise NotImplementedError("--data-generation=" + str(data_generation) + ": only 'random' is built "
                                      "(trace-driven synthetic generation is outside the hot path)")
        nbatches = int(np.ceil((data_size * 1.0) / mini_batch_size))
        if num_batches != 0:
            nbatches = num_batches
            data_size = nbatches * mini_batch_size
        self.m_den, self.ln_emb = m_den, ln_emb
        self.data_size, self.num_batches, self.mini_batch_size = data_size, nbatches, mini_batch_size
        self.num_indices_per_lookup = num_indices_per_lookup
        self.num_indices_per_lookup_fixed = num_indices_per_lookup_fixed
        self.num_targets, self.round_targets = num_targets, round_targets
        self.data_generation, self.trace_file, self.enable_padding = data_generation, trace_file, enable_padding
        self.reset_seed_on_access, self.rand_seed = reset_seed_on_access, rand_seed
        self.rand_data_dist, self.rand_data_min, self.rand_data_max = rand_data_dist, rand_data_min, rand_data_max
        self.rand_data_mu, self.rand_data_sigma = rand_data_mu, rand_data_sigma

    def reset_numpy_seed(self, numpy_rand_seed):
        np.random.seed(numpy_rand_seed)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return [self[i] for i in range(index.start or 0, index.stop or len(self), index.step or 1)]
        if self.reset_seed_on_access and index == 0:          # same samples every epoch (:832-833)
            self.reset_numpy_seed(self.rand_seed)
        n = min(self.mini_batch_size, self.data_size - (index * self.mini_batch_size))
        X, lS_o, lS_i = generate_dist_input_batch(
            self.m_den, self.ln_emb, n, self.num_indices_per_lookup, self.num_indices_per_lookup_fixed,
            rand_data_dist=self.rand_data_dist, rand_data_min=self.rand_data_min, rand_data_max=self.rand_data_max,
            rand_data_mu=self.rand_data_mu, rand_data_sigma=self.rand_data_sigma)
        T = generate_random_output_batch(n, self.num_targets, self.round_targets)
        return X, lS_o, lS_i, T

    def __len__(self):
        return self.num_batches


def collate_wrapper_random_offset(list_of_tuples):
    """The DataLoader hands over one pre-built batch (:868-874): stack the offsets, keep the index list."""
    X, lS_o, lS_i, T = list_of_tuples[0]
    return X, torch.stack(lS_o), lS_i, T


def make_random_data_and_loader(args, ln_emb, m_den, offset_to_length_converter=False):
    """:886-960.  `args` needs: data_size, num_batches, mini_batch_size, num_indices_per_lookup,
    num_indices_per_lookup_fixed, round_targets, data_generation, numpy_rand_seed and optionally
    data_trace_file, data_trace_enable_padding, rand_data_{dist,min,max,mu,sigma}, num_workers."""
    if offset_to_length_converter:
        raise NotImplementedError("the length (caffe2) batch format is not consumed by the DQRM drivers")
    opt = lambda k, d: getattr(args, k, d)

    def build():
        return RandomDataset(m_den, ln_emb, args.data_size, args.num_batches, args.mini_batch_size,
                             args.num_indices_per_lookup, args.num_indices_per_lookup_fixed, 1, args.round_targets,
                             args.data_generation, opt("data_trace_file", ""), opt("data_trace_enable_padding", False),
                             reset_seed_on_access=True, rand_data_dist=opt("rand_data_dist", "uniform"),
                             rand_data_min=opt("rand_data_min", 0), rand_data_max=opt("rand_data_max", 1),
                             rand_data_mu=opt("rand_data_mu", -1), rand_data_sigma=opt("rand_data_sigma", 1),
                             rand_seed=args.numpy_rand_seed)
    train_data, test_data = build(), build()
    kw = dict(batch_size=1, shuffle=False, num_workers=opt("num_workers", 0), collate_fn=collate_wrapper_random_offset,
              pin_memory=False, drop_last=False)
    return train_data, DataLoader(train_data, **kw), test_data, DataLoader(test_data, **kw)


# ---------------------------------------------------------------------------------------------------------------
# Criteo Kaggle / Terabyte, pre-processed files
# ---------------------------------------------------------------------------------------------------------------
class CriteoDataset(Dataset):
    """In-memory reader of a pre-processed Criteo file (:44-300, the ``memory_map=False`` branch with the
    processed file present).  Needs ``pro_data`` (npz: X_int [S,13], X_cat [S,26], y [S], counts [26]) and, next to
    ``raw_path``, ``<stem>_day_count.npz`` (total_per_file) -- both written by the reference's pre-processing.
    Items are (X_int[i], X_cat[i] (% max_ind_range), y[i]) numpy rows, as in the reference."""

    def __init__(self, dataset, max_ind_range, sub_sample_rate, randomize, split="train", raw_path="", pro_data="",
                 memory_map=False, dataset_multiprocessing=False):
        if dataset == "kaggle":
            days = 7
        elif dataset == "terabyte":
            days = 24
        else:
            raise ValueError("Data set option is not supported")
        if memory_map:
            raise NotImplementedError("memory-mapped per-day files are outside the hot path; use the processed npz")
        if not os.path.exists(str(pro_data)):
            raise NotImplementedError(f"pre-processed file {pro_data!r} not found: raw Criteo text pre-processing "
                                      "(data_utils.getCriteoAdData) is outside the hot path -- run the reference's once")
        self.max_ind_range, self.memory_map, self.split = max_ind_range, memory_map, split
        parts = raw_path.split("/")
        self.d_path = "/".join(parts[0:-1]) + "/"
        self.d_file = parts[-1].split(".")[0] if dataset == "kaggle" else parts[-1]
        with np.load(self.d_path + self.d_file + "_day_count.npz") as data:
            total_per_file = data["total_per_file"]
        self.offset_per_file = np.concatenate([[0], np.cumsum(np.asarray(total_per_file)[:days])]).astype(np.int64)
        with np.load(str(pro_data)) as data:
            X_int, X_cat, y = data["X_int"], data["X_cat"], data["y"]
            self.counts = data["counts"]
        self.m_den, self.n_emb = X_int.shape[1], len(self.counts)
        order = np.arange(len(y))
        if split == "none":
            if randomize == "total":
                order = ra.permutation(order)
            # the reference scatters (X[order] = X): sample i moves to position order[i]   (:236-238)
            self.X_int, self.X_cat, self.y = np.empty_like(X_int), np.empty_like(X_cat), np.empty_like(y)
            self.X_int[order], self.X_cat[order], self.y[order] = X_int, X_cat, y
            return
        per_day = np.array_split(order, self.offset_per_file[1:-1])
        if randomize == "day":
            for d in range(len(per_day) - 1):
                per_day[d] = ra.permutation(per_day[d])
        train_idx = np.concatenate(per_day[:-1])
        test_idx, val_idx = np.array_split(per_day[-1], 2)
        if randomize == "total":
            train_idx = ra.permutation(train_idx)
        if split not in ("train", "val", "test"):
            raise SystemExit("ERROR: dataset split is neither none, nor train or test.")
        pick = {"train": train_idx, "val": val_idx, "test": test_idx}[split]
        self.X_int, self.X_cat, self.y = X_int[pick], X_cat[pick], y[pick]

    def __getitem__(self, index):
        if isinstance(index, slice):
            return [self[i] for i in range(index.start or 0, index.stop or len(self), index.step or 1)]
        cat = self.X_cat[index] % self.max_ind_range if self.max_ind_range > 0 else self.X_cat[index]
        return self.X_int[index], cat, self.y[index]

    def __len__(self):
        return len(self.y)


def collate_wrapper_criteo_offset(list_of_tuples):
    """List of (X_int, X_cat, y) -> (X [B,13] = log(X_int + 1), lS_o [26,B] = arange(B) per table,
    lS_i [26,B] = X_cat transposed, T [B,1])  (:328-345)."""
    ints, cats, ys = zip(*list_of_tuples)
    X = torch.log(torch.tensor(np.asarray(ints), dtype=torch.float) + 1)
    cat = torch.tensor(np.asarray(cats)).type(torch.LongTensor)
    T = torch.tensor(np.asarray(ys), dtype=torch.float32).view(-1, 1)
    B, F = cat.shape
    lS_o = torch.arange(B).repeat(F, 1)
    lS_i = cat.t().contiguous()
    return X, lS_o, lS_i, T


class PackedCriteoCollate:
    """collate_fn that writes a Criteo batch straight into the packed staging layout of a GraphedTrainStep
    (X | lS_o | lS_i | T, one pinned uint8 buffer), so the step's inputs arrive with ONE H2D copy:

        collate = PackedCriteoCollate(step, depth=4)
        loader = DataLoader(train_data, batch_size=B, collate_fn=collate, drop_last=True)
        for packed in loader:
            step.load_packed(packed); step.run()

    `depth` pinned buffers are cycled; a buffer must not be refilled while its copy is still in flight, so keep
    depth >= the number of steps the host may run ahead (+ DataLoader prefetch).  num_workers must be 0 (pinned
    buffers are owned by this process)."""

    def __init__(self, step, depth=4):
        self.step, self.i = step, 0
        self.bufs = [step.pack_host(step.X.cpu(), step.lS_o.cpu(), step.lS_i.cpu(), step.T.cpu()) for _ in range(depth)]

    def __call__(self, list_of_tuples):
        X, lS_o, lS_i, T = collate_wrapper_criteo_offset(list_of_tuples)
        buf = self.bufs[self.i % len(self.bufs)]
        self.i += 1
        for (o, n, dt, shape), t in zip(self.step._layout, (X, lS_o, lS_i, T)):
            if tuple(t.shape) != shape:
                raise ValueError(f"batch shape {tuple(t.shape)} != captured static shape {shape} (use drop_last=True)")
            buf[o:o + n].view(dt).view(shape).copy_(t)
        return buf


def make_criteo_data_and_loaders(args, offset_to_length_converter=False):
    """:423-577, plain branch (no MLPerf loader, no memory map)."""
    if offset_to_length_converter:
        raise NotImplementedError("the length (caffe2) batch format is not consumed by the DQRM drivers")
    if getattr(args, "mlperf_logging", False) and args.memory_map and args.data_set == "terabyte":
        raise NotImplementedError("the MLPerf binary / per-day terabyte loaders are outside the hot path")

    def build(split):
        return CriteoDataset(args.data_set, args.max_ind_range, args.data_sub_sample_rate, args.data_randomize, split,
                             args.raw_data_file, args.processed_data_file, args.memory_map,
                             getattr(args, "dataset_multiprocessing", False))
    train_data, test_data = build("train"), build("test")
    train_loader = DataLoader(train_data, batch_size=args.mini_batch_size, shuffle=False, num_workers=args.num_workers,
                              collate_fn=collate_wrapper_criteo_offset, pin_memory=False, drop_last=False)
    test_loader = DataLoader(test_data, batch_size=args.test_mini_batch_size, shuffle=False,
                             num_workers=args.test_num_workers, collate_fn=collate_wrapper_criteo_offset,
                             pin_memory=False, drop_last=False)
    return train_data, train_loader, test_data, test_loader
