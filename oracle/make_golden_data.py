"""Generate tests/golden/data_*.npz and the tiny Criteo fixture by EXECUTING THE REFERENCE's input pipeline
(dlrm_data_pytorch.py) in this container.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_data.py

Pins deep_quantized_recommendation_model_dqrm_b200/dlrm_data_pytorch.py:
  RandomDataset / collate_wrapper_random_offset        dlrm_data_pytorch.py:773-874, 1092-1157
  CriteoDataset (processed file, in-memory) / collate_wrapper_criteo_offset   :44-345
The Criteo fixture (70 samples, 7 "days") is synthetic: the real dataset cannot be downloaded here.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
FIX = os.path.join(GOLD, "criteo_tiny")
REF = "/root/reference"

RANDOM_CASES = {
    # name: (ln_emb, m_den, mini_batch, num_batches, P, fixed, round_targets, dist kwargs, seed)
    "data_random_uniform": ([10000, 50, 3], 13, 16, 2, 10, False, True, {}, 123),
    "data_random_fixed": ([1000, 7], 4, 8, 2, 3, True, False, {}, 5),
    "data_random_gaussian": ([500, 40], 13, 8, 1, 6, False, True,
                             dict(rand_data_dist="gaussian", rand_data_min=0, rand_data_max=30, rand_data_mu=-1, rand_data_sigma=4), 9),
}
CRITEO_CASES = [("train", "total", 0), ("train", "day", 0), ("train", "none", 17), ("test", "total", 0), ("val", "none", 0)]


def gen_random(dp):
    for name, (ln_emb, m_den, mb, nb, P, fixed, rt, dist, seed) in RANDOM_CASES.items():
        ds = dp.RandomDataset(m_den, np.array(ln_emb), 0, nb, mb, P, fixed, 1, rt, "random", "", False,
                              reset_seed_on_access=True, rand_seed=seed, **dist)
        loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False, num_workers=0,
                                             collate_fn=dp.collate_wrapper_random_offset)
        rec = {}
        for j, (X, lS_o, lS_i, T) in enumerate(loader):
            rec[f"b{j}_X"], rec[f"b{j}_T"], rec[f"b{j}_lS_o"] = X.numpy(), T.numpy(), lS_o.numpy()
            for k, t in enumerate(lS_i):
                rec[f"b{j}_lS_i{k}"] = t.numpy()
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)


def write_criteo_fixture():
    os.makedirs(FIX, exist_ok=True)
    rng = np.random.RandomState(77)
    per_day = np.array([12, 9, 10, 11, 8, 10, 10])
    S = int(per_day.sum())
    X_int = rng.randint(0, 5000, size=(S, 13)).astype(np.int32)
    X_int[rng.rand(S, 13) < 0.2] = 0
    counts = np.array([1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27,
                       14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572], dtype=np.int32)
    X_cat = np.stack([rng.randint(0, c, size=S) for c in counts], axis=1).astype(np.int32)
    y = (rng.rand(S) < 0.25).astype(np.int32)
    np.savez_compressed(os.path.join(FIX, "kaggleAdDisplayChallenge_processed.npz"), X_int=X_int, X_cat=X_cat, y=y, counts=counts)
    np.savez_compressed(os.path.join(FIX, "train_day_count.npz"), total_per_file=per_day)


def gen_criteo(dp):
    raw = os.path.join(FIX, "train.txt")
    pro = os.path.join(FIX, "kaggleAdDisplayChallenge_processed.npz")
    rec = {}
    for split, randomize, mir in CRITEO_CASES:
        np.random.seed(31)
        ds = dp.CriteoDataset("kaggle", mir, 0.0, randomize, split, raw, pro, False, False)
        loader = torch.utils.data.DataLoader(ds, batch_size=8, shuffle=False, num_workers=0,
                                             collate_fn=dp.collate_wrapper_criteo_offset)
        key = f"{split}_{randomize}_{mir}"
        rec[key + "_len"] = len(ds)
        for j, (X, lS_o, lS_i, T) in enumerate(loader):
            if j >= 2:
                break
            rec[f"{key}_b{j}_X"], rec[f"{key}_b{j}_lS_o"] = X.numpy(), lS_o.numpy()
            rec[f"{key}_b{j}_lS_i"], rec[f"{key}_b{j}_T"] = lS_i.numpy(), T.numpy()
    np.savez_compressed(os.path.join(GOLD, "data_criteo_tiny.npz"), **rec)


def gen_terabyte_bin():
    """data_loader_terabyte.py: numpy_to_binary + CriteoBinDataset on two tiny 'day' files cut from the fixture."""
    import data_loader_terabyte as dlt
    with np.load(os.path.join(FIX, "kaggleAdDisplayChallenge_processed.npz")) as d:
        X_int, X_cat, y = d["X_int"], d["X_cat"], d["y"]
    days = []
    for i, (a, b) in enumerate(((0, 33), (33, 70))):
        path = os.path.join(FIX, f"day_{i}_reordered.npz")
        np.savez_compressed(path, X_int=X_int[a:b], X_cat=X_cat[a:b], y=y[a:b])
        days.append(path)
    rec = {}
    for split, files in (("train", days), ("test", days[1:]), ("val", days[1:])):
        out = os.path.join(FIX, f"{split}_data.bin")
        dlt.numpy_to_binary(files, out, split)
        rec[f"{split}_bytes"] = np.frombuffer(open(out, "rb").read(), dtype=np.uint8)
        for mir in (-1, 1000):
            ds = dlt.CriteoBinDataset(out, os.path.join(FIX, "kaggleAdDisplayChallenge_processed.npz"), batch_size=16,
                                      max_ind_range=mir)
            rec[f"{split}_{mir}_len"] = len(ds)
            for j in (0, len(ds) - 1):
                X, lS_o, lS_i, T = ds[j]
                rec[f"{split}_{mir}_b{j}_X"], rec[f"{split}_{mir}_b{j}_lS_o"] = X.numpy(), lS_o.numpy()
                rec[f"{split}_{mir}_b{j}_lS_i"], rec[f"{split}_{mir}_b{j}_T"] = lS_i.numpy(), T.numpy()
            del ds
        os.remove(out)                       # the tests rebuild it with OUR writer and compare the bytes
    np.savez_compressed(os.path.join(GOLD, "data_terabyte_bin.npz"), **rec)


def main():
    sys.path.insert(0, REF)
    import dlrm_data_pytorch as dp
    gen_random(dp)
    write_criteo_fixture()
    gen_criteo(dp)
    gen_terabyte_bin()
    print("ok")


if __name__ == "__main__":
    main()
