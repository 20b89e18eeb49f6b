"""CPU: the input side of the hot path (deep_quantized_recommendation_model_dqrm_b200/dlrm_data_pytorch.py) against
batches produced by the REFERENCE's dlrm_data_pytorch.py (oracle/make_golden_data.py): the random-data generator
must reproduce the reference's batches bit for bit from the same numpy seed; the Criteo reader + collate must
reproduce its splits, randomisation, index folding (max-ind-range) and tensors exactly."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
from conftest import load_golden
from deep_quantized_recommendation_model_dqrm_b200 import dlrm_data_pytorch as dp
import make_golden_data as G

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(G.RANDOM_CASES))
def test_random_dataset_reproduces_reference_batches(name):
    ln_emb, m_den, mb, nb, P, fixed, rt, dist, seed = G.RANDOM_CASES[name]
    g = load_golden(name)
    args = SimpleNamespace(data_size=0, num_batches=nb, mini_batch_size=mb, num_indices_per_lookup=P,
                           num_indices_per_lookup_fixed=fixed, round_targets=rt, data_generation="random",
                           numpy_rand_seed=seed, **dist)
    np.random.seed(999)                                   # the loader reseeds on access to item 0
    train_data, train_loader, _, _ = dp.make_random_data_and_loader(args, np.array(ln_emb), m_den)
    assert len(train_data) == nb
    for epoch in range(2):                                # identical samples every epoch (reset_seed_on_access)
        for j, (X, lS_o, lS_i, T) in enumerate(train_loader):
            assert X.dtype == torch.float32 and lS_o.dtype == torch.int64 and T.shape == (mb, 1)
            assert np.array_equal(X.numpy(), g[f"b{j}_X"]) and np.array_equal(T.numpy(), g[f"b{j}_T"])
            assert np.array_equal(lS_o.numpy(), g[f"b{j}_lS_o"])
            assert len(lS_i) == len(ln_emb)
            for k, t in enumerate(lS_i):
                assert t.dtype == torch.int64 and np.array_equal(t.numpy(), g[f"b{j}_lS_i{k}"]), (j, k)


@pytest.mark.parametrize("split,randomize,mir", G.CRITEO_CASES)
def test_criteo_dataset_and_collate_vs_reference(split, randomize, mir):
    g = load_golden("data_criteo_tiny")
    fix = os.path.join(GOLD, "criteo_tiny")
    np.random.seed(31)
    ds = dp.CriteoDataset("kaggle", mir, 0.0, randomize, split, os.path.join(fix, "train.txt"),
                          os.path.join(fix, "kaggleAdDisplayChallenge_processed.npz"))
    key = f"{split}_{randomize}_{mir}"
    assert len(ds) == int(g[key + "_len"]) and ds.m_den == 13 and ds.n_emb == 26
    loader = torch.utils.data.DataLoader(ds, batch_size=8, shuffle=False, collate_fn=dp.collate_wrapper_criteo_offset)
    for j, (X, lS_o, lS_i, T) in enumerate(loader):
        if j >= 2:
            break
        assert X.dtype == torch.float32 and lS_i.dtype == torch.int64 and lS_o.dtype == torch.int64
        assert np.array_equal(X.numpy(), g[f"{key}_b{j}_X"])             # log(x + 1): the same torch op
        assert np.array_equal(lS_o.numpy(), g[f"{key}_b{j}_lS_o"])
        assert np.array_equal(lS_i.numpy(), g[f"{key}_b{j}_lS_i"])
        assert np.array_equal(T.numpy(), g[f"{key}_b{j}_T"])
        if mir > 0:
            assert int(lS_i.max()) < mir


def test_unsupported_inputs_raise():
    with pytest.raises(ValueError):
        dp.CriteoDataset("avazu", 0, 0.0, "total", "train", "x/train.txt", "nope.npz")
    with pytest.raises(NotImplementedError):
        dp.CriteoDataset("kaggle", 0, 0.0, "total", "train", "x/train.txt", "missing_processed.npz")
    with pytest.raises(NotImplementedError):
        dp.RandomDataset(13, [10], 1, 1, 1, 1, False, data_generation="synthetic")


def test_packed_layout_matches_collate():
    """The staging layout used by GraphedTrainStep.load_packed (graph_step._packed_layout) holds exactly the four
    collated tensors, 16-byte aligned."""
    from deep_quantized_recommendation_model_dqrm_b200.graph_step import _packed_layout
    fix = os.path.join(GOLD, "criteo_tiny")
    np.random.seed(31)
    ds = dp.CriteoDataset("kaggle", 0, 0.0, "none", "train", os.path.join(fix, "train.txt"),
                          os.path.join(fix, "kaggleAdDisplayChallenge_processed.npz"))
    X, lS_o, lS_i, T = dp.collate_wrapper_criteo_offset(ds[0:8])
    lay, nbytes = _packed_layout(X, lS_o, lS_i, T)
    buf = torch.zeros(nbytes, dtype=torch.uint8)
    for (o, n, dt, shape), t in zip(lay, (X, lS_o, lS_i, T)):
        assert o % 16 == 0
        buf[o:o + n].view(dt).view(shape).copy_(t)
    for (o, n, dt, shape), t in zip(lay, (X, lS_o, lS_i, T)):
        assert torch.equal(buf[o:o + n].view(dt).view(shape), t)
    assert nbytes == sum((t.numel() * t.element_size() + 15) // 16 * 16 for t in (X, lS_o, lS_i, T))


@pytest.mark.parametrize("split", ["train", "test", "val"])
def test_terabyte_binary_writer_and_reader_vs_reference(split, tmp_path):
    """data_loader_terabyte: our numpy_to_binary writes the reference's bytes, and CriteoBinDataset returns the
    reference's batches (first and short last batch, with and without max-ind-range folding)."""
    from deep_quantized_recommendation_model_dqrm_b200 import data_loader_terabyte as dlt
    g = load_golden("data_terabyte_bin")
    fix = os.path.join(GOLD, "criteo_tiny")
    days = [os.path.join(fix, f"day_{i}_reordered.npz") for i in range(2)]
    out = str(tmp_path / f"{split}_data.bin")
    dlt.numpy_to_binary(days if split == "train" else days[1:], out, split)
    assert np.array_equal(np.frombuffer(open(out, "rb").read(), dtype=np.uint8), g[f"{split}_bytes"])
    counts = os.path.join(fix, "kaggleAdDisplayChallenge_processed.npz")
    for mir in (-1, 1000):
        ds = dlt.CriteoBinDataset(out, counts, batch_size=16, max_ind_range=mir)
        assert len(ds) == int(g[f"{split}_{mir}_len"]) and ds.m_den == 13 and len(ds.counts) == 26
        for j in (0, len(ds) - 1):
            X, lS_o, lS_i, T = ds[j]
            assert X.dtype == torch.float32 and lS_i.dtype == torch.int64 and lS_o.dtype == torch.int64
            assert np.array_equal(X.numpy(), g[f"{split}_{mir}_b{j}_X"])
            assert np.array_equal(lS_o.numpy(), g[f"{split}_{mir}_b{j}_lS_o"])
            assert np.array_equal(lS_i.numpy(), g[f"{split}_{mir}_b{j}_lS_i"])
            assert np.array_equal(T.numpy(), g[f"{split}_{mir}_b{j}_T"])
        with pytest.raises(IndexError):
            ds[len(ds)]
    with pytest.raises(ValueError):
        dlt.numpy_to_binary(days, out, "test")


def test_int4_table_file_roundtrip(tmp_path):
    """int4_checkpoint: tables packed by the oracle spec (the bytes dqrm_table_pack_int4 produces, pinned on the GPU
    by tests/test_gpu_parity.py) survive save -> load bit for bit, the layout is 256-byte aligned, and the serving
    forward on the loaded file equals the one on the original arrays."""
    import warnings
    from oracle import dqrm_oracle as O
    from deep_quantized_recommendation_model_dqrm_b200 import int4_checkpoint as ck, synthetic
    rng = np.random.RandomState(4)
    rows, dim = [3, 1000, 257], 16
    Ws = [synthetic.table_weights_numpy(n, dim, rng) for n in rows]
    scales = np.array([O.table_scale_spec(W, 4) for W in Ws], dtype=np.float32)
    packed = [O.pack_int4_spec(W, s) for W, s in zip(Ws, scales)]
    path = str(tmp_path / "tables.dqrm4")
    total = ck.save(path, packed, scales, dim)
    head, offs, want_total = ck.layout(rows, dim)
    assert total == want_total == os.path.getsize(path) and all(o % 256 == 0 for o in offs) and head == offs[0]
    assert total < sum(W.nbytes for W in Ws) // 7                      # ~8x smaller than fp32 (+ alignment)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                                # torch warns about the read-only map
        got, sc, got_rows, got_dim = ck.load(path)
    assert got_rows == rows and got_dim == dim and np.array_equal(sc.numpy(), scales)
    for a, b in zip(packed, got):
        assert np.array_equal(a, b.numpy())
    idx = rng.randint(0, 1000, size=40)
    off = np.arange(0, 40, 4)
    assert np.array_equal(O.embbag_forward_int4_spec(got[1].numpy(), idx, off, sc[1].item()),
                          O.embbag_forward_int4_spec(packed[1], idx, off, scales[1]))
    with open(path, "r+b") as f:
        f.write(b"NOTDQRM!")
    with pytest.raises(ValueError):
        ck.load(path)
    with pytest.raises(ValueError):
        ck.save(path, [packed[0].astype(np.int8)], scales[:1], dim)
