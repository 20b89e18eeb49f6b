"""CPU (no GPU): the C-ABI library loads and exports every symbol include/dqrm_b200.h declares, argument
validation works without a device, host-side logic (input packing, shard, arena channel table, flag
parsing), and the world_size-2 gloo path of extend_distributed / weight_syncc."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from deep_quantized_recommendation_model_dqrm_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dqrm_b200.h")).read()
    declared = set(re.findall(r"DQRM_API\s+[\w\s\*]+?\b(dqrm_\w+)\s*\(", hdr))
    assert len(declared) >= 22
    lib = _lib.load()
    import ctypes
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
        assert name in _lib.SIGNATURES, f"{name} declared in the header but not bound in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) == declared
    assert lib.dqrm_abi_version() == _lib.ABI_VERSION == 5


def test_argument_validation_without_a_device():
    from deep_quantized_recommendation_model_dqrm_b200 import _lib
    lib = _lib.load()
    assert lib.dqrm_scan_workspace_bytes(26) == 27 * 4
    assert lib.dqrm_slot_bytes(26, 128, 16, 8) == 112 + 26 * 128 * 4 + 26 * 128 * 16
    assert lib.dqrm_slot_bytes(26, 128, 16, 16) == 112 + 26 * 128 * 4 + 26 * 128 * 16 * 2
    # room for either backward path (CTA: block sums of long rows; sort kernel: keys, histograms, partial sums)
    assert lib.dqrm_bwd_workspace_bytes(26, 128, 16) >= 26 * (128 // 64 + 128 // 65 + 2) * 16 * 4
    assert lib.dqrm_bwd_workspace_bytes(26, 8192, 64) >= 26 * (8192 // 64 + 8192 // 65 + 2) * 64 * 4
    assert lib.dqrm_bwd_workspace_bytes(1, 1 << 20, 16) > 4 * (1 << 20) * 4
    # errors are returned, not thrown, and carry a message (no kernel is launched on these paths)
    rc = lib.dqrm_table_absmax_scale(0, None, None, 16, 4, 0, 1, None, None, None, None, None)
    assert rc < 0 and "num_tables" in _lib.last_error()
    rc = lib.dqrm_embbag_fwd(1, None, None, 18, None, None, None, 4, None, None, 4, None, 0, 0, None, None, None)
    assert rc < 0
    rc = lib.dqrm_scale_from_absmax(1, 1, 40, 1, None, None)
    assert rc == -22 and "bits" in _lib.last_error()
    with pytest.raises(_lib.DqrmLibraryError):
        _lib.check(rc, "dqrm_scale_from_absmax")


def test_no_cpu_fallback():
    from deep_quantized_recommendation_model_dqrm_b200 import _lib
    from deep_quantized_recommendation_model_dqrm_b200.quantization_supp import quant_modules as qm
    from deep_quantized_recommendation_model_dqrm_b200.quantization_supp import quant_utils as qu
    E = qm.QuantEmbeddingBagTwo(10, 16, 4)
    with pytest.raises(_lib.DqrmLibraryError):
        E(torch.tensor([1, 2]), torch.tensor([0, 1]))
    with pytest.raises(_lib.DqrmLibraryError):
        qu.symmetric_linear_quantization_param_two(4, torch.zeros(4, 16), None, None, None)
    # nothing in the product package imports the oracle
    pkg = os.path.join(ROOT, "deep_quantized_recommendation_model_dqrm_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_pack_inputs_and_shard():
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic
    from deep_quantized_recommendation_model_dqrm_b200.tables import EmbeddingTableGroup
    from deep_quantized_recommendation_model_dqrm_b200 import dlrm_s_pytorch_comm_grad as drv
    X, lS_o, lS_i, T = synthetic.criteo_batch([10, 20, 30], 8, seed=1)
    idx, off, ib, bags = EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cpu")
    assert bags == 8 and ib == [0, 8, 16, 24] and idx.numel() == 24 and off.shape == (3, 8)
    X, lS_o, lS_i, T = synthetic.random_batch([10, 20, 30], 5, 4, seed=2)
    idx, off, ib, bags = EmbeddingTableGroup.pack_inputs(lS_i, lS_o, "cpu")
    assert bags == 5 and ib[-1] == sum(t.numel() for t in lS_i) and torch.equal(idx[ib[1]:ib[2]], lS_i[1])
    for n, w in ((128, 8), (10, 3), (5, 8)):
        got = []
        for r in range(w):
            got += list(range(n))[drv.get_my_slice(n, w, r)]
        assert got == list(range(n))
    assert synthetic.top_mlp_sizes(26, 16, [512, 256, 1]) == [367, 512, 256, 1]
    assert synthetic.top_mlp_sizes(26, 64, [512, 512, 256, 1])[0] == 415


def test_flags_match_reference_defaults():
    from deep_quantized_recommendation_model_dqrm_b200 import dlrm_s_pytorch_comm_grad as drv
    a = drv.parse_args(["--quantization_flag", "--quantize_act_and_lin", "--embedding_bit=4", "--weight_bit=4",
                        "--linear_channel", "--quantize_activation", "--quantize_embedding_bag_gradient",
                        "--embedding_bag_gradient_bit_num=8", "-n", "1", "-g", "4", "-nr", "0"])
    assert a.quantize_activation is False           # --linear_channel forces it off (drv:1155-1156)
    assert a.world_size == 4 and a.embedding_bag_gradient_bit_num == 8
    assert drv.parse_args([]).embedding_bag_gradient_bit_num == 16


def test_dense_arena_channel_table_cpu():
    from deep_quantized_recommendation_model_dqrm_b200.dense import DenseArena
    from deep_quantized_recommendation_model_dqrm_b200.quantization_supp import quant_modules as qm
    layers = []
    for n_in, n_out in ((13, 8), (8, 4)):
        q = qm.QuantLinear(weight_bit=4, bias_bit=4, per_channel=True)
        q.set_param(torch.nn.Linear(n_in, n_out))
        layers.append(q)
    w0 = layers[0].weight.data.clone()
    a = DenseArena(layers, torch.device("cpu"))
    assert a.total == 13 * 8 + 8 + 8 * 4 + 4 and a.num_chan == 8 + 1 + 4 + 1
    cb = a.chan_begin.tolist()
    assert cb[:3] == [0, 13, 26] and cb[8] == 104 and cb[9] == 112 and cb[-1] == a.total
    assert torch.equal(layers[0].weight.data, w0) and a.intact()
    layers[0].weight.grad.add_(1.0)
    assert a.flat_grad[:104].eq(1.0).all() and a.flat_grad[104:].eq(0.0).all()


def _gloo_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    sys.path.insert(0, ROOT)
    from deep_quantized_recommendation_model_dqrm_b200 import extend_distributed as ext
    from deep_quantized_recommendation_model_dqrm_b200 import sgd_quantized_gradients_parallel_comm as sgd
    ext.init_distributed(rank=rank, local_rank=rank, size=world, use_gpu=False, backend="gloo")
    assert ext.my_rank == rank and ext.my_size == world
    sl = ext.get_my_slice(10)
    ln, splits = ext.get_split_lengths(7)
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin.weight.fill_(float(rank + 1))
        lin.bias.fill_(float(10 * (rank + 1)))

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.l = lin
    m = M()
    sgd.weight_syncc(m, world)                       # all-reduce-average of every parameter (sgd:963-970)
    g = ext.all_gather(torch.full((2,), float(rank)), None)
    ext.barrier()
    torch.save({"slice": (sl.start, sl.stop), "split": (ln, splits), "w": m.l.weight.data.clone(),
                "b": m.l.bias.data.clone(), "g": g}, out.format(rank=rank))
    torch.distributed.destroy_process_group()


def test_world2_gloo_host_logic(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "r{rank}.pt")
    mp.spawn(_gloo_worker, args=(2, 29741, out), nprocs=2, join=True)
    r0, r1 = torch.load(out.format(rank=0)), torch.load(out.format(rank=1))
    assert r0["slice"] == (0, 5) and r1["slice"] == (5, 10)
    assert r0["split"] == (4, [4, 3]) and r1["split"] == (3, [4, 3])
    for r in (r0, r1):
        assert torch.all(r["w"] == 1.5) and torch.all(r["b"] == 15.0)
        assert r["g"].tolist() == [0.0, 0.0, 1.0, 1.0]


def test_p2p_site_layout_host_functions():
    """Peer-arena site geometry (pure host code of the library): ctl | flags | slots, 16-byte aligned slots."""
    import ctypes as C
    from deep_quantized_recommendation_model_dqrm_b200 import _lib
    lib = _lib.load()
    for world in (1, 2, 8, 16):
        for slot in (1, 104, 6496, 66672, 475985):
            fo, do, st = C.c_size_t(), C.c_size_t(), C.c_size_t()
            assert lib.dqrm_p2p_site_layout(world, slot, C.byref(fo), C.byref(do), C.byref(st)) == 0
            assert fo.value == 16 and do.value % 16 == 0 and do.value >= 16 + 4 * world
            assert st.value % 16 == 0 and slot <= st.value < slot + 16
            assert lib.dqrm_p2p_site_bytes(world, slot) == do.value + world * st.value
    # argument checking happens on the host, before any launch
    assert lib.dqrm_p2p_allgather(None, 2, 0, 0, 64, None, None) == -22
    assert b"null" in lib.dqrm_last_error()


def test_exchange_backend_env(monkeypatch):
    from deep_quantized_recommendation_model_dqrm_b200 import p2p
    monkeypatch.delenv("DQRM_EXCHANGE", raising=False)
    assert p2p.backend() == "p2p"
    monkeypatch.setenv("DQRM_EXCHANGE", "NCCL")
    assert p2p.backend() == "nccl"
    monkeypatch.setenv("DQRM_EXCHANGE", "gloo")
    import pytest
    with pytest.raises(ValueError):
        p2p.backend()


def test_driver_cli_data_flags_and_cpu_refusal():
    """The reference's data / loader flags parse with its defaults (dlrm_s_pytorch_comm_grad.py:1039-1096,
    1370-1373), and the training entry point refuses to run without a GPU instead of falling back."""
    from deep_quantized_recommendation_model_dqrm_b200 import _lib
    from deep_quantized_recommendation_model_dqrm_b200 import dlrm_s_pytorch_comm_grad as drv
    a = drv.parse_args(["--arch-sparse-feature-size=16", "--arch-embedding-size=10000-10000",
                        "--arch-mlp-bot=13-512-256-64-16", "--arch-mlp-top=512-256-1", "--data-generation=random",
                        "--mini-batch-size=128", "--num-batches=3", "--quantization_flag", "--embedding_bit=4",
                        "--weight_bit=4", "--linear_channel", "--quantize_act_and_lin", "--loss-function=bce"])
    assert (a.data_set, a.data_randomize, a.memory_map, a.num_workers) == ("kaggle", "total", False, 0)
    assert a.test_mini_batch_size == 128 and a.test_num_workers == 0 and a.rand_data_dist == "uniform"
    assert a.quantize_activation is False and a.world_size == 1
    if not torch.cuda.is_available():
        import pytest
        with pytest.raises(_lib.DqrmLibraryError):
            drv.train(a)


def test_c_abi_rejects_bad_arguments_before_any_launch():
    """Error convention of the boundary (include/dqrm_b200.h): bad arguments return -errno with a message in
    dqrm_last_error() -- checked on the host, so this runs without a GPU."""
    import ctypes as C
    import errno
    from deep_quantized_recommendation_model_dqrm_b200 import _lib
    lib = _lib.load()
    one = (C.c_void_p * 1)(16)
    rows = (C.c_int64 * 1)(10)
    cases = [
        (lib.dqrm_table_absmax_scale(0, one, rows, 16, 4, 0, 1, 16, 16, 16, 16, None), -errno.E2BIG, b"num_tables"),
        (lib.dqrm_table_absmax_scale(1, one, rows, 16, 4, 2, 2, 16, 16, 16, 16, None), -errno.EINVAL, b"shard"),
        (lib.dqrm_table_absmax_scale(1, one, rows, 16, 1, 0, 1, 16, 16, 16, 16, None), -errno.EINVAL, b"bits"),
        (lib.dqrm_scale_from_absmax(0, 16, 4, 16, 16, None), -errno.EINVAL, b"scale_from_absmax"),
        (lib.dqrm_blockmax_scan(1, one, rows, 6, 64, one, 0, 1, None), -errno.EINVAL, b"multiple of 4"),
        (lib.dqrm_blockmax_reduce(1, rows, 64, one, 0, 1, 4, 16, 16, None, 16, None), -errno.EINVAL, b"scale/inv_scale"),
        (lib.dqrm_grad_pack(1, 16, None, None, None, 8, None, 1, 1, 8, None, None, None), -errno.EINVAL, b"null"),
        (lib.dqrm_grad_pack(1, 16, 16, 16, 16, 8, 16, 0, 1, 8, 16, 16, None), -errno.EINVAL, b"scale_stride"),
        (lib.dqrm_dense_grad_quant_gathered(16, 16, 4, 16, 4, 2, 16, 16, 16, None), -errno.EINVAL, b"int8"),
        (lib.dqrm_bce_loss_grad(None, None, 4, None, None, None), -errno.EINVAL, b"bce_loss_grad"),
        (lib.dqrm_p2p_allgather(one, 17, 0, 0, 64, 16, None), -errno.EINVAL, b"world"),
        # the fused backward + row update: a missing table array, then a table pointer that is not 16-byte aligned
        (lib.dqrm_embbag_bwd_sgd(1, None, rows, 16, 16, 16, rows, 4, 16, 64, 16, None, 8, 16, 16, 16,
                                 0.1, None, 1.0, None, 1e-10, 16, 16, 1024, None), -errno.EINVAL, b"embbag_bwd_sgd"),
        (lib.dqrm_embbag_bwd_sgd(1, (C.c_void_p * 1)(24), rows, 16, 16, 16, rows, 4, 16, 64, 16, None, 8, 16, 16, 16,
                                 0.1, None, 1.0, None, 1e-10, 16, 16, 1024, None), -errno.EINVAL, b"aligned"),
    ]
    for i, (rc, want, frag) in enumerate(cases):
        assert rc == want, (i, rc, want)
    assert lib.dqrm_grad_pack(1, 16, None, None, None, 8, None, 1, 1, 8, None, None, None) == -errno.EINVAL
    assert b"null" in lib.dqrm_last_error()
    assert lib.dqrm_blockmax_scan(1, one, rows, 6, 64, one, 0, 1, None) == -errno.EINVAL
    assert b"multiple of 4" in lib.dqrm_last_error()


def test_lr_policy_scheduler_vs_reference_sequence():
    """The lr handed to weight_update_parallel_comm every iteration: our stand-alone LRPolicyScheduler against 30
    steps of the reference's (torch _LRScheduler based) class, tests/golden/lr_policy.json."""
    import json
    from deep_quantized_recommendation_model_dqrm_b200 import dlrm_s_pytorch_comm_grad as drv
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lr_policy.json")))
    for name, g in gold.items():
        sch = drv.LRPolicyScheduler(*g["args"])
        got = []
        for _ in range(len(g["lrs"])):
            got.append(sch.get_lr()[-1])
            sch.step()
        assert got == g["lrs"], name
    with pytest.raises(SystemExit):
        drv.LRPolicyScheduler(0.1, 5, 2, 3)


def test_reference_api_leftovers(capsys):
    """Names the reference modules export that its drivers never reach: present, documented, loud."""
    import pytest
    from deep_quantized_recommendation_model_dqrm_b200 import sgd_quantized_gradients_parallel_comm as sgd
    from deep_quantized_recommendation_model_dqrm_b200.quantization_supp.quant_modules import QuantEmbeddingBagTwo
    with pytest.raises(NotImplementedError):
        sgd.grad_precision_and_scale(None, 2, 0)
    e = QuantEmbeddingBagTwo(10, 16)
    e.set_iteration_bound()
    assert e.iteration_bound.item() == 0                     # (counters never move in the reference: qmngq:348-361)
    e.iteration_nt += 1
    e.set_iteration_bound()
    assert e.iteration_bound.item() == 1000 and "bound increasing to 1000" in capsys.readouterr().out


def test_oracle_ref_manifest_matches_files():
    """oracle/_ref (the unmodified reference files for bench.py's reference arm, oracle/make_ref.py): when present,
    every file has the size and sha256 its MANIFEST recorded at copy time, and the product package imports none of it."""
    import hashlib, json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = os.path.join(root, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "MANIFEST.json")):
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    man = json.load(open(os.path.join(ref, "MANIFEST.json")))
    assert "dlrm_s_pytorch_comm_grad.py" in man and "sgd_quantized_gradients_parallel_comm.py" in man
    for rel, rec in man.items():
        b = open(os.path.join(ref, rel), "rb").read()
        assert len(b) == rec["bytes"] and hashlib.sha256(b).hexdigest() == rec["sha256"], rel
    pkg = os.path.join(root, "deep_quantized_recommendation_model_dqrm_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle._ref" not in src and "oracle/_ref" not in src and "from oracle" not in src, f


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (small workload, CPU only): one JSON line with the contract's keys; kind
    "reference" when oracle/_ref travelled (the unmodified reference, one Gloo rank), "port" otherwise.  Launched under
    torchrun with two ranks as the driver does for N > 1: rank 0 alone prints, the other exits 0."""
    import json, os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "small",
           "--steps", "2", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 2 and d["warmup"] == 1
    assert d["higher_is_better"] is True and d["unit"] == "samples/s" and d["value"] > 0
    assert d["config"]["global_batch"] == 256 and "small-shape" in d["config"]["workload"]
    have_ref = os.path.exists(os.path.join(root, "oracle", "_ref", "dlrm_s_pytorch_comm_grad.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
