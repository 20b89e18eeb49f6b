"""GPU: the one-kernel NVLink exchange (csrc/p2p.cu) on ONE device -- W peer arenas inside this process, the W
"ranks" on W streams (the kernels spin on each other's flags, so they must be co-resident; each is <= 64 small CTAs).
Real multi-GPU coverage: tools/p2p_check.py under torchrun (bit-identical to the NCCL form, ranks identical)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _arenas(sites, world):
    from deep_quantized_recommendation_model_dqrm_b200 import p2p
    return p2p.PeerArena.local_group(sites, world, device="cuda")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_allgather_protocol_replays(world):
    sites = {"small": 26 * 4, "big": 66672, "odd": 1624 * 4 + 4}
    arenas = _arenas(sites, world)
    status = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    g = torch.Generator(device="cuda").manual_seed(5)
    for it in range(5):                                  # sequence numbers advance on the device
        want = {}
        for name in sites:
            want[name] = []
            for r, a in enumerate(arenas):
                mine = a.my_slot(name)
                mine.copy_(torch.randint(0, 255, mine.shape, device="cuda", dtype=torch.uint8, generator=g))
                want[name].append(mine.clone())
        torch.cuda.synchronize()
        for name in sites:                               # >= 2 sites per round, same order on every rank
            for r, a in enumerate(arenas):
                with torch.cuda.stream(streams[r]):
                    a.allgather(name, status[r])
        torch.cuda.synchronize()
        for name in sites:
            for a in arenas:
                got = a.slots(name)
                for r in range(world):
                    assert torch.equal(got[r], want[name][r]), (it, name, r)
        assert all(int(s) == 0 for s in status)


def test_allgather_gives_up_when_a_peer_is_missing():
    from deep_quantized_recommendation_model_dqrm_b200 import _lib
    arenas = _arenas({"x": 64, "y": 64}, 2)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    arenas[0].allgather("x", status)                     # rank 1 never calls: DQRM_P2P_TIMEOUT_S, then the timeout bit
    torch.cuda.synchronize()
    assert int(status) == _lib.STATUS_P2P_TIMEOUT
    # the bit is fatal and sticky: the consumers that would apply the (stale) slots are no-ops while it is set ...
    lib = _lib.load()
    param = torch.ones(8, device="cuda")
    codes = torch.full((2, 16), 5, dtype=torch.int8, device="cuda")
    chan = torch.tensor([0, 8], dtype=torch.int64, device="cuda")
    mean = torch.ones(1, device="cuda")
    _lib.check(lib.dqrm_dense_apply_gathered(param.data_ptr(), codes.data_ptr(), 16, 2, chan.data_ptr(), 1, mean.data_ptr(), 0.1,
                                             None, None, None, status.data_ptr(), _lib.stream_ptr()), "apply_g")
    torch.cuda.synchronize()
    assert torch.equal(param, torch.ones(8, device="cuda"))
    # ... and the host-side poll raises without clearing it
    from deep_quantized_recommendation_model_dqrm_b200 import p2p, tables
    g = tables.EmbeddingTableGroup([torch.zeros((4, 16), device="cuda")])
    g.status = status
    for _ in range(2):
        with pytest.raises(p2p.ExchangeTimeout):
            g.check_status()
    status.zero_()
    _lib.check(lib.dqrm_dense_apply_gathered(param.data_ptr(), codes.data_ptr(), 16, 2, chan.data_ptr(), 1, mean.data_ptr(), 0.1,
                                             None, None, None, status.data_ptr(), _lib.stream_ptr()), "apply_g")
    torch.cuda.synchronize()
    assert torch.equal(param, torch.full((8,), 1.0 - 0.1 * 5.0, device="cuda"))


@pytest.mark.parametrize("world", [1, 2, 8])
def test_gathered_dense_consumers_match_allreduce_form(world):
    """quant(gathered scales) + apply(gathered int8 codes) == the all-reduce form: sum of scales in rank order,
    fp32 sum of the integer codes, dqrm_dense_apply -- parameters bit-identical."""
    from deep_quantized_recommendation_model_dqrm_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(2)
    sizes = [(64, 13), (64,), (16, 64), (16,), (1, 16), (1,)]
    chan, off = [0], 0
    for s in sizes:
        n = int(np.prod(s))
        if len(s) == 2:
            chan += [off + (r + 1) * s[1] for r in range(s[0])]
        else:
            chan.append(off + n)
        off += n
    total, C = off, len(chan) - 1
    chan_t = torch.tensor(chan, dtype=torch.int64, device="cuda")
    grads = [torch.tensor(rng.randn(total).astype(np.float32) * 10.0 ** rng.uniform(-4, 0), device="cuda") for _ in range(world)]
    param0 = torch.tensor(rng.randn(total).astype(np.float32), device="cuda")
    st = _lib.stream_ptr()
    scales = torch.zeros((world, C + 3), dtype=torch.float32, device="cuda")          # padded stride
    for r in range(world):
        _lib.check(lib.dqrm_dense_grad_scale(grads[r].data_ptr(), None, chan_t.data_ptr(), C, 8, scales[r].data_ptr(), st), "scale")
    # all-reduce form
    ssum = scales[0, :C].clone()
    for r in range(1, world):
        ssum = ssum + scales[r, :C]
    codes_f = torch.zeros((world, total), device="cuda")
    mean_a = torch.zeros(C, device="cuda")
    for r in range(world):
        _lib.check(lib.dqrm_dense_grad_quant(grads[r].data_ptr(), chan_t.data_ptr(), C, ssum.data_ptr(), 1.0 / world, 8,
                                             codes_f[r].data_ptr(), mean_a.data_ptr(), st), "quant")
    csum = codes_f[0].clone()
    for r in range(1, world):
        csum = csum + codes_f[r]
    pa = param0.clone()
    _lib.check(lib.dqrm_dense_apply(pa.data_ptr(), csum.data_ptr(), chan_t.data_ptr(), C, mean_a.data_ptr(), 1.0 / world, 0.1, None, None, None, st), "apply")
    # gathered form
    stride = (total + 15) // 16 * 16 + 16
    codes_i = torch.zeros((world, stride), dtype=torch.int8, device="cuda")
    mean_b = torch.zeros(C, device="cuda")
    for r in range(world):
        _lib.check(lib.dqrm_dense_grad_quant_gathered(grads[r].data_ptr(), chan_t.data_ptr(), C, scales.data_ptr(),
                                                      scales.stride(0), world, 8, codes_i[r].data_ptr(), mean_b.data_ptr(), st), "quant_g")
    pb = param0.clone()
    _lib.check(lib.dqrm_dense_apply_gathered(pb.data_ptr(), codes_i.data_ptr(), stride, world, chan_t.data_ptr(), C,
                                             mean_b.data_ptr(), 0.1, None, None, None, None, st), "apply_g")
    torch.cuda.synchronize()
    assert torch.equal(mean_a, mean_b)
    assert torch.equal(codes_i[:, :total].float(), codes_f)
    assert torch.equal(pa, pb)


def test_scale_from_gathered_absmax():
    from deep_quantized_recommendation_model_dqrm_b200 import _lib
    lib = _lib.load()
    world, T = 8, 26
    g = torch.rand((world, T + 6), device="cuda")
    g[3, 5] = 0.0
    am, sc, inv = (torch.zeros(T, device="cuda") for _ in range(3))
    _lib.check(lib.dqrm_scale_from_absmax_gathered(T, g.data_ptr(), g.stride(0), world, 4, am.data_ptr(), sc.data_ptr(),
                                                   inv.data_ptr(), _lib.stream_ptr()), "gathered")
    want = g[:, :T].max(dim=0)[0]
    sc2, inv2 = torch.zeros(T, device="cuda"), torch.zeros(T, device="cuda")
    _lib.check(lib.dqrm_scale_from_absmax(T, want.data_ptr(), 4, sc2.data_ptr(), inv2.data_ptr(), _lib.stream_ptr()), "plain")
    assert torch.equal(am, want) and torch.equal(sc, sc2) and torch.equal(inv, inv2)


@pytest.mark.parametrize("world,num_ctas,ec,split", [(1, 3, False, False), (2, 5, False, False), (4, 7, True, False),
                                                     (8, 4, False, False), (4, 6, True, True)])
def test_one_kernel_dense_exchange_matches_the_five_launch_form(world, num_ctas, ec, split):
    """dqrm_dense_exchange_apply (csrc/dense_xchg.cu: CTA b exchanges with CTA b of the peers, scales and int8 codes travel as
    {payload, sequence} words stored straight into the peers' arenas) == dense_grad_scale -> gather -> dense_grad_quant_gathered -> gather -> dense_apply_gathered:
    parameters, mean scales and error-compensation residuals bit-identical, over several replays (device-side
    sequence numbers) -- W arenas in this process, the W ranks on W streams.  split: the CTAs of the top layers and of
    the bottom layers as two launches (the early / late buckets of DenseArena.after_dw)."""
    import ctypes as C
    from deep_quantized_recommendation_model_dqrm_b200 import _lib, p2p
    from deep_quantized_recommendation_model_dqrm_b200.dense import DenseArena
    lib = _lib.load()
    rng = np.random.RandomState(11 + world)
    sizes = [(64, 13), (64,), (48, 37), (48,), (16, 479), (16,), (1, 16), (1,)]       # odd row lengths: unaligned runs
    chan, off = [0], 0
    for s in sizes:
        n = int(np.prod(s))
        if len(s) == 2:
            chan += [off + (r + 1) * s[1] for r in range(s[0])]
        else:
            chan.append(off + n)
        off += n
    total, nch = off, len(chan) - 1
    chan_t = torch.tensor(chan, dtype=torch.int64, device="cuda")
    plan = DenseArena.__new__(DenseArena)
    plan.chan_begin, plan.num_chan = chan_t, nch
    part = plan.xchg_partition(num_ctas, split_chan=64 + 1 + 48 + 1 if split else None)
    assert (part["split_cta"] > 0) == split
    G, elems, chans = len(part["cta_chan"]) - 1, part["elems"], part["chans"]
    cta_chan = torch.tensor(part["cta_chan"], dtype=torch.int32, device="cuda")
    cta_word = torch.tensor(part["cta_word"], dtype=torch.int32, device="cuda")
    arenas = p2p.PeerArena.local_group({"mlp_xscale": nch * 8, "mlp_xcodes": part["cta_word"][-1] * 8}, world)
    streams = [torch.cuda.Stream() for _ in range(world)]
    status = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
    seq = [torch.zeros(G, dtype=torch.int32, device="cuda") for _ in range(world)]
    param0 = torch.tensor(rng.randn(total).astype(np.float32), device="cuda")
    pa = [param0.clone() for _ in range(world)]                     # five-launch form, per rank
    pb = [param0.clone() for _ in range(world)]                     # one-kernel form
    eca = [torch.zeros(total, device="cuda") for _ in range(world)] if ec else [None] * world
    ecb = [torch.zeros(total, device="cuda") for _ in range(world)] if ec else [None] * world
    mean_a = [torch.zeros(nch, device="cuda") for _ in range(world)]
    mean_b = [torch.zeros(nch, device="cuda") for _ in range(world)]
    st = _lib.stream_ptr
    for it in range(4):
        grads = [torch.tensor(rng.randn(total).astype(np.float32) * 10.0 ** rng.uniform(-4, 0), device="cuda")
                 for _ in range(world)]
        # ---- reference: the five-launch form on plain buffers
        ga = [g.clone() for g in grads]
        scales = torch.zeros((world, nch + 3), dtype=torch.float32, device="cuda")
        stride = (total + 15) // 16 * 16 + 16
        codes = torch.zeros((world, stride), dtype=torch.int8, device="cuda")
        for r in range(world):
            _lib.check(lib.dqrm_dense_grad_scale(ga[r].data_ptr(), _lib.ptr(eca[r]), chan_t.data_ptr(), nch, 8,
                                                 scales[r].data_ptr(), st()), "scale")
        for r in range(world):
            _lib.check(lib.dqrm_dense_grad_quant_gathered(ga[r].data_ptr(), chan_t.data_ptr(), nch, scales.data_ptr(),
                                                          scales.stride(0), world, 8, codes[r].data_ptr(),
                                                          mean_a[r].data_ptr(), st()), "quant_g")
        for r in range(world):
            _lib.check(lib.dqrm_dense_apply_gathered(pa[r].data_ptr(), codes.data_ptr(), stride, world, chan_t.data_ptr(), nch,
                                                     mean_a[r].data_ptr(), 0.1, None, ga[r].data_ptr() if ec else None,
                                                     _lib.ptr(eca[r]), None, st()), "apply_g")
        torch.cuda.synchronize()
        # ---- one kernel per rank, the ranks on their own streams (they wait for each other: co-resident)
        gb = [g.clone() for g in grads]
        torch.cuda.synchronize()
        for r, a in enumerate(arenas):
            sc, co = a.sites["mlp_xscale"], a.sites["mlp_xcodes"]
            k = part["split_cta"]
            with torch.cuda.stream(streams[r]):
                for b0, b1 in ([(k, G), (0, k)] if split else [(0, G)]):
                    rc = lib.dqrm_dense_exchange_apply(a.ptrs, world, r, sc["data_off"], sc["stride"], co["data_off"],
                                                       co["stride"], pb[r].data_ptr(), gb[r].data_ptr(), _lib.ptr(ecb[r]),
                                                       chan_t.data_ptr(), cta_chan.data_ptr() + 4 * b0,
                                                       cta_word.data_ptr() + 4 * b0, b1 - b0, elems, chans, 8,
                                                       mean_b[r].data_ptr(), seq[r].data_ptr() + 4 * b0, 0.1, None,
                                                       status[r].data_ptr(), st())
                    _lib.check(rc, "dqrm_dense_exchange_apply")
        torch.cuda.synchronize()
        assert all(int(s) == 0 for s in status)
        for r in range(world):
            assert torch.equal(mean_a[r], mean_b[r]), (it, r)
            assert torch.equal(pa[r], pb[r]), (it, r)
            if ec:
                assert torch.equal(eca[r], ecb[r]) and torch.equal(ga[r], gb[r]), (it, r)
        assert all(torch.equal(pb[0], p) for p in pb[1:])
        assert all(int(s[0]) == it + 1 for s in seq)
