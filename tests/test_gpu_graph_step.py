"""GPU: the graph-replayed training step (graph_step.GraphedTrainStep: packed single-copy inputs, fused BCE loss +
gradient kernel, CUDA graphs) against the eager reference-shaped iteration (drv.train_iteration: ATen BCELoss,
autograd) that the oracle parity tests pin."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [1, 37, 128, 2048, 8192])
def test_bce_loss_grad_kernel_matches_aten(n):
    from deep_quantized_recommendation_model_dqrm_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(n)
    z = torch.rand((n, 1), device="cuda", generator=g)
    if n > 4:
        z[0], z[1], z[2] = 0.0, 1.0, 1e-30                     # the clamps: log -> -100, denominator -> 1e-12
    t = torch.round(torch.rand((n, 1), device="cuda", generator=g))
    zr = z.clone().requires_grad_(True)
    E = torch.nn.functional.binary_cross_entropy(zr, t)
    E.backward()
    loss, dz = torch.zeros((), device="cuda"), torch.empty_like(z)
    _lib.check(lib.dqrm_bce_loss_grad(z.data_ptr(), t.data_ptr(), n, loss.data_ptr(), dz.data_ptr(), _lib.stream_ptr()), "bce")
    torch.testing.assert_close(loss, E.detach(), rtol=1e-6, atol=1e-7)        # tolerance: summation order only
    assert torch.equal(dz, zr.grad)                                            # same operation order as ATen: bit-exact


@pytest.mark.parametrize("use_graph", [True, False, "overlap"])
def test_graph_step_matches_eager_iteration(use_graph, monkeypatch):
    """"overlap": the bottom MLP + weight fake-quantisation replayed as their own graph on a second stream beside the
    table scan (forced here; by default only when the scan is long enough to hide them) -- same bits."""
    from helpers import C_SMALL, build_cuda_model
    if use_graph == "overlap":
        monkeypatch.setenv("DQRM_OVERLAP_BOTTOM", "force")
        use_graph = True
    else:
        monkeypatch.setenv("DQRM_OVERLAP_BOTTOM", "0")
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic
    from deep_quantized_recommendation_model_dqrm_b200 import dlrm_s_pytorch_comm_grad as drv
    from deep_quantized_recommendation_model_dqrm_b200.graph_step import GraphedTrainStep
    B = 32
    batches = [synthetic.criteo_batch(C_SMALL["rows"], B, seed=40 + i, zipf=1.3 if i % 2 else None) for i in range(5)]
    ref = build_cuda_model(C_SMALL, seed=9)
    ref_losses = [float(drv.train_iteration(ref, *b, lr=0.2)) for b in batches]
    ref._ensure_group().check_status()

    m = build_cuda_model(C_SMALL, seed=9)
    snapshot = {k: v.detach().clone() for k, v in m.named_parameters()}
    step = GraphedTrainStep(m, *batches[0], lr=0.2, warmup=1, use_graph=use_graph)   # warm-up iterations train: rewind
    assert (step.graph_pre is not None) == (os.environ["DQRM_OVERLAP_BOTTOM"] == "force")
    with torch.no_grad():
        for k, v in m.named_parameters():
            v.copy_(snapshot[k])
    m.emb_group.scale_valid = False
    losses = []
    with torch.cuda.stream(step.stream):
        for i, b in enumerate(batches):
            if i % 2:
                step.load_packed(step.pack_host(*b))             # one pinned H2D copy
            else:
                step.load(*b)
            step.run()
            losses.append(float(step.loss))
    torch.cuda.synchronize()
    m.emb_group.check_status()
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-6)
    for a, b in zip(ref.emb_group.weights, m.emb_group.weights):
        assert torch.equal(a, b)                                 # identical gradients -> identical tables
    for (na, pa), (nb, pb) in zip(ref.named_parameters(), m.named_parameters()):
        assert torch.equal(pa, pb), na


def test_graph_replay_follows_an_lr_schedule():
    """The learning rate is read from device memory by the update kernels: GraphedTrainStep.set_lr() between replays
    of ONE captured graph must give the same model as the eager iteration called with each step's lr
    (LRPolicyScheduler warm-up + decay, dlrm_s_pytorch_comm_grad.py:221-255)."""
    import numpy as np
    from helpers import C_SMALL, build_cuda_model
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic
    from deep_quantized_recommendation_model_dqrm_b200 import dlrm_s_pytorch_comm_grad as drv
    from deep_quantized_recommendation_model_dqrm_b200.graph_step import GraphedTrainStep
    ma, mb = build_cuda_model(C_SMALL, seed=31), build_cuda_model(C_SMALL, seed=31)
    sched = drv.LRPolicyScheduler(0.4, num_warmup_steps=3, decay_start_step=4, num_decay_steps=4)
    b0 = [t.cuda() for t in synthetic.criteo_batch(C_SMALL["rows"], 32, seed=900)]
    lrs = []
    for _ in range(8):
        lrs.append(sched.get_lr()[-1])
        sched.step()
    assert len(set(lrs)) > 3
    snapshot = {k: v.detach().clone() for k, v in mb.named_parameters()}
    step = GraphedTrainStep(mb, *b0, lr=lrs[0], warmup=1, use_graph=True)     # warm-up iterations train: rewind
    with torch.no_grad():
        for k, v in mb.named_parameters():
            v.copy_(snapshot[k])
    mb.emb_group.scale_valid = False
    for i, lr in enumerate(lrs):
        bt = synthetic.criteo_batch(C_SMALL["rows"], 32, seed=901 + i, zipf=1.2)
        la = drv.train_iteration(ma, *bt, lr=lr)
        with torch.cuda.stream(step.stream):
            step.set_lr(lr)
            step.load(*[t.cuda() for t in bt])
            step.run()
        torch.cuda.synchronize()
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.equal(pa, pb), i
