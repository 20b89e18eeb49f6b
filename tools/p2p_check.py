"""Multi-GPU check of the NVLink exchange against the NCCL form (run under torchrun, N >= 2):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py
K graph-replayed steps of the Kaggle-shape-small model with DQRM_EXCHANGE=nccl, then the same with p2p (same seeds):
losses, every table and the MLP arena must be bit-identical between the two forms and between the ranks."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from deep_quantized_recommendation_model_dqrm_b200 import synthetic, dlrm_s_pytorch_comm_grad as drv, extend_distributed as ext
from deep_quantized_recommendation_model_dqrm_b200.graph_step import GraphedTrainStep

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
ext.init_distributed(rank=rank, local_rank=lrank, size=world, use_gpu=True, backend="nccl")
dev = torch.device("cuda", lrank); torch.cuda.set_device(dev)
cfg = dict(rows=[1460, 583, 1000000, 220000, 305, 24, 12517, 633, 3, 93145], dim=16, ln_bot=[13, 64, 16], ln_top_hidden=[64, 1])
ln_top = synthetic.top_mlp_sizes(len(cfg["rows"]), cfg["dim"], cfg["ln_top_hidden"])
B, K = 128, int(os.environ.get("STEPS", "12"))
policy = os.environ.get("POLICY", "full")


def run(backend):
    os.environ["DQRM_EXCHANGE"] = backend
    np.random.seed(123)
    m = drv.DLRM_Net(cfg["dim"], np.array(cfg["rows"]), np.array(cfg["ln_bot"]), np.array(ln_top), arch_interaction_op="dot",
                     sigmoid_top=len(ln_top) - 2, loss_function="bce", quantization_flag=True, embedding_bit=4, weight_bit=4,
                     quantize_act_and_lin=True, mlp_channelwise=True, device=dev, table_seed=77)
    m.shard_scan = True
    m._ensure_group().scale_policy = policy
    batches = [[t.to(dev) for t in synthetic.criteo_batch(cfg["rows"], B, seed=500 + 31 * rank + i, zipf=1.2 if i % 2 else None)]
               for i in range(4)]
    step = GraphedTrainStep(m, *batches[0], lr=0.3, world_size=world, rank=rank, grad_bits=8, warmup=2, use_graph=True)
    losses = []
    with torch.cuda.stream(step.stream):
        for i in range(K):
            step.load(*batches[i % 4]); step.run(); losses.append(step.loss.clone())
    torch.cuda.synchronize()
    m.emb_group.check_status()
    h = hashlib.sha256()
    for w in m.emb_group.weights:
        h.update(w.detach().cpu().numpy().tobytes())
    h.update(m._dense_arena.flat.cpu().numpy().tobytes())
    h.update(m.emb_group.scale.cpu().numpy().tobytes())
    used = "p2p" if m.emb_group.p2p is not None else "nccl"
    return h.hexdigest(), torch.stack(losses).cpu().numpy().tobytes(), used


res = {}
for backend in ("nccl", "p2p"):
    digest, losses, used = run(backend)
    assert used == backend, (used, backend)
    res[backend] = digest
    allh = [None] * world
    dist.all_gather_object(allh, digest)
    if rank == 0:
        print(f"{backend}: model digest {digest[:16]} identical across {world} ranks: {len(set(allh)) == 1}", flush=True)
    assert len(set(allh)) == 1, f"{backend}: replicas diverged"
ok = res["nccl"] == res["p2p"]       # same slots, same rank-ordered consumers: only the transport differs
if rank == 0:
    print(f"p2p == nccl (tables, MLP arena, scales bit-identical): {ok} (world {world})", flush=True)
dist.barrier(); torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if ok else 1)
