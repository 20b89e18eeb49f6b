"""Where one launch of the one-kernel dense exchange (csrc/dense_xchg.cu) spends its time: per-CTA %globaltimer stamps of
graph-replayed steps on every rank (torchrun, N >= 2).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/xchg_phases.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from deep_quantized_recommendation_model_dqrm_b200 import _lib, synthetic, dlrm_s_pytorch_comm_grad as drv, extend_distributed as ext
from deep_quantized_recommendation_model_dqrm_b200.graph_step import GraphedTrainStep
rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
ext.init_distributed(rank=rank, local_rank=lrank, size=world, use_gpu=True, backend="nccl")
dev = torch.device("cuda", lrank); torch.cuda.set_device(dev)
cfg = synthetic.KAGGLE
ln_top = synthetic.top_mlp_sizes(len(cfg["rows"]), cfg["dim"], cfg["ln_top_hidden"])
np.random.seed(123)
m = drv.DLRM_Net(cfg["dim"], np.array(cfg["rows"]), np.array(cfg["ln_bot"]), np.array(ln_top), arch_interaction_op="dot",
                 sigmoid_top=len(ln_top) - 2, loss_function="bce", quantization_flag=True, embedding_bit=4, weight_bit=4,
                 quantize_act_and_lin=True, mlp_channelwise=True, device=dev)
m.shard_scan = world > 1
b = [t.to(dev) for t in synthetic.criteo_batch(cfg["rows"], 128, seed=3 + rank)]
stamps = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
_lib.load().dqrm_dense_exchange_debug(stamps.data_ptr())
step = GraphedTrainStep(m, *b, lr=0.1, world_size=world, rank=rank, use_graph=True)
G = m._dense_arena._xchg_plan["num_ctas"]
with torch.cuda.stream(step.stream):
    for _ in range(10):
        step.run()
    torch.cuda.synchronize()
    dist.barrier()
    for _ in range(5):
        step.run()
    torch.cuda.synchronize()
s = stamps.cpu().numpy().reshape(148, 8)[:G, :6].astype(np.float64)
t0 = s[:, 0].min()
names = ["launch skew (CTA start - first CTA)", "1 load + absmax + scale stores", "1 signal + wait", "2 mean scale + quantise + code stores",
         "2 signal + wait", "3 sum codes + update"]
out = [f"rank {rank}: {G} CTAs, last launch of 5 replays; per-CTA phase durations in us (min / median / max); kernel = {(s[:, 5].max() - t0) / 1e3:.1f} us"]
d = np.concatenate([(s[:, :1] - t0), np.diff(s, axis=1)], axis=1) / 1e3
for k, n in enumerate(names):
    out.append(f"  {n:42s} {d[:, k].min():7.1f} {np.median(d[:, k]):7.1f} {d[:, k].max():7.1f}")
allo = [None] * world
dist.all_gather_object(allo, "\n".join(out))
if rank == 0:
    print("\n".join(allo), flush=True)
dist.barrier(); torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
