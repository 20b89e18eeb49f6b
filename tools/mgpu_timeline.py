"""Per-kernel timeline (start, duration, stream) of one graph-replayed training step on rank 0 of an N-GPU run.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/mgpu_timeline.py [--batch 128]"""
import argparse, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from deep_quantized_recommendation_model_dqrm_b200 import synthetic, dlrm_s_pytorch_comm_grad as drv, extend_distributed as ext
from deep_quantized_recommendation_model_dqrm_b200.graph_step import GraphedTrainStep
p = argparse.ArgumentParser(); p.add_argument("--batch", type=int, default=128); p.add_argument("--workload", default="kaggle")
a = p.parse_args()
rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
ext.init_distributed(rank=rank, local_rank=lrank, size=world, use_gpu=True, backend="nccl")
dev = torch.device("cuda", lrank); torch.cuda.set_device(dev)
cfg = {"kaggle": synthetic.KAGGLE, "terabyte": synthetic.TERABYTE}[a.workload]
ln_top = synthetic.top_mlp_sizes(len(cfg["rows"]), cfg["dim"], cfg["ln_top_hidden"])
np.random.seed(123)
m = drv.DLRM_Net(cfg["dim"], np.array(cfg["rows"]), np.array(cfg["ln_bot"]), np.array(ln_top), arch_interaction_op="dot",
                 sigmoid_top=len(ln_top) - 2, loss_function="bce", quantization_flag=True, embedding_bit=4, weight_bit=4,
                 quantize_act_and_lin=True, mlp_channelwise=True, device=dev)
m.shard_scan = world > 1
b = [t.to(dev) for t in synthetic.criteo_batch(cfg["rows"], a.batch, seed=3 + rank)]
step = GraphedTrainStep(m, *b, lr=0.1, world_size=world, rank=rank, use_graph=True)
with torch.cuda.stream(step.stream):
    for _ in range(6):
        step.run()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            step.run()
        torch.cuda.synchronize()
if rank == 0:
    f = tempfile.mktemp(suffix=".json"); prof.export_chrome_trace(f)
    ev = [e for e in json.load(open(f))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    scans = [e for e in ev if "table_absmax" in e["name"]]
    t0, t1 = scans[2]["ts"], scans[3]["ts"]
    print(f"# world {world} batch {a.batch}/GPU: step 3 of 4 on rank 0: start(us) dur(us) stream name; step = {t1 - t0:.1f} us")
    for e in ev:
        if t0 - 1 <= e["ts"] < t1 - 1:
            print(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f}  s{e['args'].get('stream', '?'):<4} {e['name'][:100]}")
if world > 1:
    dist.barrier()
torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
