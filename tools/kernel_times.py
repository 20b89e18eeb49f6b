"""Warm per-kernel GPU durations of one training step (torch.profiler / CUPTI), eager launches.
    python tools/kernel_times.py [--workload kaggle] [--batch 128] [--steps 10]
ncu's per-launch list is cold-cache and serialised; this is the warm steady-state view."""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from deep_quantized_recommendation_model_dqrm_b200 import synthetic, dlrm_s_pytorch_comm_grad as drv
from deep_quantized_recommendation_model_dqrm_b200.graph_step import GraphedTrainStep

p = argparse.ArgumentParser()
p.add_argument("--workload", default="kaggle"); p.add_argument("--batch", type=int, default=128)
p.add_argument("--steps", type=int, default=10); p.add_argument("--no-fuse-mlp", action="store_true")
a = p.parse_args()
cfg = {"kaggle": synthetic.KAGGLE, "terabyte": synthetic.TERABYTE, "small": synthetic.RANDOM_SMALL}[a.workload]
ln_top = synthetic.top_mlp_sizes(len(cfg["rows"]), cfg["dim"], cfg["ln_top_hidden"])
np.random.seed(123)
m = drv.DLRM_Net(cfg["dim"], np.array(cfg["rows"]), np.array(cfg["ln_bot"]), np.array(ln_top), arch_interaction_op="dot",
                 sigmoid_top=len(ln_top) - 2, loss_function="bce", quantization_flag=True, embedding_bit=4, weight_bit=4,
                 quantize_act_and_lin=True, mlp_channelwise=True, device="cuda")
m.fuse_mlp = not a.no_fuse_mlp
b = [t.cuda() for t in synthetic.criteo_batch(cfg["rows"], a.batch, seed=3)]
step = GraphedTrainStep(m, *b, lr=0.1, use_graph=False)
for _ in range(3):
    step.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        step.run()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for e in prof.events():
    if e.device_type.name != "CUDA":
        continue
    k = agg.setdefault(e.name[:100], [0, 0.0])
    k[0] += 1; k[1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"# {a.workload} batch {a.batch}: {tot / a.steps:.1f} us of kernel time per step, {sum(v[0] for v in agg.values()) / a.steps:.0f} launches per step")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{t / a.steps:9.1f} us/step {c / a.steps:5.1f}x {t / c:8.1f} us/launch  {n}")
