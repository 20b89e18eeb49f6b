"""Per-kernel timeline (start, duration, stream) of one pipelined training step, from a torch.profiler
chrome trace: shows what actually overlaps with the table rescan.
    python tools/timeline.py [--workload kaggle] [--batch 128] [--policy pipelined]"""
import argparse, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from deep_quantized_recommendation_model_dqrm_b200 import synthetic, dlrm_s_pytorch_comm_grad as drv
from deep_quantized_recommendation_model_dqrm_b200.graph_step import GraphedTrainStep

p = argparse.ArgumentParser()
p.add_argument("--workload", default="kaggle"); p.add_argument("--batch", type=int, default=128)
p.add_argument("--policy", default="pipelined"); p.add_argument("--no-graph", action="store_true")
p.add_argument("--dot", default=None, help="dump graph A as DOT to this path"); p.add_argument("--no-fuse-mlp", action="store_true"); p.add_argument("--rows", type=int, default=1000)
a = p.parse_args()
cfg = {"kaggle": synthetic.KAGGLE, "terabyte": synthetic.TERABYTE, "small": synthetic.RANDOM_SMALL}[a.workload]
ln_top = synthetic.top_mlp_sizes(len(cfg["rows"]), cfg["dim"], cfg["ln_top_hidden"])
np.random.seed(123)
m = drv.DLRM_Net(cfg["dim"], np.array(cfg["rows"]), np.array(cfg["ln_bot"]), np.array(ln_top), arch_interaction_op="dot",
                 sigmoid_top=len(ln_top) - 2, loss_function="bce", quantization_flag=True, embedding_bit=4, weight_bit=4,
                 quantize_act_and_lin=True, mlp_channelwise=True, device="cuda")
m._ensure_group().scale_policy = a.policy
m.fuse_mlp = not a.no_fuse_mlp
b = [t.cuda() for t in synthetic.criteo_batch(cfg["rows"], a.batch, seed=3)]
if a.dot:
    _orig = torch.cuda.CUDAGraph.capture_begin
    def _cb(self, *x, **k):
        self.enable_debug_mode()
        return _orig(self, *x, **k)
    torch.cuda.CUDAGraph.capture_begin = _cb
step = GraphedTrainStep(m, *b, lr=0.1, use_graph=not a.no_graph)
if a.dot and step.graph is not None:
    step.graph.debug_dump(a.dot)
with torch.cuda.stream(step.stream):
    for _ in range(5):
        step.run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            step.run()
        torch.cuda.synchronize()
f = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(f)
ev = [e for e in json.load(open(f))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
scans = [e for e in ev if "blockmax_scan" in e["name"] or "table_absmax" in e["name"]]
t0 = scans[2]["ts"] if len(scans) > 2 else ev[0]["ts"]
t1 = scans[3]["ts"] if len(scans) > 3 else ev[-1]["ts"] + 1
print(f"# step 3 of 4: start(us) dur(us) stream name")
n = 0
for e in ev:
    if t0 - 5 <= e["ts"] < t1 - 5 and n < a.rows:
        n += 1
        print(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f}  s{e['args'].get('stream', '?'):<4} {e['name'][:90]}")
