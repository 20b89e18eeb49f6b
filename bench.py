#!/usr/bin/env python
"""bench.py -- DQRM hot-path benchmark (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

Metric (BASELINE.json): train samples/sec, Criteo-Kaggle shape (26 tables at the Kaggle cardinalities,
dim 16, bot 13-512-256-64-16, top 512-256-1), INT4 embedding + MLP QAT, INT8 quantised sparse gradient
exchange, batch 128 per GPU, synthetic data.  A "step" is one full iteration of the reference hot loop
(dlrm_s_pytorch_comm_grad.py:1909-1957): scale scan of all tables, forward, loss, backward with row
de-duplication, quantised exchange, SGD update of tables and MLPs.

  value    samples/s with the step's inputs already resident in HBM (device-to-device refill of the
           static buffers + scan launch + CUDA-graph replay), CUDA events, max over ranks.
  e2e      same step through the public API with HOST (pinned) inputs: H2D copy of X/lS_o/lS_i/T and a
           D2H read of the loss inside every timed step.
  roofline the dominant kernel is the table max-abs scan (HBM-bound): algorithmic bytes = sum(rows)*D*4
           per launch (/N when row-sharded), duration = CUDA events around every scan launch inside the
           timed region.
  cpu_baseline / --impl reference: the oracle port of the reference's CPU path (torch CPU ops in the
           reference's order), all host threads, same config.
Multi-GPU: one process per GPU (torchrun), batch-sharded DP, weak scaling (128 samples per GPU); the
table scan is row-sharded 1/N per rank + MAX all-reduce (replicas are bit-identical).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

PER_GPU_BATCH = 128
LR = 0.1
METRIC = "train samples/sec (Criteo-Kaggle shape, INT4 emb+MLP QAT, INT8 sparse grad exchange)"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", type=str, default="kaggle", choices=["kaggle", "terabyte", "small"])
    p.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch")
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-steps", type=int, default=5)
    p.add_argument("--scale-policy", type=str, default="full", choices=["full", "pipelined", "incremental"],
                   help="full = rescan every table in front of every forward (reference order; headline); pipelined = "
                        "the same rescan overlapped with the step on a side stream; incremental = exact block-max "
                        "tracker (reads only the touched blocks)")
    p.add_argument("--no-extras", "--no-incremental-extra", dest="no_extras", action="store_true",
                   help="skip the additional measurements (serial rescan, incremental tracker) reported beside the headline")
    return p.parse_args()


def workload_cfg(name):
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic
    return {"kaggle": synthetic.KAGGLE, "terabyte": synthetic.TERABYTE, "small": synthetic.RANDOM_SMALL}[name]


def mlp_sizes(cfg):
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic
    return synthetic.top_mlp_sizes(len(cfg["rows"]), cfg["dim"], cfg["ln_top_hidden"])


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md).  The timed region is tens
    of milliseconds, too short for `nvidia-smi -lms`, so NVML is polled from a thread every ~2 ms."""

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.stop_flag, self.th, self.err = gpu_index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES if it remaps them
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:                                    # pragma: no cover
            self.nv, self.err = None, repr(e)

    def start(self):
        if self.nv is None:
            return
        nv = self.nv

        def pump():
            while not self.stop_flag:
                try:
                    mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    self.samples.append((time.perf_counter(), float(mhz), int(reasons)))
                except Exception as e:                            # pragma: no cover
                    self.err = repr(e)
                    return
                time.sleep(0.002)
        self.th = threading.Thread(target=pump, daemon=True)
        self.th.start()

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.th is not None:
            self.th.join(timeout=1.0)
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable: %s" % self.err]}
        nv = self.nv
        sel = [s for s in self.samples if (t0 is None or s[0] >= t0) and (t1 is None or s[0] <= t1)] or self.samples
        bits = 0
        for s in sel:
            bits |= s[2]
        names = []
        for nm, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                         ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                         ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                         ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
            flag = getattr(nv, attr, None)
            if flag is None:
                flag = getattr(nv, attr.replace("ClocksEventReason", "ClocksThrottleReason"), 0)
            if bits & flag:
                names.append(nm)
        return {"sm_mhz": float(np.median([s[1] for s in sel])), "sm_max_mhz": self.max_mhz, "reasons": names,
                "samples": len(sel)}


# --------------------------------------------------------------------------------------------
def oracle_arm(cfg, batch, steps, warmup):
    """The reference's CPU path restated (oracle/dqrm_oracle.py, torch CPU ops in the reference's op
    order), all host threads.  Returns (samples_per_s, ms_per_step, cores, build_s)."""
    from deep_quantized_recommendation_model_dqrm_b200 import synthetic
    from oracle import dqrm_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    rng = np.random.RandomState(123)
    emb = []
    for k, n in enumerate(cfg["rows"]):
        w = torch.empty((n, cfg["dim"]), dtype=torch.float32)
        synthetic.table_weights_(w, k, 1234)                 # same law as the GPU arm; host generator
        emb.append(w)
    bot = synthetic.mlp_params(cfg["ln_bot"], rng)
    top = synthetic.mlp_params(mlp_sizes(cfg), rng)
    model = O.OracleDLRM(cfg["rows"], cfg["dim"], bot, top, emb_weights=emb)
    del emb
    build_s = time.perf_counter() - t0
    times = []
    for s in range(warmup + steps):
        b = synthetic.criteo_batch(cfg["rows"], batch, seed=1000 + s)
        t1 = time.perf_counter()
        O.train_step_torch([model], [b], lr=LR)
        if s >= warmup:
            times.append(time.perf_counter() - t1)
    ms = 1000.0 * float(np.mean(times))
    return batch / (ms / 1000.0), ms, torch.get_num_threads(), build_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_cfg(args.workload)
    gbatch = args.batch * args.gpus
    steps = min(args.steps, 40)                               # ~1.4 s per step at Kaggle shape on 8 cores
    warm = min(args.warmup, 1)
    val, ms, cores, build_s = oracle_arm(cfg, gbatch, steps, warm)
    sample = (f"{steps} full train steps (of --steps {args.steps}) after {warm} warm-up, global batch {gbatch}, "
              f"oracle port of the reference CPU path, single process, {cores} threads; tables built in {build_s:.1f}s")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}-shape DQRM, {len(cfg['rows'])} tables, dim {cfg['dim']}, "
                                   f"batch {args.batch}/GPU x {args.gpus}", "global_batch": gbatch,
                       "parallelism": "cpu-1proc"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def exchange_microbench(dlrm, world, barrier, iters=20):
    """The two large exchange sites (embedding slots, MLP int8 codes) timed alone, outside the graph, with CUDA
    events: bus GB/s = payload * (N-1) / time per rank, against 900 GB/s per direction.  Runs after all step
    timings; every rank takes part (the kernels wait for each other).  Reported, never fatal."""
    try:
        g, d = dlrm.emb_group, dlrm._dense_arena
        if g.p2p is None or d.p2p is None:
            return {"transport": "nccl", "note": "NVLink peer arenas not in use"}
        sites = (("emb_slot", g.p2p, g.status), ("mlp_codes", d.p2p, d.status))

        def round_trip():
            for name, arena, status in sites:
                arena.allgather(name, status)
        for _ in range(3):
            round_trip()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            round_trip()
        e1.record()
        barrier()
        us = 1000.0 * e0.elapsed_time(e1) / iters
        payload = sum(arena.sites[name]["stride"] for name, arena, _ in sites)
        return {"transport": "nvlink peer memory", "sites": [n for n, _, _ in sites], "bytes_per_rank": payload,
                "us_per_pair_of_allgathers": us, "bus_gbs": payload * (world - 1) / (us * 1e-6) / 1e9,
                "nominal_gbs_per_direction": 900.0,
                "note": "two back-to-back one-kernel all-gathers incl. launch and flag round trip; latency-bound at these sizes"}
    except Exception as e:                                        # pragma: no cover
        return {"error": repr(e)}


def finish(world):
    """Multi-rank exit.  destroy_process_group() with NCCL collectives captured in live CUDA graphs was seen
    to hang at N=8 (the processes never exited); leave the communicator to process teardown and exit hard
    once every rank has passed the last barrier."""
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_ours(args):
    import torch.distributed as dist
    from deep_quantized_recommendation_model_dqrm_b200 import _lib, synthetic
    from deep_quantized_recommendation_model_dqrm_b200 import dlrm_s_pytorch_comm_grad as drv
    from deep_quantized_recommendation_model_dqrm_b200 import extend_distributed as ext_dist
    from deep_quantized_recommendation_model_dqrm_b200.graph_step import GraphedTrainStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    ext_dist.init_distributed(rank=rank, local_rank=local_rank, size=world, use_gpu=True, backend="nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.load()

    cfg = workload_cfg(args.workload)
    ln_top = mlp_sizes(cfg)
    B = args.batch
    np.random.seed(123)                                       # identical MLP init on every rank
    dlrm = drv.DLRM_Net(cfg["dim"], np.array(cfg["rows"]), np.array(cfg["ln_bot"]), np.array(ln_top),
                        arch_interaction_op="dot", sigmoid_bot=-1, sigmoid_top=len(ln_top) - 2, loss_function="bce",
                        quantization_flag=True, embedding_bit=4, weight_bit=4, quantize_act_and_lin=True,
                        mlp_channelwise=True, quantize_activation=False, device=dev, table_seed=1234)
    dlrm.shard_scan = world > 1
    table_bytes = sum(cfg["rows"]) * cfg["dim"] * 4

    # a pool of distinct local batches, pinned on the host and resident on the device
    pool = 8
    host, devb = [], []
    for i in range(pool):
        X, lS_o, lS_i, T = synthetic.criteo_batch(cfg["rows"], B, seed=10_000 + 97 * rank + i)
        hb = tuple(t.pin_memory() for t in (X, lS_o, lS_i, T))
        host.append(hb)
        devb.append(tuple(t.to(dev) for t in hb))
    dlrm._ensure_group().scale_policy = args.scale_policy
    step = GraphedTrainStep(dlrm, *devb[0], lr=LR, world_size=world, rank=rank, grad_bits=8, warmup=3,
                            use_graph=not args.no_graph)
    n0 = _lib.total_launches()                                # count OUR kernel launches of one step: one more
    with torch.cuda.stream(step.stream):
        step.scan(); step._body()                             # eager iteration (the graphs replay the same list)
    torch.cuda.synchronize()
    launches_per_step = _lib.total_launches() - n0
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, K, W, from_host):
        with torch.cuda.stream(step.stream):                 # high-priority stream of the step (graph_step.py)
            return timed_on_stream(step, K, W, from_host)

    def timed_on_stream(step, K, W, from_host):
        ev_scan = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # e2e: packed pinned host batches -> one H2D copy per step; the loss of every step is read back to pinned host
        # memory asynchronously and the host waits for step i-1's loss while step i runs (one step in flight)
        src = [step.pack_host(*hb) for hb in host] if from_host else [step.pack_host(*hb, pin=False).to(dev) for hb in host]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        for i in range(W):
            step.load_packed(src[i % pool]); step.run()
            if from_host:
                loss_host[i % 2].copy_(step.loss, non_blocking=True); torch.cuda.current_stream().synchronize()
        barrier()
        e0.record()
        for i in range(K):
            step.load_packed(src[(W + i) % pool])
            step.run(events=ev_scan[i])                      # scan launch (event-bracketed on ITS stream) + replay
            if from_host:                                    # the reference reads the loss every step (:1928)
                loss_host[i % 2].copy_(step.loss, non_blocking=True)
                done[i % 2].record()
                if i >= 1:
                    done[(i - 1) % 2].synchronize()
        if from_host:
            done[(K - 1) % 2].synchronize()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        scan_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_scan]))
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), scan_ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_begin = time.perf_counter()
    ms_dev, scan_ms = timed(step, args.steps, args.warmup, from_host=False)
    ms_e2e, _ = timed(step, args.steps, args.warmup, from_host=True)
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    dlrm.emb_group.check_status()
    dlrm._dense_arena.check_status()
    final_loss = float(step.loss.item())
    exchange = "single GPU: no exchange" if world == 1 else (
        "one-kernel all-gathers over NVLink peer memory (csrc/p2p.cu), 5 per step, no NCCL call in the step"
        if dlrm.emb_group.p2p is not None else "NCCL: 4 all-gathers + 1 MAX all-reduce per step (same slots and consumer kernels as the NVLink form)")
    extras = {}
    if args.scale_policy == "full" and not args.no_extras and world == 1:
        # the same step with (a) the rescan overlapped with the step (block maxima on a side stream + fix-up of the
        # updated blocks) and (b) the exact incremental tracker; reported beside the headline, never instead of it
        variants = [("pipelined_rescan", "pipelined",
                     "same full rescan (every table byte read once per step), overlapped with the step on a low-priority "
                     "stream; scales bit-identical (tests/test_gpu_tracker.py)"),
                    ("incremental_scale_tracker", "incremental",
                     "exact block-max tracker: reads only the touched blocks; scales bit-identical to the "
                     "rescan (SURVEY.md 8 f-1); not the headline because the reference rescans every step")]
        for key, policy, note in variants:
            dlrm.emb_group.scale_policy = policy
            dlrm.emb_group.scale_valid = False
            step2 = GraphedTrainStep(dlrm, *devb[0], lr=LR, world_size=world, rank=rank, grad_bits=8, warmup=3,
                                     use_graph=not args.no_graph)
            ms2, scan2 = timed(step2, args.steps, args.warmup, from_host=False)
            ms2_e2e, _ = timed(step2, args.steps, args.warmup, from_host=True)
            dlrm.emb_group.check_status()
            extras[key] = {"value": B * world * args.steps / (ms2 / 1000.0), "unit": "samples/s",
                           "ms_per_step": ms2 / args.steps,
                           "e2e_value": B * world * args.steps / (ms2_e2e / 1000.0),
                           "scan_kernel_ms": scan2, "note": note}
            del step2
        dlrm.emb_group.scale_policy = args.scale_policy
        dlrm.emb_group.scale_valid = False

    gbatch = B * world
    value = gbatch * args.steps / (ms_dev / 1000.0)
    e2e = gbatch * args.steps / (ms_e2e / 1000.0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    algo_bytes = table_bytes / world if dlrm.shard_scan else table_bytes
    achieved = algo_bytes / (scan_ms / 1000.0) / 1e9
    nvlink = exchange_microbench(dlrm, world, barrier) if world > 1 else None
    if rank != 0:
        finish(world)
        return
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}-shape DQRM: {len(cfg['rows'])} tables ({sum(cfg['rows'])} rows, "
                               f"{table_bytes / 1e9:.3f} GB fp32), dim {cfg['dim']}, INT4 emb+MLP QAT, INT8 grad exchange, "
                               f"batch {B}/GPU", "global_batch": gbatch, "parallelism": f"dp{world}",
                   "scale_scan": ({"incremental": "exact incremental block-max tracker",
                                   "full": "full rescan every step (reference period-1 semantics), serialised before the forward",
                                   "pipelined": "full rescan every step (every table byte read once per step, reference "
                                                "period-1 semantics), overlapped with the step on a low-priority stream; "
                                                "blocks holding updated rows are re-read after the update"}[args.scale_policy] +
                                  (", row-sharded 1/N + MAX all-reduce" if dlrm.shard_scan else "")),
                   "exchange": exchange,
                   "l2": f"table arena ({table_bytes / 1e9:.2f} GB) is {table_bytes / 126e6:.0f}x the 126 MB L2: inputs larger than L2, no flush needed",
                   "cuda_graph": step.graph is not None, "final_loss": final_loss},
        "e2e": {"value": e2e, "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": step.input_bytes(), "d2h_bytes_per_step": 4,
                "pipeline": "one packed pinned H2D copy per step; loss D2H asynchronous, host waits for step i-1 while "
                            "step i runs"},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": {"kernel": {"pipelined": "blockmax_scan_kernel", "full": "table_absmax_kernel",
                                "incremental": "table_absmax_kernel (block maxima)"}[args.scale_policy], "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_SCAN_TRAFFIC.get((args.workload, world, args.scale_policy)),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                     "frac_of_nominal_8TBs": achieved / 8000.0,
                     "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": scan_ms,
                     "share_of_step": scan_ms / (ms_dev / args.steps)},
        "clocks": clocks,
    }
    line.update(extras)
    if nvlink is not None:
        line["nvlink_exchange"] = nvlink
    if world == 1 and not args.no_cpu_baseline:
        del step, dlrm
        torch.cuda.empty_cache()
        val, ms, cores, build_s = oracle_arm(cfg, B, args.cpu_steps, 1)
        line["cpu_baseline"] = {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                                "ms_per_step": ms,
                                "sample": f"{args.cpu_steps} full train steps after 1 warm-up, batch {B}, same "
                                          f"{args.workload}-shape model on the host ({build_s:.0f}s to build tables)"}
    print(json.dumps(line), flush=True)
    finish(world)


# DRAM traffic of one table_absmax_kernel launch from `ncu --set full` (dram__bytes_read.sum +
# dram__bytes_write.sum, profiles/r01_scan_kernel_ncu_full.txt); valid for the full (unsharded) Kaggle scan.
NCU_SCAN_TRAFFIC = {("kaggle", 1, "full"): 2.1647e9 + 3.9e6}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
