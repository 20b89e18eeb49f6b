#!/usr/bin/env python
"""Text summary of an `ncu --set full` report for profiles/: one block per captured launch with the metrics the
roofline discussion in DESIGN.md uses (duration, DRAM bytes, L2 hit rate, sectors per request, occupancy, issue
rate, tensor-pipe activity, top warp-stall reasons).

    python tools/ncu_summary.py REPORT.ncu-rep [--title "..."] [--algo-bytes N] > profiles/rNN_x.txt
"""
import argparse
import csv
import re
import subprocess
import sys

WANT = [
    ("duration_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("cluster", "launch__cluster_size"),
    ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem B/CTA", "launch__shared_mem_per_block_dynamic"),
    ("static smem B/CTA", "launch__shared_mem_per_block_static"),
    ("waves/SM", "launch__waves_per_multiprocessor"),
    ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("dram read", "dram__bytes_read.sum"),
    ("dram write", "dram__bytes_write.sum"),
    ("dram throughput % of peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram read B/s", "dram__bytes_read.sum.per_second"),
    ("L2 sector hit rate %", "lts__t_sector_hit_rate.pct"),
    ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1 global-load sectors", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
    ("L1 global-load requests", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
    ("L1 global-store sectors", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"),
    ("L1 global-store requests", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("IPC (executed)", "sm__inst_executed.avg.per_cycle_elapsed"),
    ("tensor pipe active %", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("tensor hmma subpipe active cycles (avg/SM)", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"),
    ("elapsed cycles (avg/SM)", "sm__cycles_elapsed.avg"),
    ("tensor-memory (TMEM) active %", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("LSU pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("SM clock MHz", "sm__cycles_elapsed.avg.per_second"),
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--title", default="")
    ap.add_argument("--algo-bytes", type=float, default=0.0, help="algorithmic bytes per launch, for the traffic ratio")
    ap.add_argument("--max", type=int, default=12)
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {c: i for i, c in enumerate(hdr)}

    def get(r, name):
        i = col.get(name)
        if i is None:
            for c, j in col.items():
                if c.endswith(name):
                    i = j
                    break
        if i is None:
            return None, ""
        return r[i], units[i]

    print(f"# {a.title or a.report}")
    print(f"# source: ncu -i {a.report.split('/')[-1]} --page raw --csv ; {len(data)} captured launch(es); ncu serialises and "
          "replays each launch with cold caches -- durations here are NOT bench numbers")
    stall_cols = [c for c in hdr if re.search(r"smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio$", c)
                  or re.search(r"smsp__average_warp_latency_issue_stalled_.*\.ratio$", c)]
    for n, r in enumerate(data[: a.max]):
        name = r[col["Kernel Name"]]
        print(f"\n## launch {n}: {name[:150]}")
        for label, m in WANT:
            v, u = get(r, m)
            if v in (None, ""):
                continue
            print(f"  {label:32s} {v} {u}")
        rd, ru = get(r, "dram__bytes_read.sum")
        wr, wu = get(r, "dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        if rd and wr and a.algo_bytes:
            tot = num(rd) * scale.get(ru, 1) + num(wr) * scale.get(wu, 1)
            print(f"  {'traffic / algorithmic bytes':32s} {tot / a.algo_bytes:.3f}  ({tot / 1e6:.1f} MB / {a.algo_bytes / 1e6:.1f} MB)")
        ls, _ = get(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
        lr, _ = get(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
        if ls and lr and num(lr):
            print(f"  {'global-load sectors/request':32s} {num(ls) / num(lr):.2f}")
        st = []
        for c in stall_cols:
            v = num(r[col[c]])
            if v:
                st.append((v, re.sub(r".*issue_stalled_(.*?)(_per_issue_active)?\.ratio", r"\1", c)))
        st.sort(reverse=True)
        if st:
            print("  top warp stalls (cycles per issued instruction): " + ", ".join(f"{k} {v:.2f}" for v, k in st[:6]))


if __name__ == "__main__":
    sys.exit(main())
