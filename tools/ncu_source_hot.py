"""Hot source lines of one kernel from an ncu report: python tools/ncu_source_hot.py rep.ncu-rep [N]
(ncu -i rep --page source --csv --print-source cuda,sass; warp-stall samples aggregated per CUDA source line)"""
import csv, subprocess, sys, io
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, h, data = None, None, []
cols = ["stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_mio", "stall_lg", "stall_membar", "stall_math",
        "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_no_inst", "stall_sleep"]
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h = r; continue
    if h and r[0].isdigit():
        try: n = int(r[h.index("Warp Stall Sampling (All Samples)")])
        except ValueError: continue
        if n: data.append((n, cur, r))
tot = sum(d[0] for d in data)
print("total samples", tot)
for n, f, r in sorted(data, key=lambda x: -x[0])[:top]:
    extra = " ".join(f"{c[6:12]}={r[h.index(c)]}" for c in cols if r[h.index(c)] not in ("0", ""))
    print(f"{n:6d} {100*n/tot:5.1f}% {f}:{r[0]:>4} {r[1].strip()[:100]!r} {extra}")
