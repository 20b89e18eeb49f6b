"""One training step out of an ncu launch list (profiles/*_launches_*.csv):
    ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras
    python tools/launch_summary.py launches.csv > profiles/rNN_step_launch_summary.txt
The step = the launches from one table_absmax_kernel to the next (the second-to-last such pair in the list)."""
import csv, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
for d in r:
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    unit = d.get("Metric Unit", "ns")
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    rows.append((d["Kernel Name"], us, d.get("Grid Size", ""), d.get("Block Size", "")))
scans = [i for i, x in enumerate(rows) if "table_absmax_kernel" in x[0]]
a, b = scans[-2], scans[-1]
step = rows[a:b]
tot = sum(x[1] for x in step)
print(f"# {len(step)} launches, {tot:.1f} us serialised (cold-cache, one launch at a time: the SHARE carries over, not the absolute)")
print("      us  share  grid x block  kernel")
for name, us, g, blk in step:
    print(f"{us:8.2f} {100 * us / tot:5.1f}%  {g} x {blk}  {name[:130]}")
