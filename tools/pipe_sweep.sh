#!/bin/bash
# Sweep the resident-CTA cap of the pipelined rescan (DQRM_PIPE_CTAS_PER_SM) on one B200.
#   tools/pipe_sweep.sh [kaggle|terabyte] [batch] [steps]
W=${1:-kaggle}; B=${2:-128}; S=${3:-50}
for k in 0 1 2 3 4; do
  DQRM_PIPE_CTAS_PER_SM=$k timeout 300 python bench.py --workload $W --batch $B --steps $S --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null \
    | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'workload':'$W','ctas_per_sm':$k,'ms_per_step':round(d['ms_per_step'],4),'e2e_ms':round(d['e2e']['ms_per_step'],4),'scan_ms':round(d['roofline']['avg_launch_ms'],4),'scan_gbs':round(d['roofline']['achieved'],1)}))"
done
