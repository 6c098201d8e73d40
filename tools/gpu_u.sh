#!/bin/bash
# round-2 GPU call U: engine thresholds of the batch kernel: NS_E1_MAX (events a group may have in the one-thread-per-group pass),
# NS_HUB_MIN_DEG (flows a sender needs for its chain to get a whole CTA)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for t in "8 160" "8 96" "8 64" "8 48" "8 33" "16 64"; do
  set -- $t
  NS_E1_MAX=$1 NS_HUB_MIN_DEG=$2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/u_bench_$1_$2.json 2> gpurun_out/u_bench_$1_$2.err
  python - $1 $2 <<'P'
import json,sys
a,b=sys.argv[1:3]
try:
    d=json.loads(open(f'gpurun_out/u_bench_{a}_{b}.json').read().strip().splitlines()[-1])
    print('E1_MAX',a,'HUB_MIN_DEG',b, round(d['value']/1e6,2),'M ev/s ms/step',round(d['ms_per_step'],3),'parity',d['config']['parity']['ok'])
except Exception as e: print(a,b,'ERR',e)
P
done
