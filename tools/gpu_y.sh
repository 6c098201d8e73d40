#!/bin/bash
# sparse-regime timings after a change: lone replica, 296 replicas, p2p1k / swarm10k ensembles, default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-y}
if [ "$2" = "tests" ]; then ( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/${T}_pytest.log; fi
timeout 300 python tools/ncu_targets.py seq 2>&1 | tail -1
timeout 300 python tools/ncu_targets.py ens 2>&1 | tail -1
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --no-cpu-baseline > gpurun_out/${T}_p2p.json 2> gpurun_out/${T}_p2p.err
if [ "$2" = "tests" ]; then
timeout 900 python bench.py --workload ensemble --no-cpu-baseline > gpurun_out/${T}_swarm.json 2> gpurun_out/${T}_swarm.err
timeout 900 python bench.py --workload churn --no-cpu-baseline > gpurun_out/${T}_churn.json 2> gpurun_out/${T}_churn.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
fi
python - $T <<'P'
import json,sys,os
T=sys.argv[1]
for f in ('p2p','swarm','churn','bench'):
    if not os.path.exists(f'gpurun_out/{T}_{f}.json'): continue
    try:
        d=json.loads(open(f'gpurun_out/{T}_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3),'M ev/s  e2e',round(d['e2e']['value']/1e6,3),'ms/step',round(d['ms_per_step'],2), d['config']['parity']['ok'])
    except Exception as e: print(f,'ERR',e)
P
