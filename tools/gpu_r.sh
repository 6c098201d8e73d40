#!/bin/bash
# round-2 GPU call R: full spread out of line in the one-warp-per-simulation kernels — parity, then the sparse-regime timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r}
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/${T}_summary.log
timeout 300 python tools/ncu_targets.py seq > gpurun_out/${T}_plain_seq.log 2>&1; echo "plain seq rc=$?" >> gpurun_out/${T}_summary.log
timeout 300 python tools/ncu_targets.py ens > gpurun_out/${T}_plain_ens.log 2>&1; echo "plain ens rc=$?" >> gpurun_out/${T}_summary.log
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --no-cpu-baseline > gpurun_out/${T}_p2p.json 2> gpurun_out/${T}_p2p.err; echo "p2p ens rc=$?" >> gpurun_out/${T}_summary.log
timeout 900 python bench.py --workload ensemble --no-cpu-baseline > gpurun_out/${T}_swarm.json 2> gpurun_out/${T}_swarm.err; echo "swarm ens rc=$?" >> gpurun_out/${T}_summary.log
timeout 900 python bench.py --workload churn --no-cpu-baseline > gpurun_out/${T}_churn.json 2> gpurun_out/${T}_churn.err; echo "churn rc=$?" >> gpurun_out/${T}_summary.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> gpurun_out/${T}_summary.log
cat gpurun_out/${T}_summary.log; tail -4 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_plain_seq.log gpurun_out/${T}_plain_ens.log
python - $T <<'P'
import json,sys
T=sys.argv[1]
for f in ('p2p','swarm','churn','bench'):
    try:
        d=json.loads(open(f'gpurun_out/{T}_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3),'M ev/s  e2e',round(d['e2e']['value']/1e6,3),'ms/step',round(d['ms_per_step'],2), d['config']['parity']['ok'], 'launches', d['gpu_launches'])
    except Exception as e: print(f,'ERR',e)
P
