#!/bin/bash
# one-simulation-per-warp kernels with shared (called) per-event helpers: parity, timings, instruction-cache requests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-y}
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log | head -1
timeout 300 python tools/ncu_targets.py seq 2>&1 | tail -1
timeout 300 python tools/ncu_targets.py ens 2>&1 | tail -1
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --no-cpu-baseline > gpurun_out/${T}_p2p.json 2> gpurun_out/${T}_p2p.err
timeout 900 python bench.py --workload ensemble --no-cpu-baseline > gpurun_out/${T}_swarm.json 2> gpurun_out/${T}_swarm.err
timeout 900 python bench.py --workload churn --no-cpu-baseline > gpurun_out/${T}_churn.json 2> gpurun_out/${T}_churn.err
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --replicas 1 --no-cpu-baseline > gpurun_out/${T}_p2p1.json 2> gpurun_out/${T}_p2p1.err
python - $T <<'P'
import json,sys
T=sys.argv[1]
for f in ('p2p','swarm','churn','p2p1'):
    try:
        d=json.loads(open(f'gpurun_out/{T}_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3),'M ev/s  e2e',round(d['e2e']['value']/1e6,3),'ms/step',round(d['ms_per_step'],2), d['config']['parity']['ok'])
    except Exception as e: print(f,'ERR',e)
P
timeout 600 ncu --metrics gpu__time_duration.sum,gcc__cache_requests_type_instruction.sum,smsp__inst_executed.sum,gcc__average_cache_request_hit_rate.pct,smsp__pcsamp_warps_issue_stalled_no_instructions,smsp__pcsamp_warps_issue_stalled_long_scoreboard,smsp__pcsamp_warps_issue_stalled_wait --clock-control none -k regex:k_run_seq -c 1 --csv --log-file gpurun_out/${T}_ens_icache.csv python tools/ncu_targets.py ens > /dev/null 2>&1
grep -E "gpu__time|gcc__|inst_executed|stalled" gpurun_out/${T}_ens_icache.csv | awk -F'","' '{print $(NF-2), $NF}'
