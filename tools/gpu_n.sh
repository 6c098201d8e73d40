#!/bin/bash
# round-2 GPU call N: in-kernel record folding, module-API test module on the device core, full suite, ensemble benches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/n_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/n_summary.log
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --no-cpu-baseline > gpurun_out/n_p2p.json 2> gpurun_out/n_p2p.err; echo "p2p ens rc=$?" >> gpurun_out/n_summary.log
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --replicas 128 --no-cpu-baseline > gpurun_out/n_p2p128.json 2> gpurun_out/n_p2p128.err; echo "p2p 128 rc=$?" >> gpurun_out/n_summary.log
timeout 900 python bench.py --workload ensemble --no-cpu-baseline > gpurun_out/n_swarm.json 2> gpurun_out/n_swarm.err; echo "swarm ens rc=$?" >> gpurun_out/n_summary.log
cat gpurun_out/n_summary.log
tail -8 gpurun_out/n_pytest.log
python - <<'P'
import json
for f in ('n_p2p','n_p2p128','n_swarm'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3),'M ev/s  e2e',round(d['e2e']['value']/1e6,3),'ms/step',round(d['ms_per_step'],2), d['config']['parity']['ok'], 'launches', d['gpu_launches'])
    except Exception as e: print(f,'ERR',e)
P
