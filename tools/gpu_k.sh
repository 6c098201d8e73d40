#!/bin/bash
# round-2 GPU call K: full parity suite with the client behaviour, headline bench, p2p1k ensemble both arms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/k_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/k_summary.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?" >> gpurun_out/k_summary.log
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --no-cpu-baseline > gpurun_out/k_p2p.json 2> gpurun_out/k_p2p.err; echo "p2p ens rc=$?" >> gpurun_out/k_summary.log
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --impl reference > gpurun_out/k_p2p_ref.json 2> gpurun_out/k_p2p_ref.err; echo "p2p ref rc=$?" >> gpurun_out/k_summary.log
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --replicas 1 --no-cpu-baseline > gpurun_out/k_p2p_single.json 2> gpurun_out/k_p2p_single.err; echo "p2p single rc=$?" >> gpurun_out/k_summary.log
cat gpurun_out/k_summary.log
tail -6 gpurun_out/k_pytest.log
python - <<'P'
import json
for f in ('k_bench','k_p2p','k_p2p_ref','k_p2p_single'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3),'M ev/s  e2e',round(d['e2e']['value']/1e6,3),'ms/step',round(d['ms_per_step'],2), d.get('config',{}).get('parity'), 'frac', d.get('roofline',{}).get('frac'), 'launches', d.get('gpu_launches'))
    except Exception as e: print(f,'ERR',e)
P
tail -n 3 gpurun_out/k_p2p.err
