#!/bin/bash
# round-2 GPU call I: ensemble engine hand-over (CTA <-> warp per replica), churn bench both arms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu -k "ensemble" ) > gpurun_out/i_pytest_ens.log 2>&1
echo "pytest ens rc=$?" > gpurun_out/i_summary.log
timeout 900 python bench.py --workload ensemble --no-cpu-baseline > gpurun_out/i_ens.json 2> gpurun_out/i_ens.err; echo "ens rc=$?" >> gpurun_out/i_summary.log
timeout 600 python tests/gpu_perf_probe.py ensemble 1024 1000 100000 > gpurun_out/i_ens1k.log 2>&1; echo "ens1k rc=$?" >> gpurun_out/i_summary.log
timeout 900 python bench.py --workload churn --no-cpu-baseline > gpurun_out/i_churn.json 2> gpurun_out/i_churn.err; echo "churn rc=$?" >> gpurun_out/i_summary.log
timeout 900 python bench.py --workload churn --impl reference > gpurun_out/i_churn_ref.json 2> gpurun_out/i_churn_ref.err; echo "churn ref rc=$?" >> gpurun_out/i_summary.log
cat gpurun_out/i_summary.log
tail -4 gpurun_out/i_pytest_ens.log
python - <<'P'
import json
for f in ('i_ens','i_churn','i_churn_ref'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,2),'M ev/s  e2e',round(d['e2e']['value']/1e6,2),'ms/step',round(d['ms_per_step'],2), d.get('config',{}).get('parity'), 'frac', d.get('roofline',{}).get('frac'), 'launches', d.get('gpu_launches'), d.get('config',{}).get('parallel_fraction'))
    except Exception as e: print(f,'ERR',e)
P
tail -3 gpurun_out/i_ens.err gpurun_out/i_churn.err; tail -3 gpurun_out/i_ens1k.log
