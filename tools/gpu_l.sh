#!/bin/bash
# round-2 GPU call L: analyzer reductions on the device, full suite, p2p1k ensemble with the analysis inside
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_analysis.py -x -q -m gpu ) > gpurun_out/l_pytest_ana.log 2>&1
echo "pytest ana rc=$?" > gpurun_out/l_summary.log
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/l_summary.log
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --no-cpu-baseline > gpurun_out/l_p2p.json 2> gpurun_out/l_p2p.err; echo "p2p ens rc=$?" >> gpurun_out/l_summary.log
cat gpurun_out/l_summary.log
tail -12 gpurun_out/l_pytest_ana.log; tail -4 gpurun_out/l_pytest.log
python - <<'P'
import json
for f in ('l_p2p',):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3),'M ev/s  e2e',round(d['e2e']['value']/1e6,3),'ms/step',round(d['ms_per_step'],2), d['config']['parity']['ok'], d['config'].get('analysis_rank0'))
    except Exception as e: print(f,'ERR',e)
P
tail -n 3 gpurun_out/l_p2p.err
