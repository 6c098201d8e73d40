#!/bin/bash
# round-2 GPU call D: parity suite + bench after the extract/commit/chain changes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/d_summary.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?" >> gpurun_out/d_summary.log
timeout 600 python tests/gpu_perf_probe.py sparse > gpurun_out/d_sparse.log 2>&1
cat gpurun_out/d_summary.log
tail -5 gpurun_out/d_pytest.log
python - <<'P'
import json
for f in ('d_bench',):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,2),'M ev/s  e2e',round(d['e2e']['value']/1e6,2),'ms/step',round(d['ms_per_step'],2), d.get('config',{}).get('parity'), 'frac', d.get('roofline',{}).get('frac'))
    except Exception as e: print(f,'ERR',e)
P
tail -3 gpurun_out/d_bench.err
