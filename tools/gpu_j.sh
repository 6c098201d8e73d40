#!/bin/bash
# round-2 GPU call J: client behaviour on the device vs the compiled reference; headline bench sanity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/t_clients.py > gpurun_out/j_clients.log 2>&1; echo "clients rc=$?" > gpurun_out/j_summary.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?" >> gpurun_out/j_summary.log
cat gpurun_out/j_summary.log; tail -40 gpurun_out/j_clients.log
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/j_bench.json').read().strip().splitlines()[-1]); print('bench', round(d['value']/1e6,2),'M ev/s', d['config']['parity']['ok'])
except Exception as e: print('bench ERR', e)
P
