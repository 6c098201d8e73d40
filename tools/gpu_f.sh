#!/bin/bash
# round-2 GPU call F: wide runs in the single-warp engine — parity suite, A/B on the sparse / ensemble probes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/f_summary.log
for off in 0 4; do
  NS_FEATURES_OFF=$off timeout 600 python tests/gpu_perf_probe.py sparse > gpurun_out/f_sparse_off$off.log 2>&1; echo "sparse off=$off rc=$?" >> gpurun_out/f_summary.log
  NS_FEATURES_OFF=$off timeout 600 python tests/gpu_perf_probe.py ensemble 1024 1000 100000 > gpurun_out/f_ens_off$off.log 2>&1; echo "ens off=$off rc=$?" >> gpurun_out/f_summary.log
done
timeout 600 python tests/gpu_perf_probe.py soak > gpurun_out/f_soak.log 2>&1; echo "soak rc=$?" >> gpurun_out/f_summary.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_summary.log
cat gpurun_out/f_summary.log
tail -4 gpurun_out/f_pytest.log
for off in 0 4; do echo "== off=$off"; grep -v "^---" gpurun_out/f_sparse_off$off.log | tail -6; tail -2 gpurun_out/f_ens_off$off.log; done
tail -3 gpurun_out/f_soak.log
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
    print('bench', round(d['value']/1e6,2),'M ev/s', d['config']['parity']['ok'])
except Exception as e: print('bench ERR', e)
P
