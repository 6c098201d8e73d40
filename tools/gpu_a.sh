#!/bin/bash
# round-2 GPU call A: parity suite with the new chain/hub paths, A/B bench, window phase split
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.log 2>&1
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/a_pytest_all.log 2>&1
rc=$?
echo "pytest all-on rc=$rc" > gpurun_out/a_summary.log
if [ $rc -ne 0 ]; then
  NS_FEATURES_OFF=2 timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/a_pytest_nohub.log 2>&1; echo "pytest hub-off rc=$?" >> gpurun_out/a_summary.log
  NS_FEATURES_OFF=1 timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/a_pytest_nochain2.log 2>&1; echo "pytest chain2-off rc=$?" >> gpurun_out/a_summary.log
  NS_FEATURES_OFF=3 timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/a_pytest_off.log 2>&1; echo "pytest both-off rc=$?" >> gpurun_out/a_summary.log
fi
for off in 0 1 2 3; do
  NS_FEATURES_OFF=$off timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_off$off.json 2> gpurun_out/a_bench_off$off.err
  echo "bench off=$off rc=$?" >> gpurun_out/a_summary.log
done
timeout 900 python tests/gpu_perf_probe.py big 6000000 check > gpurun_out/a_big_check.log 2>&1; echo "big check rc=$?" >> gpurun_out/a_summary.log
NS_LIB=$PWD/netsim_b200/libnetsim_b200_prof.so timeout 600 python tests/gpu_perf_probe.py windows > gpurun_out/a_windows.log 2>&1; echo "windows rc=$?" >> gpurun_out/a_summary.log
timeout 600 python tests/gpu_perf_probe.py sparse > gpurun_out/a_sparse.log 2>&1; echo "sparse rc=$?" >> gpurun_out/a_summary.log
cat gpurun_out/a_summary.log
tail -3 gpurun_out/a_pytest_all.log
for off in 0 1 2 3; do python - <<P
import json
try:
    d=json.loads(open('gpurun_out/a_bench_off$off.json').read().strip().splitlines()[-1])
    print('off=$off', round(d['value']/1e6,1),'M ev/s  e2e',round(d['e2e']['value']/1e6,1),'ms/step',round(d['ms_per_step'],2),'hash',d['config']['record_hash'],'frac',round(d['roofline']['frac'],5))
except Exception as e: print('off=$off', 'ERR', e)
P
done
tail -12 gpurun_out/a_windows.log
