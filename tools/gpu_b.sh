#!/bin/bash
# round-2 GPU call B: parity suite (incl. the bench-scale test), bench both arms on the same event range, phase split
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/b_summary.log
timeout 900 python bench.py > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?" >> gpurun_out/b_summary.log
for off in 1 2; do
  NS_FEATURES_OFF=$off timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b_bench_off$off.json 2> gpurun_out/b_bench_off$off.err
  echo "bench off=$off rc=$?" >> gpurun_out/b_summary.log
done
timeout 900 python bench.py --impl reference > gpurun_out/b_ref.json 2> gpurun_out/b_ref.err; echo "ref rc=$?" >> gpurun_out/b_summary.log
echo skip-windows
cat gpurun_out/b_summary.log
tail -3 gpurun_out/b_pytest.log
python - <<'P'
import json
for f in ('b_bench','b_bench_off1','b_bench_off2','b_ref'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,2),'M ev/s  e2e',round(d['e2e']['value']/1e6,2),'ms/step',round(d['ms_per_step'],2), d.get('config',{}).get('parity'), 'frac', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
P
tail -4 gpurun_out/b_bench.err
tail -12 gpurun_out/b_windows.log
