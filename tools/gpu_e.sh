#!/bin/bash
# round-2 GPU call E: state check after re-entry — parity suite, both bench arms, phase split, probes, ncu list + full capture, sanitizer probe
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nproc > gpurun_out/e_host.log; lscpu | head -20 >> gpurun_out/e_host.log
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/e_summary.log
timeout 900 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?" >> gpurun_out/e_summary.log
timeout 900 python bench.py --impl reference > gpurun_out/e_ref.json 2> gpurun_out/e_ref.err; echo "ref rc=$?" >> gpurun_out/e_summary.log
NS_LIB=$PWD/netsim_b200/libnetsim_b200_prof.so timeout 600 python tests/gpu_perf_probe.py windows > gpurun_out/e_windows.log 2>&1; echo "windows rc=$?" >> gpurun_out/e_summary.log
timeout 600 python tests/gpu_perf_probe.py sparse > gpurun_out/e_sparse.log 2>&1; echo "sparse rc=$?" >> gpurun_out/e_summary.log
timeout 600 python tests/gpu_perf_probe.py ensemble 1024 1000 100000 > gpurun_out/e_ens.log 2>&1; echo "ens rc=$?" >> gpurun_out/e_summary.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/e_ncu1.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/e_summary.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_run_par -s 4 -c 1 -o gpurun_out/e_prof_r2b -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/e_ncu2.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/e_summary.log
( timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/e_sanitizer.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/e_summary.log
cat gpurun_out/e_summary.log
tail -3 gpurun_out/e_pytest.log
python - <<'P'
import json
for f in ('e_bench','e_ref'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,2),'M ev/s  e2e',round(d['e2e']['value']/1e6,2),'ms/step',round(d['ms_per_step'],2), d.get('config',{}).get('parity'), 'frac', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
P
tail -11 gpurun_out/e_windows.log
cat gpurun_out/e_sparse.log gpurun_out/e_ens.log | tail -12
tail -5 gpurun_out/e_sanitizer.log
