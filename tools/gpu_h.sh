#!/bin/bash
# round-2 GPU call H: bench --workload ensemble (both arms), default bench sanity, ncu of k_run_ens
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --workload ensemble > gpurun_out/h_ens.json 2> gpurun_out/h_ens.err; echo "ens rc=$?" > gpurun_out/h_summary.log
timeout 900 python bench.py --workload ensemble --impl reference > gpurun_out/h_ens_ref.json 2> gpurun_out/h_ens_ref.err; echo "ens ref rc=$?" >> gpurun_out/h_summary.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?" >> gpurun_out/h_summary.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/h_ens_launches.csv python bench.py --workload ensemble --replicas 296 --warmup 1 --no-cpu-baseline > gpurun_out/h_ncu1.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/h_summary.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_run_ens -s 1 -c 1 -o gpurun_out/h_prof_ens -f python bench.py --workload ensemble --replicas 296 --warmup 1 --no-cpu-baseline > gpurun_out/h_ncu2.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/h_summary.log
cat gpurun_out/h_summary.log
python - <<'P'
import json
for f in ('h_ens','h_ens_ref','h_bench'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,2),'M ev/s  e2e',round(d['e2e']['value']/1e6,2),'ms/step',round(d['ms_per_step'],2), d.get('config',{}).get('parity'), 'frac', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))
    except Exception as e: print(f,'ERR',e)
P
tail -3 gpurun_out/h_ens.err
