#!/bin/bash
# round-2 GPU call W: state of the final build — parity suite (incl. red zones, host core with p2p.c), smoke, every bench workload with
# its reference arm, launch list + full capture of k_run_par converted on the box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > $O/w_pytest.log 2>&1
echo "pytest rc=$?" > $O/w_summary.log
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/w_smoke.log 2>&1; echo "smoke rc=$?" >> $O/w_summary.log
timeout 900 python bench.py > $O/w_bench.json 2> $O/w_bench.err; echo "bench rc=$?" >> $O/w_summary.log
timeout 900 python bench.py --impl reference > $O/w_ref.json 2> $O/w_ref.err; echo "ref rc=$?" >> $O/w_summary.log
timeout 900 python bench.py --workload churn > $O/w_churn.json 2> $O/w_churn.err; echo "churn rc=$?" >> $O/w_summary.log
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --no-cpu-baseline > $O/w_p2p.json 2> $O/w_p2p.err; echo "p2p ens rc=$?" >> $O/w_summary.log
timeout 900 python bench.py --workload ensemble --no-cpu-baseline > $O/w_swarm.json 2> $O/w_swarm.err; echo "swarm ens rc=$?" >> $O/w_summary.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/w_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/w_ncu1.log 2>&1; echo "ncu list rc=$?" >> $O/w_summary.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_run_par -s 4 -c 1 -o $O/w_prof -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/w_ncu2.log 2>&1; echo "ncu full rc=$?" >> $O/w_summary.log
ncu -i $O/w_prof.ncu-rep --page raw --csv > $O/w_k_run_par_raw.csv 2>/dev/null; rm -f $O/w_prof.ncu-rep
cat $O/w_summary.log; tail -3 $O/w_pytest.log; tail -3 $O/w_smoke.log
python - <<'P'
import json
for f in ('w_bench','w_ref','w_churn','w_p2p','w_swarm'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3),'M ev/s  e2e',round(d['e2e']['value']/1e6,3),'ms/step',round(d['ms_per_step'],2), (d.get('config',{}).get('parity') or {}).get('ok'), 'frac', (d.get('roofline') or {}).get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'launches', d.get('gpu_launches'))
    except Exception as e: print(f,'ERR',e)
P
