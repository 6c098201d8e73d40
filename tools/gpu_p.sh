#!/bin/bash
# round-2 GPU call P: prefetch build — tests, p2p1k single/ensemble, churn; then ncu captures converted to CSV on the box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/p_summary.log
timeout 300 python tools/ncu_targets.py seq > gpurun_out/p_plain_seq.log 2>&1; echo "plain seq rc=$?" >> gpurun_out/p_summary.log
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --no-cpu-baseline > gpurun_out/p_p2p.json 2> gpurun_out/p_p2p.err; echo "p2p ens rc=$?" >> gpurun_out/p_summary.log
timeout 900 python bench.py --workload ensemble --ens-kind p2p1k --replicas 128 --no-cpu-baseline > gpurun_out/p_p2p128.json 2> gpurun_out/p_p2p128.err; echo "p2p 128 rc=$?" >> gpurun_out/p_summary.log
timeout 900 python bench.py --workload ensemble --no-cpu-baseline > gpurun_out/p_swarm.json 2> gpurun_out/p_swarm.err; echo "swarm ens rc=$?" >> gpurun_out/p_summary.log
timeout 900 python bench.py --workload churn --no-cpu-baseline > gpurun_out/p_churn.json 2> gpurun_out/p_churn.err; echo "churn rc=$?" >> gpurun_out/p_summary.log
cap(){ # name kernel-regex skip cmd...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none -k regex:$rx -s $skip -c 1 -o gpurun_out/p_prof_$name -f "$@" > gpurun_out/p_ncu_$name.log 2>&1
  echo "ncu $name rc=$?" >> gpurun_out/p_summary.log
  ncu -i gpurun_out/p_prof_$name.ncu-rep --page raw --csv > gpurun_out/p_raw_$name.csv 2>/dev/null
  rm -f gpurun_out/p_prof_$name.ncu-rep
}
cap pick k_pick_small 0 python tools/ncu_targets.py pick
cap seq k_run_seq 0 python tools/ncu_targets.py seq
cap ens_seq k_run_seq 0 python tools/ncu_targets.py ens
cap ens_cta k_run_ens 0 python tools/ncu_targets.py ens
cap init k_init_par 0 python bench.py --steps 1 --warmup 1 --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/p_launches_ens.csv python tools/ncu_targets.py ens > gpurun_out/p_ncu_list_ens.log 2>&1; echo "ncu list ens rc=$?" >> gpurun_out/p_summary.log
du -sh gpurun_out
cat gpurun_out/p_summary.log; tail -4 gpurun_out/p_pytest.log; cat gpurun_out/p_plain_seq.log
python - <<'P'
import json
for f in ('p_p2p','p_p2p128','p_swarm','p_churn'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3),'M ev/s  e2e',round(d['e2e']['value']/1e6,3),'ms/step',round(d['ms_per_step'],2), d['config']['parity']['ok'], 'launches', d['gpu_launches'])
    except Exception as e: print(f,'ERR',e)
P
