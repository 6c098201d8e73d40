#!/bin/bash
# round-2 GPU call O: picker tests + microbench; ncu --set full captures of k_pick_small, k_run_seq, k_run_ens, k_init_par
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -x -q -m gpu -k "pick" ) > gpurun_out/o_pytest_pick.log 2>&1; echo "pytest pick rc=$?" > gpurun_out/o_summary.log
timeout 300 python tools/pick_bench.py > gpurun_out/o_pick_bench.log 2>&1; echo "pick bench rc=$?" >> gpurun_out/o_summary.log
for t in pick seq ens; do
  timeout 300 python tools/ncu_targets.py $t > gpurun_out/o_plain_$t.log 2>&1; echo "plain $t rc=$?" >> gpurun_out/o_summary.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pick_small -c 1 -o gpurun_out/o_prof_pick -f python tools/ncu_targets.py pick > gpurun_out/o_ncu_pick.log 2>&1; echo "ncu pick rc=$?" >> gpurun_out/o_summary.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_seq -c 1 -o gpurun_out/o_prof_seq -f python tools/ncu_targets.py seq > gpurun_out/o_ncu_seq.log 2>&1; echo "ncu seq rc=$?" >> gpurun_out/o_summary.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_seq -s 0 -c 1 -o gpurun_out/o_prof_ens_seq -f python tools/ncu_targets.py ens > gpurun_out/o_ncu_ens_seq.log 2>&1; echo "ncu ens(seq phase) rc=$?" >> gpurun_out/o_summary.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_ens -s 0 -c 1 -o gpurun_out/o_prof_ens_cta -f python tools/ncu_targets.py ens > gpurun_out/o_ncu_ens_cta.log 2>&1; echo "ncu ens(cta phase) rc=$?" >> gpurun_out/o_summary.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_init_par -c 1 -o gpurun_out/o_prof_init -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/o_ncu_init.log 2>&1; echo "ncu init rc=$?" >> gpurun_out/o_summary.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/o_launches_ens.csv python tools/ncu_targets.py ens > gpurun_out/o_ncu_list_ens.log 2>&1; echo "ncu list ens rc=$?" >> gpurun_out/o_summary.log
cat gpurun_out/o_summary.log; tail -3 gpurun_out/o_pytest_pick.log; cat gpurun_out/o_pick_bench.log; tail -n 2 gpurun_out/o_plain_*.log
