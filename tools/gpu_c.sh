#!/bin/bash
# round-2 GPU call C: phase split per window (profiling build), ncu launch list + full capture of one window
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NS_LIB=$PWD/netsim_b200/libnetsim_b200_prof.so timeout 600 python tests/gpu_perf_probe.py windows > gpurun_out/c_windows.log 2>&1; echo "windows rc=$?" > gpurun_out/c_summary.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/c_plain.json 2> gpurun_out/c_plain.err; echo "plain rc=$?" >> gpurun_out/c_summary.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/c_ncu1.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/c_summary.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_run_par -s 4 -c 1 -o gpurun_out/c_prof_r2a -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/c_ncu2.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/c_summary.log
cat gpurun_out/c_summary.log
cat gpurun_out/c_windows.log | tail -11
