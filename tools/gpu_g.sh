#!/bin/bash
# round-2 GPU call G: one CTA per replica (k_run_ens) — ensemble parity tests, probes at 1k / 10k peers per replica
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu -k "ensemble" ) > gpurun_out/g_pytest_ens.log 2>&1
echo "pytest ens rc=$?" > gpurun_out/g_summary.log
timeout 600 python tests/gpu_perf_probe.py ensemble 1024 1000 100000 > gpurun_out/g_ens1k.log 2>&1; echo "ens1k rc=$?" >> gpurun_out/g_summary.log
timeout 600 python tests/gpu_perf_probe.py ensemble 128 10000 400000 0.02 > gpurun_out/g_ens10k.log 2>&1; echo "ens10k rc=$?" >> gpurun_out/g_summary.log
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/g_pytest.log 2>&1
echo "pytest all rc=$?" >> gpurun_out/g_summary.log
cat gpurun_out/g_summary.log
tail -15 gpurun_out/g_pytest_ens.log
cat gpurun_out/g_ens1k.log gpurun_out/g_ens10k.log | tail -8
tail -5 gpurun_out/g_pytest.log
