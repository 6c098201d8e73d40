#!/bin/bash
# round-2 GPU call Q: red-zone memory-safety pass of the parity workloads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1200 python tools/redzone_run.py ) > gpurun_out/q_redzone.log 2>&1
echo "redzone rc=$?" > gpurun_out/q_summary.log
( time timeout 600 python -m pytest tests/test_redzones.py -x -q -m gpu ) > gpurun_out/q_pytest_rz.log 2>&1
echo "pytest rz rc=$?" >> gpurun_out/q_summary.log
cat gpurun_out/q_summary.log; tail -5 gpurun_out/q_redzone.log
