#!/bin/bash
# round-2 GPU call M (8 GPUs): ensemble workloads sharded over 2 / 4 / 8 ranks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/m_gpus.log 2>&1
for kind in p2p1k swarm10k; do
  for N in 8 4 2; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --workload ensemble --ens-kind $kind --gpus $N --no-cpu-baseline > gpurun_out/m_${kind}_$N.json 2> gpurun_out/m_${kind}_$N.err
    echo "$kind N=$N rc=$?" >> gpurun_out/m_summary.log
  done
done
cat gpurun_out/m_summary.log
python - <<'P'
import json
for kind in ('p2p1k','swarm10k'):
    for N in (2,4,8):
        try:
            d=json.loads(open(f'gpurun_out/m_{kind}_{N}.json').read().strip().splitlines()[-1])
            print(kind, N, round(d['value']/1e6,2),'M ev/s  e2e',round(d['e2e']['value']/1e6,2),'ms',round(d['ms_per_step'],1), d['config']['parity']['ok'], d['config']['device_ms_per_rank'])
        except Exception as e: print(kind,N,'ERR',e)
P
tail -n 5 gpurun_out/m_p2p1k_8.err
