#!/bin/bash
# round-2 GPU call Z (2 GPUs): the default bench under torchrun with the final build, both arms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/z_bench_$N.json 2> gpurun_out/z_bench_$N.err; echo "ours N=$N rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/z_ref_$N.json 2> gpurun_out/z_ref_$N.err; echo "ref N=$N rc=$?"
python - $N <<'P'
import json,sys
N=sys.argv[1]
for f in (f'z_bench_{N}', f'z_ref_{N}'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'n_gpus', d['n_gpus'], round(d['value']/1e6,3),'M ev/s  e2e',round(d['e2e']['value']/1e6,3),'ms/step',round(d['ms_per_step'],2), (d.get('config',{}).get('parity') or {}).get('ok'), (d.get('cpu_baseline') or {}).get('cores'))
    except Exception as e: print(f,'ERR',e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
P
