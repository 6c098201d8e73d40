#!/bin/bash
# round-2 GPU call T: unmodified p2p.c (+ cloud_behaviour.c, client_behaviour.c, resource.c) on the host core against the
# reference run live on the same box, over the committed configurations
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
R=$PWD
mkdir -p gpurun_out
H=$R/netsim_b200/host/_build
O=$R/gpurun_out
: > $O/t_summary.log
for n in ${@:-small50 base200 push1 play2b play2c cloudb cloudc push2}; do
  seed=$(python3 -c "import json; print(json.load(open('$R/tests/golden/p2p/expected.json'))['$n']['seed'])")
  d=$(mktemp -d); mkdir $d/ref $d/dev; cp $R/tests/golden/p2p/$n/p2p.cfg $d/ref/; cp $R/tests/golden/p2p/$n/p2p.cfg $d/dev/
  ( cd $d/ref; NS_FAKE_TIME=$seed LD_PRELOAD=$R/oracle/_ref/timeshim.so timeout 300 $R/oracle/_ref/netsim $R/oracle/_ref/modules/p2p > out.txt 2> err.txt; echo "$n ref rc=$?" >> $O/t_summary.log )
  s=$(date +%s.%N)
  ( cd $d/dev; NS_FAKE_TIME=$seed LD_PRELOAD=$R/oracle/_ref/timeshim.so timeout 900 $H/netsim_dev $H/modules/p2p > out.txt 2> err.txt; echo "$n dev rc=$?" >> $O/t_summary.log )
  e=$(date +%s.%N)
  python3 - $d $n $s $e <<'P' >> $O/t_summary.log
import sys
d,n,s,e=sys.argv[1:5]
def pay(p):
    x=open(p,'rb').read(); k=int.from_bytes(x[:4],'big'); return x[4:4+k]
try:
    a=pay(d+'/ref/p2p_record'); b=pay(d+'/dev/p2p_record')
    ea=open(d+'/ref/err.txt','rb').read(); eb=open(d+'/dev/err.txt','rb').read()
    print(n, 'records', len(b), 'ref', len(a), 'IDENTICAL' if a==b else 'DIFFER', '| stderr', 'IDENTICAL' if ea==eb else 'DIFFER', '| dev wall %.1fs'%(float(e)-float(s)))
    if a!=b:
        k=next((i for i in range(min(len(a),len(b))) if a[i]!=b[i]), min(len(a),len(b)))
        print('  first record difference at byte', k, 'ref', a[max(0,k-19):k+19].hex(), 'dev', b[max(0,k-19):k+19].hex())
    if ea!=eb:
        la=ea.split(b'\n'); lb=eb.split(b'\n')
        k=next((i for i in range(min(len(la),len(lb))) if la[i]!=lb[i]), min(len(la),len(lb)))
        print('  first stderr difference at line', k, 'ref', la[k:k+2], 'dev', lb[k:k+2])
except Exception as ex: print(n,'ERR',ex)
P
  tail -3 $d/dev/err.txt > $O/t_${n}_tail.err
done
cat $O/t_summary.log
