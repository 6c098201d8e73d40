#!/bin/bash
# round-2 GPU call S: unmodified test_client2/3.c + client_behaviour.c and p2p.c + cloud_behaviour.c on the host core
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
H=$PWD/netsim_b200/host/_build
G=$PWD/tests/golden
O=$PWD/gpurun_out
: > $O/s_summary.log
for m in tc2 tc3; do
  d=$(mktemp -d); ( cd $d; timeout 120 $H/netsim_dev $H/modules/$m > $O/s_$m.out 2> $O/s_$m.err; echo "$m rc=$?" >> $O/s_summary.log )
  if cmp -s $O/s_$m.err $G/$m.stderr.txt; then echo "$m stderr identical" >> $O/s_summary.log; else echo "$m stderr DIFFERS" >> $O/s_summary.log; diff $O/s_$m.err $G/$m.stderr.txt | head -20 >> $O/s_summary.log; fi
done
d=$(mktemp -d); cp $G/p2p_small/p2p.cfg $d/
( cd $d; NS_FAKE_TIME=1 LD_PRELOAD=$PWD/../oracle/_ref/timeshim.so true; NS_FAKE_TIME=1 LD_PRELOAD=$GRAFT_REPO_ROOT/oracle/_ref/timeshim.so timeout 600 $H/netsim_dev $H/modules/p2p > $O/s_p2p.out 2> $O/s_p2p.err; echo "p2p rc=$?" >> $O/s_summary.log
  python3 - <<P >> $O/s_summary.log
import os
try:
    d=open('p2p_record','rb').read(); n=int.from_bytes(d[:4],'big'); r=d[4:4+n]
    g=open('$G/p2p_small/records.bin','rb').read()
    print('p2p records', len(r), 'golden', len(g), 'identical' if r==g else 'DIFFER')
    if r!=g:
        k=next((i for i in range(min(len(r),len(g))) if r[i]!=g[i]), min(len(r),len(g)))
        print('first difference at byte', k)
        open('$O/s_p2p_records.bin','wb').write(r)
except Exception as e: print('p2p records ERR', e)
P
)
if cmp -s $O/s_p2p.err $G/p2p_small/stderr.txt; then echo "p2p stderr identical" >> $O/s_summary.log; else echo "p2p stderr DIFFERS" >> $O/s_summary.log; diff $O/s_p2p.err $G/p2p_small/stderr.txt | head -20 >> $O/s_summary.log; fi
( time timeout 600 python -m pytest tests/test_host_dropin.py tests/test_host.py -x -q -m gpu ) > $O/s_pytest.log 2>&1; echo "pytest dropin rc=$?" >> $O/s_summary.log
cat $O/s_summary.log; tail -5 $O/s_pytest.log
