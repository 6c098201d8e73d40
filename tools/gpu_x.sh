#!/bin/bash
# host-core time split on a mid-size p2p run (NS_HOST_PROF=1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
R=$PWD; H=$R/netsim_b200/host/_build
d=$(mktemp -d); cat > $d/p2p.cfg <<C
max_rsrc=20
new_resource_handler=0
client_new_play=0
smart_cloud=0
new_connection_handler=0
use_client=0
metric=0
nsvr=4
ncld=1
nclnt=1000
end_simulation=4200
C
cd $d; s=$(date +%s.%N); NS_HOST_PROF=1 NS_FAKE_TIME=1 LD_PRELOAD=$R/oracle/_ref/timeshim.so timeout 900 $H/netsim_dev $H/modules/p2p 2> err.txt | tee $R/gpurun_out/x_hostprof.log; e=$(date +%s.%N)
python3 -c "print('wall %.1f s'%($e-$s)); d=open('p2p_record','rb').read(); print('record bytes', int.from_bytes(d[:4],'big'))" | tee -a $R/gpurun_out/x_hostprof.log
