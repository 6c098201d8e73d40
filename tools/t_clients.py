import sys,time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from netsim_b200 import scenario as S, engine as E
from oracle import refrun
import numpy as np
ok_all=True
for name,sc in (("c40", S.p2p_clients(40, 4, play_seconds=60.0)),
                ("c300_sat", S.p2p_clients(300, 4, play_seconds=200.0, arrival_window=20.0)),
                ("c200_dist", S.p2p_clients(200, 6, play_seconds=100.0, metric=0, arrival_window=5.0, arrival_quantum=0.1)),
                ("c600_end", S.p2p_clients(600, 8, play_seconds=300.0, arrival_window=10.0, end_time=80.0))):
    b=sc.to_bytes()
    r=refrun.run_reference(b, trace=True, record=True)
    assert r['returncode']==0, r['stderr'][-800:]
    tr=r['trace']
    print(name,'ref events',r['stats']['events'],'types',np.bincount(tr['type']&0xff).tolist(),'records',len(r['records']))
    for mname,mode in (("seq",E.MODE_SEQ),("auto",E.MODE_AUTO)):
        e=E.Engine(b, mode=mode, records=True, trace=True)
        st=e.run(0)
        rec=e.records(); t2=e.trace()
        same=rec==r['records']
        c=e.counters()
        print('  ',mname,'status',st,'events',c['events'],'user',c['user'],'records same',same,'trace same',t2.tobytes()==tr.tobytes(), 'par',c['par_events'])
        if not same:
            ok_all=False
            a=np.frombuffer(rec,dtype=np.uint8); bb=np.frombuffer(r['records'],dtype=np.uint8)
            n=min(len(a),len(bb)); d=np.nonzero(a[:n]!=bb[:n])[0]
            print('     first diff at byte', d[0] if len(d) else n, 'of', len(a), len(bb))
            # first differing trace row
            m=min(len(t2),len(tr))
            for i in range(m):
                if t2[i].tobytes()!=tr[i].tobytes(): print('     trace row',i,'dev',t2[i],'ref',tr[i]); break
        e.close()
    # ensemble engines on the same script
    for mname,mode in (("ens cta",E.MODE_AUTO),("ens warp",E.MODE_SEQ)):
        ens=E.Ensemble([b,b], mode=mode, record_cap=1<<20, ring_cap=1<<16)
        ens.run(0); dg=ens.digests()
        print('  ',mname,[d['status'] for d in dg],[d['events'] for d in dg],'hash ok',[d['record_hash']==E.stream_hash(r['records']) for d in dg])
        ok_all &= all(d['record_hash']==E.stream_hash(r['records']) for d in dg)
        ens.close()
print('ALL OK' if ok_all else 'MISMATCH')
