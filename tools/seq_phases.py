"""Where a timestamp's time goes in the single-warp engine (profiling build: make -C netsim_b200/csrc PROF=1, NS_LIB=...):
globaltimer deltas accumulated by lane 0 around the phases of seq_run, on the config-2-shaped client script (mean batch 7.6)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from netsim_b200 import engine as E, scenario as S
sc = S.p2p_clients(1000, 4, n_catalogue=20, play_seconds=900.0, arrival_window=0, metric=1, seed=1000, rand_seed=1)
e = E.Engine(sc.to_bytes(), mode=E.MODE_SEQ, records=True, record_cap=64 << 20)
t0 = time.time(); st = e.run(0); w = time.time() - t0
a = (C.c_uint64 * 16)(); e.L.ns_dev_phase_ns(e.h, a)
c = e.counters()
names = ["find_tmin", "sparse_check", "collect_ring", "collect_l", "pre(ops,room)", "process", "loop_head"]
print(f"status {st} events {c['events']} batches {c['batches']} device {e.device_ms:.1f} ms -> {e.device_ms*1e3/c['batches']:.2f} us/batch, {e.device_ms*1e3/c['events']:.2f} us/event")
for i, n in enumerate(names):
    print(f"  {n:14s} {a[i]/1e6:9.2f} ms  {a[i]/1e3/c['batches']:7.2f} us/batch")
