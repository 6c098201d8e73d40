"""Stand-alone GPU probe: throughput of the engine on swarm scripts + record-stream digest vs the oracle."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from netsim_b200 import engine as E
from netsim_b200 import scenario as S
from oracle import pyoracle


def probe(npeers, nhubs, max_events, modes, check=True, **kw):
    t0 = time.time()
    sc = S.swarm_fast(npeers, nhubs, **kw)
    b = sc.to_bytes()
    print(f"--- swarm {npeers}+{nhubs}: {sc.n_creates()} flows, script {len(b)>>20} MiB, gen {time.time()-t0:.1f}s", flush=True)
    ref_hash = None
    if check:
        t0 = time.time()
        o = pyoracle.Oracle(b, trace=False, records=True)
        t1 = time.time()
        o.run(max_events)
        t2 = time.time()
        co = o.counters()
        ref_hash = E.stream_hash(o.records())
        print(f"oracle: init {t1-t0:.2f}s run {t2-t1:.2f}s -> {co['events']/(t2-t1):.0f} ev/s; events {co['events']} solves {co['solves']} full {co['full_spreads']} visits {co['flow_visits']} seq {co['seq']}", flush=True)
        o.close()
    for mname, mode in modes:
        t0 = time.time()
        e = E.Engine(b, mode=mode, records=True)
        t1 = time.time()
        st = e.run(max_events)
        t2 = time.time()
        c = e.counters()
        d = e.digest()
        ok = (d["record_hash"] == ref_hash) if ref_hash is not None else None
        print(f"{mname}: status {st} create+init {t1-t0:.2f}s run wall {t2-t1:.3f}s device {e.device_ms:.1f} ms -> {c['events']/(e.device_ms/1e3):.0f} ev/s; "
              f"events {c['events']} par {c['par_events']} batches {c['batches']} maxbatch {c['max_batch']} segs {c['waves']} records {d['record_bytes']} hash_ok {ok}", flush=True)
        print('   phases(ms):', {k: round(v, 1) for k, v in e.phase_ms().items()}, flush=True)
        if check:
            for k in ("events", "solves", "full_spreads", "flow_visits", "seq", "draws"):
                if c[k] != co[k]:
                    print("   COUNTER MISMATCH", k, c[k], co[k])
        e.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    both = [("auto", E.MODE_AUTO), ("seq", E.MODE_SEQ)]
    if which == "small":
        probe(2000, 8, 500000, both, n_catalogue=16, seeds_per_peer=4)
        probe(10000, 32, 2000000, both, n_catalogue=32, seeds_per_peer=6)
    elif which == "mid":
        probe(10000, 32, 5000000, [("auto", E.MODE_AUTO)], n_catalogue=32, seeds_per_peer=6)
    elif which == "ensemble":
        nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 128
        npeers = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
        maxev = int(sys.argv[4]) if len(sys.argv) > 4 else 200000
        hub_frac = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1
        scs = [S.swarm_fast(npeers, max(2, npeers // 250), n_catalogue=16, seeds_per_peer=4, seed=100 + i, rand_seed=1 + i, hub_frac=hub_frac).to_bytes() for i in range(nrep)]
        o = pyoracle.Oracle(scs[3], trace=False); o.run(maxev)
        for mname, mode in (("cta/replica", E.MODE_AUTO), ("warp/replica", E.MODE_SEQ)):
            t0 = time.time()
            ens = E.Ensemble(scs, records=True, ring_cap=1 << 21, record_cap=8 << 20, mode=mode)
            t1 = time.time()
            ens.run(maxev)
            t2 = time.time()
            dg = ens.digests()
            ev = sum(d["events"] for d in dg)
            bad = [d["status"] for d in dg if d["status"] != 0]
            print(f"ensemble {nrep} x swarm{npeers} hub_frac {hub_frac} [{mname}]: create+init {t1-t0:.1f}s run device {ens.device_ms:.1f} ms wall {t2-t1:.2f}s -> {ev/(ens.device_ms/1e3):.0f} ev/s aggregate; "
                  f"events {ev}; launches {ens.launches}; bad status {bad[:5]}; replica 3 digest ok: {dg[3]['record_hash'] == E.stream_hash(o.records())} {dg[3]['events']} {o.counters()['events']}", flush=True)
            ens.close()
    elif which == "sparse":
        # small batches: a swarm run to completion (FLOW_DONE dominated tail) and a churn script
        for name, sc in (("complete2000", S.swarm(2000, 8, n_catalogue=8, seeds_per_peer=3, mean_strips=4.0, res_len=40000)),
                         ("churn5000", S.churn(5000, 16, rate=1000.0, mean_session=3.0, sim_end=10.0))):
            b = sc.to_bytes()
            t0 = time.time()
            o = pyoracle.Oracle(b, trace=False)
            o.run(0)
            to = time.time() - t0
            co = o.counters()
            print(f"--- {name}: events {co['events']} done {co['flow_done']} user {co['user']} oracle {to:.2f}s -> {co['events']/to:.0f} ev/s", flush=True)
            for mname, mode in both:
                e = E.Engine(b, mode=mode, records=True, record_cap=256 << 20)
                t1 = time.time()
                st = e.run(0)
                t2 = time.time()
                c = e.counters()
                print(f"{mname}: status {st} device {e.device_ms:.1f} ms -> {c['events']/(e.device_ms/1e3):.0f} ev/s; batches {c['batches']} par {c['par_events']} ok {e.records() == o.records()}", flush=True)
                e.close()
    elif which == "soak":
        # everything at once, to completion: arrivals (USER) on the speed-change lattice, full downloads
        # (FLOW_DONE, merges, closes), a record buffer that must be drained many times
        sc = S.swarm(4000, 8, n_catalogue=8, seeds_per_peer=3, mean_strips=5.0, res_len=60000, arrival_window=10.0, arrival_quantum=0.02)
        b = sc.to_bytes()
        o = pyoracle.Oracle(b, trace=False)
        o.run(0)
        co = o.counters()
        for mname, mode in both:
            e = E.Engine(b, mode=mode, records=True, record_cap=6 << 20)
            t0 = time.time()
            stream = e.run_draining(0)
            t1 = time.time()
            c = e.counters()
            bad = [k for k in ("events", "solves", "full_spreads", "flow_visits", "seq", "draws", "flow_done", "user", "closes") if c[k] != co[k]]
            print(f"soak {mname}: wall {t1-t0:.2f}s events {c['events']} done {c['flow_done']} user {c['user']} par {c['par_events']} launches {e.launches} "
                  f"records {len(stream)>>20} MB identical {stream == o.records()} counter mismatches {bad}", flush=True)
            e.close()
    elif which == "windows":
        sc = S.swarm_fast(100032, 320, n_catalogue=64, seeds_per_peer=8, hub_frac=0.02)
        e = E.Engine(sc.to_bytes(), mode=E.MODE_AUTO, records=True, record_cap=4 << 30)
        last = {k: 0 for k in e.phase_ms()}
        for i in range(1, 11):
            t0 = time.time()
            e.run(i * 2000000)
            w = time.time() - t0
            ph = e.phase_ms()
            print(f"window {i}: device {e.device_ms:.1f} ms", {k: (round(ph[k] - last[k], 1) if k.endswith('_ms') or k in ('argmin', 'extract', 'group', 'solve_warp', 'solve_thread', 'commit', 'kernel') else ph[k]) for k in ph if k not in ('sequential', 'tail', 'slowest_group_ms', 'slowest_group_degree', 'slowest_group_events')}, flush=True)
            last = ph
        e.close()
    elif which == "big":
        probe(100032, 320, int(sys.argv[2]) if len(sys.argv) > 2 else 5000000, [("auto", E.MODE_AUTO)], check=len(sys.argv) > 3, n_catalogue=64, seeds_per_peer=8, hub_frac=0.02)
