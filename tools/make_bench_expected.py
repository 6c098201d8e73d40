"""Expected record digest of bench.py's default workload, computed HERE by the CPU oracle (oracle/ns_oracle.c,
pinned to the compiled reference by tests/test_oracle_golden.py) and committed as tests/golden/bench_expected.json.
bench.py compares the device engine's digest after its last window against it (outside the timed region).
    python tools/make_bench_expected.py [ranks] [total_events]
    python tools/make_bench_expected.py ensemble [peers]        digests of sampled replicas of bench.py --workload ensemble
    python tools/make_bench_expected.py p2p1k                   digests of sampled replicas of --workload ensemble --ens-kind p2p1k (compiled reference)
    python tools/make_bench_expected.py churn [total_events]    digest of bench.py --workload churn after its last window
"""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(rank, total_events, peers, hubs, hub_frac):
    from netsim_b200 import engine as E
    from netsim_b200 import scenario as S
    from oracle import pyoracle
    sc = S.swarm_fast(peers, hubs, n_catalogue=64, seeds_per_peer=8, mean_strips=9.0, hub_frac=hub_frac, res_len=600000,
                      seed=12345 + rank, rand_seed=1 + rank)
    t0 = time.time()
    o = pyoracle.Oracle(sc.to_bytes(), trace=False, records=True)
    st = o.run(total_events)
    c = o.counters()
    rec = o.records()
    out = {"rank": rank, "status": st, "events": c["events"], "solves": c["solves"], "full_spreads": c["full_spreads"],
           "seq": c["seq"], "record_bytes": len(rec), "record_hash": E.stream_hash(rec), "oracle_seconds": round(time.time() - t0, 1)}
    o.close()
    return out


def ens_one(seed_index, peers):
    import argparse
    import bench
    from netsim_b200 import engine as E
    from oracle import pyoracle
    sc = bench.ens_script(argparse.Namespace(ens_peers=peers), seed_index)
    o = pyoracle.Oracle(sc.to_bytes(), trace=False, records=True)
    st = o.run(0)
    c = o.counters()
    rec = o.records()
    out = {"seed": seed_index, "status": st, "events": c["events"], "solves": c["solves"], "record_bytes": len(rec), "record_hash": E.stream_hash(rec)}
    o.close()
    return out


def churn_one(rank, total_events):
    import argparse
    import bench
    from netsim_b200 import engine as E
    from oracle import pyoracle
    a = argparse.Namespace(workload="churn", churn_users=1000000, churn_init=100000, churn_session=100.0, churn_end=60.0)
    sc = bench.churn_workload(a, rank)
    t0 = time.time()
    o = pyoracle.Oracle(sc.to_bytes(), trace=False, records=True)
    st = o.run(total_events)
    c = o.counters()
    rec = o.records()
    out = {"rank": rank, "status": st, "events": c["events"], "solves": c["solves"], "full_spreads": c["full_spreads"], "seq": c["seq"],
           "record_bytes": len(rec), "record_hash": E.stream_hash(rec), "oracle_seconds": round(time.time() - t0, 1)}
    o.close()
    return a, out


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "churn":
    total = int(sys.argv[2]) if len(sys.argv) > 2 else 13_000_000
    a, r = churn_one(0, total)
    path = os.path.join(ROOT, "tests", "golden", "bench_expected.json")
    cur = json.load(open(path))
    cur[f"churn,users={a.churn_users},init={a.churn_init},session={a.churn_session},end={a.churn_end},rank=0,events={total}"] = r
    print(r)
    json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    sys.exit(0)

def p2p_one(seed_index):
    """config-2-shaped replicas run the behaviour library: their digest comes from the COMPILED REFERENCE (oracle/_ref)"""
    import argparse
    import bench
    from netsim_b200 import engine as E
    from oracle import refrun
    sc = bench.ens_script(argparse.Namespace(ens_kind="p2p1k", ens_peers=0), seed_index)
    r = refrun.run_reference(sc.to_bytes(), trace=False, record=True)
    assert r["returncode"] == 0, r["stderr"]
    return {"seed": seed_index, "status": 0, "events": r["stats"]["events"], "solves": None, "record_bytes": len(r["records"]),
            "record_hash": E.stream_hash(r["records"]), "source": "compiled reference (oracle/_ref/netsim + modules/scenario.so)"}


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "p2p1k":
    seeds = [0, 1, 2, 3, 7, 100, 511, 1023]
    with ProcessPoolExecutor(max_workers=4) as ex:
        res = list(ex.map(p2p_one, seeds))
    path = os.path.join(ROOT, "tests", "golden", "bench_expected.json")
    cur = json.load(open(path))
    for r in res:
        cur[f"ensemble,kind=p2p1k,seed={r['seed']}"] = r
        print(r)
    json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "ensemble":
    peers = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    seeds = [0, 1, 2, 3, 7, 100, 511, 1023]
    with ProcessPoolExecutor(max_workers=4) as ex:
        res = list(ex.map(ens_one, seeds, [peers] * len(seeds)))
    path = os.path.join(ROOT, "tests", "golden", "bench_expected.json")
    cur = json.load(open(path))
    for r in res:
        cur[f"ensemble,peers={peers},seed={r['seed']}"] = r
        print(r)
    json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    sys.exit(0)

if __name__ == "__main__":
    ranks = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    total = int(sys.argv[2]) if len(sys.argv) > 2 else 13_000_000
    peers, hubs, hub_frac = 100032, 320, 0.02
    with ProcessPoolExecutor(max_workers=min(ranks, 4)) as ex:
        res = list(ex.map(one, range(ranks), [total] * ranks, [peers] * ranks, [hubs] * ranks, [hub_frac] * ranks))
    path = os.path.join(ROOT, "tests", "golden", "bench_expected.json")
    try:
        cur = json.load(open(path))
    except Exception:
        cur = {}
    for r in res:
        cur[f"peers={peers},hubs={hubs},hub_frac={hub_frac},rank={r['rank']},events={total}"] = r
        print(r)
    json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
