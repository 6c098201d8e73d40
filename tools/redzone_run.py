"""Memory-safety pass (compute-sanitizer is closed on the GPU pool): the parity workloads once more with NS_REDZONE=1 — every device
table sits between two 256-byte red zones — and a check after every run that no red-zone byte changed.  Run by
tests/test_redzones.py in its own process (the switch is read once per process)."""
import os, sys
os.environ["NS_REDZONE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from netsim_b200 import engine as E, scenario as S
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import golden_cases

bad = []
n_runs = 0


def one(name, b, mode, max_events=0, expect=None, **kw):
    global n_runs
    e = E.Engine(b, mode=mode, records=True, trace=True, **kw)
    st = e.run(max_events)
    v = e.check()
    assert v >= 0, "red zones are not enabled"
    if v or st != 0 or (expect is not None and e.records() != expect):
        bad.append((name, mode, st, v))
    n_runs += 1
    e.close()


for name, z in golden_cases():
    for mode in (E.MODE_SEQ, E.MODE_AUTO):
        one(name, z["scenario"].tobytes(), mode, int(z["max_events"]), z["records"].tobytes())
for seed in range(12):
    for mode in (E.MODE_SEQ, E.MODE_AUTO):
        one(f"rand{seed}", S.random_small(1000 + seed, n_nodes=6 + seed % 7, n_flows=8 + seed % 17, relay=seed % 3 == 0).to_bytes(), mode)
# the grid engine on a swarm with hubs (parallel init, radix partition, chains, hub CTAs, staged commit with the overflow arena)
b = S.swarm_fast(20000, 64, n_catalogue=32, seeds_per_peer=6, hub_frac=0.05).to_bytes()
one("swarm20k", b, E.MODE_AUTO, 3000000, record_cap=400 << 20, ring_cap=1 << 24)
one("churn", S.churn(3000, 16, rate=1000.0, mean_session=2.0, sim_end=6.0).to_bytes(), E.MODE_AUTO, 0, record_cap=256 << 20)
one("soak", S.swarm(2000, 8, n_catalogue=8, seeds_per_peer=3, mean_strips=4.0, res_len=40000, arrival_window=5.0, arrival_quantum=0.02).to_bytes(), E.MODE_AUTO, 0, record_cap=256 << 20)
one("clients", S.p2p_clients(400, 4, n_catalogue=10, play_seconds=900.0, arrival_window=0, metric=1).to_bytes(), E.MODE_AUTO, 0, record_cap=64 << 20)
# ensembles, both engines, record folding + analysis in the kernel
scs = [S.swarm(300 + 40 * i, 3, n_catalogue=6, seeds_per_peer=3, mean_strips=4.0, res_len=30000, seed=50 + i).to_bytes() for i in range(6)]
scs += [S.p2p_clients(150, 4, play_seconds=200.0, arrival_window=10.0, seed=70 + i).to_bytes() for i in range(6)]
for mode in (E.MODE_AUTO, E.MODE_SEQ):
    ens = E.Ensemble(scs, mode=mode, record_cap=1 << 19, ring_cap=1 << 18)   # (a replica's module init must fit: 500 peers x 3 creates)
    ens.analysis_begin(30.0)
    ens.run(0)
    dg = ens.digests(); ens.analysis()
    v = ens.check()
    if v or any(d["status"] for d in dg):
        bad.append(("ensemble", mode, [d["status"] for d in dg], v))
    n_runs += 1
    ens.close()
print(f"redzone pass: {n_runs} runs, violations: {bad}")
sys.exit(1 if bad else 0)
