"""Stand-alone GPU parity sweep (also used by tests/test_gpu_parity.py): CUDA engine vs CPU oracle vs
compiled reference on the restated reference tests and randomized scripts."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from netsim_b200 import engine as E
from netsim_b200 import scenario as S
from oracle import pyoracle, refrun


def compare(name, sc, mode, max_events=0, use_ref=False, verbose=True):
    b = sc.to_bytes()
    o = pyoracle.Oracle(b)
    so = o.run(max_events)
    e = E.Engine(b, mode=mode, records=True, trace=True)
    se = e.run(max_events)
    so_m = -so if so else 0
    ok = True
    msg = []
    if se != so_m:
        ok = False
        msg.append(f"status {se} vs oracle {so_m}")
    if so == 0:
        if e.records() != o.records():
            ok = False
            msg.append("records differ")
        te, to = e.trace(), o.trace()
        if te.shape != to.shape or te.tobytes() != to.tobytes():
            ok = False
            n = min(len(te), len(to))
            d = [i for i in range(n) if te[i].tobytes() != to[i].tobytes()]
            msg.append(f"trace differs (len {len(te)} vs {len(to)}, first diff {d[:1]}: {te[d[0]] if d else ''} vs {to[d[0]] if d else ''})")
        ce, co = e.counters(), o.counters()
        for k in ("events", "solves", "full_spreads", "flow_visits", "speed_change", "flow_done", "drain", "throttle",
                  "user", "creates", "closes", "draws", "seq"):
            if ce[k] != co[k]:
                ok = False
                msg.append(f"counter {k}: {ce[k]} vs {co[k]}")
        if e.digest()["record_hash"] != E.stream_hash(o.records()):
            ok = False
            msg.append("digest differs")
        if use_ref and refrun.available():
            r = refrun.run_reference(b, max_events=max_events)
            if r["returncode"] == 0 and r["records"] != e.records():
                ok = False
                msg.append("records differ from compiled reference")
    if verbose or not ok:
        print(("ok   " if ok else "FAIL ") + name, "; ".join(msg), flush=True)
    e.close()
    o.close()
    return ok


def main():
    modes = [("seq", E.MODE_SEQ), ("auto", E.MODE_AUTO)]
    bad = 0
    t0 = time.time()
    for mname, mode in modes:
        for name, sc in S.reference_tests().items():
            bad += not compare(f"{name}[{mname}]", sc, mode, use_ref=True)
        for seed in range(int(os.environ.get("NS_PARITY_SEEDS", "60"))):
            for relay in (False, True):
                sc = S.random_small(seed, n_nodes=6 + seed % 7, n_flows=8 + seed % 17, relay=relay)
                bad += not compare(f"rand{seed}{'r' if relay else ''}[{mname}]", sc, mode, verbose=False, use_ref=(seed % 10 == 0))
        sc = S.swarm(300, 4, n_catalogue=8, seeds_per_peer=3, mean_strips=4.0)
        bad += not compare(f"swarm300[{mname}]", sc, mode, max_events=60000)
    print("failures:", bad, "elapsed", round(time.time() - t0, 1))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
