"""Regenerate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref, built by `make -C oracle ref`
from /root/reference).  Each fixture holds the scenario script, the reference's record payload and
its fp64 event trace.  Run here (the container that has /root/reference); the GPU box only reads them.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from netsim_b200 import scenario as S  # noqa: E402
from oracle import refrun  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cases():
    for name, sc in S.reference_tests().items():
        yield name, sc, 0
    for seed in (1, 2, 3, 7, 11, 13, 21, 34, 55, 89):
        yield f"rand{seed}", S.random_small(seed, n_nodes=6 + seed % 7, n_flows=8 + seed % 17, relay=False), 0
    for seed in (1, 4, 6, 9, 12, 16):
        yield f"relay{seed}", S.random_small(seed, n_nodes=6 + seed % 7, n_flows=8 + seed % 17, relay=True), 0
    yield "swarm200", S.swarm(200, 3, n_catalogue=6, seeds_per_peer=3, mean_strips=3.0), 20000
    yield "churn150", S.churn(150, 4, rate=100.0, mean_session=1.5, sim_end=4.0), 0
    for name, sc in S.edge_cases().items():
        yield "edge_" + name, sc, 0


def main():
    assert refrun.available(), "build the reference first: make -C oracle ref"
    n = 0
    for name, sc, max_events in cases():
        b = sc.to_bytes()
        r = refrun.run_reference(b, max_events=max_events, trace=True, debug=False)
        if r["returncode"] != 0:
            print("skip", name, "(reference exits", r["returncode"], ")")
            continue
        np.savez_compressed(os.path.join(OUT, name + ".npz"), scenario=np.frombuffer(b, dtype=np.uint8),
                            records=np.frombuffer(r["records"], dtype=np.uint8), trace=r["trace"],
                            max_events=np.int64(max_events), events=np.int64(r["stats"]["events"]))
        n += 1
        print(name, len(b), len(r["records"]), len(r["trace"]))
    print(n, "fixtures")




def stderr_goldens():
    """stderr of the reference's own modules (what a user of `netsim modules/test0x` reads)."""
    import subprocess
    import tempfile
    for t in ("test00", "test01", "test02", "test03", "test04", "test05"):
        with tempfile.TemporaryDirectory() as d:
            p = subprocess.run([os.path.join(refrun.REF, "netsim"), os.path.join(refrun.REF, "modules", t)], cwd=d,
                               capture_output=True)
            open(os.path.join(OUT, t + ".stderr.txt"), "wb").write(p.stderr)


if __name__ == "__main__":
    main()
    stderr_goldens()
