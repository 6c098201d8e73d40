"""Small programs for the ncu captures of the kernels other than k_run_par (tools/gpu_o.sh): seq | ens | pick"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import bench
from netsim_b200 import engine as E
which = sys.argv[1]
if which == "seq":       # one config-2-shaped replica on the single-warp engine
    sc = bench.ens_script(argparse.Namespace(ens_kind="p2p1k", ens_peers=0), 0)
    e = E.Engine(sc.to_bytes(), mode=E.MODE_SEQ, records=True, record_cap=64 << 20)
    print("status", e.run(0), e.counters()["events"], e.device_ms)
elif which == "ens":     # 296 config-2-shaped replicas, CTA / warp engines in rounds
    scs = [bench.ens_script(argparse.Namespace(ens_kind="p2p1k", ens_peers=0), i).to_bytes() for i in range(296)]
    ens = E.Ensemble(scs, records=True, ring_cap=1 << 17, record_cap=2 << 20)
    ens.run(0)
    print("events", sum(d["events"] for d in ens.digests()), ens.device_ms)
elif which == "pick":
    rng = np.random.default_rng(1)
    n_req, c = 100000, 64
    off = (np.arange(n_req + 1) * c).astype(np.int32)
    ov, on, oc, ms = E.pick_batch(1, off, rng.integers(-4000, 0, n_req * c).astype(np.int32), rng.integers(0, 100000, n_req * c).astype(np.int32), np.full(n_req, 32, dtype=np.int32))
    print("pick ms", ms)
