"""k_pick / k_pick_small throughput: many simultaneous NEW_CONNECTION requests with the reference's table sizes (c = nsvr = 64 or ncld)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from netsim_b200 import engine as E
rng = np.random.default_rng(1)
for n_req, c in ((100000, 64), (100000, 16), (4096, 1500)):
    off = (np.arange(n_req + 1) * c).astype(np.int32)
    val = rng.integers(-4000, 0, n_req * c).astype(np.int32)
    node = rng.integers(0, 100000, n_req * c).astype(np.int32)
    k = np.full(n_req, max(1, c // 2), dtype=np.int32)
    for it in range(3):
        t0 = time.time(); ov, on, oc, ms = E.pick_batch(1, off, val, node, k); w = time.time() - t0
    print(f"pick n_req {n_req} c {c}: kernel {ms:.3f} ms -> {n_req*c*8/ms/1e6:.1f} GB/s of candidate bytes read ({n_req/ms/1e3:.2f} M requests/s), call wall {w*1e3:.1f} ms")
