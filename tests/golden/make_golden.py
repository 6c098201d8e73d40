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
    # behaviour library, client side (NS_SC_F_CLIENTS): the reference's OWN client_new_play1 / water-mark handlers run
    # inside the replay module (oracle/ref_scenario.c links modules/client_behaviour.c)
    yield "clients_play40", S.p2p_clients(40, 4, play_seconds=60.0), 0
    yield "clients_saturated", S.p2p_clients(90, 2, play_seconds=150.0, arrival_window=8.0, metric=1), 0
    yield "clients_distance", S.p2p_clients(70, 6, play_seconds=80.0, metric=0, arrival_window=3.0, arrival_quantum=0.1), 0
    yield "clients_cut", S.p2p_clients(120, 4, play_seconds=300.0, arrival_window=6.0, end_time=40.0), 0


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


def modx_golden():
    """tests/modules/modx.c compiled against the REFERENCE's headers and run by the reference core: its stderr and record file"""
    import subprocess
    import tempfile
    src = os.path.join(ROOT, "tests", "modules", "modx.c")
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["gcc", "-O2", "-g", "-DNDEBUG", "-ffp-contract=off", "-w", "-fPIC", "-shared", "-I/root/reference", src,
                        "-o", os.path.join(d, "modx.so")], check=True)
        p = subprocess.run([os.path.join(refrun.REF, "netsim"), os.path.join(d, "modx")], cwd=d, capture_output=True)
        assert p.returncode == 0
        open(os.path.join(OUT, "modx.stderr.txt"), "wb").write(p.stderr)
        open(os.path.join(OUT, "modx.rec"), "wb").write(refrun.record_payload(os.path.join(d, "RX")))


def pick_goldens():
    """Candidate tables through the reference's own qselect_eval (client_behaviour.h:63-65 over qselect.h:5-27)
    and fetch metrics (p2p_common.h:171-198), via oracle/ref_pick.c -> oracle/_ref/libref_pick.so."""
    import ctypes as C
    L = C.CDLL(os.path.join(refrun.REF, "libref_pick.so"))
    L.ref_pick_opt1.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.ref_metric.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    rng = np.random.default_rng(2024)
    offs, vals, nodes, ks = [0], [], [], []
    for q in range(600):
        c = int(rng.integers(0, 90)) if q % 50 else int(rng.integers(1100, 2600))   # some beyond the kernel's smem table
        # many ties ... the widest range on which the reference is DEFINED: its cmp is `a->val - b->val` in int
        # (client_behaviour.h:63), so |val| < 2^30 keeps every difference representable (beyond that the
        # reference is signed-overflow UB and gcc -O2 no longer wraps; real metrics are within +-4000)
        span = int(rng.choice([3, 30, 1000, 2 ** 30 - 1]))
        offs.append(offs[-1] + c)
        vals.extend(rng.integers(-span, span, c, dtype=np.int64).tolist())
        nodes.extend(rng.permutation(max(c, 1) * 3)[:c].tolist())
        ks.append(int(rng.integers(1, max(2, c + 3))))
    off, val, node, k = (np.array(x, dtype=np.int32) for x in (offs, vals, nodes, ks))
    ov, on, oc = val.copy(), node.copy(), np.zeros(len(ks), dtype=np.int32)
    for q in range(len(ks)):
        lo, hi = int(off[q]), int(off[q + 1])
        v, nd = val[lo:hi].copy(), node[lo:hi].copy()
        oc[q] = L.ref_pick_opt1(v.ctypes.data, nd.ctypes.data, hi - lo, int(k[q]))
        ov[lo:hi], on[lo:hi] = v, nd
    # metrics: (metric, tz_node, tz_client, total_up, max_up) -> value
    mq = np.zeros(4000, dtype=[("metric", "<i4"), ("tz_node", "<i4"), ("tz_client", "<i4"), ("total_up", "<f8"), ("max_up", "<f8"), ("out", "<i4")])
    mq["metric"] = rng.integers(0, 2, len(mq))
    mq["tz_node"] = rng.integers(0, 24, len(mq))
    mq["tz_client"] = rng.integers(0, 24, len(mq))
    mq["total_up"] = np.round(rng.random(len(mq)) * 400000, 1) * (rng.random(len(mq)) < 0.9)
    mq["max_up"] = rng.choice([8000.0, 40000.0, 160000.0, 1000.0], len(mq))
    for i in range(len(mq)):
        mq["out"][i] = L.ref_metric(int(mq["metric"][i]), int(mq["tz_node"][i]), int(mq["tz_client"][i]), float(mq["total_up"][i]), float(mq["max_up"][i]))
    np.savez_compressed(os.path.join(OUT, "pick.npz"), off=off, val=val, node=node, k=k, out_val=ov, out_node=on, out_cnt=oc, metric=mq)
    print("pick", len(ks), "tables,", len(val), "candidates,", len(mq), "metric points")


# the behaviour library's own test modules and p2p itself, on the unmodified reference core: what the host core
# (netsim_b200/host) must reproduce byte for byte when the SAME sources are compiled against its headers
P2P_CFGS = {   # name: (seed for the time() shim, p2p.cfg fields in file order)
    #            max_rsrc nrh cnp smart nch use_client metric nsvr ncld nclnt end
    "small50":       (1, (20, 0, 0, 0, 0, 0, 0, 4, 1, 50, 6000)),
    "base200":       (2, (20, 0, 0, 0, 0, 0, 0, 4, 1, 200, 9000)),
    "push1":         (5, (20, 0, 0, 0, 1, 0, 1, 8, 8, 300, 8000)),
    "play2b":        (7, (20, 0, 1, 0, 0, 1, 0, 4, 4, 150, 5000)),      # client_new_play2 with clients as sources: relays, throttles
    "play2c":        (8, (20, 0, 1, 0, 0, 0, 0, 4, 4, 150, 5000)),
    "cloudb":        (4, (20, 1, 0, 1, 0, 0, 1, 4, 8, 200, 7000)),      # smart cloud, HOUR_PASS
    "cloudc":        (9, (20, 1, 0, 0, 1, 0, 0, 4, 8, 200, 7000)),
    "push2":         (6, (20, 0, 0, 0, 2, 0, 1, 8, 8, 200, 6000)),
    "evict3":        (11, (3, 0, 0, 0, 0, 0, 0, 4, 1, 150, 9000)),     # max_rsrc exceeded: node_del_resource + auto_delete (resource.c:49-66, flow.c:278-281)
    "evict4":        (13, (4, 0, 0, 0, 1, 0, 1, 4, 4, 200, 12000)),
    "config2_seed1": (1, (20, 0, 0, 0, 0, 0, 0, 4, 1, 1000, 36000)),    # SURVEY section 8d, config 2 as specified (slow on the host core)
}
P2P_KEYS = ("max_rsrc", "new_resource_handler", "client_new_play", "smart_cloud", "new_connection_handler", "use_client", "metric",
            "nsvr", "ncld", "nclnt", "end_simulation")


def client_goldens():
    import hashlib, json, subprocess, tempfile
    ref = os.path.join(ROOT, "oracle", "_ref")
    src = "/root/reference/modules"
    lib = [os.path.join(src, f) for f in ("client_behaviour.c", "resource.c", "gaussian.c")]
    with tempfile.TemporaryDirectory() as d:
        for k in (2, 3):   # test_client1.c dereferences s->user_data == NULL on the reference itself (client_new_connection): no golden
            so = os.path.join(d, f"tc{k}.so")
            subprocess.run(["gcc", "-O2", "-DNDEBUG", "-ffp-contract=off", "-w", "-fPIC", "-shared", "-I/root/reference", "-I" + src,
                            os.path.join(src, f"test_client{k}.c")] + lib + ["-lm", "-o", so], check=True)
            p = subprocess.run([os.path.join(ref, "netsim"), so[:-3]], cwd=d, capture_output=True, check=True)
            open(os.path.join(OUT, f"tc{k}.stderr.txt"), "wb").write(p.stderr)
    exp = {}
    for name, (seed, vals) in P2P_CFGS.items():
        cd = os.path.join(OUT, "p2p", name)
        os.makedirs(cd, exist_ok=True)
        open(os.path.join(cd, "p2p.cfg"), "w").write("".join(f"{k}={v}\n" for k, v in zip(P2P_KEYS, vals)))
        with tempfile.TemporaryDirectory() as d:
            import shutil
            shutil.copy(os.path.join(cd, "p2p.cfg"), d)
            env = dict(os.environ, NS_FAKE_TIME=str(seed), LD_PRELOAD=os.path.join(ref, "timeshim.so"))
            p = subprocess.run([os.path.join(ref, "netsim"), os.path.join(ref, "modules", "p2p")], cwd=d, capture_output=True, check=True, env=env)
            data = open(os.path.join(d, "p2p_record"), "rb").read()
            n = int.from_bytes(data[:4], "big")
            exp[name] = {"seed": seed, "record_bytes": n, "record_sha256": hashlib.sha256(data[4:4 + n]).hexdigest(),
                         "stderr_sha256": hashlib.sha256(p.stderr).hexdigest()}
            if len(p.stderr) < (1 << 20):
                open(os.path.join(cd, "stderr.txt"), "wb").write(p.stderr)
            else:
                exp[name]["slow"] = True
            print("p2p", name, n, "record bytes,", p.stderr.count(b"\n"), "log lines")
    json.dump(exp, open(os.path.join(OUT, "p2p", "expected.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pick":
        pick_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "clients":
        client_goldens()
        sys.exit(0)
    main()
    stderr_goldens()
    modx_golden()
    pick_goldens()
    client_goldens()
