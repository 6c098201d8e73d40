"""tests/emul/run_checks.py — TEST INFRASTRUCTURE.  Runs inside a process that has loaded tests/emul/_build/libnetsim_emul.so in
place of the CUDA library (NS_LIB) with the AddressSanitizer runtime preloaded: the device code of the single-warp engine, the
pickers, the analyzer kernel and the host layer, and both scopes of the batch engine, compiled for the lock-step warp model (cuda_runtime.h here), against the
committed goldens of the compiled reference and against the CPU oracle.  Prints one JSON line per check; any sanitizer report
aborts the process.

    python run_checks.py goldens:test00,test05 random:0,1 live:2,5 windows ensemble ensemble_cta analysis grid picks malformed args
    (EMU_SMS=3: CTAs of the grid engine; EMU_MODE=auto: random / live on the grid engine; EMU_ORDER=rev: lanes highest-first)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from netsim_b200 import engine as E  # noqa: E402
from netsim_b200 import scenario as S  # noqa: E402
from oracle import pyoracle  # noqa: E402

KEYS = ("events", "solves", "full_spreads", "flow_visits", "speed_change", "flow_done", "drain", "throttle", "user",
        "creates", "closes", "draws", "seq")


def out(name, ok, t0, **kw):
    print(json.dumps(dict(check=name, ok=bool(ok), sec=round(time.time() - t0, 1), **kw)), flush=True)


def goldens(names):
    for n in names:
        t0 = time.time()
        z = np.load(os.path.join(ROOT, "tests", "golden", n + ".npz"))
        e = E.Engine(z["scenario"].tobytes(), mode=E.MODE_SEQ, records=True, trace=True)
        st = e.run(0)
        ok = st == 0 and e.records() == z["records"].tobytes() and ("trace" not in z.files or e.trace().tobytes() == z["trace"].tobytes())
        ok = ok and ("events" not in z.files or e.counters()["events"] == int(z["events"]))
        out("golden " + n, ok, t0, events=e.counters()["events"])
        e.close()


def random_scripts(seeds):
    """tests/test_gpu_parity.py::test_engine_matches_oracle_random, one-warp engine, on the model"""
    for seed in seeds:
        for relay in (False, True):
            t0 = time.time()
            sc = S.random_small(200 + seed, n_nodes=6 + seed % 7, n_flows=8 + seed % 17, relay=relay)
            b = sc.to_bytes()
            o = pyoracle.Oracle(b)
            so = o.run(0)
            e = E.Engine(b, mode=E.MODE_AUTO if os.environ.get("EMU_MODE") == "auto" else E.MODE_SEQ, records=True, trace=True)   # EMU_MODE=auto: the grid engine
            se = e.run(0)
            ok = se == -so
            if so == 0:
                ce, co = e.counters(), o.counters()
                ok = (ok and e.records() == o.records() and e.trace().tobytes() == o.trace().tobytes() and all(ce[k] == co[k] for k in KEYS)
                      and e.flow_state(sc.n_creates()).tobytes() == o.flow_state(sc.n_creates()).tobytes()
                      and e.node_state(sc.n_nodes).tobytes() == o.node_state(sc.n_nodes).tobytes()
                      and e.digest()["record_hash"] == E.stream_hash(o.records()))
            out(f"random {seed} relay={int(relay)}", ok, t0, status=se)
            e.close()
            o.close()


def live(seeds):
    """fresh scripts — client behaviour (client_new_play1 + the water-mark machine), churn, swarms to completion — with the COMPILED
    REFERENCE run live beside the device code on the model (EMU_MODE=auto: the grid engine): records and event counts"""
    from oracle import refrun
    if not refrun.available():
        out("live reference: oracle/_ref not built", True, time.time())
        return
    mode = E.MODE_AUTO if os.environ.get("EMU_MODE") == "auto" else E.MODE_SEQ
    for seed in seeds:
        cases = {
            "p2p_clients": S.p2p_clients(40 + 7 * seed, 2 + seed % 5, n_catalogue=3 + seed % 4, play_seconds=60.0 + 10 * seed, arrival_window=8.0,
                                         arrival_quantum=0.05 * (1 + seed % 3), metric=seed % 2, seed=100 + seed, rand_seed=1 + seed),
            "churn": S.churn(n_users=120 + 10 * seed, n_servers=3 + seed % 3, rate=200.0, mean_session=0.4, quantum=0.01, res_len=200000, seed=2024 + seed, sim_end=1.5),
            "swarm": S.swarm(50 + 5 * seed, 2, n_catalogue=4, seeds_per_peer=2, mean_strips=2.0 + seed % 3, res_len=20000, seed=77 + seed),
        }
        for name, sc in cases.items():
            t0 = time.time()
            b = sc.to_bytes()
            saved = os.environ.pop("LD_PRELOAD", None)          # the reference runs without the sanitizer runtime
            try:
                r = refrun.run_reference(b, trace=False, record=True)
            finally:
                if saved is not None:
                    os.environ["LD_PRELOAD"] = saved
            e = E.Engine(b, mode=mode, records=True)
            st = e.run(0)
            ok = r["returncode"] == 0 and st == 0 and e.records() == r["records"] and e.counters()["events"] == r["stats"]["events"]
            out(f"live reference: {name} seed {seed}", ok, t0, events=e.counters()["events"])
            e.close()


def windows():
    """stop / resume in the middle of an equal-timestamp batch (max_events), and a record buffer far smaller than the stream
    (soft stop before a batch, drain, resume): the paths through seq_restore"""
    t0 = time.time()
    sc = S.swarm(60, 2, n_catalogue=4, seeds_per_peer=2, mean_strips=2.0)
    b = sc.to_bytes()
    o = pyoracle.Oracle(b)
    assert o.run(900) == 0
    e = E.Engine(b, mode=E.MODE_SEQ, records=True, trace=True)
    ok = e.run(307) == 0 and e.counters()["events"] == 307 and e.run(601) == 0 and e.run(900) == 0
    ce, co = e.counters(), o.counters()
    ok = ok and e.records() == o.records() and e.trace().tobytes() == o.trace().tobytes() and all(ce[k] == co[k] for k in KEYS)
    out("windows: three instalments", ok, t0, events=ce["events"])
    e.close()
    t0 = time.time()
    e = E.Engine(b, mode=E.MODE_SEQ, records=True, record_cap=24 << 10)
    stream = e.run_draining(900)
    ok = stream == o.records() and len(stream) > (24 << 10) and e.counters()["events"] == o.counters()["events"]
    out("windows: bounded record buffer drained", ok, t0, stream=len(stream))
    e.close()
    o.close()


def ensemble():
    """replicas on the one-warp engine: digests from the device against single runs of the oracle; the client-behaviour fixture
    as replicas with a record buffer that is folded in place, analysed on the device (k_ana) against the numpy restatement"""
    from oracle import pyanalyze
    t0 = time.time()
    scs = [S.random_small(300 + i, n_nodes=7, n_flows=10).to_bytes() for i in range(4)]
    ens = E.Ensemble(scs, mode=E.MODE_SEQ)
    ens.run(0)
    ok = True
    for d, b in zip(ens.digests(), scs):
        o = pyoracle.Oracle(b)
        so = o.run(0)
        ok = ok and d["status"] == -so
        if so == 0:
            ok = ok and d["events"] == o.counters()["events"] and d["solves"] == o.counters()["solves"] and d["record_bytes"] == len(o.records()) \
                and d["record_hash"] == E.stream_hash(o.records())
        o.close()
    cs = ens.counters_sum()
    out("ensemble: digests of 4 replicas", ok, t0, events=cs["events"])
    ens.close()
    t0 = time.time()
    z = np.load(os.path.join(ROOT, "tests", "golden", "clients_play40.npz"))
    b, rec = z["scenario"].tobytes(), z["records"].tobytes()
    ens = E.Ensemble([b, b], mode=E.MODE_SEQ, record_cap=1 << 12, ring_cap=1 << 12)
    ens.analysis_begin(7.5)
    ens.run(0)
    dg, an = ens.digests(), ens.analysis()
    s = pyanalyze.summarize(rec, 7.5)
    ok = True
    for d, a in zip(dg, an):
        ok = ok and d["status"] == 0 and d["record_hash"] == E.stream_hash(rec) and d["record_bytes"] == len(rec)
        ok = ok and all(a[k] == s[k] for k in ("records", "usage_records", "state_records", "stale_now", "stale_max")) and a["t_last"] == s["t_last"]
        ok = ok and np.allclose(a["usage"], s["usage"], rtol=1e-9, atol=1e-6) and np.allclose(a["stale"], s["stale"], rtol=1e-9, atol=1e-9)
    out("ensemble: client fixture folded + analysed on the device", ok, t0, records=int(an[0]["records"]))
    ens.close()


def analysis():
    """tests/test_analysis.py::test_device_reductions_match_the_restatement on the model: the four client-behaviour fixtures as replicas
    of one ensemble with small record buffers (folded in the kernel many times), two periods, one-warp and CTA engines"""
    from oracle import pyanalyze
    names = ("clients_cut", "clients_distance", "clients_play40", "clients_saturated")
    zs = [np.load(os.path.join(ROOT, "tests", "golden", n + ".npz")) for n in names]
    scs, recs = [z["scenario"].tobytes() for z in zs], [z["records"].tobytes() for z in zs]
    for mode, mname in ((E.MODE_SEQ, "one warp per replica"), (E.MODE_AUTO, "one CTA per replica")):
        for period in (3600.0, 7.5):
            t0 = time.time()
            ens = E.Ensemble(scs, mode=mode, record_cap=1 << 14, ring_cap=1 << 14)
            ens.analysis_begin(period)
            ens.run(0)
            dg, an = ens.digests(), ens.analysis()
            ok = True
            for d, a, rec in zip(dg, an, recs):
                s = pyanalyze.summarize(rec, period)
                ok = ok and d["status"] == 0 and d["record_hash"] == E.stream_hash(rec)
                ok = ok and all(a[k] == s[k] for k in ("records", "usage_records", "state_records", "stale_now", "stale_max")) and a["t_last"] == s["t_last"]
                ok = ok and np.allclose(a["usage"], s["usage"], rtol=1e-9, atol=1e-6) and np.allclose(a["stale"], s["stale"], rtol=1e-9, atol=1e-9)
            out(f"analysis on the device: {mname}, period {period}", ok, t0, records=int(sum(a["records"] for a in an)))
            ens.close()


def ensemble_cta(n_events=1500):
    """the batch engine at CTA scope (k_run_ens = ns_par_kernel.inc with __syncthreads barriers: ordered extraction, radix partition by
    (node, dir) group, one thread / one warp / the whole CTA per group, staged event_adds and records, ordered commit): small replicas,
    and a 300-peer swarm whose first timestamps are batches of hundreds of events with hub spreads"""
    t0 = time.time()
    scs = [S.random_small(300 + i, n_nodes=7, n_flows=10).to_bytes() for i in range(3)]
    ens = E.Ensemble(scs, mode=E.MODE_AUTO)
    ens.run(0)
    ok = True
    for d, b in zip(ens.digests(), scs):
        o = pyoracle.Oracle(b)
        so = o.run(0)
        ok = ok and d["status"] == -so and (so != 0 or (d["events"] == o.counters()["events"] and d["record_hash"] == E.stream_hash(o.records())))
        o.close()
    out("ensemble (CTA engine): small replicas", ok, t0)
    ens.close()
    t0 = time.time()
    b = S.swarm(300, 3, n_catalogue=6, seeds_per_peer=3, mean_strips=4.0, res_len=30000, seed=50).to_bytes()
    ens = E.Ensemble([b], mode=E.MODE_AUTO, record_cap=1 << 18, ring_cap=1 << 18)
    ens.run(n_events)
    d, cs = ens.digests()[0], ens.counters_sum()
    o = pyoracle.Oracle(b, trace=False)
    ok = o.run(n_events) == 0 and d["status"] == 0 and d["events"] == o.counters()["events"] == n_events and d["record_bytes"] == len(o.records()) \
        and d["record_hash"] == E.stream_hash(o.records()) and all(cs[k] == o.counters()[k] for k in ("solves", "full_spreads", "flow_visits", "draws", "seq"))
    ok = ok and cs["par_events"] == n_events and cs["max_batch"] > 256 and cs["full_spreads"] > 100      # it really went through the parallel pieces
    out("ensemble (CTA engine): swarm batches of hundreds of events", ok, t0, max_batch=cs["max_batch"], full_spreads=cs["full_spreads"])
    ens.close()
    o.close()


def grid(n_events=0):
    """the grid-scope batch engine (k_run_par / k_run_par_cl: cooperative launch, grid barriers, EMU_SMS CTAs of 256 threads) and the
    parallel module init (k_init_par: CREATE ops in conflict-free waves): goldens with NS_MODE_AUTO, and a 300-peer swarm run to
    completion — batches of up to 552 events, hub spreads, then FLOW_DONE / close / merge pieces and the hand-over to k_run_seq"""
    for n in ("test04", "test05", "rand2", "rand55", "relay9", "edge_end_and_closes", "churn150", "clients_play40"):
        t0 = time.time()
        z = np.load(os.path.join(ROOT, "tests", "golden", n + ".npz"))
        e = E.Engine(z["scenario"].tobytes(), mode=E.MODE_AUTO, records=True, trace=True)
        ok = e.run(0) == 0 and e.records() == z["records"].tobytes() and e.trace().tobytes() == z["trace"].tobytes()
        out("grid engine: golden " + n, ok, t0)
        e.close()
    t0 = time.time()
    b = S.swarm(300, 3, n_catalogue=6, seeds_per_peer=3, mean_strips=4.0, res_len=30000, seed=50).to_bytes()
    o = pyoracle.Oracle(b)
    assert o.run(n_events) == 0
    e = E.Engine(b, mode=E.MODE_AUTO, records=True, trace=True, trace_cap=100000)
    ok = e.run(n_events) == 0
    ce, co = e.counters(), o.counters()
    ok = ok and e.records() == o.records() and e.trace().tobytes() == o.trace().tobytes() and all(ce[k] == co[k] for k in KEYS)
    ok = ok and ce["waves"] > 0 and ce["par_events"] > 5000 and ce["max_batch"] > 256 and ce["flow_done"] > 500        # parallel init, parallel pieces, completions
    out("grid engine: 300-peer swarm to completion", ok, t0, events=ce["events"], par_events=ce["par_events"], max_batch=ce["max_batch"], launches=e.launches)
    e.close()
    o.close()


def picks():
    from conftest import pick_golden
    t0 = time.time()
    z = pick_golden()
    ov, on, oc, _ = E.pick_batch(1, z["off"], z["val"], z["node"], z["k"])
    out("picks: reference golden (k_pick / k_pick_small)", (oc == z["out_cnt"]).all() and (ov == z["out_val"]).all() and (on == z["out_node"]).all(), t0,
        requests=int(len(z["k"])))
    t0 = time.time()
    rng = np.random.default_rng(5)
    L = pyoracle.lib()
    ok = True
    for lo_c, hi_c, n in ((0, 70, 40), (65, 140, 6)):          # one request per thread (<= 64 entries, >= 32 requests) and one warp per request
        offs, vals, nodes, ks = [0], [], [], []
        for _ in range(n):
            c = int(rng.integers(lo_c, hi_c))
            offs.append(offs[-1] + c)
            vals.extend(rng.integers(-30, 30, c).tolist())
            nodes.extend(rng.permutation(1000)[:c].tolist())
            ks.append(int(rng.integers(0, 40)))
        off, val, node, k = (np.array(x, dtype=np.int32) for x in (offs, vals, nodes, ks))
        for mode in (1, 2):
            ov, on, oc, _ = E.pick_batch(mode, off, val, node, k)
            for q in range(n):
                lo, hi = off[q], off[q + 1]
                v, nd = val[lo:hi].copy(), node[lo:hi].copy()
                if mode == 1:
                    cnt = L.ns_oracle_pick_opt1(v.ctypes.data, nd.ctypes.data, hi - lo, int(k[q]))
                    ok = ok and oc[q] == cnt and (ov[lo:hi] == v).all() and (on[lo:hi] == nd).all()
                else:
                    u = np.ones(hi - lo, dtype=np.uint8)
                    i = L.ns_oracle_pick_opt2(v.ctypes.data, u.ctypes.data, hi - lo)
                    ok = ok and oc[q] == (1 if i >= 0 else 0) and (i < 0 or on[lo] == nd[i])
    out("picks: random tables against the oracle", ok, t0)


def malformed():
    import malformed as M
    t0 = time.time()
    z = np.load(os.path.join(ROOT, "tests", "golden", "test04.npz"))
    base = z["scenario"].tobytes()
    ok, n = True, 0
    for rule, m in M.mutations(base):
        for make in (lambda: E.Engine(m, mode=E.MODE_SEQ), lambda: E.Ensemble([base, m], mode=E.MODE_SEQ, record_cap=1 << 16, ring_cap=1 << 12)):
            try:
                make()
                ok = False
            except E.NsError as ex:
                ok = ok and ex.code == -1
            n += 1
    out("malformed scripts refused", ok, t0, cases=n)


def args():
    """tests/test_gpu_parity.py::test_c_abi_argument_checks on the model"""
    import ctypes as C
    import malformed as M
    t0 = time.time()
    sc = S.reference_tests()["test04"]
    z = np.load(os.path.join(ROOT, "tests", "golden", "test04.npz"))
    e = E.Engine(sc.to_bytes(), mode=E.MODE_SEQ, records=True, trace=True)
    good = dict(off=[0, 2, 3], cand_node=[0, 1, 2], cand_tz=[1, 2, 3], client_tz=[0, 5])
    ov, on, oc, _ = e.pick_live(1, 0, k=[1, 1], **good)
    ok = oc.tolist() == [1, 1] and sorted(on[:2].tolist()) == [0, 1] and on[2] == 2
    for kw, mode, metric in ((dict(good, off=[0, 3, 2]), 1, 0), (dict(good, off=[-1, 2, 3]), 1, 0), (dict(good, cand_node=[0, 1, sc.n_nodes]), 1, 0), (good, 3, 0), (good, 1, 2)):
        try:
            e.pick_live(mode, metric, **kw)
            ok = False
        except E.NsError as ex:
            ok = ok and ex.code == -10
    L = E.lib()
    L.ns_dev_exec.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ns_dev_exec.restype = C.c_int
    for kind, a in ((1, 99), (1, -1), (3, -1), (9, 99), (10, -1), (11, 0), (12, 0), (99, 0), (0, 0)):
        op = np.zeros(1, dtype=M.OP)
        op["kind"], op["a"] = kind, a
        ok = ok and L.ns_dev_exec(e.h, op.ctypes.data, 1) == -10
    ok = ok and e.run(0) == 0 and e.records() == z["records"].tobytes()
    out("C ABI argument checks", ok, t0)
    e.close()


if __name__ == "__main__":
    assert "emul" in E.LIB_PATH, "run_checks.py is for the lock-step model only (NS_LIB=tests/emul/_build/libnetsim_emul.so)"
    for a in sys.argv[1:]:
        name, _, arg = a.partition(":")
        if name == "goldens":
            goldens(arg.split(","))
        elif name == "random":
            random_scripts([int(x) for x in arg.split(",")])
        elif name == "live":
            live([int(x) for x in arg.split(",")])
        else:
            {"windows": windows, "ensemble": ensemble, "ensemble_cta": ensemble_cta, "analysis": analysis, "grid": grid, "picks": picks, "malformed": malformed, "args": args}[name]()
