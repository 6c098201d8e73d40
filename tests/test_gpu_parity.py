"""GPU: the CUDA engine, called through the C ABI, against the CPU oracle on the same scripts, the
committed golden vectors from the compiled reference, and (when it travelled) the reference itself."""
import numpy as np
import pytest

from conftest import golden_cases, pick_golden
from netsim_b200 import engine as E
from netsim_b200 import scenario as S
from oracle import pyoracle, refrun

pytestmark = pytest.mark.gpu
CASES = golden_cases()
MODES = [("seq", E.MODE_SEQ), ("auto", E.MODE_AUTO)]
KEYS = ("events", "solves", "full_spreads", "flow_visits", "speed_change", "flow_done", "drain", "throttle", "user",
        "creates", "closes", "draws", "seq")


def run_engine(b, mode, max_events=0, **kw):
    e = E.Engine(b, mode=mode, records=True, trace=True, **kw)
    st = e.run(max_events)
    return e, st


@pytest.mark.parametrize("mname,mode", MODES)
@pytest.mark.parametrize("name,z", CASES, ids=[c[0] for c in CASES])
def test_engine_matches_reference_golden(name, z, mname, mode, built):
    e, st = run_engine(z["scenario"].tobytes(), mode, int(z["max_events"]))
    assert st == 0
    assert e.records() == z["records"].tobytes()          # byte-identical record stream
    assert e.trace().tobytes() == z["trace"].tobytes()    # bit-identical fp64 times / speeds
    assert e.counters()["events"] == int(z["events"])
    assert e.launches > 0
    e.close()


@pytest.mark.parametrize("mname,mode", MODES)
@pytest.mark.parametrize("seed", range(30))
def test_engine_matches_oracle_random(seed, mname, mode, built):
    for relay in (False, True):
        sc = S.random_small(200 + seed, n_nodes=6 + seed % 7, n_flows=8 + seed % 17, relay=relay)
        b = sc.to_bytes()
        o = pyoracle.Oracle(b)
        so = o.run(0)
        e, se = run_engine(b, mode)
        assert se == -so
        if so == 0:
            assert e.records() == o.records()
            assert e.trace().tobytes() == o.trace().tobytes()
            ce, co = e.counters(), o.counters()
            assert {k: ce[k] for k in KEYS} == {k: co[k] for k in KEYS}
            fs = e.flow_state(sc.n_creates())
            assert fs.tobytes() == o.flow_state(sc.n_creates()).tobytes()
            assert e.node_state(sc.n_nodes).tobytes() == o.node_state(sc.n_nodes).tobytes()
            assert e.digest()["record_hash"] == E.stream_hash(o.records())
        e.close()


@pytest.mark.parametrize("mname,mode", MODES)
def test_engine_swarm_window_and_resume(mname, mode, built):
    """A bounded window of a 1.5k-peer swarm (equal-timestamp batches of thousands of events), run in
    two instalments: stopping and resuming mid-batch must not change anything."""
    sc = S.swarm(1500, 6, n_catalogue=8, seeds_per_peer=3, mean_strips=5.0)
    b = sc.to_bytes()
    o = pyoracle.Oracle(b)
    assert o.run(150000) == 0
    e = E.Engine(b, mode=mode, records=True, trace=True, trace_cap=200000)
    assert e.run(70001) == 0
    assert e.counters()["events"] == 70001
    assert e.run(150000) == 0
    assert e.records() == o.records()
    assert e.trace().tobytes() == o.trace().tobytes()
    ce, co = e.counters(), o.counters()
    assert {k: ce[k] for k in KEYS} == {k: co[k] for k in KEYS}
    e.close()


@pytest.mark.skipif(not refrun.available(), reason="compiled reference did not travel")
@pytest.mark.parametrize("mname,mode", MODES)
def test_engine_vs_compiled_reference(mname, mode, built):
    sc = S.swarm(400, 4, n_catalogue=6, seeds_per_peer=3, mean_strips=4.0, arrival_window=2.0)
    b = sc.to_bytes()
    r = refrun.run_reference(b, max_events=50000, trace=True)
    assert r["returncode"] == 0
    e, st = run_engine(b, mode, 50000, trace_cap=100000)
    assert st == 0
    assert e.records() == r["records"]
    assert e.trace().tobytes() == r["trace"].tobytes()
    e.close()


def test_reference_abort_is_reported(built):
    """Scripts on which the reference abort()s (event.h:60-64) end with the matching error code."""
    found = 0
    for seed in range(40, 140):
        sc = S.random_small(seed, n_nodes=6 + seed % 7, n_flows=8 + seed % 17, relay=True)
        b = sc.to_bytes()
        o = pyoracle.Oracle(b)
        so = o.run(0)
        if so != 0:
            e, se = run_engine(b, E.MODE_SEQ)
            assert se == -so
            e.close()
            found += 1
            if found >= 3:
                break
    assert found >= 1


@pytest.mark.parametrize("mode", [E.MODE_AUTO, E.MODE_SEQ], ids=["cta_per_replica", "warp_per_replica"])
def test_ensemble_digests_match_single_runs(built, mode):
    scs = [S.random_small(300 + i, n_nodes=7, n_flows=10).to_bytes() for i in range(12)]
    ens = E.Ensemble(scs, mode=mode)
    ens.run(0)
    dig = ens.digests()
    for i, b in enumerate(scs):
        o = pyoracle.Oracle(b)
        so = o.run(0)
        assert dig[i]["status"] == -so
        if so == 0:
            assert dig[i]["events"] == o.counters()["events"]
            assert dig[i]["solves"] == o.counters()["solves"]
            assert dig[i]["record_bytes"] == len(o.records())
            assert dig[i]["record_hash"] == E.stream_hash(o.records())
    ens.close()


@pytest.mark.parametrize("mode", [E.MODE_AUTO, E.MODE_SEQ], ids=["cta_per_replica", "warp_per_replica"])
def test_ensemble_swarm_replicas_with_record_folding(built, mode):
    """replicas big enough for the batch engine (grouping, chains, hub spreads inside one CTA), to completion
    (FLOW_DONE, merges, closes), with a record buffer so small that it is folded into the running digest many times"""
    scs = [S.swarm(300 + 40 * i, 3, n_catalogue=6, seeds_per_peer=3, mean_strips=4.0, res_len=30000, seed=50 + i).to_bytes()
           for i in range(6)]
    ens = E.Ensemble(scs, mode=mode, record_cap=1 << 18, ring_cap=1 << 18)
    ens.run(0)
    dig = ens.digests()
    for i, b in enumerate(scs):
        o = pyoracle.Oracle(b, trace=False)
        assert o.run(0) == 0
        assert dig[i]["status"] == 0
        assert dig[i]["events"] == o.counters()["events"]
        assert dig[i]["record_bytes"] == len(o.records()) > (1 << 19)
        assert dig[i]["record_hash"] == E.stream_hash(o.records())
    ens.close()


def test_pick_batch_matches_reference_golden(built):
    """k_pick against the tables the reference's own qselect_eval left (tests/golden/pick.npz): same order"""
    z = pick_golden()
    ov, on, oc, ms = E.pick_batch(1, z["off"], z["val"], z["node"], z["k"])
    assert (oc == z["out_cnt"]).all()
    assert (ov == z["out_val"]).all() and (on == z["out_node"]).all()


def test_pick_batch_matches_qselect(built):
    rng = np.random.default_rng(11)
    offs, vals, nodes, ks = [0], [], [], []
    for _ in range(500):
        c = int(rng.integers(0, 70))
        offs.append(offs[-1] + c)
        vals.extend(rng.integers(-30, 30, c).tolist())
        nodes.extend(rng.permutation(1000)[:c].tolist())
        ks.append(int(rng.integers(1, 40)))
    big = 3000                                    # beyond the shared-memory staging size
    offs.append(offs[-1] + big)
    vals.extend(rng.integers(-1000, 1000, big).tolist())
    nodes.extend(range(big))
    ks.append(1500)
    off, val, node, k = (np.array(x, dtype=np.int32) for x in (offs, vals, nodes, ks))
    ov, on, oc, ms = E.pick_batch(1, off, val, node, k)
    for q in range(len(ks)):
        lo, hi = off[q], off[q + 1]
        v, nd = val[lo:hi].copy(), node[lo:hi].copy()
        cnt = pyoracle.lib().ns_oracle_pick_opt1(v.ctypes.data, nd.ctypes.data, hi - lo, int(k[q]))
        assert oc[q] == cnt
        assert (ov[lo:hi] == v).all() and (on[lo:hi] == nd).all()     # same order as the reference's partition
    # mode 2: first strict minimum, INT32_MAX marks unusable holders
    val2 = val.copy()
    val2[rng.random(len(val2)) < 0.3] = np.iinfo(np.int32).max
    ov, on, oc, ms = E.pick_batch(2, off, val2, node, None)
    for q in range(len(ks)):
        lo, hi = off[q], off[q + 1]
        usable = (val2[lo:hi] != np.iinfo(np.int32).max).astype(np.uint8)
        v = np.ascontiguousarray(val2[lo:hi])
        i = pyoracle.lib().ns_oracle_pick_opt2(v.ctypes.data, usable.ctypes.data, hi - lo)
        assert oc[q] == (1 if i >= 0 else 0)
        if i >= 0:
            assert on[lo] == node[lo + i]


@pytest.mark.parametrize("mname,mode", MODES)
@pytest.mark.parametrize("kw", [
    dict(n_peers=300, n_hubs=4, n_catalogue=6, seeds_per_peer=3, mean_strips=3.0, res_len=20000),
    dict(n_peers=600, n_hubs=4, n_catalogue=6, seeds_per_peer=3, mean_strips=3.0, res_len=40000, arrival_window=5.0),
], ids=["complete300", "arrivals600"])
def test_engine_runs_swarm_to_completion(kw, mname, mode, built):
    """File distribution to completion: every flow finishes (FLOW_DONE, range merges, closes with
    full spreads), arrivals come as USER events between speed-change batches."""
    sc = S.swarm(**kw)
    b = sc.to_bytes()
    o = pyoracle.Oracle(b)
    assert o.run(0) == 0
    e, st = run_engine(b, mode, 0, trace_cap=200000)
    assert st == 0
    assert e.records() == o.records()
    assert e.trace().tobytes() == o.trace().tobytes()
    ce, co = e.counters(), o.counters()
    assert {k: ce[k] for k in KEYS} == {k: co[k] for k in KEYS}
    assert ce["live_flows"] == 0 and ce["flow_done"] == co["closes"]
    e.close()


@pytest.mark.parametrize("mname,mode", MODES)
def test_engine_churn_arrivals_departures(mname, mode, built):
    """SURVEY §8d config 4 shape at test scale: Poisson arrivals/departures on a 10 ms grid —
    equal-timestamp batches mixing USER events (creates, closes) with the speed events they trigger."""
    sc = S.churn(2000, 8, rate=500.0, mean_session=4.0, sim_end=12.0)
    b = sc.to_bytes()
    o = pyoracle.Oracle(b)
    assert o.run(0) == 0
    e, st = run_engine(b, mode, 0, trace_cap=400000, record_cap=64 << 20)
    assert st == 0
    assert e.records() == o.records()
    assert e.trace().tobytes() == o.trace().tobytes()
    ce, co = e.counters(), o.counters()
    assert {k: ce[k] for k in KEYS} == {k: co[k] for k in KEYS}
    e.close()


def test_engine_mixed_batches_use_parallel_pieces(built):
    """Arrivals on a 20 ms grid land on the same timestamps as the speed-change cascades: one batch then
    holds USER events (flow creations) between long runs of SPEED_CHANGE events.  The engine must split
    it into pieces — parallel for the runs, sequential for the USER events — and stay bit-identical."""
    sc = S.swarm(1500, 6, n_catalogue=8, seeds_per_peer=3, mean_strips=5.0, arrival_window=1.0, arrival_quantum=0.02)
    b = sc.to_bytes()
    o = pyoracle.Oracle(b)
    assert o.run(400000) == 0
    e = E.Engine(b, mode=E.MODE_AUTO, records=True, trace=True, trace_cap=500000, record_cap=64 << 20)
    assert e.run(123457) == 0                      # stop inside a batch, resume
    assert e.run(400000) == 0
    assert e.records() == o.records()
    assert e.trace().tobytes() == o.trace().tobytes()
    ce, co = e.counters(), o.counters()
    assert {k: ce[k] for k in KEYS} == {k: co[k] for k in KEYS}
    assert ce["user"] > 100 and ce["par_events"] > 10000
    e.close()


@pytest.mark.parametrize("mname,mode", MODES)
def test_bounded_record_buffer_drains_without_loss(mname, mode, built):
    """A record buffer far smaller than the stream: the engine stops BEFORE a batch that might not fit
    (NS_E_RECORDS_FULL), the host drains and continues; the concatenated stream is the oracle's."""
    sc = S.swarm(1500, 6, n_catalogue=8, seeds_per_peer=3, mean_strips=5.0)
    b = sc.to_bytes()
    o = pyoracle.Oracle(b, trace=False)
    assert o.run(200000) == 0
    e = E.Engine(b, mode=mode, records=True, record_cap=3 << 20)
    stream = e.run_draining(200000)
    assert len(stream) > (3 << 20)                  # it really had to drain
    assert stream == o.records()
    assert e.counters()["events"] == 200000
    e.close()


@pytest.mark.parametrize("mname,mode", MODES)
@pytest.mark.parametrize("name", sorted(S.edge_cases()))
def test_engine_edge_cases(name, mname, mode, built):
    """Corner cases of the module API: refused flows (missing resource, existing range), an empty
    script, sim_end with flows in flight, closes of dead flows, a receiver without capacity."""
    b = S.edge_cases()[name].to_bytes()
    o = pyoracle.Oracle(b)
    assert o.run(0) == 0
    e, st = run_engine(b, mode)
    assert st == 0
    assert e.records() == o.records()
    assert e.trace().tobytes() == o.trace().tobytes()
    ce, co = e.counters(), o.counters()
    assert {k: ce[k] for k in KEYS} == {k: co[k] for k in KEYS}
    e.close()


def test_pick_live_uses_device_node_table(built):
    """Score map + selection on the device against the running simulation's node table: share_metric
    (p2p_common.h:190-198) depends on each candidate's CURRENT total_bwupbound[0]."""
    sc = S.swarm(600, 6, n_catalogue=6, seeds_per_peer=3, mean_strips=4.0)
    e = E.Engine(sc.to_bytes(), mode=E.MODE_AUTO, records=False)
    assert e.run(3000) == 0                            # stop while the flows are still up
    ns = e.node_state(sc.n_nodes)                      # usage0 usage1 total0 total1
    rng = np.random.default_rng(3)
    tz_of = rng.integers(0, 24, sc.n_nodes)
    offs, cn, ks, ctz = [0], [], [], []
    for _ in range(300):
        c = int(rng.integers(0, 40))
        offs.append(offs[-1] + c)
        cn.extend(rng.permutation(sc.n_nodes)[:c].tolist())
        ks.append(int(rng.integers(1, 20)))
        ctz.append(int(rng.integers(0, 24)))
    off, cn, ks, ctz = (np.array(x, dtype=np.int32) for x in (offs, cn, ks, ctz))
    ct = tz_of[cn].astype(np.int32)
    L = pyoracle.lib()
    for metric in (0, 1):
        exp_val = np.array([L.ns_oracle_metric(metric, int(ct[i]), int(ctz[q]), float(ns[cn[i], 2]), float(sc.max_up[cn[i]]))
                            for q in range(len(ks)) for i in range(off[q], off[q + 1])], dtype=np.int32)
        ov, on, oc, ms = e.pick_live(1, metric, off, cn, ct, ctz, k=ks)
        for q in range(len(ks)):
            lo, hi = off[q], off[q + 1]
            v, nd = exp_val[lo:hi].copy(), cn[lo:hi].copy()
            cnt = L.ns_oracle_pick_opt1(v.ctypes.data, nd.ctypes.data, hi - lo, int(ks[q]))
            assert oc[q] == cnt
            assert (ov[lo:hi] == v).all() and (on[lo:hi] == nd).all()
        usable = (rng.random(len(cn)) < 0.7).astype(np.uint8)
        ov, on, oc, ms = e.pick_live(2, metric, off, cn, ct, ctz, usable=usable)
        for q in range(len(ks)):
            lo, hi = off[q], off[q + 1]
            v = np.ascontiguousarray(exp_val[lo:hi])
            u = np.ascontiguousarray(usable[lo:hi])
            i = L.ns_oracle_pick_opt2(v.ctypes.data, u.ctypes.data, hi - lo)
            assert oc[q] == (1 if i >= 0 else 0)
            if i >= 0:
                assert on[lo] == cn[lo + i]
    assert (ns[:, 2] > 0).any()                        # the table really is live (flows add to total_bwupbound)
    e.close()


@pytest.mark.parametrize("mname,mode", MODES)
def test_engine_config1_shape_1k_clients_one_seeder(mname, mode, built):
    """BASELINE.json configs[1] in script form: 1 000 clients + 1 cloud seeder, arrivals spread over 30 s,
    every download runs to completion (234 k events, 4 036 flows, the seeder's 1 000-flow full spreads)."""
    sc = S.swarm(1000, 1, n_catalogue=4, seeds_per_peer=1, mean_strips=3.0, hub_frac=0.3, res_len=30000,
                 arrival_window=30.0, arrival_quantum=0.1)
    b = sc.to_bytes()
    o = pyoracle.Oracle(b, trace=False)
    assert o.run(0) == 0
    e = E.Engine(b, mode=mode, records=True, record_cap=64 << 20)
    assert e.run(0) == 0
    assert e.records() == o.records()
    ce, co = e.counters(), o.counters()
    assert {k: ce[k] for k in KEYS} == {k: co[k] for k in KEYS}
    assert ce["live_flows"] == 0 and ce["flow_done"] == 4036
    if refrun.available():
        r = refrun.run_reference(b, trace=False)
        assert r["returncode"] == 0 and r["records"] == e.records()
    e.close()


def test_engine_bench_scale_matches_oracle(built):
    """BASELINE.json configs[2] at the bench's own scale: the 100 032-peer / ~1M-flow script bench.py times
    (same generator arguments), first 6 M events in three instalments — lock-step batches of ~1 M events, the
    register-resident chains, the CTA-wide hub spreads and the parallel init all see their real sizes here.
    Record stream, digest and every counter must equal the CPU oracle's."""
    sc = S.swarm_fast(100032, 320, n_catalogue=64, seeds_per_peer=8, mean_strips=9.0, hub_frac=0.02, res_len=600000,
                      seed=12345, rand_seed=1)
    b = sc.to_bytes()
    o = pyoracle.Oracle(b, trace=False, records=True)
    assert o.run(6_000_000) == 0
    e = E.Engine(b, mode=E.MODE_AUTO, records=True, record_cap=(6_000_000 + 4 * sc.n_creates()) * 48 + (64 << 20), ring_cap=1 << 26)
    for upto in (2_000_000, 4_000_000, 6_000_000):
        assert e.run(upto) == 0
    ce, co = e.counters(), o.counters()
    assert {k: ce[k] for k in KEYS} == {k: co[k] for k in KEYS}
    assert ce["par_events"] == ce["events"]               # the whole range ran on the grid-parallel path
    ref = o.records()
    d = e.digest()
    assert d["record_bytes"] == len(ref)
    assert d["record_hash"] == E.stream_hash(ref)
    assert e.records() == ref
    assert e.node_state(sc.n_nodes).tobytes() == o.node_state(sc.n_nodes).tobytes()
    assert e.flow_state(sc.n_creates()).tobytes() == o.flow_state(sc.n_creates()).tobytes()
    e.close()
    o.close()


# ---------------------------------------------------------------- behaviour library, client side, on the device
CLIENT_CASES = [c for c in CASES if c[0].startswith("clients_")]


@pytest.mark.parametrize("mode", [E.MODE_AUTO, E.MODE_SEQ], ids=["cta_per_replica", "warp_per_replica"])
def test_client_behaviour_in_ensembles_matches_reference_fixtures(built, mode):
    """the client fixtures (records of the reference running its own modules/client_behaviour.c) as replicas of one ensemble"""
    scs = [z["scenario"].tobytes() for _, z in CLIENT_CASES]
    ens = E.Ensemble(scs, mode=mode, record_cap=1 << 16, ring_cap=1 << 16)       # record buffers folded several times
    ens.run(0)
    for (name, z), d in zip(CLIENT_CASES, ens.digests()):
        assert d["status"] == 0, name
        assert d["events"] == int(z["events"]), name
        assert d["record_bytes"] == len(z["records"]), name
        assert d["record_hash"] == E.stream_hash(z["records"].tobytes()), name
    ens.close()


@pytest.mark.parametrize("mname,mode", MODES)
@pytest.mark.parametrize("cfg", [dict(n_clients=300, n_servers=4, play_seconds=200.0, arrival_window=20.0, metric=1),
                                 dict(n_clients=200, n_servers=6, play_seconds=100.0, arrival_window=5.0, arrival_quantum=0.1, metric=0),
                                 dict(n_clients=600, n_servers=8, play_seconds=300.0, arrival_window=10.0, end_time=80.0),
                                 dict(n_clients=400, n_servers=4, n_catalogue=10, play_seconds=900.0, arrival_window=0, metric=1)],
                         ids=["saturated_share", "distance_metric", "cut_by_sim_end", "gauss_arrivals_config2_shape"])
def test_client_behaviour_matches_live_reference(built, cfg, mname, mode):
    """mid-size p2p client scripts (client_new_play1 + water-mark machine + node-state records on the device) against a LIVE
    run of the compiled reference, which executes the reference's own client_behaviour.c: records and fp64 trace identical"""
    if not refrun.available():
        pytest.skip("oracle/_ref not shipped")
    b = S.p2p_clients(**cfg).to_bytes()
    r = refrun.run_reference(b, trace=True, record=True)
    assert r["returncode"] == 0, r["stderr"]
    e, st = run_engine(b, mode, trace_cap=len(r["trace"]) + 16, record_cap=len(r["records"]) + (4 << 20))
    assert st == 0
    assert e.records() == r["records"]
    assert e.trace().tobytes() == r["trace"].tobytes()
    c = e.counters()
    assert c["events"] == r["stats"]["events"] and c["user"] == r["stats"]["user"] and c["user"] > cfg["n_clients"]
    e.close()


def test_pick_small_tables_one_request_per_thread(built):
    """the reference's real table sizes (c <= 64): k_pick_small, one quick-select per THREAD, against the same goldens"""
    z = pick_golden()
    off = z["off"]
    keep = [q for q in range(len(off) - 1) if off[q + 1] - off[q] <= 64]
    assert len(keep) > 300
    sel = np.concatenate([np.arange(off[q], off[q + 1]) for q in keep]).astype(np.int64)
    noff = np.concatenate([[0], np.cumsum([off[q + 1] - off[q] for q in keep])]).astype(np.int32)
    ov, on, oc, ms = E.pick_batch(1, noff, z["val"][sel], z["node"][sel], z["k"][keep])
    assert (oc == z["out_cnt"][keep]).all()
    assert (ov == z["out_val"][sel]).all() and (on == z["out_node"][sel]).all()
    # mode 2 (first strict minimum) on the same tables against the definition
    ov2, on2, oc2, _ = E.pick_batch(2, noff, z["val"][sel], z["node"][sel])
    v, nd = z["val"][sel], z["node"][sel]
    for i, q in enumerate(keep):
        lo, hi = noff[i], noff[i + 1]
        if hi == lo:
            assert oc2[i] == 0
            continue
        j = int(np.argmin(v[lo:hi]))
        assert oc2[i] == 1 and on2[lo] == nd[lo + j] and ov2[lo] == v[lo + j]


@pytest.mark.gpu
def test_host_view_matches_the_table_read_backs(built):
    """ns_dev_host_view (what the host core's mirrors are refreshed from, one kernel into pinned host memory) against the oracle-checked
    read-backs: live flows with their speeds by creation ordinal, node usage / total_bwupbound, one destination range per live flow."""
    sc = S.swarm(300, 3, n_catalogue=6, seeds_per_peer=3, mean_strips=4.0)
    b = sc.to_bytes()
    o = pyoracle.Oracle(b, trace=False)
    assert o.run(20000) == 0
    e = E.Engine(b, mode=E.MODE_SEQ, records=True)
    assert e.run(20000) == 0
    v = e.host_view()
    fs = o.flow_state(sc.n_creates())                       # per ordinal {alive, speed0, speed1, drng len}
    alive = np.nonzero(fs[:, 0] != 0)[0]
    fl = v["flows"][np.argsort(v["flows"]["ord"])]
    assert fl["ord"].tolist() == alive.tolist()
    assert fl["speed"].tobytes() == np.ascontiguousarray(fs[alive, 1:3]).tobytes()
    assert v["nodes"].tobytes() == o.node_state(sc.n_nodes).tobytes()
    rg = {int(r["slot"]): r for r in v["ranges"]}
    for f in fl:
        r = rg[int(f["drng"])]
        assert int(r["producer_ord"]) == int(f["ord"]) and float(r["len"]) == float(fs[int(f["ord"]), 3])
    e.close()
    o.close()


def test_malformed_scenarios_are_rejected_before_use(built):
    """ns_dev_create / ns_ens_create answer NS_E_BAD_SCENARIO for every malformed script of tests/malformed.py (counts, offsets, operands
    and tables the device would index with) and the library is unharmed afterwards."""
    import malformed as M
    z = dict(CASES)["test04"]
    base = z["scenario"].tobytes()
    for rule, m in M.mutations(base):
        for mode in (E.MODE_SEQ, E.MODE_AUTO):
            with pytest.raises(E.NsError) as ei:
                E.Engine(m, mode=mode)
            assert ei.value.code == -1, rule
        with pytest.raises(E.NsError) as ei:
            E.Ensemble([base, m], record_cap=1 << 16, ring_cap=1 << 12)
        assert ei.value.code == -1, rule
    e, st = run_engine(base, E.MODE_AUTO)
    assert st == 0 and e.records() == z["records"].tobytes()
    e.close()


def test_c_abi_argument_checks(built):
    """what the host hands to ns_dev_pick_live / ns_dev_pick_batch / ns_dev_exec is checked before a kernel indexes with it
    (NS_E_ARG), and a refused call leaves the simulation usable"""
    import ctypes as C
    import malformed as M
    sc = S.reference_tests()["test04"]
    z = dict(CASES)["test04"]
    e = E.Engine(sc.to_bytes(), mode=E.MODE_SEQ, records=True, trace=True)
    good = dict(off=[0, 2, 3], cand_node=[0, 1, 2], cand_tz=[1, 2, 3], client_tz=[0, 5])
    ov, on, oc, _ = e.pick_live(1, 0, k=[1, 1], **good)
    assert oc.tolist() == [1, 1] and sorted(on[:2].tolist()) == [0, 1] and on[2] == 2
    for kw, mode, metric in ((dict(good, off=[0, 3, 2]), 1, 0), (dict(good, off=[-1, 2, 3]), 1, 0), (dict(good, cand_node=[0, 1, sc.n_nodes]), 1, 0),
                             (dict(good, cand_node=[0, -1, 2]), 1, 0), (good, 3, 0), (good, 1, 2)):
        with pytest.raises(E.NsError) as ei:
            e.pick_live(mode, metric, **kw)
        assert ei.value.code == -10
    for mode, off in ((1, [0, 3, 2]), (3, [0, 2, 3])):
        with pytest.raises(E.NsError) as ei:
            E.pick_batch(mode, off, [5, 1, 3], [0, 1, 2], k=[1, 1])
        assert ei.value.code == -10
    L = E.lib()
    L.ns_dev_exec.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ns_dev_exec.restype = C.c_int
    for kind, a in ((1, 99), (1, -1), (3, -1), (9, 99), (10, -1), (11, 0), (12, 0), (99, 0), (0, 0)):   # CREATE, CLOSE, CREATE_RID, RANGE_SET, DELRES / AUTODEL without a host core, unknown kinds
        op = np.zeros(1, dtype=M.OP)
        op["kind"], op["a"] = kind, a
        assert L.ns_dev_exec(e.h, op.ctypes.data, 1) == -10, (kind, a)
    op = np.zeros(1, dtype=M.OP)
    assert L.ns_dev_exec(e.h, op.ctypes.data, -1) == -10 and L.ns_dev_exec(e.h, None, 1) == -10 and L.ns_dev_exec(e.h, op.ctypes.data, 0) == 0
    assert e.run(0) == 0 and e.records() == z["records"].tobytes()
    e.close()
