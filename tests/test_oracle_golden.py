"""CPU: the oracle restatement (oracle/ns_oracle.c) against the golden vectors generated from the
compiled reference (tests/golden/make_golden.py), and — when oracle/_ref exists — live differential
runs against the reference itself.  Bit-exact: record bytes and fp64 trace rows."""
import numpy as np
import pytest

from conftest import golden_cases, pick_golden
from netsim_b200 import scenario as S
from oracle import pyoracle, refrun

ALL_CASES = golden_cases()
# scripts with client behaviour are outside the CPU restatement: their oracle is the compiled reference itself
CASES = [c for c in ALL_CASES if not c[0].startswith("clients_")]
CLIENT_CASES = [c for c in ALL_CASES if c[0].startswith("clients_")]


@pytest.mark.parametrize("name,z", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_golden(name, z, built):
    o = pyoracle.Oracle(z["scenario"].tobytes())
    st = o.run(int(z["max_events"]))
    assert st == 0
    assert o.records() == z["records"].tobytes()
    assert o.trace().tobytes() == z["trace"].tobytes()
    assert o.counters()["events"] == int(z["events"])


def test_golden_test05_is_the_reference_fixture():
    """test05's record payload is 837 bytes / 47 records with these ids."""
    z = dict(CASES)["test05"]
    rec = z["records"].tobytes()
    assert len(rec) == 837
    # first record: R_NODE_CREATE (minor 8) of node 1804289383 at t=0
    assert rec[0] == 0 and rec[1] == 8 and int.from_bytes(rec[2:6], "big") == 1804289383
    n, p = 0, 0
    while p < len(rec):
        b = rec[p + 12]
        p += 13 + (1 if b == 1 else 4 if b == 4 else 6)
        n += 1
    assert n == 47 and p == len(rec)


def test_restated_modules_match_known_answers(built):
    """Known answers of the reference's own modules: completion times."""
    exp = {"test00": [10000.2], "test01": [5000.2], "test02": [5000.2, 10000.2], "test03": [8000.2, 10000.3],
           "test04": [5000.2, 5020.4, 10000.65, 10010.55]}
    for name, times in exp.items():
        o = pyoracle.Oracle(S.reference_tests()[name].to_bytes())
        assert o.run(0) == 0
        tr = o.trace()
        done = tr["time"][(tr["type"] & 0xFF) == 1]
        assert np.allclose(done, times, rtol=0, atol=1e-6), (name, done)


@pytest.mark.skipif(not refrun.available(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("seed", range(24))
def test_oracle_vs_live_reference(seed, built):
    for relay in (False, True):
        sc = S.random_small(100 + seed, n_nodes=6 + seed % 7, n_flows=8 + seed % 17, relay=relay)
        b = sc.to_bytes()
        r = refrun.run_reference(b, trace=True)
        o = pyoracle.Oracle(b)
        st = o.run(0)
        if r["returncode"] != 0:
            # the reference abort()s exactly where the oracle reports it would (event.h:60-70)
            assert st != 0
            continue
        assert st == 0
        assert o.records() == r["records"]
        assert o.trace().tobytes() == r["trace"].tobytes()
        assert o.counters()["events"] == r["stats"]["events"]


def test_glibc_random_restatement(built):
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.random.restype = ctypes.c_long
    for seed in (1, 2, 12345, 4294967295):
        libc.srandom(seed)
        exp = [libc.random() for _ in range(500)]
        assert exp == list(pyoracle.glibc_random(seed, 500))


def test_qselect_restatement_against_definition(built):
    """qselect leaves the k smallest in front (any order) — the property qselect.h promises —
    and the oracle's order is deterministic."""
    rng = np.random.default_rng(5)
    for _ in range(200):
        n = int(rng.integers(1, 40))
        k = int(rng.integers(1, n + 1))
        val = rng.integers(-50, 50, n).astype(np.int32)
        node = np.arange(n, dtype=np.int32)
        v, nd = pyoracle.qselect(val, node, k)
        assert sorted(v[:k]) == sorted(val)[:k]
        assert sorted(nd) == list(range(n))
        assert all(val[nd[i]] == v[i] for i in range(n))


def test_picker_oracle_pinned_to_reference_golden(built):
    """oracle/ns_oracle.c's picker functions against what the reference's own qselect_eval (modules/qselect.h:5-27
    as instantiated at client_behaviour.h:63-65) and fetch metrics (p2p_common.h:171-198) produced
    (tests/golden/pick.npz, generated through oracle/ref_pick.c): same kept count, same ORDER."""
    z = pick_golden()
    L = pyoracle.lib()
    off, val, node, k = z["off"], z["val"], z["node"], z["k"]
    for q in range(len(k)):
        lo, hi = int(off[q]), int(off[q + 1])
        v, nd = val[lo:hi].copy(), node[lo:hi].copy()
        cnt = L.ns_oracle_pick_opt1(v.ctypes.data, nd.ctypes.data, hi - lo, int(k[q]))
        assert cnt == int(z["out_cnt"][q])
        assert (v == z["out_val"][lo:hi]).all() and (nd == z["out_node"][lo:hi]).all()
    import ctypes as C
    L.ns_oracle_metric.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    for r in z["metric"]:
        assert L.ns_oracle_metric(int(r["metric"]), int(r["tz_node"]), int(r["tz_client"]), float(r["total_up"]), float(r["max_up"])) == int(r["out"])


@pytest.mark.skipif(not refrun.available(), reason="compiled reference (oracle/_ref) not present")
def test_picker_oracle_vs_live_reference(built):
    """fresh random tables through oracle/_ref/libref_pick.so (the reference's code) and the restatement"""
    import ctypes as C
    import os
    R = C.CDLL(os.path.join(refrun.REF, "libref_pick.so"))
    L = pyoracle.lib()
    rng = np.random.default_rng(77)
    for _ in range(400):
        c = int(rng.integers(1, 200))
        k = int(rng.integers(1, c + 4))
        val = rng.integers(-40, 40, c).astype(np.int32)
        node = rng.permutation(c).astype(np.int32)
        v1, n1, v2, n2 = val.copy(), node.copy(), val.copy(), node.copy()
        assert R.ref_pick_opt1(v1.ctypes.data, n1.ctypes.data, c, k) == L.ns_oracle_pick_opt1(v2.ctypes.data, n2.ctypes.data, c, k)
        assert (v1 == v2).all() and (n1 == n2).all()


@pytest.mark.parametrize("name,z", CLIENT_CASES, ids=[c[0] for c in CLIENT_CASES])
def test_client_fixtures_are_what_the_reference_does(name, z, built):
    """the committed client-behaviour fixtures against a LIVE run of the compiled reference (its own
    modules/client_behaviour.c inside oracle/ref_scenario.c), optimised and assert-enabled builds"""
    if not refrun.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for debug in (False, True):
        r = refrun.run_reference(z["scenario"].tobytes(), max_events=int(z["max_events"]), trace=True, debug=debug)
        assert r["returncode"] == 0, r["stderr"]
        assert r["records"] == z["records"].tobytes()
        assert r["trace"].tobytes() == z["trace"].tobytes()
    # the fixtures do exercise the state machine: node-state records (minor 7) beyond the servers' N_SERVER ones
    rec = z["records"].tobytes()
    n_state, p = 0, 0
    while p < len(rec):
        minor, nb = rec[p + 1], int.from_bytes(rec[p + 12:p + 13], "big", signed=True)
        n_state += minor == 7
        p += 13 + (1 if nb == 1 else 4 if nb == 4 else 6)
    assert p == len(rec) and n_state > 20


def test_oracle_rejects_client_scripts(built):
    o = pyoracle.Oracle(CLIENT_CASES[0][1]["scenario"].tobytes())
    assert o.run(0) != 0
