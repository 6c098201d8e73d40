"""Analyzer reductions: the numpy restatement oracle/pyanalyze.py against the reference's own `analyze`
binary (CPU, needs oracle/_ref), and the device reductions of netsim_b200/csrc/ns_ana.cu against the restatement (GPU)."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import golden_cases
from netsim_b200 import scenario as S
from oracle import pyanalyze, refrun

CLIENT_CASES = [c for c in golden_cases() if c[0].startswith("clients_")]


def _analyze(payload: bytes, *args: str) -> str:
    with tempfile.TemporaryDirectory(prefix="nsana") as d:
        path = os.path.join(d, "R")
        with open(path, "wb") as f:
            f.write(struct.pack(">I", len(payload)) + payload + b"\0" * 64)
        return subprocess.run([os.path.join(refrun.REF, "analyze"), path, *args], capture_output=True, check=True).stdout.decode()


def _long_run():
    """a config-2-shaped replica spanning two hours of simulated time (the analyzer's per-period reports use one hour)"""
    b = S.p2p_clients(300, 4, n_catalogue=10, play_seconds=900.0, arrival_window=0, metric=1, seed=1003, rand_seed=4).to_bytes()
    r = refrun.run_reference(b, trace=False, record=True)
    assert r["returncode"] == 0
    return b, r["records"]


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_restatement_matches_the_reference_analyzer(built):
    _, rec = _long_run()
    s = pyanalyze.summarize(rec, 3600.0)
    # node classes: `analyze R list_nodes` (analyzers.c:121-140)
    names = {"Server": 0, "Cloud": 1, "Client": 2}
    got = {int(a): names[b] for a, b in (ln.split() for ln in _analyze(rec, "list_nodes").splitlines())}
    assert got == s["classes"]
    # stale clients: `analyze R stale_client` prints (time, count) whenever the count changed and time moved on (analyzers.c:400-431)
    ser = [(float(a), int(b)) for a, b in (ln.split() for ln in _analyze(rec, "stale_client").splitlines())]
    assert max(c for _, c in ser) <= s["stale_max"]                  # their print skips changes inside one timestamp
    assert ser[-1][1] == s["stale_series"][-1][1] == s["stale_now"]
    mine = dict()
    for t, c in s["stale_series"]:
        mine[round(t, 3)] = c                                        # last value per timestamp
    for t, c in ser[1:]:
        assert round(t, 3) in mine
    # per-class bandwidth per hour: `analyze R node_type_speed <class> <dir> -h` prints "k average" (analyzers.c:153-180, :285-297)
    for cls_name, ci in (("server", 0), ("client", 2)):
        for dname, di in (("out", 0), ("in", 1)):
            rows = [ln.split() for ln in _analyze(rec, "node_type_speed", cls_name, dname, "-h").splitlines()]
            hours = {int(k): float(v) for k, v in rows}
            if not hours:
                continue
            for k in sorted(hours)[:-1]:                             # complete hours (the last line is the running partial one)
                assert hours[k] == pytest.approx(s["usage"][ci, di, k] / 3600.0, rel=1e-6, abs=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("mode_name", ["cta_per_replica", "warp_per_replica"])
def test_device_reductions_match_the_restatement(built, mode_name):
    from netsim_b200 import engine as E
    mode = E.MODE_AUTO if mode_name == "cta_per_replica" else E.MODE_SEQ
    scs = [z["scenario"].tobytes() for _, z in CLIENT_CASES]
    recs = [z["records"].tobytes() for _, z in CLIENT_CASES]
    scs.append(S.p2p_clients(300, 4, n_catalogue=10, play_seconds=900.0, arrival_window=0, metric=1, seed=1003, rand_seed=4).to_bytes())
    recs.append(None)
    for period in (3600.0, 7.5):
        ens = E.Ensemble(scs, mode=mode, record_cap=1 << 17, ring_cap=1 << 16)     # small record buffers: folded many times
        ens.analysis_begin(period)
        ens.run(0)
        dg = ens.digests()
        an = ens.analysis()
        for i, a in enumerate(an):
            assert dg[i]["status"] == 0
            rec = recs[i]
            if rec is None:
                if not refrun.available():
                    continue
                rec = refrun.run_reference(scs[i], trace=False, record=True)["records"]
            assert dg[i]["record_hash"] == E.stream_hash(rec)
            s = pyanalyze.summarize(rec, period)
            for k in ("records", "usage_records", "state_records", "stale_now", "stale_max"):
                assert a[k] == s[k], (i, k)
            assert a["t_last"] == s["t_last"]
            np.testing.assert_allclose(a["usage"], s["usage"], rtol=1e-9, atol=1e-6)
            np.testing.assert_allclose(a["stale"], s["stale"], rtol=1e-9, atol=1e-9)
            assert s["state_records"] > 0 and (period < 3600.0 or s["usage"].sum() > 0)
        ens.close()
