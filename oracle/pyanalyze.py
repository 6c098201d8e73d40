"""TEST INFRASTRUCTURE: the reference analyzer's reductions restated in numpy/python on a record payload —
node classes (analyzer/analyzers.c:80-101), per-class bandwidth integrals per period (:226-297 with :153-180) and the
stale-client count (:384-431) — to check the device reductions (netsim_b200/csrc/ns_ana.cu) against.  Pinned to the
reference's own `analyze` binary by tests/test_analysis.py.  Only tests may import this."""
from __future__ import annotations

import numpy as np


def parse(payload: bytes):
    """record_reader.c:28-77 -> list of (minor, id, time, value) for major == 0"""
    out = []
    p, n = 0, len(payload)
    while p < n:
        major, minor = payload[p], payload[p + 1]
        rid = int.from_bytes(payload[p + 2:p + 6], "big")
        t = int.from_bytes(payload[p + 6:p + 10], "big") + int.from_bytes(payload[p + 10:p + 12], "big") / 1000.0
        nb = int.from_bytes(payload[p + 12:p + 13], "big", signed=True)
        if nb == 1:
            v, sz = payload[p + 13], 14
        elif nb == 4:
            v, sz = int.from_bytes(payload[p + 13:p + 17], "big"), 17
        elif nb == 2:
            v, sz = int.from_bytes(payload[p + 13:p + 15], "big"), 15
        else:
            v, sz = int.from_bytes(payload[p + 13:p + 17], "big") + int.from_bytes(payload[p + 17:p + 19], "big") / 1000.0, 19
        if major == 0:
            out.append((minor, rid, t, v))
        p += sz
    assert p == n
    return out


def _spread(bucket, period, a, b, v):
    if not (b > a) or v == 0.0:
        return
    k = int(a / period)
    while a < b and k < len(bucket):
        e = min(b, (k + 1) * period)
        bucket[k] += v * (e - a)
        a = e
        k += 1


def summarize(payload: bytes, period: float, n_buckets: int = 24):
    recs = parse(payload)
    cls = {}                      # node id -> 0 server / 1 cloud / 2 client
    last_u = {}                   # (id, dir) -> (time, value)
    state = {}                    # id -> (state, since)
    usage = np.zeros((3, 2, n_buckets))
    stale = np.zeros(n_buckets)
    stale_now = stale_max = 0
    n_us = n_st = 0
    t_last = 0.0
    series = []                   # (time, stale count) whenever it changes, like sc_finish prints
    for minor, rid, t, v in recs:
        t_last = max(t_last, t)
        if minor == 8:
            cls[rid] = 2
        elif minor == 7:
            if v == 6:
                cls[rid] = 1
            elif v == 5:
                cls[rid] = 0
            if cls[rid] == 2:
                n_st += 1
                ps, since = state.get(rid, (-1, 0.0))
                if ps == 1:
                    _spread(stale, period, since, t, 1.0)
                if ps != v:
                    d = (v == 1) - (ps == 1)
                    if d:
                        stale_now += d
                        stale_max = max(stale_max, stale_now)
                        series.append((t, stale_now))
                state[rid] = (v, t)
        elif minor in (0, 1):
            n_us += 1
            pt, pv = last_u.get((rid, minor), (0.0, 0.0))
            _spread(usage[cls[rid], minor], period, pt, t, pv)
            last_u[(rid, minor)] = (t, v)
    for (rid, d), (pt, pv) in last_u.items():
        _spread(usage[cls[rid], d], period, pt, t_last, pv)
    for rid, (ps, since) in state.items():
        if cls[rid] == 2 and ps == 1:
            _spread(stale, period, since, t_last, 1.0)
    return {"records": len(recs), "usage_records": n_us, "state_records": n_st, "t_last": t_last, "usage": usage, "stale": stale,
            "stale_now": stale_now, "stale_max": stale_max, "classes": cls, "stale_series": series}
