"""Scenario scripts: the workload description shared by the reference core, the CPU oracle and the
CUDA engine (binary layout: include/ns_scenario.h).

A scenario writes down, as data, what a netsim behaviour module does to the core through its API:
nodes and their capacities (modules/p2p.c:144-170), resources (sim.c:129-154), flows established at
init or from USER events (sim.c:42-94, modules/test02.c:31-33), flows started when another flow
completes (modules/test04.c:23-35), explicit closes (flow.c:241) and the end of simulation
(modules/p2p.c:50-53).  `reference_tests()` restates the reference's own scenario modules
test00..test05 in this form; `swarm()`/`churn()` are the synthetic configs of SURVEY.md §8d.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

MAGIC = 0x4353534E
VERSION = 1
OP_CREATE, OP_SCHED, OP_CLOSE, OP_END, OP_PLAY1 = 1, 2, 3, 4, 5
F_CLIENTS = 1
CK_PLAIN, CK_CLIENT, CK_SERVER = 0, 1, 2

OP_DTYPE = np.dtype([("kind", "<i4"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4"),
                     ("start", "<i8"), ("x", "<f8"), ("y", "<f8")])
RES_DTYPE = np.dtype([("owner", "<i4"), ("bit_rate", "<i4"), ("len", "<i8")])
COPY_DTYPE = np.dtype([("node", "<i4"), ("res", "<i4")])
ONDONE_DTYPE = np.dtype([("flow_ord", "<i4"), ("group", "<i4")])
CLIENT_DTYPE = np.dtype([("kind", "<i4"), ("time_zone", "<i4"), ("lowwm", "<i4"), ("highwm", "<i4")])


def create(src: int, dst: int, res: int, start: int, ub: float, delay: float):
    return (OP_CREATE, src, dst, res, start, float(ub), float(delay))


def sched(group: int, time: float, relative: bool = False):
    return (OP_SCHED, group, 0, 1 if relative else 0, 0, float(time), 0.0)


def close(flow_ord: int):
    return (OP_CLOSE, flow_ord, 0, 0, 0, 0.0, 0.0)


def end():
    return (OP_END, 0, 0, 0, 0, 0.0, 0.0)


def play1(client: int, res: int, metric: int = 0):
    """client_new_play1(rid[res], node[client]) with fetch metric 0 = distance, 1 = share (needs Scenario.clients)"""
    return (OP_PLAY1, client, metric, res, 0, 0.0, 0.0)


@dataclass
class Scenario:
    max_up: np.ndarray
    max_down: np.ndarray
    resources: np.ndarray                      # RES_DTYPE
    copies: np.ndarray                         # COPY_DTYPE
    groups: List[Sequence[tuple]] = field(default_factory=list)   # group 0 = init
    ondone: Dict[int, int] = field(default_factory=dict)          # flow ordinal -> group
    seed: int = 1
    # behaviour library, client side (NS_SC_F_CLIENTS): one CLIENT_DTYPE row per node, or None
    clients: np.ndarray | None = None
    # optional pre-packed op table (big generators build numpy directly)
    _ops: np.ndarray | None = None
    _group_off: np.ndarray | None = None

    @property
    def n_nodes(self) -> int:
        return int(self.max_up.shape[0])

    def packed_ops(self) -> Tuple[np.ndarray, np.ndarray]:
        if self._ops is not None:
            return self._ops, self._group_off
        off = np.zeros(len(self.groups) + 1, dtype="<i4")
        rows = []
        for g, ops in enumerate(self.groups):
            rows.extend(ops)
            off[g + 1] = len(rows)
        arr = np.zeros(len(rows), dtype=OP_DTYPE)
        for i, r in enumerate(rows):
            arr[i] = r
        return arr, off

    def n_creates(self) -> int:
        ops, _ = self.packed_ops()
        return int((ops["kind"] == OP_CREATE).sum())

    def to_bytes(self) -> bytes:
        ops, off = self.packed_ops()
        od = np.zeros(len(self.ondone), dtype=ONDONE_DTYPE)
        for i, (k, v) in enumerate(sorted(self.ondone.items())):
            od[i] = (k, v)
        n_groups = len(off) - 1
        flags = F_CLIENTS if self.clients is not None else 0
        hdr = struct.pack("<IIII6i2i", MAGIC, VERSION, self.seed & 0xFFFFFFFF, flags,
                          self.n_nodes, len(self.resources), len(self.copies), n_groups,
                          len(ops), len(od), 0, 0)
        return b"".join([
            hdr,
            np.ascontiguousarray(self.max_up, dtype="<f8").tobytes(),
            np.ascontiguousarray(self.max_down, dtype="<f8").tobytes(),
            np.ascontiguousarray(self.resources, dtype=RES_DTYPE).tobytes(),
            np.ascontiguousarray(self.copies, dtype=COPY_DTYPE).tobytes(),
            np.ascontiguousarray(off, dtype="<i4").tobytes(),
            np.ascontiguousarray(ops, dtype=OP_DTYPE).tobytes(),
            od.tobytes(),
            np.ascontiguousarray(self.clients, dtype=CLIENT_DTYPE).tobytes() if self.clients is not None else b"",
        ])

    def write(self, path: str) -> None:
        with open(path, "wb") as f:
            f.write(self.to_bytes())


def _mk(max_bw: Sequence[Tuple[float, float]], resources, copies=(), groups=(), ondone=None, seed=1):
    mu = np.array([m[0] for m in max_bw], dtype="<f8")
    md = np.array([m[1] for m in max_bw], dtype="<f8")
    res = np.zeros(len(resources), dtype=RES_DTYPE)
    for i, r in enumerate(resources):
        res[i] = (r[0], r[2] if len(r) > 2 else 0, r[1])
    cp = np.zeros(len(copies), dtype=COPY_DTYPE)
    for i, c in enumerate(copies):
        cp[i] = c
    return Scenario(mu, md, res, cp, [list(g) for g in groups], dict(ondone or {}), seed)


# --------------------------------------------------------------------------------------------
# The reference's own scenario modules, restated (node order == sim_create_node call order).
# --------------------------------------------------------------------------------------------
def reference_tests() -> Dict[str, Scenario]:
    L = 5000000
    t = {}
    # modules/test00.c:22-40 — server(5000) -> client(1000), ub 500, delay .2
    t["test00"] = _mk([(5000, 5000), (1000, 1000)], [(0, L)], groups=[[create(0, 1, 0, 0, 500, 0.2)]])
    # modules/test01.c:22-41 — same with ub 500000 (node-limited)
    t["test01"] = _mk([(5000, 5000), (1000, 1000)], [(0, L)], groups=[[create(0, 1, 0, 0, 500000, 0.2)]])
    # modules/test02.c:38-63 — second flow (USER @5000) after the first has closed
    t["test02"] = _mk([(5000, 5000), (1000, 1000), (1000, 1000)], [(0, L)],
                      groups=[[create(0, 1, 0, 0, 500000, 0.2), sched(1, 5000)],
                              [create(0, 2, 0, 0, 500000, 0.2)]])
    # modules/test03.c:38-63 — second flow (USER @2000) throttles the first (server max 1000)
    t["test03"] = _mk([(1000, 1000)] * 3, [(0, L)],
                      groups=[[create(0, 1, 0, 0, 500000, 0.2), sched(1, 2000)],
                              [create(0, 2, 0, 0, 500000, 0.2)]])
    # modules/test04.c:23-73 / test05.c:24-75 — s1->s2; when it completes: s1->c1 @0 and a USER
    # event .1 s later that starts s1->c2 @0 and s2->c1 @10000 (relay source, range merge).
    # nodes: 0=s1 1=s2 2=c1 3=c2
    t["test04"] = _mk([(1000, 1000)] * 4, [(0, L)],
                      groups=[[create(0, 1, 0, 0, 500000, 0.2)],
                              [create(0, 2, 0, 0, 500000, 0.2), sched(2, 0.1, relative=True)],
                              [create(0, 3, 0, 0, 500000, 0.2), create(1, 2, 0, 10000, 500000, 0.2)]],
                      ondone={0: 1})
    t["test05"] = t["test04"]
    return t


# --------------------------------------------------------------------------------------------
# Synthetic generators
# --------------------------------------------------------------------------------------------
def _tz(i: np.ndarray, n: int) -> np.ndarray:
    """time zone of node i of n, as modules/p2p.c:146,155,171 (24*i/n, integer division)."""
    return (24 * i.astype(np.int64)) // max(n, 1)


def _ub_delay(tz_a: np.ndarray, tz_b: np.ndarray):
    """distance_based_bw / distance_based_delay, modules/p2p_common.h:181-189,214-221."""
    d = np.abs(tz_a - tz_b)
    return (4000 - 100 * d).astype("<f8"), 0.02 * d.astype("<f8") + 0.1


def swarm(n_peers: int, n_hubs: int, n_catalogue: int = 16, seeds_per_peer: int = 4,
          mean_strips: float = 9.0, max_strips: int = 32, hub_frac: float = 0.1,
          res_len: int = 600000, arrival_window: float = 0.0, arrival_quantum: float = 0.01,
          seed: int = 12345, end_time: float | None = None, rearm: bool = False) -> Scenario:
    """SURVEY.md §8d config 3 as a script: `n_hubs` seeders (160000/160000 Kbit/s, hold the whole
    catalogue) and `n_peers` peers (8000 up / 40000 down) that each seed `seeds_per_peer` catalogue
    entries and fetch one they do not hold, in 1+Poisson(mean_strips) (<= max_strips) strips from
    distinct holders (Zipf(1) over hubs with probability `hub_frac`, uniform seeding peers
    otherwise).  ub/delay follow p2p_common.h distance rules.  Arrivals are at t=0 (init) when
    arrival_window == 0, else USER events on an `arrival_quantum` grid inside the window.
    """
    rng = np.random.default_rng(seed)
    n = n_hubs + n_peers
    idx = np.arange(n)
    is_hub = idx < n_hubs
    tz = np.where(is_hub, _tz(idx, n_hubs), _tz(idx - n_hubs, n_peers))
    mu = np.where(is_hub, 160000.0, 8000.0)
    md = np.where(is_hub, 160000.0, 40000.0)
    res = np.zeros(n_catalogue, dtype=RES_DTYPE)
    res["owner"] = rng.integers(0, n_hubs, n_catalogue)
    res["len"] = res_len
    res["bit_rate"] = 2000
    # catalogue copies: every hub holds everything; each peer seeds a few
    holders: List[List[int]] = [[] for _ in range(n_catalogue)]
    cp = []
    for r in range(n_catalogue):
        for h in range(n_hubs):
            if h != res["owner"][r]:
                cp.append((h, r))
    seeds = np.stack([rng.permutation(n_catalogue)[:seeds_per_peer + 1] for _ in range(n_peers)]) \
        if n_peers else np.zeros((0, seeds_per_peer + 1), dtype=np.int64)
    for p in range(n_peers):
        for r in seeds[p, :seeds_per_peer]:
            cp.append((n_hubs + p, int(r)))
            holders[int(r)].append(n_hubs + p)
    holders_np = [np.array(h, dtype=np.int64) for h in holders]
    zipf_w = 1.0 / np.arange(1, n_hubs + 1)
    zipf_w /= zipf_w.sum()
    ops = []
    per_peer_ops: List[List[tuple]] = []
    for p in range(n_peers):
        node = n_hubs + p
        want = int(seeds[p, seeds_per_peer])          # a catalogue entry this peer does not seed
        k = int(min(max_strips, 1 + rng.poisson(mean_strips)))
        srcs: List[int] = []
        tries = 0
        while len(srcs) < k and tries < 8 * k:
            tries += 1
            if rng.random() < hub_frac or len(holders_np[want]) == 0:
                s = int(rng.choice(n_hubs, p=zipf_w))
            else:
                s = int(holders_np[want][rng.integers(0, len(holders_np[want]))])
            if s != node and s not in srcs:
                srcs.append(s)
        k = len(srcs)
        strip = res_len // k
        ub, dl = _ub_delay(tz[np.array(srcs)], np.full(k, tz[node]))
        per_peer_ops.append([create(srcs[i], node, want, i * strip, ub[i], dl[i]) for i in range(k)])
    groups: List[List[tuple]] = [[]]
    if arrival_window <= 0:
        for o in per_peer_ops:
            groups[0].extend(o)
    else:
        t_arr = np.floor(rng.random(n_peers) * arrival_window / arrival_quantum) * arrival_quantum
        order = np.argsort(t_arr, kind="stable")
        for p in order:
            groups.append(per_peer_ops[p])
            groups[0].append(sched(len(groups) - 1, float(t_arr[p])))
    if end_time is not None:
        groups.append([end()])
        groups[0].append(sched(len(groups) - 1, end_time))
    sc = Scenario(mu.astype("<f8"), md.astype("<f8"), res, np.array(cp, dtype=COPY_DTYPE).reshape(-1),
                  groups, {}, seed=1)
    return sc


def random_small(seed: int, n_nodes: int = 8, n_flows: int = 12, relay: bool = False,
                 closes: bool = True) -> Scenario:
    """Small randomized scripts for differential testing: a few sources holding full copies and
    "downloads" (one destination fetching one resource in 1..4 strips from distinct sources, all
    strips established together like client_new_play1, client_behaviour.c:382-385) started at
    init, from USER events on a coarse time grid (equal-time batches) or when another flow
    completes; optional explicit closes from USER events; optional relay (a download whose source
    is itself still being filled, exercising FLOW_SPEED_THROTTLE, flow.c:408-423)."""
    rng = np.random.default_rng(seed)
    caps = [(float(rng.choice([500, 1000, 2000, 8000])), float(rng.choice([500, 1000, 4000, 40000])))
            for _ in range(n_nodes)]
    n_src = max(2, n_nodes // 3)
    L = int(rng.choice([20000, 100000, 1000000]))
    n_res = 3
    resources = [(int(rng.integers(0, n_src)), L) for _ in range(n_res)]
    copies = [(s, r) for r in range(n_res) for s in range(n_src) if s != resources[r][0]]
    groups: List[List[tuple]] = [[]]
    ondone: Dict[int, int] = {}
    used_pairs = set()
    have = set()                   # (dst, res) already being fetched
    first_strip: List[Tuple[int, int, int]] = []   # (dst, res, flow ordinal of the strip at 0)
    ord_ = 0

    def download():
        nonlocal ord_
        for _ in range(50):
            r = int(rng.integers(0, n_res))
            d = int(rng.integers(n_src, n_nodes))
            if (d, r) in have:
                continue
            k = int(rng.integers(1, 5))
            srcs = [int(x) for x in rng.permutation(n_src)[:k] if (int(x), d) not in used_pairs]
            if not srcs:
                continue
            k = len(srcs)
            have.add((d, r))
            strip = L // k
            out = []
            for i, s in enumerate(srcs):
                used_pairs.add((s, d))
                ub = float(rng.choice([300, 700, 1500, 4000, 500000]))
                dl = float(rng.choice([0.1, 0.12, 0.2, 0.34]))
                out.append(create(s, d, r, i * strip, ub, dl))
                if i == 0:
                    first_strip.append((d, r, ord_))
                ord_ += 1
            return out
        return []

    groups[0].extend(download())
    grid = [float(x) for x in rng.choice([1.0, 2.5, 10.0, 10.0, 40.0, 100.0], size=8)]
    guard = 0
    while ord_ < n_flows and guard < 40:
        guard += 1
        g_ops = download()
        if not g_ops:
            continue
        by_done = rng.random() < 0.35 and ord_ - len(g_ops) > 0
        if by_done:
            k = int(rng.integers(0, ord_ - len(g_ops)))
            if k in ondone:
                by_done = False
        if not by_done and closes and ord_ > 2 and rng.random() < 0.3:
            g_ops.append(close(int(rng.integers(0, ord_))))
        groups.append(g_ops)
        g = len(groups) - 1
        if by_done:
            ondone[k] = g
        else:
            groups[0].append(sched(g, grid[int(rng.integers(0, len(grid)))]))
    if relay and first_strip:
        # a peer that is still downloading [0, ...) of r becomes the (only) source of another peer
        d0, r0, _ = first_strip[0]
        for d in range(n_src, n_nodes):
            if d != d0 and (d0, d) not in used_pairs and (d, r0) not in have:
                have.add((d, r0))
                used_pairs.add((d0, d))
                groups.append([create(d0, d, r0, 0, float(rng.choice([700, 4000, 500000])), 0.2)])
                groups[0].append(sched(len(groups) - 1, float(rng.choice([0.05, 1.0, 3.0]))))
                break
    return _mk(caps, resources, copies, groups, ondone, seed=1 + seed % 7)


def swarm_fast(n_peers: int = 100032, n_hubs: int = 320, n_catalogue: int = 64, seeds_per_peer: int = 8,
               mean_strips: float = 9.0, max_strips: int = 32, hub_frac: float = 0.1, res_len: int = 600000,
               seed: int = 12345, rand_seed: int = 1) -> Scenario:
    """Vectorised `swarm()` for the 100k-peer / ~1M-flow config: same shape
    (hubs hold the whole catalogue, peers seed `seeds_per_peer` entries and fetch one they lack in
    1+Poisson(mean_strips) strips from distinct holders, ub/delay by time-zone distance), all flows
    established at init like modules/p2p.c's init loop does for its first connections."""
    rng = np.random.default_rng(seed)
    n = n_hubs + n_peers
    idx = np.arange(n)
    is_hub = idx < n_hubs
    tz = np.where(is_hub, _tz(idx, n_hubs), _tz(idx - n_hubs, n_peers))
    mu = np.where(is_hub, 160000.0, 8000.0)
    md = np.where(is_hub, 160000.0, 40000.0)
    res = np.zeros(n_catalogue, dtype=RES_DTYPE)
    res["owner"] = rng.integers(0, n_hubs, n_catalogue)
    res["len"] = res_len
    res["bit_rate"] = 2000
    # every peer: a random permutation prefix of the catalogue -> seeds + the wanted entry
    perm = np.argsort(rng.random((n_peers, n_catalogue)), axis=1)[:, :seeds_per_peer + 1]
    seeds, want = perm[:, :seeds_per_peer], perm[:, seeds_per_peer]
    hub_cp = np.array([(h, r) for r in range(n_catalogue) for h in range(n_hubs) if h != res["owner"][r]],
                      dtype=np.int64).reshape(-1, 2)
    peer_cp = np.stack([np.repeat(np.arange(n_peers) + n_hubs, seeds_per_peer), seeds.reshape(-1)], axis=1)
    cp = np.zeros(len(hub_cp) + len(peer_cp), dtype=COPY_DTYPE)
    cp["node"] = np.concatenate([hub_cp[:, 0], peer_cp[:, 0]])
    cp["res"] = np.concatenate([hub_cp[:, 1], peer_cp[:, 1]])
    # holders per catalogue entry
    order = np.argsort(peer_cp[:, 1], kind="stable")
    hold_nodes = peer_cp[order, 0]
    hold_off = np.searchsorted(peer_cp[order, 1], np.arange(n_catalogue + 1))
    k = np.minimum(max_strips, 1 + rng.poisson(mean_strips, n_peers))
    kmax = int(k.max())
    zipf_w = 1.0 / np.arange(1, n_hubs + 1)
    zipf_w /= zipf_w.sum()
    use_hub = rng.random((n_peers, kmax)) < hub_frac
    hub_pick = rng.choice(n_hubs, size=(n_peers, kmax), p=zipf_w)
    cnt = (hold_off[want + 1] - hold_off[want])
    r_idx = (rng.random((n_peers, kmax)) * np.maximum(cnt, 1)[:, None]).astype(np.int64)
    peer_pick = hold_nodes[np.minimum(hold_off[want][:, None] + r_idx, len(hold_nodes) - 1)]
    src = np.where(use_hub | (cnt == 0)[:, None], hub_pick, peer_pick)
    me = (np.arange(n_peers) + n_hubs)[:, None]
    valid = (np.arange(kmax)[None, :] < k[:, None]) & (src != me)
    # drop duplicate sources within a row (keep first occurrence)
    srt = np.argsort(src, axis=1, kind="stable")
    ss = np.take_along_axis(src, srt, axis=1)
    dup_sorted = np.zeros_like(valid)
    dup_sorted[:, 1:] = ss[:, 1:] == ss[:, :-1]
    dup = np.zeros_like(valid)
    np.put_along_axis(dup, srt, dup_sorted, axis=1)
    # a duplicate is only a problem if the earlier twin is itself valid; being conservative is fine
    valid &= ~dup
    kk = valid.sum(axis=1)
    rows, cols = np.nonzero(valid)
    strip_no = (np.cumsum(valid, axis=1) - 1)[rows, cols]
    strip = res_len // np.maximum(kk, 1)
    ops = np.zeros(len(rows), dtype=OP_DTYPE)
    ops["kind"] = OP_CREATE
    ops["a"] = src[rows, cols]
    ops["b"] = rows + n_hubs
    ops["c"] = want[rows]
    ops["start"] = strip_no * strip[rows]
    ub, dl = _ub_delay(tz[ops["a"]], tz[ops["b"]])
    ops["x"] = ub
    ops["y"] = dl
    sc = Scenario(mu.astype("<f8"), md.astype("<f8"), res, cp, [], {}, seed=rand_seed)
    sc._ops = ops
    sc._group_off = np.array([0, len(ops)], dtype="<i4")
    return sc


def churn(n_users: int = 5000, n_servers: int = 16, rate: float = 1000.0, mean_session: float = 6.0,
          quantum: float = 0.01, res_len: int = 4000000, seed: int = 2024, sim_end: float | None = None) -> Scenario:
    """SURVEY §8d config 4 (test_client3-style churn) as a script: `n_users` clients arrive as a Poisson
    process of `rate`/s, each fetching the file from one of `n_servers` servers (server_picker1,
    client_behaviour.c:333-345: first server it is not connected to, here round-robin over zones) and
    leaving after an exponential session (its flow is closed if still alive).  Timestamps are quantised
    to `quantum` so that equal-timestamp batches hold several arrivals/departures plus the speed events
    they trigger (modules/test_client3.c:39-45 is one such arrival)."""
    rng = np.random.default_rng(seed)
    n = n_servers + n_users
    idx = np.arange(n)
    is_srv = idx < n_servers
    tz = np.where(is_srv, _tz(idx, n_servers), _tz(idx - n_servers, n_users))
    mu = np.where(is_srv, 160000.0, 8000.0)
    md = np.where(is_srv, 160000.0, 40000.0)
    res = np.zeros(1, dtype=RES_DTYPE)
    res["owner"] = 0
    res["len"] = res_len
    res["bit_rate"] = 2000
    cp = np.zeros(n_servers - 1, dtype=COPY_DTYPE)
    cp["node"] = np.arange(1, n_servers)
    cp["res"] = 0
    t_arr = np.floor(np.cumsum(rng.exponential(1.0 / rate, n_users)) / quantum) * quantum
    t_dep = np.floor((t_arr + rng.exponential(mean_session, n_users)) / quantum + 1) * quantum
    srv = rng.integers(0, n_servers, n_users)
    ub, dl = _ub_delay(tz[srv], tz[n_servers + np.arange(n_users)])
    groups: List[List[tuple]] = [[]]
    events = [(float(t_arr[u]), 0, u) for u in range(n_users)] + [(float(t_dep[u]), 1, u) for u in range(n_users)]
    events.sort()
    for t, kind, u in events:
        if kind == 0:
            groups.append([create(int(srv[u]), n_servers + u, 0, 0, float(ub[u]), float(dl[u]))])
        else:
            groups.append([close(u)])          # flow ordinal == arrival index (arrivals are created in time order)
        groups[0].append(sched(len(groups) - 1, t))
    if sim_end is not None:
        groups.append([end()])
        groups[0].append(sched(len(groups) - 1, sim_end))
    return Scenario(mu.astype("<f8"), md.astype("<f8"), res, cp, groups, {}, seed=1)


def p2p_clients(n_clients: int = 200, n_servers: int = 4, n_catalogue: int = 6, play_seconds: float = 120.0, bit_rate: int = 2000,
                arrival_window: float = 30.0, arrival_quantum: float = 0.05, metric: int = 1, lowwm: int = 0, highwm: int = 25000,
                seed: int = 7, rand_seed: int = 1, end_time: float | None = None) -> Scenario:
    """The client side of modules/p2p.c as a script (BASELINE.json configs[1] shape: servers + clients in server-only mode,
    client_new_play = 0): `n_servers` servers (160000/160000 Kbit/s, p2p.c:144-145) hold the whole catalogue, `n_clients`
    clients (8000 up / 40000 down, water marks 0 / 25000, p2p.c:166-173) each start ONE play at a USER event
    (NS_OP_PLAY1 = client_new_play1: server_picker_opt1 over the servers with the share or distance metric, nsvr/2 strips,
    client_start_play), then run the water-mark state machine of client_behaviour.c on every speed change until the
    resource has been played (DONE_PLAY).  Time zones 24*i/n as p2p.c:146,171."""
    rng = np.random.default_rng(seed)
    n = n_servers + n_clients
    idx = np.arange(n)
    is_srv = idx < n_servers
    tz = np.where(is_srv, _tz(idx, n_servers), _tz(idx - n_servers, n_clients))
    mu = np.where(is_srv, 160000.0, 8000.0)
    md = np.where(is_srv, 160000.0, 40000.0)
    res = np.zeros(n_catalogue, dtype=RES_DTYPE)
    res["owner"] = rng.integers(0, n_servers, n_catalogue)
    res["len"] = (play_seconds * (0.5 + rng.random(n_catalogue))).astype(np.int64) * bit_rate     # rlen*rbr, resource.c:41
    res["bit_rate"] = bit_rate
    cp = np.array([(s_, r) for r in range(n_catalogue) for s_ in range(n_servers) if s_ != res["owner"][r]], dtype=COPY_DTYPE).reshape(-1)
    cl = np.zeros(n, dtype=CLIENT_DTYPE)
    cl["kind"] = np.where(is_srv, CK_SERVER, CK_CLIENT)
    cl["time_zone"] = tz
    cl["lowwm"] = lowwm
    cl["highwm"] = highwm
    if arrival_window > 0:
        t_arr = np.floor(rng.random(n_clients) * arrival_window / arrival_quantum) * arrival_quantum
    else:                                         # modules/p2p.c:178: gaussian_noise_nz(900, 3600), i.e. N(3600 s, 900 s) clipped at 0, not quantised
        t_arr = np.maximum(0.0, rng.normal(3600.0, 900.0, n_clients))
    want = rng.integers(0, n_catalogue, n_clients)
    groups: List[List[tuple]] = [[]]
    for c_ in np.argsort(t_arr, kind="stable"):
        groups.append([play1(n_servers + int(c_), int(want[c_]), metric)])
        groups[0].append(sched(len(groups) - 1, float(t_arr[c_])))
    if end_time is not None:
        groups.append([end()])
        groups[0].append(sched(len(groups) - 1, end_time))
    return Scenario(mu.astype("<f8"), md.astype("<f8"), res, cp, groups, {}, seed=rand_seed, clients=cl)


def churn_fast(n_users: int = 1000000, n_servers: int = 1024, rate: float = 1000.0, mean_session: float = 600.0,
               quantum: float = 0.01, res_len: int = 4000000, seed: int = 2024, sim_end: float | None = None,
               init_users: int = 0) -> Scenario:
    """Vectorised `churn()` for SURVEY §8d config 4 at scale (1M users, 1024 servers, Poisson arrivals of `rate`/s,
    exponential sessions): the same script, built as numpy tables.  Group 0 schedules one USER event per arrival and per
    departure (in time order, arrivals before departures at equal times, like `churn()`); group k holds the one
    sim_establish_flow / flow_close of timeline entry k."""
    rng = np.random.default_rng(seed)
    n = n_servers + n_users
    idx = np.arange(n)
    is_srv = idx < n_servers
    tz = np.where(is_srv, _tz(idx, n_servers), _tz(idx - n_servers, n_users))
    mu = np.where(is_srv, 160000.0, 8000.0)
    md = np.where(is_srv, 160000.0, 40000.0)
    res = np.zeros(1, dtype=RES_DTYPE)
    res["owner"] = 0
    res["len"] = res_len
    res["bit_rate"] = 2000
    cp = np.zeros(n_servers - 1, dtype=COPY_DTYPE)
    cp["node"] = np.arange(1, n_servers)
    cp["res"] = 0
    # `init_users` users are already connected when the simulation starts (a warm swarm: their flows are established by the
    # module's init, like modules/p2p.c's first connections); the others arrive as a Poisson process
    I = int(init_users)
    t_arr = np.zeros(n_users)
    t_arr[I:] = np.floor(np.cumsum(rng.exponential(1.0 / rate, n_users - I)) / quantum) * quantum
    t_dep = np.floor((t_arr + rng.exponential(mean_session, n_users)) / quantum + 1) * quantum
    srv = rng.integers(0, n_servers, n_users)
    ub, dl = _ub_delay(tz[srv], tz[n_servers + np.arange(n_users)])
    if sim_end is not None:                       # nothing after the end of the simulation needs to be in the script
        n_a = int((t_arr < sim_end).sum())
    else:
        n_a = n_users
    users = np.arange(n_a)
    dep_ok = np.ones(n_a, dtype=bool) if sim_end is None else (t_dep[:n_a] < sim_end)
    arr_u = users[I:]
    ev_t = np.concatenate([t_arr[I:n_a], t_dep[:n_a][dep_ok]])
    ev_kind = np.concatenate([np.zeros(n_a - I, dtype=np.int64), np.ones(int(dep_ok.sum()), dtype=np.int64)])
    ev_u = np.concatenate([arr_u, users[dep_ok]])
    order = np.lexsort((ev_u, ev_kind, ev_t))     # (time, kind, user): the order of sorted((t, kind, u)) in churn()
    ev_t, ev_kind, ev_u = ev_t[order], ev_kind[order], ev_u[order]
    n_ev = len(ev_t)
    n_end = 1 if sim_end is not None else 0
    n0 = I + n_ev + n_end
    ops = np.zeros(n0 + n_ev + n_end, dtype=OP_DTYPE)
    # group 0: the warm swarm's flows (creation ordinal == user index), then the schedule
    ops["kind"][:I] = OP_CREATE
    ops["a"][:I] = srv[:I]
    ops["b"][:I] = n_servers + np.arange(I)
    ops["x"][:I] = ub[:I]
    ops["y"][:I] = dl[:I]
    ops["kind"][I:I + n_ev] = OP_SCHED
    ops["a"][I:I + n_ev] = 1 + np.arange(n_ev)
    ops["x"][I:I + n_ev] = ev_t
    if n_end:
        ops[I + n_ev] = (OP_SCHED, n_ev + 1, 0, 0, 0, float(sim_end), 0.0)
    # groups 1..n_ev: one op each
    body = ops[n0:n0 + n_ev]
    is_a = ev_kind == 0
    body["kind"] = np.where(is_a, OP_CREATE, OP_CLOSE)
    body["a"] = np.where(is_a, srv[ev_u], ev_u)   # close: flow ordinal == arrival index (arrivals are created in time order)
    body["b"] = np.where(is_a, n_servers + ev_u, 0)
    body["x"] = np.where(is_a, ub[ev_u], 0.0)
    body["y"] = np.where(is_a, dl[ev_u], 0.0)
    if n_end:
        ops[n0 + n_ev] = (OP_END, 0, 0, 0, 0, 0.0, 0.0)
    off = np.concatenate([[0], n0 + np.arange(n_ev + n_end + 1)]).astype("<i4")
    sc = Scenario(mu.astype("<f8"), md.astype("<f8"), res, cp, [], {}, seed=1)
    sc._ops = ops
    sc._group_off = off
    return sc


def edge_cases() -> Dict[str, Scenario]:
    """Hand-written corner cases of the module API (sim.c:42-67 error returns, main.c:50-52 exits)."""
    L = 100000
    t = {}
    # nothing to do: no flows, no events (main.c:50 loop never runs)
    t["empty"] = _mk([(1000, 1000), (1000, 1000)], [(0, L)], groups=[[]])
    # the source does not hold the resource: sim_establish_flow returns NULL (sim.c:49-53), no draw burned;
    # the second create (valid) must get the ids the reference gives it
    t["missing_resource"] = _mk([(1000, 1000), (1000, 1000), (1000, 1000)], [(0, L)],
                                groups=[[create(1, 2, 0, 0, 500, 0.1), create(0, 2, 0, 0, 500, 0.1)]])
    # a second flow into a range that already exists (sim.c:62-67) is refused
    t["existing_range"] = _mk([(4000, 4000), (4000, 4000), (1000, 1000)], [(0, L)], copies=[(1, 0)],
                              groups=[[create(0, 2, 0, 0, 500, 0.1), create(1, 2, 0, 0, 700, 0.2)]])
    # sim_end from a USER event while flows are still running; a relative USER event; closes of dead and
    # never-created flows are no-ops
    t["end_and_closes"] = _mk([(4000, 4000), (1000, 1000), (1000, 1000)], [(0, L)],
                              groups=[[create(0, 1, 0, 0, 500, 0.1), create(0, 2, 0, 0, 300, 0.2), sched(1, 5.0), sched(3, 50.0)],
                                      [close(0), close(0), close(7), sched(2, 2.5, relative=True)],
                                      [create(0, 1, 0, 50000, 800, 0.1)],
                                      [end()]])
    # zero-capacity receiver: the flow never gets a speed, no completion is ever scheduled
    t["zero_capacity"] = _mk([(1000, 1000), (0, 0)], [(0, L)], groups=[[create(0, 1, 0, 0, 500, 0.1)]])
    return t
