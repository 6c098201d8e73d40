"""ctypes binding of the device library (include/netsim_b200.h).  This is plumbing only: the product
is netsim_b200/libnetsim_b200.so (CUDA, sm_100a).  There is no CPU fallback — if the library or a
GPU is missing every call raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NS_LIB") or os.path.join(HERE, "libnetsim_b200.so")

COUNTERS = ["events", "solves", "full_spreads", "flow_visits", "speed_change", "flow_done", "drain",
            "throttle", "user", "creates", "closes", "draws", "seq", "live_flows", "peak_flows",
            "batches", "par_events", "max_batch", "waves", "retimes"]
MODE_AUTO, MODE_SEQ = 0, 1
TRACE_DTYPE = np.dtype([("type", "<u4"), ("ord", "<u4"), ("time", "<f8"), ("v0", "<f8"), ("v1", "<f8")])


class NsError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what}: {strerror(code)} ({code})")
        self.code = code


class Opts(C.Structure):
    _fields_ = [("device", C.c_int), ("mode", C.c_int), ("want_records", C.c_int), ("want_trace", C.c_int),
                ("ring_cap", C.c_uint64), ("record_cap", C.c_uint64), ("trace_cap", C.c_uint64),
                ("cap_flows", C.c_int), ("features_off", C.c_int), ("cap_ord", C.c_int), ("cap_ranges", C.c_int),
                ("cap_users", C.c_int), ("reserved", C.c_int * 1)]


class Digest(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("events", C.c_uint64), ("solves", C.c_uint64),
                ("record_bytes", C.c_uint64), ("record_hash", C.c_uint64), ("t_end", C.c_double),
                ("status", C.c_int64)]


ANA_BUCKETS = 24


class AnaSummary(C.Structure):
    _fields_ = [("off", C.c_uint64), ("records", C.c_uint64), ("usage_records", C.c_uint64), ("state_records", C.c_uint64),
                ("t_last", C.c_double), ("usage", C.c_double * ANA_BUCKETS * 2 * 3), ("stale", C.c_double * ANA_BUCKETS),
                ("stale_now", C.c_int32), ("stale_max", C.c_int32)]


class RandState(C.Structure):
    _fields_ = [("st", C.c_int32 * 31), ("f", C.c_int32), ("r", C.c_int32)]


_lib = None


def build(verbose: bool = False) -> None:
    subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NsError(-9, f"{LIB_PATH} missing: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        L.ns_dev_count.restype = C.c_int
        L.ns_dev_create.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t, C.POINTER(Opts)]
        L.ns_dev_run.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_float)]
        L.ns_dev_status.argtypes = [C.c_void_p]
        L.ns_dev_now.restype = C.c_double
        L.ns_dev_now.argtypes = [C.c_void_p]
        L.ns_dev_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.ns_dev_records.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.ns_dev_records_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.ns_dev_records_wait.argtypes = [C.c_void_p]
        L.ns_dev_records_drain.argtypes = [C.c_void_p]
        L.ns_dev_digest.argtypes = [C.c_void_p, C.POINTER(Digest)]
        L.ns_dev_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ns_dev_flow_state.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ns_dev_node_state.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ns_dev_phase_ns.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.ns_dev_host_view.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.ns_dev_check.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ns_ens_check.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ns_dev_launches.restype = C.c_uint64
        L.ns_dev_launches.argtypes = [C.c_void_p]
        L.ns_dev_destroy.argtypes = [C.c_void_p]
        L.ns_ens_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.POINTER(Opts)]
        L.ns_ens_run.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_float)]
        L.ns_ens_digests.argtypes = [C.c_void_p, C.POINTER(Digest), C.c_int]
        L.ns_ens_counters_sum.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.ns_ens_analysis_begin.argtypes = [C.c_void_p, C.c_double]
        L.ns_ens_analysis.argtypes = [C.c_void_p, C.POINTER(AnaSummary), C.c_int]
        L.ns_ens_launches.restype = C.c_uint64
        L.ns_ens_launches.argtypes = [C.c_void_p]
        L.ns_ens_destroy.argtypes = [C.c_void_p]
        L.ns_dev_pick_batch.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.POINTER(C.c_float)]
        L.ns_dev_pick_live.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 9 + [C.POINTER(C.c_float)]
        L.ns_rand_seed_host.argtypes = [C.POINTER(RandState), C.c_uint32]
        L.ns_rand_next_host.restype = C.c_uint32
        L.ns_rand_next_host.argtypes = [C.POINTER(RandState)]
        L.ns_rand_jump_host.argtypes = [C.POINTER(RandState), C.c_uint64]
        L.ns_strerror.restype = C.c_char_p
        L.ns_strerror.argtypes = [C.c_int]
        _lib = L
    return _lib


def strerror(code: int) -> str:
    try:
        return lib().ns_strerror(code).decode()
    except Exception:
        return "?"


def device_count() -> int:
    return lib().ns_dev_count()


class Engine:
    """One simulation resident on one GPU (ns_dev_*)."""

    def __init__(self, scenario_bytes: bytes, device: int = 0, mode: int = MODE_AUTO, records: bool = True,
                 trace: bool = False, ring_cap: int = 0, record_cap: int = 0, trace_cap: int = 0, cap_flows: int = 0):
        self.L = lib()
        o = Opts(device=device, mode=mode, want_records=int(records), want_trace=int(trace),
                 ring_cap=ring_cap, record_cap=record_cap, trace_cap=trace_cap, cap_flows=cap_flows)
        self.h = C.c_void_p()
        rc = self.L.ns_dev_create(C.byref(self.h), scenario_bytes, len(scenario_bytes), C.byref(o))
        if rc:
            self.h = None
            raise NsError(rc, "ns_dev_create")
        self.device_ms = 0.0

    def run(self, max_events: int = 0) -> int:
        ms = C.c_float()
        rc = self.L.ns_dev_run(self.h, max_events, C.byref(ms))
        self.device_ms = float(ms.value)
        if rc > 0:
            raise NsError(rc, "ns_dev_run")
        return rc

    @property
    def status(self) -> int:
        return self.L.ns_dev_status(self.h)

    @property
    def now(self) -> float:
        return self.L.ns_dev_now(self.h)

    def counters(self) -> dict:
        a = (C.c_uint64 * len(COUNTERS))()
        rc = self.L.ns_dev_counters(self.h, a)
        if rc:
            raise NsError(rc, "ns_dev_counters")
        return dict(zip(COUNTERS, [int(x) for x in a]))

    def records(self, offset: int = 0) -> bytes:
        n, tot = C.c_size_t(), C.c_size_t()
        self.L.ns_dev_records(self.h, None, 0, offset, C.byref(n), C.byref(tot))
        want = max(0, tot.value - offset)
        buf = C.create_string_buffer(max(1, want))
        rc = self.L.ns_dev_records(self.h, buf, want, offset, C.byref(n), C.byref(tot))
        if rc:
            raise NsError(rc, "ns_dev_records")
        return buf.raw[:n.value]

    def run_draining(self, max_events: int = 0) -> bytes:
        """run to the end (or max_events) with a bounded record buffer: every NS_E_RECORDS_FULL stop is
        answered by copying the records out and ns_dev_records_drain; returns the whole stream"""
        out, off = [], 0
        last = None
        while True:
            rc = self.run(max_events)
            chunk = self.records(off)
            out.append(chunk)
            off += len(chunk)
            if rc != -8:
                if rc != 0:
                    raise NsError(rc, "ns_dev_run")
                return b"".join(out)
            # a soft stop must have made progress since the previous one, or record_cap can never hold one batch
            progress = (self.counters()["events"], off)
            if progress == last:
                raise NsError(-13, "run_draining: no progress between two record-buffer stops (record_cap too small for one batch)")
            last = progress
            rc2 = self.L.ns_dev_records_drain(self.h)
            if rc2:
                raise NsError(rc2, "ns_dev_records_drain")

    def digest(self) -> dict:
        d = Digest()
        rc = self.L.ns_dev_digest(self.h, C.byref(d))
        if rc:
            raise NsError(rc, "ns_dev_digest")
        return {k: getattr(d, k) for k, _ in Digest._fields_}

    def trace(self) -> np.ndarray:
        rows = C.c_size_t()
        self.L.ns_dev_trace(self.h, None, 0, C.byref(rows))
        out = np.zeros(rows.value, dtype=TRACE_DTYPE)
        if rows.value:
            rc = self.L.ns_dev_trace(self.h, out.ctypes.data, rows.value, C.byref(rows))
            if rc:
                raise NsError(rc, "ns_dev_trace")
        return out

    def flow_state(self, max_flows: int) -> np.ndarray:
        out = np.zeros((max(1, max_flows), 4), dtype="<f8")
        n = C.c_size_t()
        rc = self.L.ns_dev_flow_state(self.h, out.ctypes.data, max_flows, C.byref(n))
        if rc:
            raise NsError(rc, "ns_dev_flow_state")
        return out[:n.value]

    VIEW_FLOW = np.dtype([("ord", "<i4"), ("slot", "<i4"), ("src", "<i4"), ("dst", "<i4"), ("srng", "<i4"), ("drng", "<i4"), ("drain_type", "<i4"), ("pad", "<i4"),
                          ("flow_id", "<u4"), ("pad2", "<u4"), ("speed", "<f8", 2), ("done_t", "<f8"), ("drain_t", "<f8")])
    VIEW_RANGE = np.dtype([("slot", "<i4"), ("res", "<i4"), ("producer_ord", "<i4"), ("next", "<i4"), ("prev", "<i4"), ("pad", "<i4"),
                           ("start", "<i8"), ("total_len", "<i8"), ("len", "<f8"), ("lenc", "<f8"), ("last_update", "<f8")])

    def host_view(self) -> dict:
        """ns_dev_host_view: what the host core's mirrors are refreshed from (live flows, live ranges, node usage)"""
        hdr, nd, fl, rg = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        rc = self.L.ns_dev_host_view(self.h, C.byref(hdr), C.byref(nd), C.byref(fl), C.byref(rg))
        if rc:
            raise NsError(rc, "ns_dev_host_view")
        h = np.ctypeslib.as_array((C.c_int32 * 4).from_address(hdr.value)).copy()
        hi_flow, hi_range, n_nodes = int(h[0]), int(h[1]), int(h[2])
        flows = np.frombuffer((C.c_char * (self.VIEW_FLOW.itemsize * hi_flow)).from_address(fl.value), dtype=self.VIEW_FLOW).copy() if hi_flow else np.zeros(0, self.VIEW_FLOW)
        ranges = np.frombuffer((C.c_char * (self.VIEW_RANGE.itemsize * hi_range)).from_address(rg.value), dtype=self.VIEW_RANGE).copy() if hi_range else np.zeros(0, self.VIEW_RANGE)
        nodes = np.frombuffer((C.c_char * (32 * n_nodes)).from_address(nd.value), dtype="<f8").reshape(n_nodes, 4).copy() if n_nodes else np.zeros((0, 4))
        return {"flows": flows[flows["slot"] >= 0], "ranges": ranges[ranges["slot"] >= 0], "nodes": nodes}

    def node_state(self, n_nodes: int) -> np.ndarray:
        out = np.zeros((max(1, n_nodes), 4), dtype="<f8")
        n = C.c_size_t()
        rc = self.L.ns_dev_node_state(self.h, out.ctypes.data, n_nodes, C.byref(n))
        if rc:
            raise NsError(rc, "ns_dev_node_state")
        return out[:n.value]

    def check(self) -> int:
        """red-zone violations (NS_REDZONE=1): 0 = clean, -1 = not enabled"""
        v = C.c_int()
        rc = self.L.ns_dev_check(self.h, C.byref(v))
        if rc:
            raise NsError(rc, "ns_dev_check")
        return int(v.value)

    def phase_ms(self) -> dict:
        a = (C.c_uint64 * 16)()
        self.L.ns_dev_phase_ns(self.h, a)
        ms = ["argmin", "extract", "sequential", "group", "solve_warp", "commit", "tail", "kernel", "solve_thread"]
        out = {n: a[i] / 1e6 for i, n in enumerate(ms)}
        # profiling aid of the warp pass: the slowest (node,dir) group seen in one segment
        out.update({"slowest_group_ms": a[10] / 1e6, "slowest_group_degree": int(a[11]), "slowest_group_events": int(a[12]),
                    "sum_slowest_ms": a[13] / 1e6, "deferred_groups": int(a[14]),
                    "sum_slowest_table_ms": a[9] / 1e6, "sum_slowest_hub_ms": a[15] / 1e6,
                    "hub_jobs": int(a[11]), "sum_last_hub_end_ms": a[12] / 1e6, "hub_cta_total_ms": a[6] / 1e6})
        return out

    def pick_live(self, mode: int, metric: int, off, cand_node, cand_tz, client_tz, k=None, usable=None):
        """ns_dev_pick_live: score map + selection on the device against this simulation's node table"""
        off = np.ascontiguousarray(off, dtype=np.int32)
        cn = np.ascontiguousarray(cand_node, dtype=np.int32)
        ct = np.ascontiguousarray(cand_tz, dtype=np.int32)
        cl = np.ascontiguousarray(client_tz, dtype=np.int32)
        n_req = len(off) - 1
        kk = np.ascontiguousarray(k if k is not None else np.zeros(n_req), dtype=np.int32)
        us = np.ascontiguousarray(usable, dtype=np.uint8) if usable is not None else None
        ov = np.zeros(max(1, len(cn)), dtype=np.int32)
        on = np.zeros(max(1, len(cn)), dtype=np.int32)
        oc = np.zeros(max(1, n_req), dtype=np.int32)
        ms = C.c_float()
        rc = self.L.ns_dev_pick_live(self.h, mode, metric, n_req, off.ctypes.data, cn.ctypes.data, ct.ctypes.data, cl.ctypes.data,
                                     us.ctypes.data if us is not None else None, kk.ctypes.data, ov.ctypes.data, on.ctypes.data,
                                     oc.ctypes.data, C.byref(ms))
        if rc:
            raise NsError(rc, "ns_dev_pick_live")
        return ov[:len(cn)], on[:len(cn)], oc[:n_req], float(ms.value)

    @property
    def launches(self) -> int:
        return int(self.L.ns_dev_launches(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.ns_dev_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Ensemble:
    """Independent replicas side by side on one GPU (ns_ens_*): one CTA per replica (mode AUTO, the batch engine at
    CTA scope) or one warp per replica (mode SEQ)."""

    def __init__(self, scenarios, device: int = 0, records: bool = True, ring_cap: int = 0, record_cap: int = 0,
                 mode: int = MODE_AUTO, cap_flows: int = 0):
        self.L = lib()
        n = len(scenarios)
        self._keep = [bytes(s) for s in scenarios]
        arr = (C.c_char_p * n)(*self._keep)
        lens = (C.c_size_t * n)(*[len(s) for s in self._keep])
        o = Opts(device=device, mode=mode, want_records=int(records), want_trace=0, ring_cap=ring_cap,
                 record_cap=record_cap, cap_flows=cap_flows)
        self.h = C.c_void_p()
        self.n = n
        rc = self.L.ns_ens_create(C.byref(self.h), arr, lens, n, C.byref(o))
        if rc:
            self.h = None
            raise NsError(rc, "ns_ens_create")
        self.device_ms = 0.0

    def run(self, max_events_each: int = 0) -> None:
        ms = C.c_float()
        rc = self.L.ns_ens_run(self.h, max_events_each, C.byref(ms))
        self.device_ms = float(ms.value)
        if rc:
            raise NsError(rc, "ns_ens_run")

    def digests(self):
        d = (Digest * self.n)()
        rc = self.L.ns_ens_digests(self.h, d, self.n)
        if rc:
            raise NsError(rc, "ns_ens_digests")
        return [{k: getattr(x, k) for k, _ in Digest._fields_} for x in d]

    def analysis_begin(self, period: float) -> None:
        """analyzer reductions on the device while the replicas run (ns_ana.cu); call before run()"""
        rc = self.L.ns_ens_analysis_begin(self.h, float(period))
        if rc:
            raise NsError(rc, "ns_ens_analysis_begin")

    def analysis(self):
        a = (AnaSummary * self.n)()
        rc = self.L.ns_ens_analysis(self.h, a, self.n)
        if rc:
            raise NsError(rc, "ns_ens_analysis")
        out = []
        for x in a:
            out.append({"records": int(x.records), "usage_records": int(x.usage_records), "state_records": int(x.state_records),
                        "t_last": float(x.t_last), "usage": np.ctypeslib.as_array(x.usage).copy(), "stale": np.ctypeslib.as_array(x.stale).copy(),
                        "stale_now": int(x.stale_now), "stale_max": int(x.stale_max)})
        return out

    def check(self) -> int:
        v = C.c_int()
        rc = self.L.ns_ens_check(self.h, C.byref(v))
        if rc:
            raise NsError(rc, "ns_ens_check")
        return int(v.value)

    def counters_sum(self) -> dict:
        out = (C.c_uint64 * len(COUNTERS))()
        rc = self.L.ns_ens_counters_sum(self.h, out)
        if rc:
            raise NsError(rc, "ns_ens_counters_sum")
        return {k: int(out[i]) for i, k in enumerate(COUNTERS)}

    @property
    def launches(self) -> int:
        return int(self.L.ns_ens_launches(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.ns_ens_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pick_batch(mode: int, off: np.ndarray, val: np.ndarray, node: np.ndarray, k: np.ndarray | None = None, device: int = 0):
    """ns_dev_pick_batch: host CSR candidate tables in, permuted tables + kept counts out."""
    off = np.ascontiguousarray(off, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.int32)
    node = np.ascontiguousarray(node, dtype=np.int32)
    n_req = len(off) - 1
    kk = np.ascontiguousarray(k if k is not None else np.zeros(n_req), dtype=np.int32)
    ov = np.zeros(max(1, len(val)), dtype=np.int32)
    on = np.zeros(max(1, len(val)), dtype=np.int32)
    oc = np.zeros(max(1, n_req), dtype=np.int32)
    ms = C.c_float()
    rc = lib().ns_dev_pick_batch(device, mode, n_req, off.ctypes.data, val.ctypes.data, node.ctypes.data,
                                 kk.ctypes.data, ov.ctypes.data, on.ctypes.data, oc.ctypes.data, C.byref(ms))
    if rc:
        raise NsError(rc, "ns_dev_pick_batch")
    return ov[:len(val)], on[:len(val)], oc[:n_req], float(ms.value)


def stream_hash(payload: bytes) -> int:
    """Host mirror of k_stream_hash (the per-replica record digest)."""
    n = len(payload)
    pad = (-n) % 8
    w = np.frombuffer(payload + b"\0" * pad, dtype="<u8").astype(np.uint64)
    j = np.arange(1, len(w) + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = w + np.uint64(0x9E3779B97F4A7C15) * j
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        acc = np.uint64(z.sum(dtype=np.uint64)) if len(w) else np.uint64(0)
        f = acc ^ np.uint64(n)
        f = (f ^ (f >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        f = (f ^ (f >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        f = f ^ (f >> np.uint64(31))
    return int(f)
