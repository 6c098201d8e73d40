"""TEST INFRASTRUCTURE: ctypes binding of oracle/_build/libns_oracle.so (the CPU restatement).
Only tests/, bench.py's cpu_baseline/reference legs and __graft_entry__.smoke() may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libns_oracle.so")
TRACE_DTYPE = np.dtype([("type", "<u4"), ("ord", "<u4"), ("time", "<f8"), ("v0", "<f8"), ("v1", "<f8")])
COUNTERS = ["events", "solves", "full_spreads", "flow_visits", "speed_change", "flow_done", "drain",
            "throttle", "user", "creates", "closes", "draws", "seq", "live_flows", "peak_flows", "pending"]
_lib = None


def build() -> None:
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.ns_oracle_create.restype = C.c_void_p
        L.ns_oracle_create.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int]
        L.ns_oracle_run.argtypes = [C.c_void_p, C.c_uint64]
        L.ns_oracle_status.argtypes = [C.c_void_p]
        L.ns_oracle_now.restype = C.c_double
        L.ns_oracle_now.argtypes = [C.c_void_p]
        L.ns_oracle_records.restype = C.POINTER(C.c_uint8)
        L.ns_oracle_records.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.ns_oracle_trace.restype = C.c_void_p
        L.ns_oracle_trace.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.ns_oracle_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.ns_oracle_flow_state.restype = C.c_size_t
        L.ns_oracle_flow_state.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ns_oracle_node_state.restype = C.c_size_t
        L.ns_oracle_node_state.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ns_oracle_destroy.argtypes = [C.c_void_p]
        L.ns_rand_seed.argtypes = [C.c_void_p, C.c_uint32]
        L.ns_rand_next.restype = C.c_uint32
        L.ns_rand_next.argtypes = [C.c_void_p]
        L.ns_oracle_qselect.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
        L.ns_oracle_pick_opt1.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ns_oracle_pick_opt2.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ns_oracle_metric.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        _lib = L
    return _lib


class Oracle:
    def __init__(self, scenario_bytes: bytes, trace: bool = True, records: bool = True):
        self.L = lib()
        self.h = self.L.ns_oracle_create(scenario_bytes, len(scenario_bytes), int(trace), int(records))

    def run(self, max_events: int = 0) -> int:
        return self.L.ns_oracle_run(self.h, max_events)

    @property
    def status(self) -> int:
        return self.L.ns_oracle_status(self.h)

    @property
    def now(self) -> float:
        return self.L.ns_oracle_now(self.h)

    def records(self) -> bytes:
        n = C.c_size_t()
        p = self.L.ns_oracle_records(self.h, C.byref(n))
        return C.string_at(p, n.value) if n.value else b""

    def trace(self) -> np.ndarray:
        n = C.c_size_t()
        p = self.L.ns_oracle_trace(self.h, C.byref(n))
        if not n.value:
            return np.zeros(0, dtype=TRACE_DTYPE)
        return np.frombuffer(C.string_at(p, n.value * TRACE_DTYPE.itemsize), dtype=TRACE_DTYPE).copy()

    def counters(self) -> dict:
        a = (C.c_uint64 * len(COUNTERS))()
        self.L.ns_oracle_counters(self.h, a)
        return dict(zip(COUNTERS, [int(x) for x in a]))

    def flow_state(self, max_flows: int) -> np.ndarray:
        out = np.zeros((max_flows, 4), dtype="<f8")
        n = self.L.ns_oracle_flow_state(self.h, out.ctypes.data, max_flows)
        return out[:n]

    def node_state(self, n_nodes: int) -> np.ndarray:
        out = np.zeros((n_nodes, 4), dtype="<f8")
        n = self.L.ns_oracle_node_state(self.h, out.ctypes.data, n_nodes)
        return out[:n]

    def close(self):
        if self.h:
            self.L.ns_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def glibc_random(seed: int, n: int) -> np.ndarray:
    L = lib()
    st = C.create_string_buffer(31 * 4 + 8)
    L.ns_rand_seed(st, seed)
    return np.array([L.ns_rand_next(st) for _ in range(n)], dtype=np.uint32)


def qselect(val: np.ndarray, node: np.ndarray, k: int):
    v = np.ascontiguousarray(val, dtype=np.int32).copy()
    nd = np.ascontiguousarray(node, dtype=np.int32).copy()
    lib().ns_oracle_qselect(v.ctypes.data, nd.ctypes.data, len(v), k)
    return v, nd
