"""TEST INFRASTRUCTURE: run the compiled reference (oracle/_ref/netsim + modules/scenario.so, built
by oracle/Makefile from /root/reference) on a scenario script and collect its record stream.
Only tests/, bench.py's reference arm and __graft_entry__.smoke() may import this."""
from __future__ import annotations

import json
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
TRACE_DTYPE = np.dtype([("type", "<u4"), ("ord", "<u4"), ("time", "<f8"), ("v0", "<f8"), ("v1", "<f8")])


def available() -> bool:
    return os.path.exists(os.path.join(REF, "netsim")) and os.path.exists(os.path.join(REF, "modules", "scenario.so"))


def record_payload(path: str) -> bytes:
    """record.c:111-115,149-151: first 4 bytes = big-endian payload length, payload follows."""
    with open(path, "rb") as f:
        n = struct.unpack(">I", f.read(4))[0]
        return f.read(n)


def run_reference(scenario_bytes: bytes, max_events: int = 0, trace: bool = True, debug: bool = False,
                  record: bool = True, timeout: float = 600.0) -> dict:
    exe = os.path.join(REF, "netsim_dbg" if debug else "netsim")
    mod = os.path.join(REF, "modules_dbg" if debug else "modules", "scenario")
    with tempfile.TemporaryDirectory(prefix="nsref") as d:
        sc = os.path.join(d, "sc.bin")
        with open(sc, "wb") as f:
            f.write(scenario_bytes)
        env = dict(os.environ, NS_SCENARIO=sc, NS_RECORD=os.path.join(d, "rec") if record else "-",
                   NS_STATS=os.path.join(d, "stats.json"))
        if trace:
            env["NS_TRACE"] = os.path.join(d, "trace.bin")
        if max_events:
            env["NS_MAX_EVENTS"] = str(max_events)
        p = subprocess.run([exe, mod], cwd=d, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                           timeout=timeout)
        out = {"returncode": p.returncode, "stderr": p.stderr.decode(errors="replace")[-2000:]}
        if p.returncode != 0:
            return out
        out["stats"] = json.load(open(os.path.join(d, "stats.json")))
        if record:
            out["records"] = record_payload(os.path.join(d, "rec"))
        if trace:
            out["trace"] = np.fromfile(os.path.join(d, "trace.bin"), dtype=TRACE_DTYPE)
        return out


def analyze_dump(payload: bytes) -> str:
    """Decode a record payload with the reference's own analyzer (`analyze <file> test`)."""
    with tempfile.TemporaryDirectory(prefix="nsana") as d:
        path = os.path.join(d, "R")
        with open(path, "wb") as f:
            f.write(struct.pack(">I", len(payload)) + payload + b"\0" * 64)
        return subprocess.run([os.path.join(REF, "analyze"), path, "test"], capture_output=True,
                              check=True).stdout.decode()
