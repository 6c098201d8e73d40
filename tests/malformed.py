"""Malformed scenario scripts for the C ABI's validation (ns_dev_create must answer NS_E_BAD_SCENARIO before it indexes anything with
them): byte surgery on a valid script, plus a checker that restates the rules in numpy so that the CPU suite can prove every
mutation breaks exactly the rule it names (tests/test_host.py) before the GPU suite feeds it to the library (tests/test_gpu_parity.py)."""
import numpy as np

HDR = np.dtype([("magic", "<u4"), ("version", "<u4"), ("seed", "<u4"), ("flags", "<u4"), ("n_nodes", "<i4"), ("n_resources", "<i4"),
                ("n_copies", "<i4"), ("n_groups", "<i4"), ("n_ops", "<i4"), ("n_ondone", "<i4"), ("pad", "<i4", 2)])
RES = np.dtype([("owner", "<i4"), ("bit_rate", "<i4"), ("len", "<i8")])
COPY = np.dtype([("node", "<i4"), ("res", "<i4")])
OP = np.dtype([("kind", "<i4"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4"), ("start", "<i8"), ("x", "<f8"), ("y", "<f8")])
OND = np.dtype([("flow_ord", "<i4"), ("group", "<i4")])
MAGIC, VERSION = 0x4353534E, 1
OP_CREATE, OP_SCHED, OP_CLOSE = 1, 2, 3


def layout(b: bytes):
    """offsets of the arrays of a script without client behaviour (include/ns_scenario.h)"""
    h = np.frombuffer(b[:48], dtype=HDR)[0]
    o = {"nodes": 48}
    o["res"] = o["nodes"] + int(h["n_nodes"]) * 16
    o["copies"] = o["res"] + int(h["n_resources"]) * 16
    o["goff"] = o["copies"] + int(h["n_copies"]) * 8
    o["ops"] = o["goff"] + (int(h["n_groups"]) + 1) * 4
    o["ondone"] = o["ops"] + int(h["n_ops"]) * 40
    o["end"] = o["ondone"] + int(h["n_ondone"]) * 8
    return h, o


def violations(b: bytes):
    """the rules of build_sim (netsim_b200/csrc/ns_engine.cu), restated: names of the rules this script breaks"""
    if len(b) < 48:
        return ["short header"]
    h = np.frombuffer(b[:48], dtype=HDR)[0]
    bad = []
    if h["magic"] != MAGIC or h["version"] != VERSION:
        bad.append("magic")
    if int(h["flags"]) & ~1:
        bad.append("flags")
    if min(int(h[k]) for k in ("n_nodes", "n_resources", "n_copies", "n_ops", "n_ondone")) < 0 or h["n_groups"] < 1:
        return bad + ["counts"]
    _, o = layout(b)
    if len(b) < o["end"]:
        return bad + ["truncated"]
    nn, nr, ng, nops = int(h["n_nodes"]), int(h["n_resources"]), int(h["n_groups"]), int(h["n_ops"])
    goff = np.frombuffer(b[o["goff"]:o["ops"]], dtype="<i4")
    if goff[0] < 0 or goff[-1] > nops or (np.diff(goff) < 0).any():
        bad.append("group_off")
    ops = np.frombuffer(b[o["ops"]:o["ondone"]], dtype=OP)
    cr = ops[ops["kind"] == OP_CREATE]
    if ((cr["a"] < 0) | (cr["a"] >= nn) | (cr["b"] < 0) | (cr["b"] >= nn) | (cr["c"] < 0) | (cr["c"] >= nr)).any():
        bad.append("create operands")
    sc = ops[ops["kind"] == OP_SCHED]
    if ((sc["a"] < 0) | (sc["a"] >= ng)).any():
        bad.append("sched group")
    if (ops[ops["kind"] == OP_CLOSE]["a"] < 0).any():
        bad.append("close ordinal")
    od = np.frombuffer(b[o["ondone"]:o["end"]], dtype=OND)
    if (np.diff(od["flow_ord"]) <= 0).any():
        bad.append("ondone order")
    if ((od["group"] < 0) | (od["group"] >= ng)).any():
        bad.append("ondone group")
    rs = np.frombuffer(b[o["res"]:o["copies"]], dtype=RES)
    if ((rs["owner"] < 0) | (rs["owner"] >= nn)).any():
        bad.append("resource owner")
    cp = np.frombuffer(b[o["copies"]:o["goff"]], dtype=COPY)
    if ((cp["node"] < 0) | (cp["node"] >= nn) | (cp["res"] < 0) | (cp["res"] >= nr)).any():
        bad.append("copy operands")
    return bad


def _poke(b: bytes, off: int, fmt: str, v) -> bytes:
    return b[:off] + np.array([v], dtype=fmt).tobytes() + b[off + np.dtype(fmt).itemsize:]


def mutations(b: bytes):
    """(name of the broken rule, mutated script) for a valid script with >= 2 groups, a CREATE op, a SCHED op or more than one op, and
    an on-done entry"""
    h, o = layout(b)
    ops = np.frombuffer(b[o["ops"]:o["ondone"]], dtype=OP)
    i_cr = int(np.nonzero(ops["kind"] == OP_CREATE)[0][0])
    op_at = lambda i, field: o["ops"] + 40 * i + OP.fields[field][1]
    out = [("short header", b[:40]),
           ("truncated", b[:-1]),
           ("magic", _poke(b, 0, "<u4", 0x12345678)),
           ("magic", _poke(b, 4, "<u4", 2)),
           ("flags", _poke(b, 12, "<u4", 2)),
           ("counts", _poke(b, 32, "<i4", -1)),                                   # n_ops
           ("counts", _poke(b, 28, "<i4", 0)),                                    # n_groups
           ("truncated", _poke(b, 32, "<i4", int(h["n_ops"]) + 1000)),            # more ops than bytes
           ("group_off", _poke(b, o["goff"] + 4, "<i4", int(h["n_ops"]) + 1)),    # group_off[1] beyond group_off[2] and the table
           ("group_off", _poke(b, o["goff"], "<i4", -1)),
           ("group_off", _poke(b, o["goff"] + 4 * int(h["n_groups"]), "<i4", int(h["n_ops"]) + 1)),
           ("create operands", _poke(b, op_at(i_cr, "a"), "<i4", int(h["n_nodes"]))),
           ("create operands", _poke(b, op_at(i_cr, "b"), "<i4", -1)),
           ("create operands", _poke(b, op_at(i_cr, "c"), "<i4", int(h["n_resources"]))),
           ("close ordinal", _poke(_poke(b, op_at(int(h["n_ops"]) - 1, "kind"), "<i4", OP_CLOSE), op_at(int(h["n_ops"]) - 1, "a"), "<i4", -1)),
           ("sched group", _poke(_poke(b, op_at(int(h["n_ops"]) - 1, "kind"), "<i4", OP_SCHED), op_at(int(h["n_ops"]) - 1, "a"), "<i4", int(h["n_groups"]))),
           ("ondone group", _poke(b, o["ondone"] + 4, "<i4", int(h["n_groups"]))),
           ("ondone order", _poke(b, 36, "<i4", int(h["n_ondone"]) + 1) + b[o["end"] - 8:o["end"]]),   # the last entry twice
           ("resource owner", _poke(b, o["res"], "<i4", int(h["n_nodes"])))]
    return out
