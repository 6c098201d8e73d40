"""CPU: host-side logic and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from netsim_b200 import engine as E
from netsim_b200 import scenario as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "netsim_b200.h")).read()
    names = set(re.findall(r"\b(ns_(?:dev|ens|rand|strerror)\w*)\s*\(", hdr))
    assert len(names) >= 20
    L = ctypes.CDLL(E.LIB_PATH)
    for n in sorted(names):
        assert hasattr(L, n), n


def test_rand_host_matches_libc_and_jump(built):
    L = E.lib()
    libc = ctypes.CDLL("libc.so.6")
    libc.random.restype = ctypes.c_long
    for n in (0, 1, 95, 96, 1000, 250001):
        libc.srandom(7)
        for _ in range(n):
            libc.random()
        exp = [libc.random() for _ in range(64)]
        g = E.RandState()
        L.ns_rand_seed_host(ctypes.byref(g), 7)
        L.ns_rand_jump_host(ctypes.byref(g), n)
        assert exp == [L.ns_rand_next_host(ctypes.byref(g)) for _ in range(64)]


def test_scenario_layout_roundtrip():
    sc = S.reference_tests()["test04"]
    b = sc.to_bytes()
    magic, ver, seed, flags, nn, nr, nc, ng, nops, nod = np.frombuffer(b[:40], dtype="<u4")
    assert magic == S.MAGIC and ver == 1 and nn == 4 and nr == 1 and ng == 3 and nops == 5 and nod == 1
    assert len(b) == 48 + nn * 16 + nr * 16 + nc * 8 + (ng + 1) * 4 + nops * 40 + nod * 8


def test_stream_hash_is_order_sensitive():
    a = bytes(range(40))
    assert E.stream_hash(a) != E.stream_hash(a[8:16] + a[:8] + a[16:])
    assert E.stream_hash(a) != E.stream_hash(a + b"\0")
    assert E.stream_hash(b"") == E.stream_hash(b"")


def test_no_cpu_fallback(built):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    if E.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(E.NsError) as ei:
        E.Engine(S.reference_tests()["test00"].to_bytes())
    assert ei.value.code == -9
    with pytest.raises(E.NsError):
        E.pick_batch(1, np.array([0, 2]), np.array([1, 2]), np.array([0, 1]), np.array([1]))


@pytest.mark.gpu
def test_bad_scenario_rejected(built):
    """(the device check comes first, so this needs a GPU; the full set of malformed scripts is in tests/test_gpu_parity.py)"""
    with pytest.raises(E.NsError) as ei:
        E.Engine(b"\0" * 64)
    assert ei.value.code == -1


def test_swarm_generator_is_valid_script():
    sc = S.swarm(50, 2, n_catalogue=4, seeds_per_peer=2, mean_strips=2.0)
    ops, off = sc.packed_ops()
    cr = ops[ops["kind"] == S.OP_CREATE]
    assert len(cr) >= 50
    assert (cr["a"] != cr["b"]).all()
    pairs = set(zip(cr["a"].tolist(), cr["b"].tolist()))
    assert len(pairs) == len(cr)       # never two flows between the same pair (flow.c:300)
    assert ((cr["x"] >= 1700) & (cr["x"] <= 4000)).all()
    assert ((cr["y"] >= 0.1 - 1e-12) & (cr["y"] <= 0.56 + 1e-12)).all()


def test_host_core_containers():
    """CPU: the list / hash / skip-list headers modules compile against (netsim_b200/host/include) keep the reference's semantics
    modules rely on: insertion-order iteration, deletion inside HASH_ITER, sorted skip lists, newest-first lists."""
    import subprocess, tempfile
    inc = os.path.join(ROOT, "netsim_b200", "host", "include")
    with tempfile.TemporaryDirectory(prefix="nscont") as d:
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-O2", "-Wall", "-I" + inc, os.path.join(ROOT, "tests", "modules", "t_containers.c"), "-o", exe], check=True)
        p = subprocess.run([exe], capture_output=True, text=True)
        assert p.returncode == 0, p.stdout
        assert "iter in order 1, left 500" in p.stdout and "skip sorted 1 n 200 ge 18 le 17" in p.stdout and "list 31 13 empty 0" in p.stdout


def test_host_core_containers_randomized():
    """CPU: 60 000 random operations each on the hash (adds / finds / deletes, deletion inside HASH_ITER, iteration == insertion order)
    and on the skip list (inserts / deletes / find_ge / find_le, level-0 next / prev walks) against plain arrays."""
    import subprocess, tempfile
    inc = os.path.join(ROOT, "netsim_b200", "host", "include")
    with tempfile.TemporaryDirectory(prefix="nsfuzz") as d:
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-O2", "-I" + inc, os.path.join(ROOT, "tests", "modules", "t_containers_fuzz.c"), "-o", exe], check=True)
        p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert p.returncode == 0 and "containers fuzz: ok" in p.stdout, p.stdout


def _parse_records(payload: bytes):
    """record boundaries of a stream in the reference's layout (record.c:53-106, analyzer/record_reader.c:29-77)"""
    at, out = 0, []
    while at < len(payload):
        b = payload[at + 12]
        n = 13 + {1: 1, 2: 2, 4: 4, 255: 6}[b]
        out.append(payload[at:at + n])
        at += n
    assert at == len(payload)
    return out


def test_record_file_split_at_record_boundaries():
    """CPU: SURVEY §8 f-1, the 4 GiB limit of the record file's uint32 length field lifted behind a flag (NS_RECORD_SPLIT):
    the host core's writer (netsim_b200/host/recfile.c) cuts the stream into <path>, <path>.1, ... at record boundaries, every
    piece a complete file in the reference's layout (record.c:111-133); the unmodified analyzer reads each piece, and with
    NS_RECORD_SPLIT_PREAMBLE its node-class reports run on each piece alone."""
    import subprocess, tempfile
    host = os.path.join(ROOT, "netsim_b200", "host")
    payload = np.load(os.path.join(ROOT, "tests", "golden", "clients_saturated.npz"))["records"].tobytes()
    recs = _parse_records(payload)
    assert len(recs) > 1000
    with tempfile.TemporaryDirectory(prefix="nsrec") as d:
        so = os.path.join(d, "librecfile.so")
        subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-shared", "-fPIC", os.path.join(host, "recfile.c"), "-o", so], check=True)
        L = ctypes.CDLL(so)
        L.ns_recfile_save.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.ns_recfile_split_from_env.restype = ctypes.c_size_t
        n = ctypes.c_int(0)

        def pieces(base, count):
            out = []
            for k in range(count):
                raw = open(base + (".%d" % k if k else ""), "rb").read()
                ln = int.from_bytes(raw[:4], "big")
                size = 1 << 20
                while 4 + ln + 19 >= size:
                    size <<= 1
                assert len(raw) == size and not any(raw[4 + ln:])                # 1 MiB, doubled while the records do not fit
                out.append(raw[4:4 + ln])
            assert not os.path.exists(base + ".%d" % count)
            return out

        # flag off: one file, the reference's layout
        one = os.path.join(d, "R")
        assert L.ns_recfile_save(one.encode(), payload, len(payload), 0, 0, ctypes.byref(n)) == 0 and n.value == 1
        assert pieces(one, 1) == [payload]
        # flag on, with a small limit so that the cut is exercised: whole records only, nothing lost, nothing reordered
        limit = 4096
        base = os.path.join(d, "S")
        assert L.ns_recfile_save(base.encode(), payload, len(payload), limit, 0, ctypes.byref(n)) == 0
        assert n.value >= len(payload) // limit
        ps = pieces(base, n.value)
        for k, pc in enumerate(ps):
            assert 0 < len(pc) <= limit and (len(pc) > limit - 19 or k == n.value - 1)   # as full as whole records allow
            _parse_records(pc)                                                    # ends on a record boundary
        assert b"".join(ps) == payload
        # a limit below one record, and bytes that are not records, are refused
        assert L.ns_recfile_save(base.encode(), payload, len(payload), 12, 0, ctypes.byref(n)) == -2
        assert L.ns_recfile_save(base.encode(), b"\x01" * 64, 64, limit, 0, ctypes.byref(n)) == -3
        # an empty stream is still one (empty) file
        e = os.path.join(d, "E")
        assert L.ns_recfile_save(e.encode(), b"", 0, limit, 0, ctypes.byref(n)) == 0 and n.value == 1 and pieces(e, 1) == [b""]
        # the flags: unset / 0 = off, 1 = 2^32-1 bytes per piece, anything else = bytes
        for v, want in ((None, 0), ("0", 0), ("1", 0xFFFFFFFF), ("65536", 65536)):
            os.environ.pop("NS_RECORD_SPLIT", None)
            if v is not None:
                os.environ["NS_RECORD_SPLIT"] = v
            assert L.ns_recfile_split_from_env() == want
        os.environ.pop("NS_RECORD_SPLIT", None)
        assert L.ns_recfile_flags_from_env() == 0
        # with the preamble every continuation piece = the node-identity records of the earlier pieces + its own share
        ident = lambda r: r[0] == 0 and (r[1] == 8 or (r[1] == 7 and r[13] in (5, 6)))
        pre = os.path.join(d, "P")
        assert L.ns_recfile_save(pre.encode(), payload, len(payload), 1 << 14, 1, ctypes.byref(n)) == 0 and n.value > 3
        pp, seen, rest = pieces(pre, n.value), b"", b""
        for pc in pp:
            assert len(pc) <= 1 << 14 and pc[:len(seen)] == seen
            own = pc[len(seen):]
            rest += own
            seen += b"".join(r for r in _parse_records(own) if ident(r))
        assert rest == payload and len(seen) >= 14 * 92
        assert L.ns_recfile_save(pre.encode(), payload, len(payload), 600, 1, ctypes.byref(n)) == -4      # the preamble fills a piece
        # the reference's own analyzer (where the compiled reference is present): `test` prints every record of a piece;
        # `list_nodes` needs each node's creation record first (analyzers.c:106-110), which the preamble supplies
        ana = os.path.join(ROOT, "oracle", "_ref", "analyze")
        if os.path.exists(ana):
            run = lambda path, what: subprocess.run([ana, path, what], capture_output=True, text=True)
            whole = run(one, "test")
            assert whole.returncode == 0 and whole.stdout.count("\n") >= len(recs)
            cat = ""
            for k in range(len(ps)):
                p = run(base + (".%d" % k if k else ""), "test")
                assert p.returncode == 0 and "Corrupted" not in p.stderr
                cat += p.stdout
            assert cat == whole.stdout
            nodes = run(one, "list_nodes").stdout.splitlines()
            assert len(nodes) == 92
            last = run(pre + ".%d" % (len(pp) - 1), "list_nodes")
            assert last.returncode == 0 and sorted(last.stdout.splitlines()) == sorted(nodes)


def test_malformed_scenarios_break_the_rule_they_name():
    """CPU: the byte surgery of tests/malformed.py does what it says (the GPU suite feeds the same scripts to ns_dev_create), and
    the numpy restatement of build_sim's rules accepts every valid script of the test suite."""
    import malformed as M
    from conftest import golden_cases
    base = S.reference_tests()["test04"].to_bytes()
    assert M.violations(base) == []
    muts = M.mutations(base)
    assert len(muts) >= 18 and len({m for _, m in muts}) == len(muts)
    for rule, m in muts:
        assert rule in M.violations(m), rule
    for name, sc in S.reference_tests().items():
        assert M.violations(sc.to_bytes()) == [], name
    for name, z in golden_cases():
        b = z["scenario"].tobytes()
        if np.frombuffer(b[:48], dtype=M.HDR)[0]["flags"] == 0:
            assert M.violations(b) == [], name
    assert M.violations(S.swarm(60, 2, n_catalogue=4, seeds_per_peer=2, mean_strips=2.0).to_bytes()) == []
