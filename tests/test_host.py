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


def test_bad_scenario_rejected(built):
    if E.device_count() == 0:
        pytest.skip("needs a device to get past the device check")
    with pytest.raises(E.NsError):
        E.Engine(b"\0" * 64)


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
