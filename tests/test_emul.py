"""CPU: the DEVICE code of the single-warp engine (ns_seq.cu + ns_model.cuh), of the pickers (ns_pick.cu), of the analyzer kernel
(ns_ana.cu) and the host layer (ns_engine.cu), compiled as plain C++ against a deterministic lock-step warp model (tests/emul) and run
under AddressSanitizer + UBSan with poisoned "device memory": byte-identical to the goldens of the compiled reference and to the CPU
oracle.  compute-sanitizer is closed on the GPU pool; this is the memory-safety and collective-discipline pass the device code
gets instead (out-of-bounds reads and writes, reads of never-written memory that matter, collectives not every lane reaches or
that lanes reach from different call sites).  TEST INFRASTRUCTURE: the model is not a fallback of the product — the product
library fails with NS_E_NO_DEVICE without a GPU (tests/test_host.py::test_no_cpu_fallback) — and is ~10^4 times slower than a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")
LIB = os.path.join(EMUL, "_build", "libnetsim_emul.so")


@pytest.fixture(scope="module")
def model(built):
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("no AddressSanitizer runtime next to gcc")
    subprocess.run(["make", "-s", "-C", EMUL], check=True)
    return dict(os.environ, NS_LIB=LIB, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")


def run(env, *checks, order=None, sms=None, auto=False, timeout=900):
    env = dict(env)
    if order:
        env["EMU_ORDER"] = order
    if sms:
        env["EMU_SMS"] = str(sms)
    if auto:
        env["EMU_MODE"] = "auto"
    p = subprocess.run([sys.executable, os.path.join(EMUL, "run_checks.py"), *checks], env=env, capture_output=True, text=True, timeout=timeout)
    err = "\n".join(ln for ln in p.stderr.splitlines() if "doesn't fully support makecontext" not in ln)
    assert p.returncode == 0 and "AddressSanitizer" not in err and "runtime error" not in err and "emu:" not in err, err[-3000:]
    res = [json.loads(ln) for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert res and all(r["ok"] for r in res), [r for r in res if not r["ok"]]
    return res


# every job of every test below starts when the first test asks for its results: the tests only wait for their own jobs, the
# machine's cores are busy with all of them at once (the whole file: about a minute on 8 cores)
JOBS = {
    "goldens": [(["goldens:test00,test01,test02,test03,test04,test05,edge_empty,edge_end_and_closes,edge_existing_range,edge_missing_resource,edge_zero_capacity,rand1,rand21,rand34"], {}),
                (["goldens:rand11,rand13,rand2,rand3"], {}), (["goldens:rand55,rand7,rand89,relay1,relay16"], {}), (["goldens:relay12,relay4,relay6,relay9"], {}),
                (["goldens:churn150"], {}), (["goldens:clients_play40"], {})],
    "long": [(["goldens:swarm200"], dict(timeout=3000)), (["goldens:clients_cut"], dict(timeout=3000)), (["goldens:clients_distance"], dict(timeout=3000)),
             (["goldens:clients_saturated"], dict(timeout=3000))],
    "lanes": [(["goldens:test04,test05,rand2,rand7"], dict(order="rev")), (["goldens:rand55,relay9,relay1"], dict(order="rev")), (["random:1,4"], dict(order="rev")),
              (["ensemble", "analysis"], dict(order="rev"))],
    "paths": [(["windows"], {}), (["ensemble"], {}), (["analysis"], {}), (["picks", "malformed", "args"], {}), (["random:0,3,7"], {}), (["random:12,20,29"], {})],
    "cta": [(["ensemble_cta"], {})],
    "live": [(["live:2,5,9"], {}), (["live:3,6,10", "random:2,9,17"], dict(auto=True, sms=3))],
    "grid": [(["grid"], dict(sms=3)), (["grid"], dict(sms=1)), (["grid"], dict(sms=4, order="rev"))],
}
HOST = os.path.join(ROOT, "netsim_b200", "host", "_build")
GOLD = os.path.join(ROOT, "tests", "golden")
HOST_MODS = ["test00", "test01", "test02", "test03", "test04", "test05", "tc2", "tc3"]
HOST_CFGS = ["small50", "evict3", "play2c", "cloudc", "push2"]


def _netsim_dev(model, module, cwd, extra_preload=(), env_extra=None, timeout=900):
    """netsim_dev (the host core, linked against the CUDA library) with the model's C ABI preloaded in front of it"""
    env = dict(model, **(env_extra or {}))
    env.pop("NS_LIB")
    env["LD_PRELOAD"] = " ".join([model["LD_PRELOAD"], *extra_preload, LIB])
    p = subprocess.run([os.path.join(HOST, "netsim_dev"), os.path.join(HOST, "modules", module)], cwd=cwd, env=env, capture_output=True, timeout=timeout)
    err = b"".join(ln for ln in p.stderr.splitlines(True) if b"doesn't fully support makecontext" not in ln)
    assert p.returncode == 0 and b"AddressSanitizer" not in err and b"runtime error" not in err and b"emu:" not in err, err[-3000:].decode(errors="replace")
    return err


def _host_module(model, mod):
    import hashlib
    import tempfile
    with tempfile.TemporaryDirectory(prefix="nsemu") as d:
        err = _netsim_dev(model, mod, d)
        assert err == open(os.path.join(GOLD, mod + ".stderr.txt"), "rb").read(), mod
        if mod == "test05":
            assert hashlib.sha256(open(os.path.join(d, "R"), "rb").read()).hexdigest() == "8180090a224a219634a6004be53879144c5edb2a4951c91c691039aab3135d8f"
    return mod


def _host_p2p(model, cfg):
    import hashlib
    import shutil
    import tempfile
    exp = json.load(open(os.path.join(GOLD, "p2p", "expected.json")))[cfg]
    shim = os.path.join(ROOT, "oracle", "_ref", "timeshim.so")
    with tempfile.TemporaryDirectory(prefix="nsemu") as d:
        shutil.copy(os.path.join(GOLD, "p2p", cfg, "p2p.cfg"), d)
        err = _netsim_dev(model, "p2p", d, extra_preload=(shim,), env_extra={"NS_FAKE_TIME": str(exp["seed"])}, timeout=3000)
        data = open(os.path.join(d, "p2p_record"), "rb").read()
        n = int.from_bytes(data[:4], "big")
        assert n == exp["record_bytes"] and hashlib.sha256(data[4:4 + n]).hexdigest() == exp["record_sha256"], cfg
        assert hashlib.sha256(err).hexdigest() == exp["stderr_sha256"], cfg
    return cfg


def _host_ready():
    return (os.path.exists(os.path.join(HOST, "netsim_dev")) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "timeshim.so"))
            and all(os.path.exists(os.path.join(HOST, "modules", m + ".so")) for m in HOST_MODS + ["p2p"]))


@pytest.fixture(scope="module")
def started(model):
    from concurrent.futures import ThreadPoolExecutor
    ex = ThreadPoolExecutor(max_workers=max(4, os.cpu_count() or 4))
    fut = {}
    # longest first: the pool is never left waiting for one late straggler
    if _host_ready():
        fut["host"] = [ex.submit(_host_p2p, model, c) for c in reversed(HOST_CFGS)] + [ex.submit(_host_module, model, m) for m in HOST_MODS]
    for k in ("grid", "lanes", "paths", "long", "cta", "live", "goldens"):
        fut[k] = [ex.submit(run, model, *c, **kw) for c, kw in JOBS[k]]
    yield fut
    ex.shutdown(wait=False, cancel_futures=True)


def results(fut, key):
    return [r for f in fut[key] for r in f.result()]


def test_goldens_on_the_lockstep_model(started):
    """the reference's own test00-05, the edge cases, random scripts, relay topologies (throttles, merges), churn and the client
    behaviour fixture: records, fp64 traces and event counts of the compiled reference, from the device code under ASan + UBSan"""
    assert len(results(started, "goldens")) == 33 - 4            # all but swarm200 and the three long client fixtures (next test)


def test_long_goldens_on_the_lockstep_model(started):
    """swarm200 (3 933 events, the one-event-per-lane path on batches of dozens of events) and the three long client-behaviour
    fixtures (client_new_play1 with the picker in the loop, the water-mark machine, node-state records): 1.3 - 3.9 k events each"""
    assert len(results(started, "long")) == 4


def test_lane_schedule_does_not_matter(started):
    """the same device code with the lanes taking their turn highest-first instead of lowest-first between two rendezvous: code
    whose lanes only communicate through collectives / __syncwarp gives the same bytes (the fp64 traces included); the analyzer
    kernel too, since its gather / update split"""
    results(started, "lanes")


def test_engine_paths_on_the_lockstep_model(started):
    """stop / resume inside a batch, soft stop + drain of a small record buffer, ensembles with in-kernel record folding and the
    analyzer kernel, the pickers against the reference's own qselect tables, random scripts (relays included, and the scripts on
    which the reference aborts) against the oracle, the C ABI's refusals"""
    results(started, "paths")


def test_fresh_scripts_against_the_live_reference_on_the_lockstep_model(started):
    """scripts nobody has stored — client behaviour, churn, swarms to completion, random scripts — with the compiled reference
    (oracle/_ref) run live beside the device code: the one-warp engine and the grid engine with 3 CTAs"""
    assert len(results(started, "live")) == 9 + 9 + 6


def test_batch_engine_at_cta_scope_on_the_lockstep_model(started):
    """k_run_ens — the batch engine of ns_par_kernel.inc with one CTA per replica (the code the grid-scope k_run_par shares: ordered
    extraction, radix partition by (node, dir) group, thread / warp / CTA solves, staged event_adds and records, ordered commit) —
    with 256 fibers per CTA under ASan + UBSan: a 300-peer swarm whose first timestamps are batches of up to 552 events with 789
    full spreads, digest and counters equal to the oracle's"""
    results(started, "cta")


def test_grid_engine_on_the_lockstep_model(started):
    """k_run_par / k_run_par_cl (cooperative launch, grid barriers) and k_init_par (the parallel module init) with 1, 3 and 4 CTAs of 256
    threads — one OS thread per CTA, so the CTAs really run side by side and only meet at grid.sync() — under ASan + UBSan: goldens
    with NS_MODE_AUTO, and a 300-peer swarm to completion (parallel init in waves, batches of up to 552 events through the radix
    partition / thread, warp and CTA solves / ordered commit, then completions, closes, merges and the hand-over to the one-warp
    engine): records, fp64 traces and counters equal to the oracle's, whatever the number of CTAs and the lane order"""
    assert len(results(started, "grid")) == 3 * 9


def test_unmodified_reference_modules_through_the_host_core_on_the_model(started):
    """the reference's own test00-05.c, test_client2/3.c + client_behaviour.c and p2p.c + cloud_behaviour.c (tests/golden/p2p: servers only,
    resource eviction, client-to-client relays, smart cloud, cloud pushes), compiled unchanged against netsim_b200/host/include, with
    every device call of the host core (k_exec_ops, k_run_seq in host-callback mode, the mirror view, the shared random() stream)
    answered by the device code on the lock-step model under ASan + UBSan: stderr and record files byte-identical to the reference's"""
    if "host" not in started:
        pytest.skip("reference modules were not built against the host core (needs /root/reference at build time)")
    done = [f.result() for f in started["host"]]
    assert sorted(done) == sorted(HOST_MODS + HOST_CFGS)
