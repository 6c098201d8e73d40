#!/usr/bin/env python
"""bench.py — netsim flow/event hot path on B200: simulated events/s and rate solves/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2]): a p2p swarm of 100 032 peers + 320 hubs with ~1M concurrent
flows, one simulation per GPU (generator netsim_b200/scenario.py:swarm_fast).
A "step" is one window of `--events-per-step` simulated events handled by ONE launch of the
persistent engine kernel (ns_dev_run through the C ABI).  Both arms time the SAME absolute event range of
the same script: events [W*E, (W+K)*E) after W warm-up windows of E events each.  With N>1 every rank simulates its own
seeded replica of the swarm (weak scaling, no data-path collective: a single simulation is globally
coupled); the per-replica digests are gathered with one NCCL all_gather.

--impl reference times the UNMODIFIED reference core (oracle/_ref, compiled from /root/reference
by oracle/Makefile) on the same script and the same event range, one process per replica on
min(N, host cores) cores (one simulation is single-threaded by design).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "simulated events/s (p2p swarm, 100k peers / ~1M flows, one simulation per GPU)"
UNIT = "events/s"

# algorithmic bytes per unit of work (SURVEY.md §8d byte model)
B_SOLVE = 56          # node {total,max,usage} + flow {ub,speed} read, speed+usage written
B_VISIT = 20          # per flow visited by a full-spread pass: ub 8 + speed 8 + index 4 (lower bound)
B_INSERT = 32         # one emitted event {time, seq, flow, dir, speed}
B_RETIME = 92 + 120   # completion re-time of one flow + the byte-progress update of its two ranges


def workload(args, rank):
    if args.workload == "churn":
        return churn_workload(args, rank)
    from netsim_b200 import scenario as S
    return S.swarm_fast(args.peers, args.hubs, n_catalogue=64, seeds_per_peer=8, mean_strips=9.0,
                        hub_frac=args.hub_frac, res_len=600000, seed=12345 + rank, rand_seed=1 + rank)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region.  The windows are tens of milliseconds, so
    the sampler polls NVML in-process every few ms (nvidia-smi takes longer to start than a window);
    nvidia-smi is the fallback."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def _nvml_sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(name)

    def _smi_sample(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [x.strip() for x in out.split(",")]
        if len(r) >= 6:
            self.sm.append(float(r[0]))
            self.mx.append(float(r[1]))
            for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                if r[2 + i].lower() == "active":
                    self.reasons.add(name)

    def sample(self):
        try:
            if self.nv is not None:
                self._nvml_sample()
            else:
                self._smi_sample()
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.005 if self.nv is not None else 0.05)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml" if self.nv is not None else "nvidia-smi"}


def run_reference(args, sc_bytes, steps, warmup, events_per_step):
    """The reference core on this script, one process / one core; wall clock per window from the
    replay module (oracle/ref_scenario.c, NS_STEP_EVENTS)."""
    from oracle import refrun
    if not refrun.available():
        return None
    with tempfile.TemporaryDirectory(prefix="nsbench") as d:
        path = os.path.join(d, "sc.bin")
        with open(path, "wb") as f:
            f.write(sc_bytes)
        env = dict(os.environ, NS_SCENARIO=path, NS_RECORD=os.path.join(d, "rec"), NS_STATS=os.path.join(d, "stats.json"),
                   NS_MAX_EVENTS=str((steps + warmup) * events_per_step), NS_STEP_EVENTS=str(events_per_step))
        t0 = time.time()
        p = subprocess.run([os.path.join(refrun.REF, "netsim"), os.path.join(refrun.REF, "modules", "scenario")],
                           cwd=d, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=args.ref_timeout)
        wall = time.time() - t0
        if p.returncode != 0:
            return {"error": f"reference exited {p.returncode}: {p.stderr.decode(errors='replace')[-200:]}"}
        st = json.load(open(os.path.join(d, "stats.json")))
        rec_bytes = os.path.getsize(os.path.join(d, "rec"))
    ts = st["step_times"]
    done = len(ts)
    if done <= warmup:
        return {"error": f"reference handled only {st['events']} events"}
    t_w = ts[warmup - 1] if warmup else 0.0
    k = done - warmup
    sec = ts[-1] - t_w
    return {"events": k * events_per_step, "seconds": sec, "events_per_s": k * events_per_step / sec, "steps": k,
            "init_seconds": wall - st["loop_seconds"], "record_file_bytes": rec_bytes, "total_events": st["events"]}


def emit(line: dict) -> None:
    """the ONE JSON line, on the process's real stdout"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# libraries (NCCL's version banner, nvcc/make chatter) must not reach stdout: main() keeps the real stdout for the
# JSON line and points fd 1 at stderr for everything else
_REAL_STDOUT = 1


def _guard_stdout():
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


CHURN_METRIC = "simulated events/s (churn: 1M users + 1024 servers, Poisson arrivals/departures, one simulation per GPU)"


def churn_workload(args, rank):
    """BASELINE.json configs[3]: 1M users + 1024 servers; a warm swarm of --churn-init users is connected by
    the module's init, the others arrive as a Poisson process of 1000/s; every user leaves after an exponential session (its
    flow is closed); timestamps on a 10 ms grid, so a batch holds ~10 arrivals + ~10 departures and the speed events they trigger"""
    from netsim_b200 import scenario as S
    return S.churn_fast(args.churn_users, 1024, rate=1000.0, mean_session=args.churn_session, sim_end=args.churn_end,
                        init_users=args.churn_init, seed=2024 + rank)


def churn_config(args):
    return {"workload": f"BASELINE.json configs[3]: churn script, {args.churn_users} users + 1024 servers, {args.churn_init} users connected at init, Poisson "
                        f"arrivals 1000/s, exponential sessions of mean {args.churn_session}s (SURVEY §8d says 600 s from a cold start: a warm start with shorter "
                        f"sessions reaches the saturated steady state — ~100 flows per server, every arrival/departure a full spread — that "
                        f"the reference arm can still set up), 10 ms timestamp grid, {args.churn_end} sim-s; records on; step = window of "
                        f"{args.events_per_step} events; timed range = events [{args.warmup * args.events_per_step}, "
                        f"{(args.warmup + args.steps) * args.events_per_step}) of the script, the same for both arms",
            "users": args.churn_users, "servers": 1024, "init_users": args.churn_init, "events_per_step": args.events_per_step,
            "working_set": "1M-node tables + flows + pending-event ring exceed the 126 MB L2; no explicit L2 flush"}


def config_for(args):
    if args.workload == "churn":
        return churn_config(args)
    flows_k = args.peers * 10 // 1000
    return {"workload": f"BASELINE.json configs[2]: p2p swarm script, {args.peers} peers + {args.hubs} hubs, ~{flows_k}k flows established at init, "
                        f"records on; step = window of {args.events_per_step} events; timed range = events "
                        f"[{args.warmup * args.events_per_step}, {(args.warmup + args.steps) * args.events_per_step}) of the script, the same for both arms. "
                        f"hub_frac {args.hub_frac} of the sources are Zipf hubs (SURVEY §8d says 0.1: at 0.1 the top hub carries ~16k flows and the REFERENCE's "
                        f"module init alone queues 1.8e8 events — each flow_create on a saturated hub re-notifies every earlier flow, flow.c:157-200 — "
                        f"which does not fit the reference arm's few minutes; 0.02 keeps both arms runnable on the same script)",
            "peers": args.peers, "hubs": args.hubs, "hub_frac": args.hub_frac, "events_per_step": args.events_per_step,
            "working_set": "node/flow/range tables + pending-event ring (several GB) exceed the 126 MB L2; no explicit L2 flush"}


def reference_arm(args, n_proc):
    """n_proc independent reference processes side by side (one seeded replica each, like our ranks), every one
    on its own core; value = all timed events / the slowest process's timed seconds"""
    from concurrent.futures import ThreadPoolExecutor
    scs = [workload(args, r).to_bytes() for r in range(n_proc)]
    with ThreadPoolExecutor(max_workers=n_proc) as ex:
        rs = list(ex.map(lambda sb: run_reference(args, sb, args.steps, args.warmup, args.events_per_step), scs))
    if any(r is None for r in rs):
        return None
    bad = [r["error"] for r in rs if "error" in r]
    if bad:
        return {"error": bad[0]}
    sec = max(r["seconds"] for r in rs)
    ev = sum(r["events"] for r in rs)
    return {"events": ev, "seconds": sec, "events_per_s": ev / sec, "steps": rs[0]["steps"], "init_seconds": max(r["init_seconds"] for r in rs),
            "processes": n_proc}


def expected_digest(args, rank, total_events):
    """tests/golden/bench_expected.json: the CPU oracle's digest of this script after total_events events
    (tools/make_bench_expected.py, computed where /root/reference pins the oracle)"""
    try:
        tab = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_expected.json")))
    except Exception:
        return None
    if args.workload == "churn":
        return tab.get(f"churn,users={args.churn_users},init={args.churn_init},session={args.churn_session},end={args.churn_end},rank={rank},events={total_events}")
    return tab.get(f"peers={args.peers},hubs={args.hubs},hub_frac={args.hub_frac},rank={rank},events={total_events}")


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="swarm100k", choices=list(WORKLOADS))
    ap.add_argument("--peers", type=int, default=100032)
    ap.add_argument("--hubs", type=int, default=320)
    ap.add_argument("--hub-frac", type=float, default=0.02)
    ap.add_argument("--events-per-step", type=int, default=1000000)
    ap.add_argument("--ref-timeout", type=float, default=1500.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicas", type=int, default=1024, help="--workload ensemble: replicas per ensemble (all ranks together)")
    ap.add_argument("--ens-peers", type=int, default=10000, help="--workload ensemble: peers per replica")
    ap.add_argument("--ens-kind", default="swarm10k", choices=["swarm10k", "p2p1k"], help="--workload ensemble: the replicas' script (config 5 / config 2 shape)")
    ap.add_argument("--churn-users", type=int, default=1000000)
    ap.add_argument("--churn-init", type=int, default=100000)
    ap.add_argument("--churn-session", type=float, default=100.0)
    ap.add_argument("--churn-end", type=float, default=60.0)
    args = ap.parse_args()
    if args.workload == "ensemble" and "--steps" not in sys.argv:
        args.steps = 1                                       # one ensemble of --replicas replicas IS the config

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if WORKLOADS[args.workload] is not None:
        return WORKLOADS[args.workload](args, rank, local_rank, world)
    metric = CHURN_METRIC if args.workload == "churn" else METRIC
    config = config_for(args)
    eps = args.events_per_step
    total_steps = args.steps + args.warmup

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        n_proc = max(1, min(args.gpus, os.cpu_count() or 1))
        r = reference_arm(args, n_proc)
        if r is None or "error" in r:
            emit({"impl": "reference", "unavailable": (r or {}).get("error", "oracle/_ref not built")})
            return 0
        val = r["events_per_s"]
        line = {"metric": metric, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
                "ms_per_step": 1e3 * r["seconds"] / r["steps"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": n_proc, "kind": "reference",
                                 "sample": f"{n_proc} process(es) of the unmodified reference core (single-threaded by design), one seeded replica each "
                                           f"like our ranks; each: {r['steps']} windows of {eps} events after {args.warmup} warm-up windows of the same "
                                           f"script = the same absolute event range as our arm; init {r['init_seconds']:.1f}s untimed"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from netsim_b200 import engine as E

    if E.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sc = workload(args, rank)
    sc_bytes = sc.to_bytes()
    ring_cap = 1 << 28
    # the pending-event population of this script grows by ~2.6 events per handled event in this phase: refuse a
    # run that cannot stay inside the event ring instead of silently shrinking the windows
    if total_steps * eps * 4 > ring_cap:
        raise SystemExit(f"bench.py: {total_steps} windows of {eps} events do not fit the {ring_cap}-entry event ring; lower --steps or --events-per-step")
    t0 = time.time()
    rec_cap = int((total_steps * eps + 4 * sc.n_creates()) * 48 + (64 << 20))
    eng = E.Engine(sc_bytes, device=local_rank, mode=E.MODE_AUTO, records=True, record_cap=rec_cap, ring_cap=ring_cap)
    create_s = time.time() - t0
    c_init = eng.counters()
    import ctypes as C
    # pinned record ring: two slots, the copy of window k (side stream) overlaps the run of window k+1
    ring = [torch.empty(int(eps * 48 + (16 << 20)), dtype=torch.uint8).pin_memory() for _ in range(2)]
    rec_off = eng.digest()["record_bytes"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        """one window through the C ABI with host buffers: control word in, the window's records out"""
        nonlocal rec_off
        t_a = time.perf_counter()
        rc = eng.run((i + 1) * eps)
        if rc != 0:
            raise SystemExit(f"engine status {rc}: {E.strerror(rc)}")
        dev_ms = eng.device_ms
        n, tot = C.c_size_t(), C.c_size_t()
        slot = ring[i & 1]
        E.lib().ns_dev_records_wait(eng.h)        # slot i&1 was filled two windows ago; make sure that copy has landed
        E.lib().ns_dev_records_async(eng.h, C.c_void_p(slot.data_ptr()), slot.numel(), rec_off, C.byref(n), C.byref(tot))
        got = n.value
        rec_off += got
        return dev_ms, time.perf_counter() - t_a, got

    sampler = ClockSampler(local_rank)
    sampler.start()                      # the windows are milliseconds long: sample from the warm-up on
    for i in range(args.warmup):
        step(i)
    c0 = eng.counters()
    launches0 = eng.launches
    barrier()
    dev_ms_sum, wall_sum, d2h = 0.0, 0.0, 0
    t_region = time.perf_counter()
    for i in range(args.warmup, total_steps):
        dm, w, got = step(i)
        sampler.sample()                 # at least one reading per window, taken right after its kernel
        dev_ms_sum += dm
        wall_sum += w
        d2h += got
    t_w = time.perf_counter()
    E.lib().ns_dev_records_wait(eng.h)            # the last window's records must be on the host inside the timed region
    wall_sum += time.perf_counter() - t_w
    barrier()
    t_region = time.perf_counter() - t_region
    sampler.stop_flag = True
    launches1 = eng.launches
    c1 = eng.counters()
    dig = eng.digest()
    d = {k: c1[k] - c0[k] for k in c1}
    events = d["events"]
    # parity inside the bench (outside the timed region): the digest of everything simulated so far against the CPU
    # oracle's (committed for the default arguments; tools/make_bench_expected.py)
    exp = expected_digest(args, rank, total_steps * eps)
    parity = {"checked": exp is not None}
    if exp is not None:
        parity.update({"ok": bool(exp["record_hash"] == dig["record_hash"] and exp["record_bytes"] == dig["record_bytes"] and
                                  exp["events"] == c1["events"] and exp["solves"] == c1["solves"] and exp["seq"] == c1["seq"]),
                       "expected_record_hash": exp["record_hash"], "against": "CPU oracle (oracle/ns_oracle.c, pinned to the compiled reference)"})
        if not parity["ok"]:
            raise SystemExit(f"bench.py: PARITY FAILURE on rank {rank}: digest {dig} counters {c1} expected {exp}")

    # max over ranks of the timed quantities, sum of the work
    vals = torch.tensor([dev_ms_sum, wall_sum, t_region, float(events), float(d["solves"])], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        digs = torch.zeros(world, 4, dtype=torch.int64, device="cuda")
        mine = torch.tensor([dig["events"], dig["record_bytes"], dig["record_hash"] & 0x7FFFFFFFFFFFFFFF, dig["status"]],
                            dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(digs, mine)          # the only cross-GPU exchange: per-replica digests
        dev_ms_max, wall_max, region_max = mx[0].item(), mx[1].item(), mx[2].item()
        events_all, solves_all = sm[3].item(), sm[4].item()
    else:
        dev_ms_max, wall_max, region_max = dev_ms_sum, wall_sum, t_region
        events_all, solves_all = float(events), float(d["solves"])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = events_all / (dev_ms_max / 1e3)
    alg_bytes = B_SOLVE * d["solves"] + B_VISIT * d["flow_visits"] + B_INSERT * d["seq"] + d2h + B_RETIME * d["retimes"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (dev_ms_sum / 1e3) / 1e9
    traffic = None                       # dram bytes per launch from the committed ncu --set full capture
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if args.workload == "swarm100k" and tr.get("peers") == args.peers and tr.get("events_per_step") == eps:   # (the capture is of this workload's kernel)
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    config.update({"flows": c_init["creates"], "create_init_s": round(create_s, 2),
                   "pending_events_at_end": c1["seq"] - c1["events"], "batches_timed": d["batches"], "max_batch": c1["max_batch"],
                   "parallel_fraction": d["par_events"] / max(1, events), "solves_per_s": solves_all / (dev_ms_max / 1e3),
                   "full_spreads": d["full_spreads"], "flow_visits": d["flow_visits"], "record_hash": dig["record_hash"], "parity": parity,
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"})
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "ours", "config": config,
            "clocks": sampler.summary(),
            "e2e": {"value": events_all / wall_max, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": d2h // max(1, args.steps),
                    "note": "wall clock per window through ns_dev_run + ns_dev_records_async (pinned two-slot record ring, side stream; the last "
                            "copy is waited for inside the timed region); script upload and init (create_init_s) happen once before "
                            "the windows, like the reference's untimed module init"},
            "gpu_launches": launches1 - launches0,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "k_run_par (persistent, one launch per window)", "algorithmic_bytes_per_launch": alg_bytes / args.steps}}
    if not args.no_cpu_baseline and world == 1:
        r = reference_arm(args, 1)       # the same windows as the reference arm (same absolute event range)
        if r and "error" not in r:
            line["cpu_baseline"] = {"value": r["events_per_s"], "unit": UNIT, "cores": 1, "kind": "reference",
                                    "sample": f"unmodified reference core (oracle/_ref) on the same script, the same windows as --impl reference: "
                                              f"{r['steps']} windows of {eps} events after {args.warmup} warm-up windows; init {r['init_seconds']:.1f}s untimed"}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": (r or {}).get("error", "oracle/_ref missing")}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------
# --workload ensemble: BASELINE.json configs[4] — R seeded replicas of a 10k-peer swarm, sharded over the ranks
# ---------------------------------------------------------------------------------------------------------------
ENS_METRIC = "simulated events/s (ensemble: seeded replicas of a swarm run to completion, sharded over the GPUs)"


def ens_script(args, seed_index):
    """replica `seed_index`: 10k peers + 16 servers + 4 cloud seeders (SURVEY §8d config 5 shape), every peer fetches one
    catalogue entry in 1+Poisson(2) strips; run to completion (every flow done and closed)"""
    from netsim_b200 import scenario as S
    if getattr(args, "ens_kind", "swarm10k") == "p2p1k":
        # BASELINE.json configs[1] shape: 1k clients + 4 servers, arrivals N(3600 s, 900 s) (modules/p2p.c:178), every client plays one
        # 900 s / 2000 Kbit/s resource fetched with client_new_play1 (share metric) and runs the water-mark state machine of
        # modules/client_behaviour.c on the device until it has played it; the reference arm runs the reference's own client_behaviour.c
        return S.p2p_clients(1000, 4, n_catalogue=20, play_seconds=900.0, arrival_window=0, metric=1, seed=1000 + seed_index, rand_seed=1 + seed_index)
    return S.swarm_fast(args.ens_peers, 20, n_catalogue=32, seeds_per_peer=6, mean_strips=2.0, hub_frac=0.02, res_len=600000,
                        seed=5000 + seed_index, rand_seed=1 + seed_index)


def ens_config(args, world):
    if args.ens_kind == "p2p1k":
        return {"workload": f"BASELINE.json configs[1] as an ensemble: {args.replicas} seeded replicas of a p2p swarm of 1000 clients + 4 servers, to completion: arrivals "
                            f"N(3600 s, 900 s) as modules/p2p.c:178, every client fetches one 900 s resource with client_new_play1 (server_picker_opt1 + qselect, share "
                            f"metric) and runs the water-mark state machine of modules/client_behaviour.c ON THE DEVICE (node-state records included) until DONE_PLAY; "
                            f"~313k events in ~41k timestamps per replica (mean batch 7.6: the reference's real regime); replica i -> rank i mod N; the reference arm "
                            f"runs the reference's own client_behaviour.c through the replay module",
                "replicas": args.replicas, "clients_per_replica": 1000, "engine": "one warp per replica (k_run_seq; NS_ENS_MODE=auto adds the one-CTA-per-replica engine for dense phases, which these replicas do not have)",
                "working_set": "per replica ~2 MB of tables; the ensemble's tables + event rings exceed the 126 MB L2; no explicit L2 flush"}
    return {"workload": f"BASELINE.json configs[4]: ensemble of {args.replicas} seeded replicas of a {args.ens_peers}-peer swarm script (20 seeders, ~3 flows "
                        f"per peer, file distribution to completion: every flow runs to FLOW_DONE and is closed), replica i -> rank i mod N; "
                        f"step = one ensemble ({args.replicas} replicas over all ranks); per-replica 56-byte digests gathered with one NCCL all_reduce",
            "replicas": args.replicas, "peers_per_replica": args.ens_peers, "engine": "one CTA per replica (k_run_ens), the hardware block scheduler is the replicas' work queue",
            "working_set": "per replica ~50 MB of tables + event ring; the rank's replicas together exceed the 126 MB L2; no explicit L2 flush"}


def ens_expected():
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "bench_expected.json")))
    except Exception:
        return {}


def ens_reference(args, n_proc, per_proc):
    """min(replicas, cores) reference processes side by side, each running `per_proc` replicas one after the other (the unmodified
    reference core on the same scripts, to completion); value = all events / the slowest process's loop time"""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import refrun
    if not refrun.available():
        return None

    def worker(w):
        ev, sec = 0, 0.0
        for j in range(per_proc):
            r = refrun.run_reference(ens_script(args, w * per_proc + j).to_bytes(), trace=False, record=True, timeout=args.ref_timeout)
            if r["returncode"] != 0:
                return {"error": r["stderr"][-200:]}
            ev += r["stats"]["events"]
            sec += r["stats"]["loop_seconds"]
        return {"events": ev, "seconds": sec}
    with ThreadPoolExecutor(max_workers=n_proc) as ex:
        rs = list(ex.map(worker, range(n_proc)))
    bad = [r["error"] for r in rs if "error" in r]
    if bad:
        return {"error": bad[0]}
    ev = sum(r["events"] for r in rs)
    sec = max(r["seconds"] for r in rs)
    return {"events": ev, "seconds": sec, "events_per_s": ev / sec, "processes": n_proc, "replicas": n_proc * per_proc}


def ensemble_main(args, rank, local_rank, world):
    config = ens_config(args, world)
    if args.impl == "reference":
        if rank != 0:
            return 0
        n_proc = max(1, min(args.replicas, os.cpu_count() or 1))
        per = max(1, min(4, args.replicas // n_proc))
        r = ens_reference(args, n_proc, per)
        if r is None or "error" in r:
            emit({"impl": "reference", "unavailable": (r or {}).get("error", "oracle/_ref not built")})
            return 0
        val = r["events_per_s"]
        emit({"metric": ENS_METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": 1e3 * r["seconds"],
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
              "cpu_baseline": {"value": val, "unit": UNIT, "cores": n_proc, "kind": "reference",
                               "sample": f"{r['replicas']} of the {args.replicas} replicas (seeds 0..{r['replicas'] - 1}), {n_proc} processes of the unmodified "
                                         f"reference core side by side, {per} replicas each, every one to completion; module init untimed"},
              "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from netsim_b200 import engine as E
    if E.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(base, n_total):
        """replicas base+rank, base+rank+world, ... : host scripts -> device tables (H2D), run, digests (D2H), gather"""
        mine = list(range(rank, n_total, world))
        t0 = time.perf_counter()
        scs = [ens_script(args, base + i).to_bytes() for i in mine]
        t_gen = time.perf_counter() - t0
        t1 = time.perf_counter()
        # per-replica capacities (overridable for experiments): the event ring and the record buffer are cyclic / folded in place
        ring_log2 = int(os.environ.get("NS_ENS_RING_LOG2", 17 if args.ens_kind == "p2p1k" else 20))
        rec_kb = int(os.environ.get("NS_ENS_REC_KB", (2 if args.ens_kind == "p2p1k" else 8) << 10))
        # config-2-shaped replicas never see a batch that is worth a CTA (mean 7.6 events per timestamp): the one-warp engine alone, no rounds
        ens_mode = {"auto": E.MODE_AUTO, "seq": E.MODE_SEQ}[os.environ.get("NS_ENS_MODE", "seq" if args.ens_kind == "p2p1k" else "auto")]
        ens = E.Ensemble(scs, device=local_rank, records=True, ring_cap=1 << ring_log2, record_cap=rec_kb << 10, mode=ens_mode)
        if args.ens_kind == "p2p1k":
            ens.analysis_begin(3600.0)                       # analyzer reductions on the device, from the records while they are produced
        ens.run(0)
        dg = ens.digests()
        ana = ens.analysis() if args.ens_kind == "p2p1k" else None
        tab = torch.zeros(n_total, 6, dtype=torch.int64)
        for i, d in zip(mine, dg):
            tab[i] = torch.tensor([d["events"], d["solves"], d["record_bytes"], d["record_hash"] & 0x7FFFFFFFFFFFFFFF, d["status"],
                                   int(np.float64(d["t_end"]).view(np.int64))], dtype=torch.int64)
        tab = tab.cuda()
        if world > 1:
            dist.all_reduce(tab)                             # the only cross-GPU exchange: every rank ends up with every replica's digest
        torch.cuda.synchronize()
        wall = time.perf_counter() - t1
        cs = ens.counters_sum()
        summary = None
        if ana:                                              # what a sweep would keep per replica instead of its records
            summary = {"stale_client_seconds": float(sum(a["stale"].sum() for a in ana)), "max_stale_clients_in_a_replica": int(max(a["stale_max"] for a in ana)),
                       "server_out_Kbit": float(sum(a["usage"][0, 0].sum() for a in ana)), "client_in_Kbit": float(sum(a["usage"][2, 1].sum() for a in ana)),
                       "state_records": int(sum(a["state_records"] for a in ana)), "usage_records": int(sum(a["usage_records"] for a in ana))}
        out = {"device_ms": ens.device_ms, "wall": wall, "gen": t_gen, "launches": ens.launches, "digests": tab.cpu(), "counters": cs, "analysis": summary,
               "h2d": sum(len(b) for b in scs), "d2h": 56 * len(mine)}
        ens.close()
        return out

    sampler = ClockSampler(local_rank)
    sampler.start()
    for w in range(args.warmup):                              # small untimed ensembles on seeds outside the timed range (kernel load, allocator)
        one_pass(10_000_000 + 1000 * w, max(world * 8, 16))
    barrier()
    t_region = time.perf_counter()
    passes = [one_pass(args.replicas * k, args.replicas) for k in range(args.steps)]
    barrier()
    t_region = time.perf_counter() - t_region
    sampler.stop_flag = True

    dev_ms = sum(p["device_ms"] for p in passes)
    wall = sum(p["wall"] for p in passes)
    dig = passes[0]["digests"]
    events_all = float(sum(int(p["digests"][:, 0].sum()) for p in passes))
    solves_all = float(sum(int(p["digests"][:, 1].sum()) for p in passes))
    bad = int((passes[0]["digests"][:, 4] != 0).sum())
    if bad:
        raise SystemExit(f"bench.py: {bad} replicas ended with a non-zero status: {passes[0]['digests'][:, 4].tolist()[:16]}")
    cnt = {k: sum(p["counters"][k] for p in passes) for k in passes[0]["counters"]}
    vals = torch.tensor([dev_ms, wall, t_region] + [float(cnt[k]) for k in ("solves", "flow_visits", "seq", "retimes")], dtype=torch.float64, device="cuda")
    mx, sm = vals.clone(), vals.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        per_rank = torch.zeros(world, dtype=torch.float64, device="cuda")
        per_rank[rank] = dev_ms
        dist.all_reduce(per_rank)
        per_rank = per_rank.tolist()
    else:
        per_rank = [dev_ms]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    # parity (outside the timed region): the digests of sampled seeds against the CPU oracle's committed values
    exp = ens_expected()
    checked, ok = 0, True
    for i in range(args.replicas):
        e = exp.get(f"ensemble,kind=p2p1k,seed={i}" if args.ens_kind == "p2p1k" else f"ensemble,peers={args.ens_peers},seed={i}")
        if e is None:
            continue
        checked += 1
        row = dig[i].tolist()
        if row[0] != e["events"] or (e.get("solves") is not None and row[1] != e["solves"]) or row[2] != e["record_bytes"] or row[3] != (e["record_hash"] & 0x7FFFFFFFFFFFFFFF):
            ok = False
            raise SystemExit(f"bench.py: PARITY FAILURE replica {i}: {row} expected {e}")
    dev_ms_max, wall_max = mx[0].item(), mx[1].item()
    rec_bytes = float(sum(int(p["digests"][:, 2].sum()) for p in passes))
    alg_bytes = B_SOLVE * sm[3].item() + B_VISIT * sm[4].item() + B_INSERT * sm[5].item() + B_RETIME * sm[6].item() + rec_bytes
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / world / (dev_ms_max / 1e3) / 1e9          # per GPU: the kernel of the slowest rank against one GPU's HBM
    config.update({"events": int(events_all), "solves_per_s": solves_all / (dev_ms_max / 1e3), "record_bytes_hashed_on_device": int(rec_bytes),
                   "analysis_rank0": passes[0]["analysis"],
                   "device_ms_per_rank": [round(x, 1) for x in per_rank], "script_generation_s": round(sum(p["gen"] for p in passes), 2),
                   "parity": {"checked": checked > 0, "ok": ok, "replicas_checked": checked,
                              "against": "digests of the compiled reference / the CPU oracle pinned to it (tests/golden/bench_expected.json, tools/make_bench_expected.py)"},
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"})
    line = {"metric": ENS_METRIC, "value": events_all / (dev_ms_max / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "ours", "config": config, "clocks": sampler.summary(),
            "e2e": {"value": events_all / wall_max, "unit": UNIT, "h2d_bytes_per_step": passes[0]["h2d"], "d2h_bytes_per_step": passes[0]["d2h"],
                    "note": "wall clock from the replicas' scripts in host memory to the gathered digests: table build + upload (ns_ens_create), "
                            "init and run on the device (ns_ens_run), digest kernel + copy back, NCCL all_reduce of the digest table; script "
                            "generation (numpy) is outside"},
            "gpu_launches": sum(p["launches"] for p in passes),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "k_run_ens (one CTA per replica)", "algorithmic_bytes_per_launch": alg_bytes / world / max(1, args.steps)}}
    if not args.no_cpu_baseline and world == 1:
        n_proc = max(1, min(args.replicas, os.cpu_count() or 1))
        r = ens_reference(args, n_proc, 1)
        if r and "error" not in r:
            line["cpu_baseline"] = {"value": r["events_per_s"], "unit": UNIT, "cores": n_proc, "kind": "reference",
                                    "sample": f"{r['replicas']} of the replicas (seeds 0..{r['replicas'] - 1}) on the unmodified reference core, {n_proc} processes side by side, to completion"}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": n_proc, "kind": "reference", "sample": (r or {}).get("error", "oracle/_ref missing")}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


WORKLOADS = {"swarm100k": None, "churn": None, "ensemble": ensemble_main}


if __name__ == "__main__":
    sys.exit(main())
