"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py — per-rank seeded replicas of the
workload, MAX-reduced timing, SUM-reduced work, all_gather of per-replica digests (the only
cross-rank exchange; SURVEY §8e "replicas only")."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from netsim_b200 import engine as E
    from netsim_b200 import scenario as S
    from oracle import pyoracle
    # each rank owns the replica with its own seed (bench.py: workload(args, rank)); the oracle
    # stands in for the device here because this test runs without a GPU
    sc = S.swarm_fast(300, 4, n_catalogue=8, seeds_per_peer=3, mean_strips=3.0, seed=12345 + rank, rand_seed=1 + rank)
    o = pyoracle.Oracle(sc.to_bytes(), trace=False)
    assert o.run(20000) == 0
    c = o.counters()
    h = E.stream_hash(o.records())
    vals = torch.tensor([float(10 + rank), float(c["events"])], dtype=torch.float64)
    mx, sm = vals.clone(), vals.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    mine = torch.tensor([c["events"], len(o.records()), h & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64)
    digs = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(digs, mine)
    if rank == 0:
        out.put((mx.tolist(), sm.tolist(), [d.tolist() for d in digs]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replicas_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    mx, sm, digs = q.get(timeout=120)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert mx[0] == 11.0                       # time = max over ranks
    assert sm[1] == digs[0][0] + digs[1][0]    # work = sum over ranks
    assert digs[0][2] != digs[1][2]            # different seeds -> different record streams
    # a replica's digest does not depend on which rank ran it
    sys.path.insert(0, ROOT)
    from netsim_b200 import engine as E
    from netsim_b200 import scenario as S
    from oracle import pyoracle
    sc = S.swarm_fast(300, 4, n_catalogue=8, seeds_per_peer=3, mean_strips=3.0, seed=12346, rand_seed=2)
    o = pyoracle.Oracle(sc.to_bytes(), trace=False)
    o.run(20000)
    assert digs[1][2] == E.stream_hash(o.records()) & 0x7FFFFFFFFFFFFFFF


def _ens_worker(rank, world, port, out):
    """bench.py --workload ensemble, host logic: replica i -> rank i mod N, every rank fills its rows of the digest table,
    one all_reduce leaves the whole table on every rank (oracle instead of the device: no GPU here)"""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import argparse
    import bench
    from netsim_b200 import engine as E
    from oracle import pyoracle
    n_total = 5
    args = argparse.Namespace(ens_peers=200)
    tab = torch.zeros(n_total, 4, dtype=torch.int64)
    for i in range(rank, n_total, world):
        o = pyoracle.Oracle(bench.ens_script(args, i).to_bytes(), trace=False)
        assert o.run(0) == 0
        tab[i] = torch.tensor([o.counters()["events"], o.counters()["solves"], len(o.records()), E.stream_hash(o.records()) & 0x7FFFFFFFFFFFFFFF])
    dist.all_reduce(tab)
    if rank == 0:
        out.put(tab.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ensemble_digest_table():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_ens_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    tab = q.get(timeout=120)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    # the gathered table equals what ONE process computes for all replicas: sharding does not change a digest
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    from netsim_b200 import engine as E
    from oracle import pyoracle
    for i, row in enumerate(tab):
        o = pyoracle.Oracle(bench.ens_script(argparse.Namespace(ens_peers=200), i).to_bytes(), trace=False)
        o.run(0)
        assert row == [o.counters()["events"], o.counters()["solves"], len(o.records()), E.stream_hash(o.records()) & 0x7FFFFFFFFFFFFFFF]
    assert len({r[3] for r in tab}) == len(tab)
