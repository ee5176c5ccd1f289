"""N > 1 path on CPU: two gloo ranks each solve their contiguous shard of a Monte-Carlo batch (host emulator backend),
results are gathered and counters reduced, and the union equals the single-process run."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import sharding  # noqa: E402


def test_shard_ranges_cover_everything():
    for n in (1, 7, 64, 100, 10000):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(n, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            for a, b in zip(r, r[1:]):
                assert a[1] == b[0]
            assert all(hi - lo <= -(-n // world) for lo, hi in r)


def _solve_shard(rank, world, n_total):
    import emuapi
    import pe_b200 as pe
    import workloads as wl

    nl, info = wl.diode_resistor(n_diodes=2, v=3.0)
    rng = np.random.default_rng(42)
    r_all = wl.sweep_values(rng, 1e3, n_total, 0.9, 1.1)
    lo, hi = sharding.shard_range(n_total, rank, world)
    c = pe.Circuit(nl, emuapi.emulator())
    c.set_analyze_type(pe.OP)
    b = c.batch(hi - lo)
    b.set_param(info["R"], "r", sharding.shard_values(r_all, rank, world))
    ok = b.analyze()
    return ok, b.solution(), int(b.total_solves)


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok, x, solves = _solve_shard(rank, world, n_total)
    full = sharding.gather_instances(x, n_total, dist)
    total, failed = sharding.reduce_counters(solves, 0 if ok else 1, dist)
    if rank == 0:
        q.put((full, total, failed))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_equals_single_process():
    import torch.multiprocessing as mp

    n_total = 45
    ok, x1, solves1 = _solve_shard(0, 1, n_total)
    assert ok
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, total, failed = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert failed == 0
    assert total == solves1
    assert full.shape == x1.shape
    assert (full == x1).all()  # same kernels, same inputs: bit-identical regardless of the sharding


def test_ac_sweep_slices_reassemble_the_sweep():
    # one rank's shard of an AC sweep = a contiguous block of the reference's cumulative-product omega table; the blocks of all
    # ranks put together are the single-process sweep, bit for bit
    import emuapi
    import pe_b200 as pe
    import workloads as wl

    abi = emuapi.emulator()
    nl, _ = wl.rlc_ladder(6)
    points, world = 203, 4

    def run(first, count):
        c = pe.Circuit(nl, abi)
        c.set_analyze_type(pe.AC)
        b = c.batch(1)
        b.set_ac_sweep(pe.SWEEP_LOG, 1e3, 1e9, points)
        b.set_ac_slice(first, count)
        assert b.analyze(), abi.last_error()
        return b.ac_omegas(), b.ac_solution()[0]

    om, x = run(0, 0)
    assert len(om) == points
    parts = [run(*[(lo, hi - lo) for lo, hi in [sharding.shard_range(points, r, world)]][0]) for r in range(world)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), om)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), x)
