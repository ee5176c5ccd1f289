#!/usr/bin/env python
"""Config D of BASELINE.json across the GPUs of one box: RLC ladder (N = 64: 194 complex unknowns) AC log sweep of 1 000 000
frequency points, sharded in contiguous blocks of points, one process per GPU, no data-path collective; at the
end the complex solution of every point is gathered on every rank with ONE NCCL all-gather over NVLink and timed on the device.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/config_d_multi.py
  python tools/config_d_multi.py            # one GPU, all points

Prints one JSON line on rank 0: points/s over all ranks (max over ranks of the device time of analyze()), the gather's time,
bytes and bus bandwidth, and a parity spot check of the gathered result against the compiled reference (when present)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pe_b200 as pe  # noqa: E402
import sharding  # noqa: E402
import workloads as wl  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    points = int(os.environ.get("PE_CFG_POINTS", "1000000"))
    n_sections = int(os.environ.get("PE_CFG_SECTIONS", "64"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    nl, info = wl.rlc_ladder(n_sections)
    c = pe.Circuit(nl)
    c.set_analyze_type(pe.AC)
    b = c.batch(1)
    b.set_device(local)
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    b.set_ac_sweep(pe.SWEEP_LOG, 1e3, 1e10, points)
    lo, hi = sharding.shard_range(points, rank, world)
    b.set_ac_slice(lo, hi - lo)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    b.analyze()  # compile + warm-up
    sync()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if not b.analyze():
            raise SystemExit("analyze failed: " + c.abi.last_error())
    e1.record()
    sync()
    ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    n = b.n_unknowns()
    # ---- the one collective of the path: gather the sweep results (complex solution of every point) on every rank ----
    x_local = b.ac_solution()[0]  # [points_local, n] complex128 (host)
    per = -(-points // world)
    blk = torch.zeros((per, n, 2), dtype=torch.float64, device=dev)
    blk[: hi - lo] = torch.from_numpy(np.ascontiguousarray(np.stack((x_local.real, x_local.imag), axis=-1))).to(dev)
    full = torch.empty((world * per, n, 2), dtype=torch.float64, device=dev)
    gather_ms, bus = 0.0, 0.0
    if world > 1:
        dist.all_gather_into_tensor(full, blk)  # warm-up (communicator set-up)
        sync()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(reps):
            dist.all_gather_into_tensor(full, blk)
        g1.record()
        sync()
        tg = torch.tensor([g0.elapsed_time(g1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather_ms = float(tg.item())
        bus = blk.numel() * 8 * (world - 1) / (gather_ms * 1e-3) / 1e9  # bytes every rank receives / time
    else:
        full[:per] = blk
    solves, failed = sharding.reduce_counters(b.total_solves, int((b.status() != 0).sum()), dist if world > 1 else None, dev)
    if rank == 0:
        parity = "unchecked"
        try:
            import refapi

            if os.path.exists(refapi.REF_LIB):
                # the reference on a 257-point sweep of the same ladder; the sharded run of the same sweep must agree
                r = refapi.RefCircuit(nl)
                r.set_analyze_type(pe.AC)
                r.set_ac_sweep(pe.SWEEP_LOG, 1e3, 1e10, 257)
                r.analyze_counted()
                om, xr = r.ac_results()
                c2 = pe.Circuit(nl)
                c2.set_analyze_type(pe.AC)
                b2 = c2.batch(1)
                b2.set_ac_sweep(pe.SWEEP_LOG, 1e3, 1e10, 257)
                b2.set_ac_slice(100, 57)
                b2.analyze()
                got = b2.ac_solution()[0]
                okk = bool((b2.ac_omegas() == om[100:157]).all()) and bool((np.abs(got - xr[100:157]) <= 1e-12 + 1e-9 * np.maximum(np.abs(got), np.abs(xr[100:157]))).all())
                parity = "a sliced 257-point sweep matches the reference (omegas bit-identical, 1e-9 / 1e-12)" if okk else "MISMATCH"
        except Exception as ex:  # noqa: BLE001
            parity = f"check failed: {ex}"
        st = b.stats(pe.MODE_AC)
        bpp = 16 * (3 * st["nnz_a"] + 2 * st["n_unknowns"])
        pps = points / (ms_max * 1e-3)
        full_host_bytes = full.numel() * 8
        print(json.dumps({"case": f"D rlc_ladder N={n_sections} AC log sweep sharded over {world} GPU(s)", "points": points, "points_per_rank": per, "ms_analyze_max_over_ranks": ms_max,
                          "points_per_s": pps, "hbm_line_frac_of_6548GBs_per_gpu": pps * bpp / world / 6548.5e9, "solves": solves, "failed": failed,
                          "gather": {"collective": "ncclAllGather (torch.distributed.all_gather_into_tensor)", "ms": gather_ms, "bytes_gathered_per_rank": full_host_bytes,
                                     "bus_GBps_per_rank": bus}, "gather_share_of_run": gather_ms / (gather_ms + ms_max) if world > 1 else 0.0, "kernel": b.resident_info(pe.MODE_AC),
                          "parity": parity, "checksum": float(full.abs().sum().item())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
