#!/usr/bin/env python
"""bench.py — circuit-solves/s of the batched MNA hot path on N B200s (BASELINE.json metric).

Workload (BASELINE.json configs[1]): RC ladder, 1000 sections (1002 unknowns, nnz(A) = 3003),
transient, t_step 1e-8 / t_stop 1e-6 (100 time steps), 10 000 instances per GPU with element-wise per-instance R_i, C_i
draws (a batched parameter sweep).  One "step" = one analyze() of that batch = 1e6 solve_once-equivalents
(stamp + numeric LU refactor + forward/back substitution) per GPU.  Instances shard across ranks with no data-path
collective (weak scaling: 10 000 instances per GPU).

  python bench.py --gpus 1 --steps K --warmup W            # this framework (CUDA kernels through the C ABI)
  python bench.py --impl reference ...                     # the reference's own CPU solver (oracle/_ref), host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))

METRIC = "circuit-solves/sec"
UNIT = "solves/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sections", type=int, default=1000)
    ap.add_argument("--instances", type=int, default=10000, help="instances per GPU")
    ap.add_argument("--time-steps", type=int, default=100)
    ap.add_argument("--subtree-warps", type=int, default=0, help="0 = library default")
    ap.add_argument("--resident", default="0,0,0", help="streams,instances_per_cta,instances_per_thread of the resident kernel (0 = automatic, streams -1 = HBM-streaming kernel)")
    ap.add_argument("--workspace", type=int, default=0, help="0 = automatic, 1 = shared memory (resident kernel), 2 = HBM (tree-streaming kernel)")
    ap.add_argument("--chunks", type=int, default=0, help="tree-streaming kernel: chunks of the time loop (0 = automatic, 1 = static scheduling)")
    ap.add_argument("--tuning", type=int, default=0, help="circuit_batch_set_tuning flags (bit 0/1 L2 prefetch, bit 2 no L1 re-fetch, bit 3 fused elimination steps, bit 4 require / bit 5 forbid the specialised kernel)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e leg on ONE batch handle / stream: H2D -> solve -> D2H strictly in sequence (default: two batch handles on two host "
                                                              "threads / streams, so that a step's copies overlap the other step's solve)")
    return ap.parse_args()


def workload(args, seed):
    import workloads as wl

    nl, info = wl.rc_ladder(args.sections)
    rng = np.random.default_rng(seed)
    items = [(e, "r") for e in info["R"]] + [(e, "c") for e in info["C"]]
    nominal = np.array([1e3] * len(info["R"]) + [1e-9] * len(info["C"]))
    return nl, info, items, nominal, rng


def config_of(args):
    ws = "1.3"  # 8 B x instances x (6002 persistent + 10 007 workspace rows) at the default size; the same text in both arms
    return {
        "workload": f"RC ladder {args.sections} sections ({args.sections + 2} unknowns) transient, {args.time_steps} time steps, "
                    f"{args.instances} instances/GPU batched R_i,C_i parameter sweep (BASELINE.json configs[1])",
        "instances_per_gpu": args.instances,
        "time_steps": args.time_steps,
        "t_step": 1e-8,
        "l2_policy": f"inputs larger than L2: the per-GPU working set (about {ws} GB of per-instance parameters, state and LU workspace at 10 000 instances) is streamed every time step "
                     "and is ~10x the 126 MB L2; no explicit flush",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(args, n_inst, seed=1234, threads=None):
    """The reference's own CPU implementation (Eigen SparseLU behind circult::analyze) on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refapi
    import pe_b200 as pe

    fast = os.path.exists(refapi.REF_LIB_FAST)
    nl, info, items, nominal, rng = workload(args, seed)
    vals = nominal[:, None] * rng.uniform(0.8, 1.2, size=(len(items), n_inst))
    over = [(e, name, vals[k]) for k, (e, name) in enumerate(items)]
    threads = threads or os.cpu_count() or 1
    t0 = time.perf_counter()
    r = refapi.run_batch(nl, pe.TR, n_inst, over, t_step=1e-8, t_stop=1e-8 * (args.time_steps - 0.5), threads=threads, n_unknowns=args.sections + 2, fast=fast)
    wall = time.perf_counter() - t0
    solves = int(r["solves"].sum())
    return {"solves": solves, "wall_s": wall, "analyze_s": r["seconds"], "threads": threads, "ok": bool((r["ok"] == 1).all()),
            "build": "oracle/_ref/libpe_ref_fast.so (-O3 -march=x86-64-v3)" if fast else "oracle/_ref/libpe_ref.so (-O2)"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # each step = a bounded sample of the workload (~3-5 s of CPU work on all host cores): 48 instances per core
    n_inst = args.cpu_sample or 48 * cores
    for _ in range(args.warmup):
        cpu_reference_run(args, max(1, 4 * cores))
    # the clock is the harness' own: the slowest worker's time inside circult::analyze() (netlist construction and the
    # name-scanned parameter overrides of every fresh circuit are not the path under test and are left out, as in
    # oracle/ref_harness.cpp:399-401)
    t_total, t_wall, solves = 0.0, 0.0, 0
    for _ in range(args.steps):
        r = cpu_reference_run(args, n_inst)
        t_total += r["analyze_s"]
        t_wall += r["wall_s"]
        solves += r["solves"]
    v = solves / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["threads"], "kind": "reference",
                         "sample": f"{n_inst} instances x {args.time_steps} time steps per step, {r['build']}, one circult per instance, {r['threads']} worker threads, clock = time inside "
                                   f"analyze() of the slowest worker ({t_total:.2f} s of {t_wall:.2f} s wall)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun children of every rank) write
    to file descriptor 1 too, so everything that is not the line goes to stderr: fd 1 is pointed at fd 2 for the run and the
    line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import pe_b200 as pe

    if not torch.cuda.is_available() or pe.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION in this image) goes there
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    nl, info, items, nominal, rng = workload(args, 1000 + rank)
    n_inst = args.instances
    P = len(items)
    host_vals = torch.empty((P, n_inst), dtype=torch.float64).pin_memory()
    host_vals.numpy()[:] = nominal[:, None] * rng.uniform(0.8, 1.2, size=(P, n_inst))
    n_unk = args.sections + 2
    host_x = torch.empty((n_unk, n_inst), dtype=torch.float64).pin_memory()

    c = pe.Circuit(nl)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-8, 1e-8 * (args.time_steps - 0.5))
    b = c.batch(n_inst)
    b.set_device(local_rank)
    stream = torch.cuda.current_stream()
    b.set_stream(stream.cuda_stream)
    if args.subtree_warps:
        b.set_subtree_warps(args.subtree_warps)
    b.set_resident(*[int(v) for v in args.resident.split(",")])
    b.set_workspace(args.workspace)
    b.set_chunks(args.chunks)
    b.set_tuning(args.tuning)
    table = b.param_table(items)
    b.set_params(table, host_vals.data_ptr())  # first time: host copies + layout
    b.prepare()
    st = b.stats(pe.MODE_TR)
    # s * (nnz(A) + 2 nnz(L+U) + 2 n) with the fill-free factor of the natural (chain) order,
    # nnz(L+U) = nnz(A): 88 104 B for the 1000-section ladder, whatever ordering the schedule itself uses
    bytes_per_solve = 8 * (st["nnz_a"] + 2 * st["nnz_a"] + 2 * st["n_unknowns"])
    rinfo = b.resident_info(pe.MODE_TR)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def step_resident():
        b.reset_state()
        if not b.analyze():
            raise SystemExit("bench.py: analyze failed: " + c.abi.last_error())
        return b.total_solves

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    pe.kernel_timing(True)
    pe.kernel_ms()
    l0 = pe.launch_count()
    a0 = pe.aux_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    solves = 0
    for _ in range(args.steps):
        solves += step_resident()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = pe.launch_count() - l0  # solve-kernel launches (the roofline's unit)
    aux_launches = pe.aux_launch_count() - a0  # + one status-reduction kernel per analyze()
    kernel_ms = pe.kernel_ms()
    pe.kernel_timing(False)
    clocks = sampler.stop() if sampler else None
    ms_max = max_over_ranks(ms)
    total_solves = sum_over_ranks(float(solves))
    value = total_solves / (ms_max * 1e-3)

    # ---- end to end through the C ABI with host buffers ----
    # Every step copies its inputs from pinned host memory (H2D), solves, and reads the final state of every instance back
    # (D2H), all inside the timed region.  Consecutive steps are independent batches, so two host threads drive two batch
    # handles on two CUDA streams (the reference's threading rule: one circuit per thread, dll_main.cpp has no locks): the
    # copies of one step overlap the solve of the other (the second handle costs its own 1.3 GB of the 180 GB).
    # --e2e-serial: one handle, H2D -> solve -> D2H strictly in sequence.
    e2e = None
    if not args.no_e2e:
        import threading

        n_pipe = 1 if args.e2e_serial else 2
        if n_pipe > 1:
            # the legacy default stream synchronises with every other stream: the pipelined legs run on streams of their own
            stream = torch.cuda.Stream(device=dev)
            b.set_stream(stream.cuda_stream)
        lanes = [(c, b, table, host_vals, host_x, stream)]
        for _ in range(n_pipe - 1):
            c2 = pe.Circuit(nl)
            c2.set_analyze_type(pe.TR)
            c2.set_tr(1e-8, 1e-8 * (args.time_steps - 0.5))
            b2 = c2.batch(n_inst)
            b2.set_device(local_rank)
            st2 = torch.cuda.Stream(device=dev)
            b2.set_stream(st2.cuda_stream)
            if args.subtree_warps:
                b2.set_subtree_warps(args.subtree_warps)
            b2.set_resident(*[int(v) for v in args.resident.split(",")])
            b2.set_workspace(args.workspace)
            b2.set_chunks(args.chunks)
            b2.set_tuning(args.tuning)
            hv2 = torch.empty((P, n_inst), dtype=torch.float64).pin_memory()
            hv2.copy_(host_vals)
            hx2 = torch.empty((n_unk, n_inst), dtype=torch.float64).pin_memory()
            t2 = b2.param_table(items)
            b2.set_params(t2, hv2.data_ptr())
            b2.prepare()
            lanes.append((c2, b2, t2, hv2, hx2, st2))

        def e2e_steps(lane, count, out):
            cc, bb, tt, hv, hx, _ = lane
            torch.cuda.set_device(local_rank)
            n = 0
            for _ in range(count):
                bb.set_params(tt, hv.data_ptr())  # H2D of this step's inputs (pinned)
                bb.reset_state()
                if not bb.analyze():
                    out.append("analyze failed: " + cc.abi.last_error())
                    return
                bb.solution_soa_into(hx.data_ptr())  # D2H of the step's result (final state of every instance)
                n += bb.total_solves
            out.append(n)

        def e2e_run(total):
            outs = [[] for _ in lanes]
            counts = [total // len(lanes) + (1 if k < total % len(lanes) else 0) for k in range(len(lanes))]
            th = [threading.Thread(target=e2e_steps, args=(lanes[k], counts[k], outs[k])) for k in range(len(lanes))]
            for t in th:
                t.start()
            for t in th:
                t.join()
            torch.cuda.synchronize()
            for o in outs:
                if not o or isinstance(o[0], str):
                    raise SystemExit("bench.py: e2e " + (o[0] if o else "thread died"))
            return sum(o[0] for o in outs)

        e2e_run(max(len(lanes), args.warmup // 2))
        barrier()
        t0 = time.perf_counter()
        s2 = e2e_run(args.steps)
        t1 = time.perf_counter()
        dt = max_over_ranks(t1 - t0)
        e2e = {"value": sum_over_ranks(float(s2)) / dt, "unit": UNIT, "h2d_bytes_per_step": int(P * n_inst * 8), "d2h_bytes_per_step": int(n_unk * n_inst * 8),
               "pipeline": f"{len(lanes)} batch handle(s) / host thread(s) / stream(s); every step's H2D and D2H are inside the timed region"}
        if len(lanes) > 1 and not np.array_equal(lanes[0][4].numpy(), lanes[1][4].numpy()):
            raise SystemExit("bench.py: the two e2e pipelines disagree on identical inputs")

    checksum = float(np.abs(host_x.numpy()).sum()) if e2e else None

    # ---- roofline of the dominant kernel (the solve kernel; one launch per analyze) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    rinfo = b.resident_info(pe.MODE_TR)
    specialised = b.last_kernel() == 1  # the run-time specialised tree-streaming kernel (host/jit.cpp) ran
    kernel_name = ("pe_b200_tree_kernel" if rinfo["hbm"] else "pe_b200_resident_kernel") if rinfo["resident"] else "pe_b200_solve_kernel"
    if specialised:
        kernel_name = "pe_b200_jit_kernel"
    streamed = b.last_kernel() == 2  # the stream kernel (host/stream.cpp, csrc/pe_b200_stream.cu) ran
    if streamed:
        kernel_name = "pe_b200_stream_kernel"
    solves_per_launch = solves / max(launches, 1)
    # DRAM traffic of that kernel from the committed ncu --set full capture (profiles/r01_traffic.json), scaled to one launch
    traffic, traffic_src = None, None
    tname = "r02_traffic_stream.json" if streamed else ("r01_traffic_jit.json" if specialised else "r01_traffic.json")
    tpath = os.path.join(ROOT, "profiles", tname)
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("kernel", "").startswith(kernel_name) and (rinfo["last_S"], rinfo["last_J"]) == (tj.get("S", 32), tj.get("J", 2)):
            traffic = tj["dram_bytes_per_solve"] * solves_per_launch
            traffic_src = f"profiles/{tname}: " + tj["capture"]
    avg_launch_ms = kernel_ms / max(launches, 1)
    achieved = bytes_per_solve * solves_per_launch / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": kernel_name, "avg_launch_ms": avg_launch_ms, "bytes_per_solve": bytes_per_solve, "solves_per_launch": solves_per_launch,
                "peak_source": peak_src, "kernel_share_of_step": kernel_ms / ms if ms > 0 else None}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or 192 * cores  # ~10-30 s of CPU work on all host cores
        try:
            r = cpu_reference_run(args, sample)
            cpu = {"value": r["solves"] / r["analyze_s"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                   "sample": f"{sample} instances x {args.time_steps} time steps, {r['build']}, {r['threads']} worker threads, {r['analyze_s']:.2f} s inside analyze() "
                             f"(slowest worker; wall {r['wall_s']:.2f} s)"}
        except Exception as ex:  # the compiled reference is test infrastructure; its absence must not break the bench line
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"unavailable: {ex}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args), "working_set_gb": 8e-9 * n_inst * (st["n_inst_slots"] + (rinfo["smem_slots"] if rinfo["hbm"] else st["n_lane_slots"])),
            "e2e": e2e, "gpu_launches": int(launches + aux_launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "program": st, "resident": b.resident_info(pe.MODE_TR), "specialised_kernel": bool(specialised), "stream_kernel": b.stream_info(pe.MODE_TR) if streamed else None, "checksum": checksum,
        }
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
