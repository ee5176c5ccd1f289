#!/usr/bin/env python
"""Secondary measurements (not the BENCH line): BASELINE.json configs C and D at full size on one GPU.

  C  Newton operating point Monte Carlo, 100 000 instances: (c1) V-R-diode of test/0011 (3 unknowns), (c4) 16-stage diode
     ladder (18 unknowns).  solves = sum of Newton iterations.
  D  RLC ladder AC log sweep, N = 64 sections (194 complex unknowns), 1 000 000 frequency points (one lane per point).
Prints one JSON line per case: solves/s on the device (CUDA events around analyze()), parity spot-check against the
compiled reference on a small sample when oracle/_ref is present.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pe_b200 as pe  # noqa: E402
import workloads as wl  # noqa: E402


def timed(fn, reps=3):
    import torch

    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def ref_check(nl, at, over, got_x, got_it, n_check, **kw):
    try:
        import refapi

        if not os.path.exists(refapi.REF_LIB):
            return "reference absent"
        sub = [(e, a, v[:n_check]) for e, a, v in over]
        w = refapi.run_batch(nl, at, n_check, sub, **kw)
        err = np.abs(got_x[:n_check] - w["x"].real)
        tol = 1e-12 + 1e-9 * np.maximum(np.abs(got_x[:n_check]), np.abs(w["x"].real))
        good = w["ok"] == 1
        ok = bool((err[good] <= tol[good]).all()) and bool((got_it[:n_check] == w["solves"]).all())
        return "matches reference on %d instances" % n_check if ok else "MISMATCH"
    except Exception as ex:  # noqa: BLE001
        return f"check failed: {ex}"


def case_newton(name, nl, over, n_inst):
    import torch

    c = pe.Circuit(nl)
    c.set_analyze_type(pe.OP)
    b = c.batch(n_inst)
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    for e, a, v in over:
        b.set_param(e, a, v)

    def run():
        b.reset_state()
        b.analyze()

    ms = timed(run)
    solves = b.total_solves
    x, it = b.solution(), b.newton_iters()
    ri = b.resident_info(pe.MODE_DC)
    t0 = time.perf_counter()
    cpu = None
    try:
        import refapi

        if os.path.exists(refapi.REF_LIB_FAST):
            ns = min(n_inst, 20000)
            w = refapi.run_batch(nl, pe.OP, ns, [(e, a, v[:ns]) for e, a, v in over], fast=True)
            cpu = {"solves_per_s": float(w["solves"].sum()) / (time.perf_counter() - t0), "threads": w["threads"], "sample": ns}
    except Exception:  # noqa: BLE001
        pass
    print(json.dumps({"case": name, "instances": n_inst, "solves": int(solves), "ms": ms, "solves_per_s": solves / (ms * 1e-3), "mean_newton_iters": float(it.mean()),
                      "failed": int((b.status() != 0).sum()), "kernel": ri, "parity": ref_check(nl, pe.OP, over, x, it, 64), "cpu_reference": cpu}), flush=True)


def case_ac(n_sections, points):
    import torch

    nl, info = wl.rlc_ladder(n_sections)
    c = pe.Circuit(nl)
    c.set_analyze_type(pe.AC)
    b = c.batch(1)
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    b.set_ac_sweep(pe.SWEEP_LOG, 1e3, 1e10, points)
    ms = timed(lambda: b.analyze(), reps=2)
    ri = b.resident_info(pe.MODE_AC)
    st = b.stats(pe.MODE_AC)
    bytes_per_point = 16 * (st["nnz_a"] + 2 * st["nnz_a"] + 2 * st["n_unknowns"])
    parity = "unchecked"
    try:
        import refapi

        if os.path.exists(refapi.REF_LIB):
            c2 = pe.Circuit(nl)
            c2.set_analyze_type(pe.AC)
            b2 = c2.batch(1)
            b2.set_ac_sweep(pe.SWEEP_LOG, 1e3, 1e10, 97)
            b2.analyze()
            r = refapi.RefCircuit(nl)
            r.set_analyze_type(pe.AC)
            r.set_ac_sweep(pe.SWEEP_LOG, 1e3, 1e10, 97)
            r.analyze_counted()
            om, xr = r.ac_results()
            got = b2.ac_solution()[0]
            err = np.abs(got - xr)
            parity = "matches reference on a 97-point sweep" if bool((err <= 1e-12 + 1e-9 * np.maximum(np.abs(got), np.abs(xr))).all()) else "MISMATCH"
    except Exception as ex:  # noqa: BLE001
        parity = f"check failed: {ex}"
    print(json.dumps({"case": f"D rlc_ladder N={n_sections} AC", "points": points, "ms": ms, "points_per_s": points / (ms * 1e-3), "unknowns": st["n_unknowns"],
                      "algorithmic_bytes_per_point": bytes_per_point, "hbm_line_frac_of_6548GBs": points / (ms * 1e-3) * bytes_per_point / 6548.5e9,
                      "failed": int((b.status() != 0).sum()), "kernel": ri, "parity": parity}), flush=True)


def case_series_parallel(n_ring, n_merge, batches):
    """Config A: benchmark/series_parallel.cpp (seeded clone), DC operating point through the reduce-and-core path, batch =
    Monte-Carlo draws of every resistance.  Device time per phase from CUDA events inside the library; FP64 figure = the dense
    LU of the core (2/3 ld^3 flops) over its own device time."""
    import torch

    nl, info = wl.series_parallel(n_ring, n_merge, seed=1)
    ref_s, ref_x = None, None
    try:
        import refapi

        if os.path.exists(refapi.REF_LIB_FAST) and os.environ.get("PE_CFG_A_NO_REF") is None:
            r = refapi.RefCircuit(nl, fast=True)
            r.set_analyze_type(pe.DC)
            t0 = time.perf_counter()
            ok, _ = r.analyze_counted()
            ref_s = time.perf_counter() - t0
            ref_x = r.solution().real if ok else None
    except Exception:  # noqa: BLE001
        pass
    rng = np.random.default_rng(2)
    res, res_values = info["res"], info["res_values"]
    for batch in batches:
        c = pe.Circuit(nl)
        c.set_analyze_type(pe.DC)
        b = c.batch(batch)
        b.set_stream(torch.cuda.current_stream().cuda_stream)
        if batch > 1:
            # instance 0 keeps the netlist's values, the others are fresh draws of every resistance
            vals = rng.uniform(1e-5, 1e5, (len(res), batch))
            vals[:, 0] = res_values
            vals = np.ascontiguousarray(vals)
            b.set_params(b.param_table([(e, "r") for e in res]), vals.ctypes.data)
        b.analyze()  # symbolic phase + first solve
        ms = timed(lambda: b.analyze(), reps=2)
        fi = b.frontal_info()
        ld = fi["ld_core"]
        flops = 2.0 / 3.0 * ld ** 3 * batch
        parity = "unchecked"
        if ref_x is not None:
            x = b.solution()[0]
            err = np.abs(x - ref_x)
            scale = float(np.abs(ref_x).max())
            big = np.abs(ref_x) >= 1e-3 * scale
            rel_big = float((err[big] / np.abs(ref_x[big])).max())
            parity = ("matches reference: max |err| %.2e = %.1e of the largest |x| (bar 1e-9; kappa ~ 1e10), worst relative error on the %d unknowns above 1e-3 of it: %.1e (bar 1e-8)"
                      % (float(err.max()), float(err.max()) / scale, int(big.sum()), rel_big)) if float(err.max()) <= 1e-9 * scale and rel_big <= 1e-8 else "MISMATCH max abs err %.3e, worst relative %.3e" % (float(err.max()), rel_big)
        print(json.dumps({"case": f"A series_parallel ring {n_ring} merges {n_merge} DC", "batch": batch, "unknowns": fi["unknowns"], "ms_per_analyze": ms, "solves_per_s": batch / (ms * 1e-3),
                          "device_ms": {"reduce": fi["reduce_us"] / 1e3, "core_lu": fi["lu_us"] / 1e3, "substitutions": fi["subst_us"] / 1e3}, "plan": fi,
                          "core_lu_tflops": flops / (fi["lu_us"] * 1e-6) / 1e12 if fi["lu_us"] else None, "fp64_peak_note": "measured on this pool (tools/micro/fp64_rate.cu): DMMA.8x8x4 35-37 TFLOP/s, DFMA 32-34 TFLOP/s",
                          "failed": int((b.status() != 0).sum()), "parity": parity,
                          "cpu_reference": {"seconds": ref_s, "threads": 1, "build": "oracle/_ref/libpe_ref_fast.so (-O3 -march=x86-64-v3)"} if ref_s else None}), flush=True)


def main():
    if os.environ.get("PE_CFG_ONLY_A"):
        case_series_parallel(int(os.environ.get("PE_CFG_A_RING", "100000")), int(os.environ.get("PE_CFG_A_MERGES", "9000")), [int(v) for v in os.environ.get("PE_CFG_A_BATCH", "1,8").split(",")])
        return
    # PE_CFG_PATH = "streams,I,J,subtree_warps,workspace,tuning" forces a solve path (phy_engine_b200_set_default_path)
    if os.environ.get("PE_CFG_PATH"):
        abi = pe.bind_full_abi(pe.product())
        assert abi.lib.phy_engine_b200_set_default_path(*[int(v) for v in os.environ["PE_CFG_PATH"].split(",")]) == 0
    if os.environ.get("PE_CFG_ONLY_AC"):
        case_ac(64, int(os.environ.get("PE_CFG_POINTS", "1000000")))
        return
    n = int(os.environ.get("PE_CFG_INSTANCES", "100000"))
    rng = np.random.default_rng(1)
    nl, info = wl.diode_resistor()
    over = [(info["R"], "r", wl.sweep_values(rng, 1e3, n, 0.95, 1.05)), (info["D"][0], "Is", 1e-14 * np.exp(0.3 * rng.standard_normal(n))),
            (info["D"][0], "N", rng.uniform(1.0, 1.2, n))]
    case_newton("C1 V-R-diode OP Monte Carlo (test/0011 netlist)", nl, over, n)
    nl, info = wl.diode_ladder(16)
    over = [(e, "r", wl.sweep_values(rng, 1e3, n, 0.95, 1.05)) for e in info["R"]]
    case_newton("C4 16-stage diode ladder OP Monte Carlo", nl, over, n)
    case_ac(64, int(os.environ.get("PE_CFG_POINTS", "1000000")))


if __name__ == "__main__":
    main()
