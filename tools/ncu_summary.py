#!/usr/bin/env python
"""Summary of an ncu report (no GPU needed): usage ncu_summary.py report.ncu-rep [solves_in_capture]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
solves = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    def g(k):
        return d.get(k, ("", ""))
    print("kernel:", g("Kernel Name")[0], "grid", g("Grid Size")[0], "block", g("Block Size")[0])
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
            "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_st.sum"]
    for k in keys:
        if k in d:
            print(f"  {k}: {d[k][0]} {d[k][1]}")
    st = [(float(d[h][0]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h][0]]
    for v, h in sorted(st, reverse=True)[:7]:
        print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}: {v:.2f}")
    if solves:
        def gb(k):
            v, u = d[k]
            return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
        r, w = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
        t = float(d["gpu__time_duration.sum"][0]) * {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}[d["gpu__time_duration.sum"][1]]
        print(f"  per solve: DRAM read {r / solves / 1e3:.1f} KB + write {w / solves / 1e3:.1f} KB = {(r + w) / solves / 1e3:.1f} KB; DRAM rate {(r + w) / t / 1e12:.2f} TB/s; {solves / t / 1e6:.2f} M solves/s in the capture")
