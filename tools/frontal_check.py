import sys, time
sys.path.insert(0,'tests'); sys.path.insert(0,'phy-engine_b200')
import numpy as np, pe_b200 as pe, workloads as wl, emuapi, refapi
n_ring = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_merge = int(sys.argv[2]) if len(sys.argv) > 2 else 180
use = sys.argv[3] if len(sys.argv) > 3 else "emu"
abi = emuapi.emulator() if use == "emu" else pe.product()
abi.lib.phy_engine_b200_set_frontal_min.argtypes = [__import__('ctypes').c_size_t]
abi.lib.phy_engine_b200_set_frontal_min(100)
nl, info = wl.series_parallel(n_ring, n_merge, seed=3)
t0 = time.time()
rc = refapi.RefCircuit(nl); rc.set_analyze_type(pe.DC); ok, n = rc.analyze_counted()
print("reference ok", ok, "solves", n, "%.2f s" % (time.time() - t0))
xr = rc.solution().real
c = pe.Circuit(nl, abi); c.set_analyze_type(pe.DC)
b = c.batch(1)
t0 = time.time(); ok = b.analyze(); print("product ok", ok, "%.2f s" % (time.time()-t0), abi.last_error() if not ok else "", "kernel", b.last_kernel(), b.frontal_info())
x = b.solution()[0]
err = np.abs(x - xr); tol = 1e-12 + 1e-9 * np.maximum(np.abs(x), np.abs(xr))
print("n", len(x), "max abs err %.3e" % err.max(), "max rel err %.3e" % (err / np.maximum(np.abs(xr), 1e-300)).max(), "violations", int((err > tol).sum()), "max|x| %.3g" % np.abs(xr).max())
