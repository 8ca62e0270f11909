import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from gxbeam_b200 import workloads
from oracle import pyoracle as O
nb = 1024
w = workloads.wind_turbine_blade_batch(nb, 200)
pk = O.Packed(w["start"], w["stop"], w["elems"], w["points_b"], w["pd"], w["pl"], w["bc"])
dyns = [O.make_dyn(vb=w["motion"][b, 0:3], wb=w["motion"][b, 3:6]) for b in range(nb)]
t0 = time.time()
ref = O.steady_solve_batch(pk, nb, dyns, force_scaling=w["force_scaling"], iterations=50)
print("time", time.time() - t0, "converged", ref["converged"].sum(), "iters hist", np.bincount(ref["iters"]))
bad = np.nonzero(~ref["converged"])[0]
print("bad", bad, "speeds idx", bad % 16, "designs", bad // 16)
print("resid of bad", ref["resid_inf"][bad] if "resid_inf" in ref else None)
print(ref.keys())
conv = ref["converged"].astype(bool)
bad = np.nonzero(~conv)[0]
print("bad", bad, "speed idx", bad % 16, "design", bad // 16)
print("resid", ref["resid_inf"][bad])
ref2 = O.steady_solve_batch(pk, nb, dyns, force_scaling=w["force_scaling"], iterations=1000)
conv2 = ref2["converged"].astype(bool)
print("1000 its: converged", conv2.sum(), "iters of formerly bad", ref2["iters"][bad], ref2["resid_inf"][bad])
