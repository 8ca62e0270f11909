import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from gxbeam_b200 import api, workloads
nb, nt = 512, 401
rng = np.random.default_rng(1238)
w = workloads.rotating_blade_batch(nb)
tvec = np.linspace(0.0, 0.5 * (nt - 1) / 2000, nt)
amp = 100 * rng.uniform(0.5, 1.5, nb)
tf1 = np.where(tvec < 0.2, 0.1 * (1 - np.sin(2 * np.pi * (tvec / 0.4 + 0.25))), 0.0)
series = (amp[:, None] * tf1[None, :])[:, :, None]
ctx = api.Context(0)
ctx.topology(w["start"], w["stop"], w["pd"], w["pl"], kind=1)
ctx.set_batch(nb)
ctx.set_elements(w["elems"]); ctx.set_points(w["points_b"]); ctx.set_bc(w["bc"]); ctx.set_body(None, w["force_scaling"])
ctx.set_motion(w["motion"])
for prof in (False, True):
    opts = api.make_opts(structural_damping=True, persistent=2, profile=prof)
    ctx.set_bc(w["bc"])
    ini = ctx.initial_condition(opts=opts)
    ctx.set_bc_series([21], [7], series)
    res = ctx.newmark_run(tvec, opts=opts, want_history=False)
    st = ctx.stats()
    print(prof, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()}, flush=True)
    if prof:
        print("avg us: factor %.1f assemble %.1f update %.1f" % (1e3 * st["ms_factor"] / st["n_factor"], 1e3 * st["ms_assemble"] / st["n_assemble"], 1e3 * st["ms_update"] / st["n_update"]))
ctx.close()
