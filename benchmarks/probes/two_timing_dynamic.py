import sys
sys.path.insert(0, "/root/repo")
from gxbeam_b200 import api, workloads
for nb in (1, 512):
    w = workloads.rotating_blade_batch(nb)
    ctx = api.Context(0)
    ctx.topology(w["start"], w["stop"], w["pd"], w["pl"], kind=1)
    ctx.set_batch(nb)
    ctx.set_elements(w["elems"]); ctx.set_points(w["points_b"]); ctx.set_bc(w["bc"]); ctx.set_body(None, w["force_scaling"])
    ctx.set_motion(w["motion"])
    print("== dynamic nb", nb, flush=True)
    ctx.steady_solve_resident(api.make_opts(iterations=2, profile=True))
    print(ctx.stats(), flush=True)
    ctx.close()
