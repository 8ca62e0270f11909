import sys
sys.path.insert(0, "/root/repo")
from gxbeam_b200 import api, workloads
for nb in (1, 512):
    w = workloads.cantilever_tipmoment_batch(nb, 100)
    ctx = api.Context(0)
    ctx.topology(w["start"], w["stop"], w["pd"], w["pl"])
    ctx.set_batch(nb)
    ctx.set_elements(w["elems"]); ctx.set_bc(w["bc"]); ctx.set_body(None, w["force_scaling"])
    print("== static nb", nb, flush=True)
    ctx.static_solve(api.make_opts(iterations=1, lu_parts=2 if False else 0))
    ctx.close()
