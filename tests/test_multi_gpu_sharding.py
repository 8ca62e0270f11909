"""CPU test (gloo, world_size 2) of the multi-GPU host logic used by bench.py: instances are sharded by rank, there is no
data-path collective, the job-level numbers are a SUM of units and a MAX of times over ranks."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from gxbeam_b200 import workloads
    from oracle import pyoracle as O
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # the same sharding rule as bench.py: every rank owns its own seeded shard of independent instances
    w = workloads.cantilever_tipmoment_batch(6, 8, seed=1236 + 1000 * rank)
    pk = O.Packed(w["start"], w["stop"], w["elems"], w["points"], w["pd"], w["pl"], w["bc"])
    r = O.static_solve_batch(pk, 6, force_scaling=w["force_scaling"], nthreads=1)
    units = torch.tensor([float(r["iters"].sum())], dtype=torch.float64)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)           # pretend per-rank time
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # strong scaling (bench.py's default): every rank generates the SAME seeded batch and keeps its contiguous shard; the shards' Newton
    # iteration counts, gathered, are the full batch's
    nb = 5
    wfull = workloads.cantilever_tipmoment_batch(nb * world, 8, seed=1236)
    sl = slice(rank * nb, (rank + 1) * nb)
    pks = O.Packed(wfull["start"], wfull["stop"], wfull["elems"][sl], wfull["points"], wfull["pd"], wfull["pl"], wfull["bc"][sl])
    rs = O.static_solve_batch(pks, nb, force_scaling=wfull["force_scaling"][sl], nthreads=1)
    mine = torch.tensor(rs["iters"].astype(np.float64))
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    strong_ok = None
    if rank == 0:
        pkf = O.Packed(wfull["start"], wfull["stop"], wfull["elems"], wfull["points"], wfull["pd"], wfull["pl"], wfull["bc"])
        rf = O.static_solve_batch(pkf, nb * world, force_scaling=wfull["force_scaling"], nthreads=1)
        strong_ok = bool(np.array_equal(torch.cat(parts).numpy(), rf["iters"].astype(np.float64)))
    chk = torch.tensor([float(np.abs(w["bc"][:, -1, 11]).sum())], dtype=torch.float64)
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    if rank == 0:
        print(json.dumps({"units": units.item(), "tmax": t.item(), "own_iters": int(r["iters"].sum()),
                          "shards_differ": bool(gathered[0].item() != gathered[1].item()), "strong_ok": strong_ok}))
    dist.destroy_process_group()
""") % ROOT


def test_sharded_job_aggregates_with_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["tmax"] == 2.0                      # max over ranks
    assert d["units"] > d["own_iters"]           # sum over ranks
    assert d["shards_differ"]                    # weak mode: each rank solved different instances
    assert d["strong_ok"] is True                # strong mode: the gathered shards are the full seeded batch


def test_api_shard_bounds():
    # gxbeam_b200.api.static_analysis_batch partitions [0, nbatch) contiguously over the devices
    for nbatch, ndev in ((4096, 8), (10, 3), (5, 8)):
        bounds = np.linspace(0, nbatch, ndev + 1).astype(int)
        sizes = np.diff(bounds)
        assert sizes.sum() == nbatch and sizes.max() - sizes.min() <= 1
