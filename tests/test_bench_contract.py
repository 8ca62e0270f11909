"""bench.py's reference arm runs on the host cores, so its side of the driver contract can be checked without a GPU: exactly one JSON line on
stdout, the keys the driver reads, rank 0 only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, extra_args=()):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", *extra_args],
                          env=env, capture_output=True, text=True, check=True, cwd=ROOT).stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run({})
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "newton_iters/s" and d["dtype"] == "f64"
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0


def test_reference_arm_other_ranks_print_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""


def test_reference_arm_at_eight_ranks_keeps_the_workload_and_declares_strong_scaling():
    """Strong scaling: the reference arm solves the SAME 4096 instances whatever N is, with the workload string the GPU arm prints when its N
    shards add up to that batch (bench.py: `WORKLOAD if scaling == strong and nb * world == NBATCH`)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert '"workload": WORKLOAD if (args.scaling == "strong" and nb * world == NBATCH)' in src
    d = json.loads(_run({"RANK": "0", "WORLD_SIZE": "8", "LOCAL_RANK": "0"}, ("--gpus", "8")).strip())
    assert d["config"]["workload"] == bench.WORKLOAD and d["scaling"] == "strong" and d["n_gpus"] == 8


def test_dominant_kernel_follows_the_library_selection_rule():
    """bench.dominant_kernel mirrors launch_factor() of gxb_lib.cu (148 SMs): partitions for <= 444 instances, the two-sided kernel for
    445..888, one warp per instance on the tensor cores above; every capture it names is committed together with its meta record."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.dominant_kernel(64)[0] == "k_factor_solve_part<8>" and bench.dominant_kernel(256)[0] == "k_factor_solve_part<6>"
    assert bench.dominant_kernel(444)[0] == "k_factor_solve_part<4>"
    assert bench.dominant_kernel(445)[0] == bench.dominant_kernel(512)[0] == bench.dominant_kernel(888)[0] == "k_factor_solve_two<2,2,2,1>"
    assert bench.dominant_kernel(889)[0] == bench.dominant_kernel(4096)[0] == "k_factor_solve_mma<2,2>"
    assert bench.dominant_kernel(512, forced=3)[0] == "k_factor_solve_part<3>" and bench.dominant_kernel(4096, forced=2)[0].startswith("k_factor_solve_two")
    meta = json.load(open(os.path.join(ROOT, "profiles", "r2_prof_meta.json")))
    for nb in (512, 4096):
        kernel, prof = bench.dominant_kernel(nb)
        assert os.path.exists(os.path.join(ROOT, "profiles", prof)) and meta[prof]["kernel"] == kernel and meta[prof]["instances_in_launch"] == nb
        t, gbs, note = bench.capture_traffic(prof, kernel, nb)
        assert t and t > 0 and "unchanged" in note, note
