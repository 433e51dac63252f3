"""The reference arm of bench.py (`--impl reference`) runs on CPU, so its side of the measurement contract can be
checked here: one JSON line on stdout with the agreed keys, rank 0 only under torchrun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# bench.py run in-process by a wrapper that afterwards reports (on stderr) whether libhesic_b200.so got mapped
WRAP = ("import runpy, sys; sys.argv = ['bench.py'] + sys.argv[1:]\n"
        "try:\n    runpy.run_path('bench.py', run_name='__main__')\n"
        "finally:\n    sys.stderr.write('NATIVE_SO_MAPPED=%d\\n' % ('libhesic_b200' in open('/proc/self/maps').read()))\n")


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, "-c", WRAP, "--impl", "reference", "--gpus", "2", "--steps", "1",
                           "--warmup", "1", "--batch", "2"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_contract_line_on_rank_0():
    r = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("stereo pairs/sec @512x512")
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["n_gpus"] == 2 and d["steps"] == 1 and d["warmup"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == "pairs/s" and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] * 1e-3 - 2.0) < 1e-6     # --batch 2: 2 pairs per step
    assert d["config"]["pairs_per_gpu"] == 2 and d["gpu_launches"] == 0
    # the CPU arm builds its weights from the reference's key table: no model of this repository, no CUDA library
    assert "NATIVE_SO_MAPPED=0" in r.stderr, r.stderr[-500:]


def test_both_arms_describe_the_same_workload():
    sys.path.insert(0, ROOT)
    import bench
    for model, B in (("hesic", 16), ("hesic_plus", 16), ("dsic", 8)):
        c = bench.make_config(model, B, 1)
        assert "workload" in c and "model" not in c and c["pairs_per_gpu"] == B and c["gflop_per_pair"] == bench.GFLOP_PER_PAIR[model]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
