"""bench.py on a box without a GPU: the reference arm (the unmodified reference on the host cores) prints the contract's
JSON line, both arms describe the workload with the same `config` object, and the product arm refuses to run without a
CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=env,
                          cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    if not (os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libopal_ref.so")) or os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so"))):
        pytest.skip("no CPU library built (run __graft_entry__.build() first)")
    r = _run("--impl", "reference", "--workload", "config2", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "GCUPS" and line["unit"] == "GCUPS"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same `config` object as the product arm's (same function, same workload): the driver compares them
    import bench
    w = bench.make_workload("config2", 0, 1)
    assert line["config"] == bench.config_of(w)
    assert set(line["config"]) == {"workload", "modes", "search", "query_lengths", "matrix", "gap_open", "gap_ext", "db_sequences"}


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "config2"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    r = _run("--workload", "config2", "--steps", "1", "--warmup", "1", timeout=300)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) or "no CPU path" in (r.stderr + r.stdout)
