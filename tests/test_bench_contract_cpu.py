"""bench.py's contract, checked on the arm that runs without a GPU (`--impl reference`: the CPU oracle timed on the host cores):
exactly one line on stdout, valid JSON, the keys and meanings the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                          "--cpu-sample", "128"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "columns/sec" and d["unit"] == "columns/s"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "MLP_v1" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "128 columns" in cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_under_a_non_zero_rank_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_profiles_named_in_the_readme_exist():
    import re
    text = open(os.path.join(ROOT, "profiles", "README.md")).read()
    names = set(re.findall(r"`(r\d\d_[A-Za-z0-9_]+\.(?:csv|json|txt))`", text))
    assert len(names) >= 8
    for n in names:
        assert os.path.exists(os.path.join(ROOT, "profiles", n)), n
    traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
    for kind in ("gemm_tn_fwd", "gemm_tn_dgrad", "gemm_nt_wgrad"):
        assert traffic[kind]["dram_bytes_per_launch"] > 1e7
