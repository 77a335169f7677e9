"""bench.py's contract, checked on the arm that runs without a GPU (`--impl reference`: the CPU oracle timed on the host cores):
exactly one line on stdout, valid JSON, the keys and meanings the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                          "--cpu-sample", "128", "--extras", "none", "--no-legs"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "columns/sec" and d["unit"] == "columns/s"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "MLP_v1" in d["config"]["workload"] and "model" not in d["config"]
    # the arm says what it RAN: 128 columns per step on the CPU, not the GPU arm's 65 536
    assert d["config"]["columns_per_step"] == d["config"]["columns_per_gpu_per_step"] == d["config"]["global_batch"] == 128
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "128 columns" in cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_covers_every_baseline_configuration_and_the_baseline_md_legs():
    """Default reference run: the headline line plus `workloads` (cnn / hsr / ed CPU samples) and BASELINE.md section 3's legs
    (B = 1024 and 3072, forward-only and training, cores stated)."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                          "--cpu-sample", "256"], capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads(res.stdout.strip())
    assert set(d["workloads"]) == {"cnn", "hsr", "ed"}
    for name, w in d["workloads"].items():
        assert w["value"] > 0 and w["cpu_baseline"]["cores"] >= 1 and w["config"]["columns_per_step"] > 0, name
    assert "Dropout 0.175" in d["workloads"]["cnn"]["config"]["workload"] and "TWO" in d["workloads"]["hsr"]["config"]["workload"]
    legs = d["baseline_md_legs"]
    assert set(legs) == {"B1024_train", "B1024_fwd", "B3072_train", "B3072_fwd"}
    assert all(v["columns_per_s"] > 0 and v["cores"] >= 1 for v in legs.values())
    assert legs["B1024_fwd"]["columns_per_s"] > legs["B1024_train"]["columns_per_s"]


def test_flop_models_match_the_survey():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.FLOP_TRAIN == 10_309_632
    assert 2 * sum(k * n for k, n in b.hsr_dims()) * 2 == 13_615_104                 # HSR (both nets) fwd 13.6 MFLOP/col
    assert abs(2 * sum(k * n for k, n in b.ed_dims()) - 1.66e6) < 0.01e6             # ED fwd 1.66 MFLOP/col
    assert abs(b.cnn_train_flops() - 4.75e9) < 0.01e9                                # CNN train 4.75 GFLOP/col
    # burst vs sustained is decided by the run's own clock record
    peaks = {"tf_sustained": 1396.8, "tf_burst": 1667.7, "src": "measured"}
    assert b.choose_peak(peaks, {"reasons": [], "sm_mhz": 1965.0, "sm_max_mhz": 1965.0}, 0.016)[0] == 1667.7
    assert b.choose_peak(peaks, {"reasons": ["sw_power_cap"], "sm_mhz": 1500.0, "sm_max_mhz": 1965.0}, 0.5)[0] == 1396.8
    assert b.choose_peak(peaks, {"reasons": [], "sm_mhz": 1400.0, "sm_max_mhz": 1965.0}, 0.5)[0] == 1396.8


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
