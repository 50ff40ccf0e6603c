"""CPU-only check of the driver-facing contract of bench.py's reference arm (the oracle timed on host cores)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-budget-s", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("EKF steps/sec at N=1024")
    # the reference's own translation unit where oracle/_ref is built (here), the C port elsewhere
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    if os.path.isdir("/root/reference"):
        assert d["cpu_baseline"]["kind"] == "reference"
    assert abs(d["cpu_baseline"]["per_core_steps_s"] * d["cpu_baseline"]["cores"] - d["value"]) < 1e-9
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
    # the line reports the steps it actually ran, and the timed region it claims fits inside the run
    assert d["steps"] == 1 and d["warmup"] >= 1
    assert d["steps"] * d["ms_per_step"] * 1e-3 <= d["wall_s"]
    assert abs(d["value"] - d["cpu_baseline"]["cores"] * d["steps"] / (d["steps"] * d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]


def test_both_arms_describe_the_same_workload():
    """The reference arm runs on the B200 arm's `config` (the driver compares the two dicts)."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        a = bench.parse_args()
    finally:
        sys.argv = argv
    c = bench.config_dict(a, 1)
    assert c["sessions_per_gpu"] == 8 and c["n"] == 2051 and c["r"] == 200 and "C3" in c["workload"]
    assert bench.metric_name(a) == "EKF steps/sec at N=1024 landmarks (100 observed/step)"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, env=env, cwd=ROOT)
    assert res.returncode == 0 and res.stdout.strip() == ""
