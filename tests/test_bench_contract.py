"""The `bench.py --impl reference` arm runs without a GPU (the reference's CPU algorithm restated by the
oracle, timed on the host cores): check here that it honours the JSON contract of the task."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--n", "2048", "--tile", "256", "--cpu-budget", "1.0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("effective FP64 TFLOP/s") and d["dtype"] == "f64" and d["steps"] == 1 and d["warmup"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_every_config():
    """`bench.py --impl reference --config Cx`: the CPU arm exists for every BASELINE config (block-sparse pattern of
    config 3 incl. the masked variant, the T,T op flags of config 4 as the reference plans it, config 5's fused 1024^3
    tile GEMMs) and reports the sample it timed."""
    for cfg, word in (("C3", "block-sparse"), ("C3m", "block-sparse"), ("C5", "permuted 4-index"), ("C1", "dense")):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", cfg, "--steps", "1",
                              "--warmup", "0", "--cpu-budget", "0.5"], capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert out.returncode == 0, (cfg, out.stderr[-2000:])
        d = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][0])
        assert d["impl"] == "reference" and d["value"] > 0 and word in d["config"]["workload"] and cfg in d["config"]["workload"]
        assert "tile pairs" in d["cpu_baseline"]["sample"]
