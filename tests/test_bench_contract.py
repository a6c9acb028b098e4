"""bench.py's CPU-runnable parts: the reference arm prints one JSON line with the contract's keys (it is what the driver
runs as `bench.py --impl reference`), and the bookkeeping helpers agree with SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "patches/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["config"]["oracle_engine"] in ("port", "kymatio")


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--config",
                        "cfg1", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_flops_and_bytes_model_match_the_survey_table():
    sys.path.insert(0, ROOT)
    import bench
    # flops_alg / patch and K per config
    assert bench.num_coefficients(4, 8, 2) == 417 and bench.num_coefficients(5, 8, 2) == 681
    assert abs(bench.flops_model(128, 4, 8, 3) / 410.6e6 - 1) < 2e-3
    assert abs(bench.flops_model(64, 3, 8, 3) / 81.5e6 - 1) < 5e-3
    assert abs(bench.flops_model(32, 2, 8, 3) / 13.6e6 - 1) < 5e-3
    assert abs(bench.flops_model(512, 5, 8, 4) / 9322e6 - 1) < 2e-3
