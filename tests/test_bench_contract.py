"""bench.py's JSON contract (the driver parses these lines): the reference arm is run here on a tiny
sample (CPU only), the B200 arm is checked on the committed line of the last measurement pass."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def _check_common(d):
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and d["dtype"] in ("f64", "f32")
    assert isinstance(d["config"].get("workload"), str) and "model" not in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["unit"] == d["unit"]
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference") and c["cores"] >= 1


def test_reference_arm_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--envs-per-gpu", "2048"], capture_output=True, text=True, env=env,
                         timeout=300, check=True).stdout.strip().splitlines()
    assert len(out) == 1, out                      # exactly ONE JSON line
    d = json.loads(out[0])
    _check_common(d)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_b200_line_has_the_contract_keys():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01*_bench_1gpu.json")))
    assert files, "no committed bench line under profiles/"
    d = json.load(open(files[-1]))
    _check_common(d)
    assert d["n_gpus"] == 1 and d["gpu_launches"] == d["steps"] > 0 and d["warmup"] >= 3
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["traffic"] is None or isinstance(r["traffic"], (int, float))
    assert d["config"]["envs_per_gpu"] == 65536 and d["config"]["substeps"] == 16       # BASELINE configs[1]
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"]          # the host-buffer path cannot repeat the device-timed number
