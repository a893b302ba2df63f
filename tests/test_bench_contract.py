"""bench.py's reference arm runs without a GPU (it times the C oracle port on the host cores): its JSON line carries the
keys the driver reads.  The GPU arm's line is checked on the B200 (-m gpu)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-envs", "32768", "--ref-python-seconds", "2")
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "env-steps/s" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    rp = cb["reference_python"]   # the unmodified Python reference under the SubprocVecEnv clone (staged by build() in this container)
    if "unavailable" not in rp:
        assert rp["kind"] == "reference" and rp["value"] > 0 and rp["cores"] >= 1 and rp["serial_one_process"] > 0


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run("--steps", "40", "--warmup", "5", "--e2e-steps", "3", "--rollout-steps", "0", "--post-steps", "5", "--fused", "0",
             "--cpu-seconds", "2", "--ref-python-seconds", "2")
    assert BASE_KEYS <= set(d) and d.get("impl") != "reference"
    assert d["n_gpus"] == 1 and d["gpu_launches"] >= 40 and d["dtype"] == "f32" and d["scaling"] == "weak"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.5 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == d["config"]["envs_per_gpu"] * 24 and e["d2h_bytes_per_step"] > d["config"]["envs_per_gpu"] * 104
    assert e["value"] < d["value"]   # host buffers and PCIe inside the timed region
    assert d["cpu_baseline"]["kind"] == "port"
    # the timed region runs at the steady state: finished episodes and in-kernel auto-resets inside it
    assert d["episode_stats"]["episodes"] > 0 and d["done_episodes_per_step"] > 100
    assert r["layout_bytes_per_env_step"] == 550 and 0 < r["frac_layout"] < r["frac"]
    assert d["strong"]["total_envs"] == 1 << 20
    for k in ("cfg2@4096", "cfg3@262144"):
        c = d["configs"][k]
        assert c["value"] > 0 and c["us_per_tick_graph"] > 0 and c["us_per_tick_eager"] > 0
    assert e["obs17"]["obs_dim"] == 17 and e["obs17"]["d2h_bytes_per_step"] < e["d2h_bytes_per_step"]
    c = d["clocks"]
    assert "reasons" in c and (c["sm_max_mhz"] is None or (c["sm_max_mhz"] > 0 and c["sm_mhz"] > 0))   # None only without nvidia-smi
