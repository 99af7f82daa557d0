"""bench.py's reference arm runs on the CPU (it times the oracle port): check its JSON line against the driver's contract here; the GPU arm's line is checked on the GPU box."""
import json, os, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "3", "--cpu-steps", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/s" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 4 and d["warmup"] == 3 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "20k" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_gpu_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "40", "--warmup", "3", "--settle", "200", "--e2e-steps", "5", "--cpu-steps", "5", "--n", "200000",
                        "--cfg3-n", "200000", "--cfg4-n", "100000", "--cfg5-n", "100000", "--cfg-settle", "300", "--cfg1-cpu-steps", "100"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e",
              "gpu_launches", "clocks"):
        assert k in d, k
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1.2 and d["roofline"]["peak"] > 1000
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and d["cpu_baseline"]["kind"] == "port"
    assert "200000" in d["config"]["workload"] and d["config"]["n_spheres_total"] == 200000          # the label follows --n
    assert set(d["configs"]) == {"cfg1_deposition_10k", "cfg3_hertz_4M", "cfg4_periodic_shear", "cfg5_drum_2M"}
    for c in d["configs"].values():
        assert c["value"] > 0 and c["n_spheres"] > 0 and "workload" in c
    assert d["configs"]["cfg1_deposition_10k"]["same_window"]["cpu_port_value"] > 0
