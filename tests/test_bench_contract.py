"""The bench JSON lines committed under profiles/ (produced by `python bench.py` and `python bench.py --impl reference` on a
B200, tools/round_profiles.sh) carry every key of the measurement contract; bench.py's CPU arm runs here on a tiny sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


def test_committed_bench_line_has_the_contract_keys():
    d = _line("r1_bench_1gpu.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "flow_residual_constraints_per_sec_per_gn_iteration" and d["unit"] == "constraints/s" and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "traffic" in r
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] > 0
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"]) and not [x for x in d["clocks"]["reasons"] if "thermal" in x or "hw_slowdown" in x]
    assert abs(d["value"] - d["config"]["constraints"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]


def test_committed_reference_arm_line():
    d = _line("r1_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == _line("r1_bench_1gpu.json")["metric"] and d["unit"] == "constraints/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_the_host_cores(tmp_path):
    """`bench.py --impl reference` needs no GPU: smallest workload, one step."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "config1_8f_128x96_grid4x4_sep10", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["cores"] >= 1 and d["config"]["workload"].startswith("config1")
