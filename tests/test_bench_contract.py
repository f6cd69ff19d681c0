"""CPU-only: the reference arm of bench.py runs without a GPU (it times the CPU oracle) and prints
one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--recordings", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "rips_h0h1_diagrams_per_sec"
    assert line["unit"] == "diagrams/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["n_gpus"] == 1 and line["steps"] == 1


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1", "--recordings", "2"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_bench_lines_carry_the_contract_keys():
    """profiles/r01_bench.json (and the 2-GPU line) are what bench.py printed on the B200: one JSON
    object with every key the contract names, internally consistent."""
    for name, gpus in (("r01_bench.json", 1), ("r01_scale_n2_24warps.json", 2)):
        text = open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()
        line = json.loads(text[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
            assert k in line, (name, k)
        assert line["n_gpus"] == gpus and line["scaling"] == "weak" and line["vs_baseline"] is None
        assert line["warmup"] >= 3 and line["gpu_launches"] > 0 and "workload" in line["config"]
        per_gpu = line["config"]["diagrams_per_gpu"]
        assert abs(line["value"] - gpus * per_gpu / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
        rf = line["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in rf, (name, k)
        assert rf["bound"] == "hbm" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
        e2e = line["e2e"]
        assert e2e["unit"] == line["unit"] and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
        assert e2e["value"] < line["value"] and e2e["matches_device_path"] is True
        assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if gpus == 1:
            cb = line["cpu_baseline"]
            assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == line["unit"] and "sample" in cb
