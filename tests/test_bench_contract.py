"""bench.py contract (no GPU needed): the reference arm times the CPU restatement of the reference formulation on a bounded
sample and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-graphs", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "edges/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("cfg4") and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--ref-graphs", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_committed_product_lines_carry_the_contract_keys():
    """The bench lines committed under profiles/ (written by bench.py on a B200) carry every key of the measurement contract and
    are consistent with themselves: value = edges / step time, roofline.frac = achieved / peak, the dominant kernel's time is
    below the step time, and e2e declares its copies."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_cfg4*final4*.json")))
    assert paths, "no committed final bench line"
    for p in paths:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity"):
            assert k in d, (p, k)
        assert d["unit"] == "edges/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
        assert d["config"]["workload"].startswith("cfg4") and "model" not in d["config"]
        edges = 4096 * 512 * d["n_gpus"]
        assert abs(d["value"] - edges / (d["ms_per_step"] * 1e-3)) <= 1e-3 * d["value"]
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and 0 < r["frac"] < 1.05
        assert r["avg_launch_ms"] * r["launches_per_step"] < d["ms_per_step"]
        assert r["traffic"] is None or r["traffic"] >= 0.9 * r["alg_bytes_per_launch"]
        e = d["e2e"]
        assert e["unit"] == "edges/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
        assert d["gpu_launches"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert max(d["parity"]["max_norm_rel_err"].values()) <= d["parity"]["tolerance"]
