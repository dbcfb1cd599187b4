"""bench.py's reference arm runs on the CPU: check here that it prints exactly ONE JSON line on stdout with the keys the
driver reads (the CUDA arm prints the same keys plus roofline / clocks; it is exercised on the GPU box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                        "0", "--crop", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "SEAN-generator 256x256 images/s" and d["unit"] == "images/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    # "reference": the unmodified reference tree staged under baseline/_ref (baseline/make_ref.py); "port": the oracle
    from baseline import ref_runner
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_runner.ref_root() else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                        "1", "--warmup", "0", "--crop", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
