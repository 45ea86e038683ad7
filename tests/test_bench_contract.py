"""CPU: the reference arm of bench.py (the oracle's C engine on the host cores) prints ONE JSON line with the keys the driver
reads; the GPU arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    # a wrapper that records which shared libraries the arm maps: it must never load the product library
    code = ("import runpy, sys, os; sys.argv = ['bench.py'] + sys.argv[1:]\n"
            "try:\n    runpy.run_path(os.path.join(%r, 'bench.py'), run_name='__main__')\n"
            "except SystemExit as e:\n    rc = e.code or 0\n"
            "maps = open('/proc/self/maps').read()\n"
            "sys.stderr.write('MAPPED_PRODUCT=%%d MAPPED_ORACLE=%%d\\n' %% ('libgripper_sim_b200' in maps, 'liboracle' in maps))\nsys.exit(rc)\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code, "--impl", "reference", "--steps", "3", "--warmup", "1", "--preroll", "4", "--cpu-seconds", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "MAPPED_PRODUCT=0 MAPPED_ORACLE=1" in out.stderr, out.stderr[-2000:]
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["metric"].startswith("env-steps/sec") and "substeps" in d["unit"] and d["value"] > 0 and d["n_gpus"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle/engine.c" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("acorn_env")
    # --steps / --warmup are honoured (transitions per environment), and ms_per_step is per transition of the sample: the
    # timed region the driver derives (steps x ms_per_step) is the time actually spent in the timed windows
    assert d["steps"] == 3 and d["warmup"] == 1 and d["preroll"] == 4
    assert "3 timed transitions (after 5 untimed from reset)" in cb["sample"]
    assert 0 < d["steps"] * d["ms_per_step"] / 1e3 < 60
    assert "oracle/mjcf.py" in cb["sample"]


def test_action_tapes():
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    u = np.random.default_rng(0).uniform(-1, 1, (1000, 6))
    a = bench.shape_actions(u.copy(), "contact")  # biased towards +x / close
    assert a[:, 0].min() >= 0.2 and a[:, 0].max() <= 1.0 and a[:, 5].max() <= 0.0 and a[:, 5].min() >= -1.0 and np.abs(a[:, 1:5]).max() <= 0.3
    np.testing.assert_array_equal(bench.shape_actions(u.copy(), "uniform"), u)


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
