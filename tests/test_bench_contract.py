"""bench.py's contract with the driver, checked without a GPU: both arms name the workload identically (the driver
divides one line by the other only if `metric`, `unit` and `config.workload` agree), the reference arm runs on every
host core whatever OMP_NUM_THREADS says, and ranks other than 0 do no work there."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_both_arms_use_the_same_names():
    import bench
    for precision in ("f64", "f32"):
        for workload in ("direct", "bh"):
            assert bench.metric_name(workload, precision) == bench.metric_name(workload, precision)
            assert "N=" in bench.metric_name(workload, precision)
    assert bench.workload_name("direct", bench.N_DIRECT, 10.0).startswith("direct all-pairs fcompute N=1048576")
    assert bench.workload_name("bh", bench.N_BH, 10.0).startswith("Barnes-Hut heap_stackless fcompute N=4194304 ratio 10")
    src = open(os.path.join(ROOT, "bench.py")).read()
    # the two arms build their lines from the same two functions, never from literals of their own
    assert src.count('"metric": metric_name(') + src.count('"metric": blk["metric"]') >= 3
    assert src.count("workload_name(") >= 4


def test_reference_arm_line_and_thread_count(ref64, monkeypatch):
    import bench
    monkeypatch.setattr(bench, "N_CPU_STEP", 4096)          # a sample that takes milliseconds
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0")
    code = ("import bench, sys; bench.N_CPU_STEP = 4096; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'direct', "
            "'--steps', '2', '--warmup', '1']; sys.exit(bench.main())")
    res = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.metric_name("direct", "f64")
    assert line["config"]["workload"] == bench.workload_name("direct", bench.N_DIRECT, 10.0)
    assert line["unit"] == "pair interactions/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["cores"] == bench.host_cores() and line["cpu_baseline"]["kind"] == "reference"
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # any other rank prints nothing and exits 0
    res = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(env, RANK="3"), capture_output=True, text=True, timeout=60)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_overlay_patches_apply_to_the_reference(tmp_path):
    import shutil
    ref = os.environ.get("NB200_REFERENCE", "/root/reference/nbody")
    if not os.path.isdir(ref) or shutil.which("patch") is None:
        pytest.skip("needs the reference checkout and patch(1)")
    (tmp_path / "nbody").mkdir()
    for name in ("nbody_data.cpp", "nbody_engines.cpp", "nbody_engines.h"):
        shutil.copy(os.path.join(ref, name), tmp_path / "nbody" / name)
    for patch in ("nbody_data.patch", "nbody_engines.patch"):
        with open(os.path.join(ROOT, "integration", patch)) as f:
            res = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch"], stdin=f, cwd=tmp_path, capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
    text = (tmp_path / "nbody" / "nbody_data.cpp").read_text()
    assert "b200->statistics(b200->get_y()" in text and text.count("summed_on_device") == 6
