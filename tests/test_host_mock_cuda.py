"""Host logic of libnb200 on a machine without a GPU: the library runs against a stand-in CUDA runtime
(tests/mock_cuda/mock_cudart.c, preloaded into a subprocess) in which copies are real, kernels are only counted and a
captured graph replays its copies and counts its kernels. That pins down WHEN the library issues what -- the shard
layout behind write/read_buffer, the argument checks, and the step table of nb200_stepgraph.cuh: recorded, captured and
replayed launches per solver step, deferral until the boundary, fall-back on a read in the middle of a step, fmaxabs
as a segment border, periodic patterns (Bulirsch-Stoer-like sub-steps, tree_build_rate), table limits.
The numerical side of the same calls is the GPU suite's business (-m gpu); nothing here is loaded by the product."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "mock_cuda")


def cuda_include():
    for base in (os.environ.get("CUDA_HOME"), os.environ.get("CUDA_PATH"), "/usr/local/cuda"):
        if base and os.path.exists(os.path.join(base, "include", "cuda_runtime_api.h")):
            return os.path.join(base, "include")
    return None


@pytest.fixture(scope="module")
def mock_runtime(tmp_path_factory):
    from nbody_b200 import build
    inc, gcc = cuda_include(), shutil.which("gcc")
    if inc is None or gcc is None:
        pytest.skip("needs gcc and the CUDA runtime headers")
    if not os.path.exists(build.lib_path("f64")):
        pytest.skip("libnb200_f64.so not built")
    out = str(tmp_path_factory.mktemp("mock") / "mock_cudart.so")
    subprocess.run([gcc, "-shared", "-fPIC", "-O1", "-I" + inc, "-o", out, os.path.join(HERE, "mock_cudart.c")], check=True)
    return out


def test_host_logic_scenarios_under_the_mock_runtime(mock_runtime):
    env = dict(os.environ, LD_PRELOAD=mock_runtime, NBREF_QUIET="1")
    res = subprocess.run([sys.executable, os.path.join(HERE, "drive.py"), mock_runtime], env=env, capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    for name in ("buffers_and_shards", "fixed_step_replay", "deferral_is_real", "error_norm_border", "read_in_the_middle",
                 "periodic_patterns", "never_repeating_and_table_limit", "buffers_change_empties_the_table", "tree_build_rate",
                 "direct_path_selection", "bodies_and_statistics_are_host_visible"):
        assert "ok " + name in res.stdout, res.stdout
    assert "all host-logic scenarios passed" in res.stdout


def test_every_reference_solver_is_replayed_under_the_mock_runtime(mock_runtime):
    """The reference's own solver classes on the C++ adapter: euler ... rkfeagin14 (one table entry), Adams (its history
    buffers rotate: 10 entries), Bulirsch-Stoer (12 distinct sub-steps) all end up as graph launches only."""
    from nbody_b200 import build
    from oracle import refharness as R
    if not R.available("f64") or not os.path.exists(build.adapter_path("f64")):
        pytest.skip("needs oracle/_ref and the C++ adapter (built where /root/reference is mounted)")
    env = dict(os.environ, LD_PRELOAD=mock_runtime, NBREF_QUIET="1")
    res = subprocess.run([sys.executable, os.path.join(HERE, "drive_adapter.py"), mock_runtime], env=env, capture_output=True,
                         text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "all reference solvers replayed" in res.stdout and res.stdout.count("ok ") == 19, res.stdout


def test_without_the_mock_there_is_still_no_device(mock_runtime):
    """The stand-in only exists inside the subprocess above: a plain process on this machine gets no context."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from nbody_b200 import device_count\n"
            "print('devices', device_count('f64'))\n" % ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    if "devices 0" not in res.stdout:
        pytest.skip("this machine has a GPU")
