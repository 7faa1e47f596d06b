"""The drop-in registration, checked for real: integration/nbody_engines.patch is applied to a scratch copy of the
reference's nbody/nbody_engines.{h,cpp}, the patched factory is compiled (Qt shim, -DHAVE_B200) and linked against the
C++ adapter, and the reference's own nbody_create_engine then returns nbody_engine_b200 for --engine=b200 / b200_bh,
still returns its own engines for the other aliases, and rejects what the cuda aliases reject. Runs on a machine
without a GPU under the stand-in CUDA runtime of tests/mock_cuda. Needs the reference checkout (/root/reference)."""
import os
import shutil
import subprocess
import sys

import pytest

from test_host_mock_cuda import mock_runtime  # noqa: F401  (fixture)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("NB200_REFERENCE", "/root/reference/nbody")


def test_patched_reference_factory_creates_b200_engines(mock_runtime, tmp_path):  # noqa: F811
    from nbody_b200 import build
    from oracle import refharness as R
    if not os.path.isdir(REFERENCE) or shutil.which("patch") is None:
        pytest.skip("needs the reference checkout and patch(1)")
    if not R.available("f64") or not os.path.exists(build.adapter_path("f64")):
        pytest.skip("needs oracle/_ref and the C++ adapter")
    scratch = tmp_path / "nbody"
    scratch.mkdir()
    for name in ("nbody_engines.cpp", "nbody_engines.h"):
        shutil.copy(os.path.join(REFERENCE, name), scratch / name)
    with open(os.path.join(ROOT, "integration", "nbody_engines.patch")) as patch:
        subprocess.run(["patch", "-p1", "--no-backup-if-mismatch"], stdin=patch, cwd=tmp_path, check=True, capture_output=True)
    assert "nbody_create_engine_b200" in (scratch / "nbody_engines.cpp").read_text()
    flags = ["/usr/bin/g++", "-std=gnu++17", "-O1", "-fPIC", "-fopenmp", "-w", "-DNB_COORD_PRECISION=2", "-DNB200_PRECISION=2",
             "-DHAVE_B200", "-Dnbody_create_engine=nbody_create_engine_with_b200",    # libnbref already holds the unpatched one
             "-I" + str(scratch), "-I" + os.path.join(ROOT, "oracle", "qtshim"), "-I" + REFERENCE,
             "-I" + os.path.join(ROOT, "nbody_b200", "host"), "-I" + os.path.join(ROOT, "include")]
    out = str(tmp_path / "libfactory.so")
    subprocess.run(flags + ["-shared", "-o", out, str(scratch / "nbody_engines.cpp"),
                            os.path.join(ROOT, "tests", "mock_cuda", "factory_wrapper.cpp"),
                            "-L" + os.path.join(ROOT, "oracle", "_ref"), "-lnbref_f64",
                            "-L" + os.path.join(ROOT, "nbody_b200", "host"), "-lnbody_engine_b200_f64",
                            "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_ref"),
                            "-Wl,-rpath," + os.path.join(ROOT, "nbody_b200", "host")], check=True)
    env = dict(os.environ, LD_PRELOAD=mock_runtime, NBREF_QUIET="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mock_cuda", "drive_factory.py"), mock_runtime, out],
                         env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "patched factory ok" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
