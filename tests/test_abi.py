"""CPU-side checks of the drop-in boundary: the C-ABI libraries build, load and export every symbol
include/nb200.h declares, and the product path fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nb200_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    from nbody_b200 import build
    return {p: build.build_kernels(p) for p in ("f64", "f32")}


def test_header_declares_the_engine_surface():
    names = declared_symbols()
    for need in ("nb200_create", "nb200_alloc", "nb200_free", "nb200_read", "nb200_write", "nb200_copy", "nb200_fill",
                 "nb200_fcompute_direct", "nb200_fcompute_bh", "nb200_fmadd_inplace", "nb200_fmadd", "nb200_fmaddn",
                 "nb200_fmaddn_inplace", "nb200_fmaddn_corr", "nb200_fmaxabs", "nb200_clamp"):
        assert need in names


@pytest.mark.parametrize("precision,size", [("f64", 8), ("f32", 4)])
def test_library_exports_every_declared_symbol(built, precision, size):
    lib = ctypes.CDLL(built[precision])
    for name in declared_symbols():
        assert hasattr(lib, name), "%s missing from %s" % (name, built[precision])
    assert lib.nb200_real_size() == size


def test_library_is_sm100a_only(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", built["f64"]], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu(built):
    """Without a CUDA device the engine refuses to construct; nothing routes to the CPU oracle."""
    from nbody_b200 import Engine, device_count
    if device_count("f64") > 0:
        pytest.skip("a GPU is present")
    with pytest.raises((RuntimeError, ValueError)):
        Engine(devices=[0])


def test_product_does_not_import_the_oracle():
    """Only tests/, bench.py's cpu_baseline legs and __graft_entry__.smoke() may touch oracle/: the product
    package must not import, link or execute the CPU restatement or the compiled reference."""
    pkg = os.path.join(ROOT, "nbody_b200")
    banned = ("import oracle", "from oracle", "liboracle", "nbody_oracle", "refharness", "orc_", "nbref_")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                for word in banned:
                    if f == "build.py" and word == "nbref_":
                        continue    # the C++ adapter LINKS the shim-built reference for nbody_engine's base class
                    assert word not in text, "%s mentions %r" % (f, word)


def test_parse_devices_mirrors_select_devices():
    """select_devices (nbody_engine_cuda.cpp:576-616): "", "a", "0,a", "-1", ">= count" are rejected."""
    from nbody_b200 import parse_devices
    assert parse_devices("0", 1) == [0]
    assert parse_devices("0,0,0,0", 1) == [0, 0, 0, 0]
    assert parse_devices("0,1", 2) == [0, 1]
    for bad in ("", "a", "0,a", "-1", "9999", ","):
        assert parse_devices(bad, 8) is None
