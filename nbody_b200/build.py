"""Build recipe for the nb200 native libraries (sm_100a only, in-tree outputs).

    libnb200_f64.so / libnb200_f32.so   CUDA kernels + C ABI (include/nb200.h)
    libnbody_engine_b200_f{64,32}.so    C++ adapter class for the reference's
                                        nbody_engine API; needs the reference
                                        headers, so it is only (re)built where
                                        /root/reference is mounted

nvcc cross-compiles without a GPU. Outputs are git-ignored but travel to the
GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
REFERENCE = os.environ.get("NB200_REFERENCE", "/root/reference/nbody")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-cudart", "shared",
]
PRECISIONS = {"f64": 2, "f32": 1}


def lib_path(precision, variant=None):
    """The product library; `variant` (or the environment variable NB200_LIB_VARIANT) names an A/B build made with
    build_variant() -- measurement only, never shipped."""
    variant = variant if variant is not None else os.environ.get("NB200_LIB_VARIANT", "")
    return os.path.join(CSRC, "libnb200_%s%s.so" % (precision, "_" + variant if variant else ""))


def adapter_path(precision):
    return os.path.join(HOST, "libnbody_engine_b200_%s.so" % precision)


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(folder, exts):
    return sorted(os.path.join(folder, f) for f in os.listdir(folder) if f.endswith(exts))


def build_kernels(precision, force=False, verbose=False):
    out = lib_path(precision, "")
    deps = _sources(CSRC, (".cu", ".cuh")) + [os.path.join(ROOT, "include", "nb200.h")]
    if not force and _newer(out, deps):
        return out
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DNB200_PRECISION=%d" % PRECISIONS[precision], "-o", out,
                                    os.path.join(CSRC, "nb200_api.cu"), "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return out


def build_variant(precision, variant, defines):
    """A/B build of the kernel library with extra -D flags, e.g. build_variant("f64", "minb6", ["NB200_BHG_MINB=6"])."""
    out = lib_path(precision, variant)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DNB200_PRECISION=%d" % PRECISIONS[precision]] + ["-D" + d for d in defines] + [
        "-o", out, os.path.join(CSRC, "nb200_api.cu"), "-ldl"]
    subprocess.run(cmd, check=True)
    return out


def build_adapter(precision, force=False):
    """C++ nbody_engine subclass; compiled against the reference's headers (never copied)."""
    out = adapter_path(precision)
    if not os.path.isdir(REFERENCE):
        return out if os.path.exists(out) else None
    deps = _sources(HOST, (".cpp", ".h")) + [os.path.join(ROOT, "include", "nb200.h")]
    if not force and _newer(out, deps):
        return out
    cmd = ["/usr/bin/g++", "-std=gnu++17", "-O2", "-fPIC", "-shared", "-fopenmp", "-w",
           "-DNB_COORD_PRECISION=%d" % PRECISIONS[precision], "-DNB200_PRECISION=%d" % PRECISIONS[precision],
           "-I" + os.path.join(ROOT, "oracle", "qtshim"), "-I" + REFERENCE, "-I" + os.path.join(ROOT, "include"),
           "-o", out] + _sources(HOST, (".cpp",)) + [
               # base-class symbols come from the shim-built reference, kernels from libnb200
               "-L" + os.path.join(ROOT, "oracle", "_ref"), "-lnbref_%s" % precision,
               "-L" + CSRC, "-lnb200_%s" % precision,
               "-Wl,-rpath,$ORIGIN/../../oracle/_ref", "-Wl,-rpath,$ORIGIN/../csrc", "-ldl"]
    subprocess.run(cmd, check=True)
    return out


def build_all(force=False, verbose=False):
    built = []
    for p in PRECISIONS:
        built.append(build_kernels(p, force=force, verbose=verbose))
    if os.path.isdir(HOST) and _sources(HOST, (".cpp",)):
        for p in PRECISIONS:
            r = build_adapter(p, force=force)
            if r:
                built.append(r)
    return built


if __name__ == "__main__":
    for path in build_all(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(path)
