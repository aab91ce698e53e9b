"""Builds libkompass_b200.so (sm_100a) in-tree with nvcc. No GPU needed to build.

    python kompass-core_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libkompass_b200.so")
SOURCES = ["kc_api.cu", "kc_planner.cu", "kc_mapper.cu", "kc_dwa.cu"]
HEADERS = ["kc_common.cuh", "kc_host_math.h", "kc_hostcopy.h", "kc_libm_compat.cuh", "kc_planner_kernels.cuh",
           os.path.join("..", "..", "include", "kompass_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    # numerics contract: no FMA contraction anywhere (device and host), IEEE div/sqrt
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=default",
    "-cudart", "static",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    for s in SOURCES:
        o = os.path.join(LIB_DIR, s.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(o)
    cmd = [_nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", LIB] + objs
    subprocess.check_call(cmd)
    return LIB


BIND_SRC = os.path.join(HERE, "bindings", "kompass_cpp_bindings.cpp")
BIND_DIR = os.path.join(HERE, "bindings", "_build")


def bindings_path():
    import sysconfig
    return os.path.join(BIND_DIR, "kompass_cpp" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_bindings(force=False):
    """The `kompass_cpp` extension module for the hot-path classes (pybind11 over the C++ mirror and
    libkompass_b200.so). Host-only compile: no CUDA code in this TU."""
    import sysconfig

    import pybind11

    out = bindings_path()
    deps = [BIND_SRC, os.path.join(HERE, "host", "kompass_b200.hpp"),
            os.path.join(HERE, "..", "include", "kompass_b200.h"), build()]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    os.makedirs(BIND_DIR, exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-fvisibility=hidden", BIND_SRC, "-o", out,
           "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"], "-L" + LIB_DIR,
           "-lkompass_b200", "-Wl,-rpath,$ORIGIN/../../lib"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_bindings(force="--force" in sys.argv))
