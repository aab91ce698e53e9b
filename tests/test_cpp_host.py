"""Builds the C++ host-API test (reference Boost cases restated against the drop-in classes) and,
on a GPU box, runs it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_api")


def build(pkg):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", SRC, "-o", OUT, f"-L{libdir}",
                           "-lkompass_b200", f"-Wl,-rpath,{libdir}"])
    return OUT


def test_cpp_host_classes_compile_and_link(pkg):
    assert os.path.exists(build(pkg))


@pytest.mark.gpu
def test_cpp_host_classes_run(pkg):
    exe = build(pkg)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASSED" in r.stdout


def test_copy_pool_host_only():
    """The staging copy pool (kc_hostcopy.h) is plain host code: hammer it here on the CPU."""
    src = os.path.join(ROOT, "tests", "cpp", "test_copy_pool.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_copy_pool")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-pthread", src, "-o", exe])
    for env_extra in ({}, {"KOMPASS_B200_COPY_THREADS": "3", "KOMPASS_B200_COPY_SPIN_US": "0"},
                      {"KOMPASS_B200_COPY_THREADS": "0"}):
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env={**os.environ, **env_extra})
        assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout + r.stderr


def test_launch_geometry_host_only():
    """class_ctas_for (classifying CTAs of k_scatter) covers every tile of every query window of the grid
    and is monotone in the cell count (a batch launches for its largest window). Host code of the kernel
    header, compiled with nvcc, run on the CPU."""
    src = os.path.join(ROOT, "tests", "cpp", "test_launch_geometry.cu")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_launch_geometry")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-cudart", "static", src, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout + r.stderr
