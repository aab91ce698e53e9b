"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/kompass_b200.h declares, and fails LOUDLY (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "kompass_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(pkg):
    decl = declared_symbols()
    assert len(decl) >= 35
    L = pkg.lib()
    for s in decl:
        assert hasattr(L, s), f"{s} declared in include/kompass_b200.h but not exported"
    assert sorted(pkg.ABI_SYMBOLS) == decl


def test_library_is_sm100a_only(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_version_and_error_text(pkg):
    assert b"sm_100a" in pkg.lib().kc_version()


def test_no_cpu_fallback_without_device(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.KompassB200Error, match="no usable CUDA device"):
        pkg.Planner(pkg.planner_config())
    with pytest.raises(pkg.KompassB200Error, match="no usable CUDA device"):
        pkg.LocalMapperGPU(10, 10, 0.1, (0, 0, 0), 0.0, False, 10, 0.1, 1.0, 0.0, 10.0)


def test_invalid_arguments_are_rejected_before_device_use(pkg):
    cfg = pkg.planner_config(time_step=0.0)
    with pytest.raises(IndexError):  # out-of-range parameter -> std::out_of_range in the reference
        pkg.Planner(cfg)
    cfg = pkg.planner_config(control_type=7)
    with pytest.raises(ValueError):
        pkg.Planner(cfg)


def test_product_does_not_reference_oracle():
    """The product path must never import, link or execute anything under oracle/."""
    pk = os.path.join(ROOT, "kompass-core_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "kompass_oracle" not in txt and "oracle/" not in txt.replace("the oracle", ""), f
    out = subprocess.run(["ldd", os.path.join(pk, "lib", "libkompass_b200.so")], capture_output=True,
                         text=True).stdout
    assert "oracle" not in out


def test_libm_compatible_atan2f(pkg):
    """The device atan2f source, built for the host, must match this box's libm bit for bit
    (pointcloud.h:244 bins by the float atan2)."""
    L = pkg.lib()
    rng = np.random.default_rng(7)
    n = 4_000_000
    y = rng.uniform(-12, 12, n).astype(np.float32)
    x = rng.uniform(-12, 12, n).astype(np.float32)
    # mix in extreme ratios, zeros, infinities
    x[::17] = np.ldexp(x[::17], rng.integers(-40, 40, len(x[::17]))).astype(np.float32)
    y[::13] = np.ldexp(y[::13], rng.integers(-40, 40, len(y[::13]))).astype(np.float32)
    x[:8] = [0.0, -0.0, 1.0, np.inf, -np.inf, 0.0, 5.0, -5.0]
    y[:8] = [0.0, 0.0, 0.3, np.inf, 1.0, -2.0, 0.0, -0.0]
    out = np.zeros(n, np.float32)
    L.kc_debug_atan2f_array(y.ctypes.data_as(C.POINTER(C.c_float)), x.ctypes.data_as(C.POINTER(C.c_float)),
                            out.ctypes.data_as(C.POINTER(C.c_float)), C.c_int64(n))
    libm = C.CDLL("libm.so.6")
    libm.atan2f.restype = C.c_float
    libm.atan2f.argtypes = [C.c_float, C.c_float]
    ref = np.arctan2(y, x)  # numpy float32 arctan2 -> may use SIMD loops; verify a slice via libm
    idx = rng.integers(0, n, 20000)
    for i in list(range(8)) + list(idx):
        r = np.float32(libm.atan2f(float(y[i]), float(x[i])))
        assert r.tobytes() == out[i].tobytes(), (y[i], x[i], r, out[i])
    # numpy agrees with libm almost everywhere; use it as a wide net with a tiny tolerance
    bad = np.flatnonzero(ref.view(np.uint32) != out.view(np.uint32))
    for i in bad[:2000]:
        r = np.float32(libm.atan2f(float(y[i]), float(x[i])))
        assert r.tobytes() == out[i].tobytes(), (y[i], x[i], r, out[i])
