"""CPU: the streaming kernels of the host path (deodr_b200/csrc/host_simd.cpp: fp64 <-> fp32 conversions with
non-temporal stores, accumulate, copy, zero fill, mirror comparison) against numpy, bit for bit, on every alignment and
on sizes around the vector width.  Host-only code of the product, so it is tested here without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-C", os.path.join(HERE, "emul"), "libsimd.so"], check=True, capture_output=True)
    so = C.CDLL(os.path.join(HERE, "emul", "libsimd.so"))
    so.hook_equal_f32.restype = C.c_int
    for name in ("hook_f64_to_f32", "hook_f32_to_f64", "hook_f32_add_f64", "hook_copy", "hook_equal_f32"):
        getattr(so, name).argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    so.hook_zero.argtypes = [C.c_void_p, C.c_long]
    return so


SIZES = [0, 1, 7, 8, 15, 16, 17, 63, 64, 65, 1000, 4099]
OFFSETS = [0, 1, 3, 5, 8, 13]


def values(rng, n):
    v = rng.normal(size=n) * rng.choice([1e-30, 1e-3, 1.0, 1e3, 1e30], size=n)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-320, 3.4e38, 3.5e38, 1 + 2.0**-24, 1 + 2.0**-25])
    k = min(n, special.size)
    v[:k] = special[:k]
    return v


@pytest.mark.parametrize("n", SIZES)
def test_conversions_match_numpy_bit_for_bit(lib, n):
    rng = np.random.default_rng(n)
    for off_src in OFFSETS[:3]:
        for off_dst in OFFSETS:
            src64 = np.zeros(n + 16)[off_src:off_src + n]
            src64[:] = values(rng, n)
            dst32 = np.full(n + 32, 7.0, np.float32)
            view32 = dst32[off_dst:off_dst + n]
            lib.hook_f64_to_f32(view32.ctypes.data, src64.ctypes.data, n)
            with np.errstate(over="ignore", invalid="ignore"):
                want = src64.astype(np.float32)
            assert np.array_equal(view32.view(np.uint32), want.view(np.uint32))
            assert np.all(dst32[:off_dst] == 7.0) and np.all(dst32[off_dst + n:] == 7.0)  # nothing outside

            dst64 = np.full(n + 32, 7.0)
            view64 = dst64[off_dst:off_dst + n]
            lib.hook_f32_to_f64(view64.ctypes.data, want.ctypes.data, n)
            assert np.array_equal(view64.view(np.uint64), want.astype(np.float64).view(np.uint64))
            assert np.all(dst64[:off_dst] == 7.0) and np.all(dst64[off_dst + n:] == 7.0)

            acc = rng.normal(size=n + 32)
            before = acc.copy()
            finite = np.nan_to_num(want, nan=1.0, posinf=2.0, neginf=-2.0)
            lib.hook_f32_add_f64(acc[off_dst:off_dst + n].ctypes.data, finite.ctypes.data, n)
            assert np.array_equal(acc[off_dst:off_dst + n], before[off_dst:off_dst + n] + finite.astype(np.float64))
            assert np.array_equal(acc[:off_dst], before[:off_dst])


@pytest.mark.parametrize("n", SIZES)
def test_copy_zero_and_mirror_comparison(lib, n):
    rng = np.random.default_rng(100 + n)
    for off in OFFSETS:
        src = rng.integers(0, 256, size=n + 16, dtype=np.uint8)[off:off + n]
        dst = np.full(n + 80, 9, np.uint8)
        lib.hook_copy(dst[off:off + n].ctypes.data, src.ctypes.data, n)
        assert np.array_equal(dst[off:off + n], src) and np.all(dst[:off] == 9) and np.all(dst[off + n:] == 9)
        lib.hook_zero(dst[off:off + n].ctypes.data, n)
        assert np.all(dst[off:off + n] == 0) and np.all(dst[:off] == 9) and np.all(dst[off + n:] == 9)

        user = np.zeros(n + 8)[off % 4:off % 4 + n]
        user[:] = values(rng, n)
        with np.errstate(over="ignore", invalid="ignore"):
            mirror = user.astype(np.float32)
        assert lib.hook_equal_f32(user.ctypes.data, mirror.ctypes.data, n) == 1
        if n:
            for k in {0, n // 2, n - 1}:
                changed = user.copy()
                changed[k] = 12345.678 if not np.isclose(user[k], 12345.678) else 1.0
                assert lib.hook_equal_f32(changed.ctypes.data, mirror.ctypes.data, n) == 0
                # a change below fp32 resolution is NOT a change of the device copy
                tiny = user.copy()
                if np.isfinite(tiny[k]) and abs(tiny[k]) > 1e-30 and abs(tiny[k]) < 1e30:
                    as32 = np.float32(tiny[k])
                    tiny[k] = float(as32) * (1 + 2.0**-40)
                    with np.errstate(over="ignore"):
                        same = np.float32(tiny[k]) == as32
                    assert lib.hook_equal_f32(tiny.ctypes.data, mirror.ctypes.data, n) == int(same and float(as32) == float(mirror[k]))
