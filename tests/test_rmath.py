"""CPU: the division-free floor(RN(a/b)) of deodr_b200/csrc/rmath.h (candidate from an fp32 quotient, settled by an
exact FMA sign test, real division only inside the half-ulp band) is bit-identical to the reference's
`(short)floor(a / b)` / `(short)ceil(a / b)` with clamping and robust fall-back (DifferentiableRenderer.h:440-519)."""
import ctypes as C

import numpy as np
import pytest
from canon import Emulator


@pytest.fixture(scope="module")
def lib():
    e = Emulator().lib
    e.emul_div_mismatches.restype = C.c_long
    e.emul_div_mismatches.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_int]
    e.emul_tri_span_mismatches.restype = C.c_long
    e.emul_tri_span_mismatches.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int]
    e.emul_span_mismatches.restype = C.c_long
    e.emul_span_mismatches.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int]
    return e


def mismatches(lib, a, b, lo=-1, hi=32767):
    a = np.ascontiguousarray(np.broadcast_to(a, np.broadcast(a, b).shape), dtype=np.float64).ravel()
    b = np.ascontiguousarray(np.broadcast_to(b, a.shape) if np.ndim(b) == 0 else b, dtype=np.float64).ravel()
    return lib.emul_div_mismatches(a.ctypes.data, b.ctypes.data, a.size, lo, hi)


def test_generic_quotients(lib):
    rng = np.random.default_rng(0)
    n = 500_000
    assert mismatches(lib, rng.normal(size=n) * 1000, rng.normal(size=n)) == 0
    assert mismatches(lib, rng.uniform(-40000, 40000, size=n), np.ones(n)) == 0
    assert mismatches(lib, rng.uniform(-40000, 40000, size=n) * 1e-3, np.full(n, 1e-3), lo=3, hi=700) == 0


def test_quotients_at_and_around_integers(lib):
    """a = x*b exactly, a few ulps around it, and relative perturbations around the half-ulp rounding band."""
    rng = np.random.default_rng(1)
    n = 300_000
    b = rng.normal(size=n) * rng.choice([1e-3, 1.0, 1e3], size=n)
    x = rng.integers(-3000, 3000, size=n).astype(np.float64)
    a = x * b
    bad = mismatches(lib, a, b)
    up, down = a.copy(), a.copy()
    for _ in range(5):
        up, down = np.nextafter(up, np.inf), np.nextafter(down, -np.inf)
        bad += mismatches(lib, up, b) + mismatches(lib, down, b)
    for eps in (2.0**-54, 2.0**-53, 2.0**-52, 2.0**-51, 2.0**-50, 1e-15, 1e-14, 1e-12):
        bad += mismatches(lib, a * (1 - eps), b) + mismatches(lib, a * (1 + eps), b)
    assert bad == 0


def test_special_operands(lib):
    sp = np.array([0.0, -0.0, 1e-320, -1e-320, 1e-300, 1e-45, 3e-45, 1e-40, 1e-30, 1e30, 1e300, np.inf, -np.inf,
                   np.nan, 1.0, -1.0, 32766.5, 32767.0, -32767.5, 0.5, -0.5])
    a, b = np.meshgrid(sp, sp)
    assert mismatches(lib, a.ravel(), b.ravel()) == 0
    assert mismatches(lib, a.ravel(), b.ravel(), lo=5, hi=9) == 0


def test_edge_row_span_matches_line_by_line_formulation(lib):
    """edge_row_span computes the four half-plane quotients side by side before the sequential clamps; it must give
    the spans of the formulation that follows DifferentiableRenderer.h:2620-2648 line by line, including the
    degenerate half-planes (zero / tiny / huge coefficients) that take the incremental fall-back."""
    rng = np.random.default_rng(7)
    n = 400_000
    ineq = rng.normal(size=(n, 12)) * rng.choice([1e-6, 1e-2, 1.0, 50.0], size=(n, 1))
    # constant terms of the size of pixel coordinates; some exactly-integer crossings; some degenerate rows
    ineq[:, 2::3] *= rng.choice([1.0, 100.0, 2000.0], size=(n, 4))
    k = rng.integers(0, n, size=n // 10)
    ineq[k, 0] = rng.choice([0.0, 1e-300, -1e-300, 1e-20, 1.0, -1.0], size=k.size)
    k = rng.integers(0, n, size=n // 10)
    ineq[k, 3] = np.round(ineq[k, 3])
    ineq[k, 4] = np.round(ineq[k, 4])
    ineq[k, 5] = np.round(ineq[k, 5])
    y = rng.integers(0, 2048, size=n).astype(np.int32)
    ineq = np.ascontiguousarray(ineq)
    assert lib.emul_span_mismatches(ineq.ctypes.data, y.ctypes.data, n, 2048) == 0
    assert lib.emul_span_mismatches(ineq.ctypes.data, y.ctypes.data, n, 37) == 0


@pytest.mark.parametrize("strict", [1, 0])
def test_triangle_row_spans_match_line_by_line_formulation(lib, strict):
    """tri_half_span computes the left and right bounds side by side; same spans as the line-by-line formulation of
    DifferentiableRenderer.h:864-906 on random, sliver, axis-aligned and integer-vertex triangles."""
    rng = np.random.default_rng(11 + strict)
    n = 60_000
    c = rng.uniform(-20, 276, size=(n, 1, 2))
    V = c + rng.normal(size=(n, 3, 2)) * rng.choice([0.3, 2.0, 15.0, 120.0], size=(n, 1, 1))
    k = rng.integers(0, n, size=n // 8)
    V[k] = np.round(V[k])                      # vertices on pixel centres: quotients exactly at integers
    k = rng.integers(0, n, size=n // 8)
    V[k, 1, 1] = V[k, 0, 1]                    # horizontal edge
    k = rng.integers(0, n, size=n // 8)
    V[k, 2, 0] = V[k, 0, 0]                    # vertical edge
    k = rng.integers(0, n, size=n // 16)
    V[k, 2] = (V[k, 0] + V[k, 1]) / 2 + 1e-9   # slivers
    V = np.ascontiguousarray(V.reshape(n, 6))
    assert lib.emul_tri_span_mismatches(V.ctypes.data, n, 256, 256, strict) == 0
