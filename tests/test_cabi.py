"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/deodr_b200.h declares, and fails
loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from conftest import ROOT

from deodr_b200 import _cabi


def test_library_exports_every_declared_symbol(build_native):
    header = open(os.path.join(ROOT, "include", "deodr_b200.h")).read()
    declared = set(re.findall(r"\b(deodr_b200_\w+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/deodr_b200.h but not exported"
    assert declared == {s[0] for s in _cabi.SYMBOLS}
    assert b"sm_100a" in lib.deodr_b200_version()


def test_struct_layouts_match_header(build_native):
    # 13 pointers + 13 int32 (padded to 8) ; 5 pointers ; reference struct Scene layout (DR.h:56-90)
    assert C.sizeof(_cabi.SceneView) == 13 * 8 + 56
    assert C.sizeof(_cabi.Grads) == 5 * 8
    assert C.sizeof(_cabi.HostScene) == 200
    assert C.sizeof(_cabi.ViewIO) == 9 * 8
    assert C.sizeof(_cabi.Camera) == (12 + 9 + 5) * 8 + 8
    assert C.sizeof(_cabi.MeshTopology) == 6 * 8 + 16


def test_flag_values_match_header():
    """The Python constants of the *_views flags and the error codes are the header's enum values."""
    header = open(os.path.join(ROOT, "include", "deodr_b200.h")).read()
    values = {name: int(v) for name, v in re.findall(r"\b(DEODR_B200_[A-Z_]+)\s*=\s*(-?\d+)", header)}
    assert values["DEODR_B200_ANTIALIASE_ERROR"] == _cabi.ANTIALIASE_ERROR
    assert values["DEODR_B200_ERROR_ADJOINT_COMPLETE"] == _cabi.ERROR_ADJOINT_COMPLETE
    assert values["DEODR_B200_FORWARD_GEOMETRY"] == _cabi.FORWARD_GEOMETRY
    assert values["DEODR_B200_FORWARD_RESUME"] == _cabi.FORWARD_RESUME
    flags = [values[k] for k in ("DEODR_B200_ANTIALIASE_ERROR", "DEODR_B200_ERROR_ADJOINT_COMPLETE",
                                 "DEODR_B200_FORWARD_GEOMETRY", "DEODR_B200_FORWARD_RESUME")]
    assert all(f & (f - 1) == 0 for f in flags) and len(set(flags)) == 4  # distinct single bits
    assert values["DEODR_B200_ECUDA"] == _cabi.ECUDA and values["DEODR_B200_EINVAL"] == _cabi.EINVAL


def test_no_cpu_fallback_without_gpu(build_native):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _cabi.load()
    ws = C.c_void_p()
    rc = lib.deodr_b200_workspace_create(C.byref(ws), 0)
    assert rc == _cabi.ECUDA
    assert b"no CPU fallback" in lib.deodr_b200_last_error()
    from deodr_b200.renderer import Renderer

    with pytest.raises(RuntimeError):
        Renderer(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deodr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text and "libemul" not in text, f


# ---- the header seen by a C compiler, the boundary called from plain C (tests/ffi/abi_probe.c) -------------------------


def _abi_probe(tmp_path):
    import subprocess

    exe = str(tmp_path / "abi_probe")
    libdir = os.path.join(ROOT, "deodr_b200")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-o", exe, os.path.join(ROOT, "tests", "ffi", "abi_probe.c"), "-L", libdir, "-ldeodr_b200",
                    f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
    return lambda mode: subprocess.run([exe, mode], check=True, capture_output=True, text=True, timeout=300).stdout


def test_header_is_plain_c_and_the_ctypes_mirror_has_its_layout(build_native, tmp_path):
    """include/deodr_b200.h compiles as C99 (-pedantic -Werror: no C++ needed), and every field of every struct sits
    at the offset the ctypes description gives it (a cgo / JNI / Cython binding sees the same layout)."""
    lines = _abi_probe(tmp_path)("layout").split("\n")
    seen = {}
    for line in lines:
        if line.strip():
            name, value = line.split()
            seen[name] = int(value)
    mirrors = {"DeodrSceneView": _cabi.SceneView, "DeodrGrads": _cabi.Grads, "DeodrHostScene": _cabi.HostScene,
               "DeodrViewIO": _cabi.ViewIO, "DeodrCamera": _cabi.Camera, "DeodrMeshTopology": _cabi.MeshTopology}
    for cname, mirror in mirrors.items():
        assert seen[cname] == C.sizeof(mirror), cname
    checked = 0
    for key, offset in seen.items():
        if "." in key:
            cname, field = key.split(".")
            assert getattr(mirrors[cname], field).offset == offset, key
            checked += 1
    for cname in ("DeodrSceneView", "DeodrGrads", "DeodrHostScene", "DeodrViewIO"):  # every field was probed
        assert sum(k.startswith(cname + ".") for k in seen) == len(mirrors[cname]._fields_), cname
    assert checked == 26 + 5 + 31 + 9


def _two_triangles():
    """The scene tests/ffi/abi_probe.c renders."""
    from deodr_b200.scenes import SceneArrays

    return SceneArrays(
        faces=np.arange(6, dtype=np.uint32).reshape(2, 3), faces_uv=np.zeros((2, 3), dtype=np.uint32),
        ij=np.array([[1, 1], [12, 2], [3, 10], [4, 3], [14, 5], [6, 11]], dtype=np.float64),
        depths=np.array([1, 1, 1, 2, 2, 2], dtype=np.float64), textured=np.zeros(2, dtype=bool), uv=np.zeros((1, 2)),
        shade=np.ones(6), colors=np.array([[0.1], [0.2], [0.3], [0.7], [0.8], [0.9]]), shaded=np.zeros(2, dtype=bool),
        edgeflags=np.ones((2, 3), dtype=bool), height=12, width=16, nb_colors=1, texture=np.zeros((1, 1, 1)),
        background_image=None, background_color=np.array([0.5]), clockwise=False, backface_culling=False,
        strict_edge=True, perspective_correct=False, integer_pixel_centers=True)


def test_plain_c_caller_fails_loudly_without_a_gpu(build_native, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = _abi_probe(tmp_path)("call")
    assert out.startswith("workspace_create 3 ") and "no CPU fallback" in out, out


@pytest.mark.gpu
def test_plain_c_caller_renders_through_the_host_entry_point(build_native, checker, tmp_path):
    """deodr_b200_render_host called from a C program (no Python, no torch in the process) against the oracle."""
    image, z = checker.render(_two_triangles(), 1.0)
    out = _abi_probe(tmp_path)("call").splitlines()
    assert out[0] == "render_host 0 ok", out
    covered, total = int(out[1].split()[1]), float(out[1].split()[3])
    assert covered == int(np.isfinite(z).sum()) and covered > 0
    assert abs(total - float(image.sum())) <= 192 * 1e-6 + 1e-5  # (IMAGE_TOL per pixel, 6 printed decimals)
