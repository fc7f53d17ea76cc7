"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/deodr_b200.h declares, and fails
loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re

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
