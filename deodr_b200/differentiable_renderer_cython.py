"""Drop-in for the reference FFI module ``deodr.differentiable_renderer_cython``.

Exports the two names of deodr/differentiable_renderer_cython.pyx with the same signatures, argument meaning,
in-place numpy semantics and error behaviour:

* ``renderSceneCpp(scene, sigma, image, z_buffer, antialiase_error=0, obs=None, err_buffer=None, check_valid=1)``
  (pyx:50-202) - fully overwrites ``image`` / ``z_buffer``;
* ``renderSceneBCpp(scene, sigma, image, z_buffer, image_b=None, antialiase_error=0, obs=None, err_buffer=None,
  err_buffer_b=None, check_valid=1)`` (pyx:206-410) - accumulates into copies of ``scene.*_b`` and rebinds them.

``scene`` is duck-typed exactly like in the pyx (numpy arrays or torch CPU tensors).  The work is done by the sm_100a
library through ``deodr_b200_render_host`` / ``deodr_b200_render_b_host``; a maintainer of the reference can make the
whole ``deodr`` package run on the GPU with ``sys.modules['deodr.differentiable_renderer_cython'] = this module``
(see INTEGRATION.md).
"""

from __future__ import annotations

import numpy as np

from . import _cabi
from .renderer import default_renderer


def _call(fn, check_valid, *args):
    """Runs a host entry point; an out-of-range face index surfaces as AssertionError when ``check_valid`` (the pyx
    asserts, pyx:76-77) and as the library's error otherwise (the C core throws, DR.h:2703-2714)."""
    try:
        fn(*args)
    except _cabi.DeodrB200Error as exc:
        if check_valid and exc.code == _cabi.EINVAL and "greater than scene.nb_" in exc.message:
            raise AssertionError(exc.message) from None
        raise


def zero_arrays(arrays) -> None:
    """Zero-fills C-contiguous numpy arrays in place with the library's copy threads (host memory only)."""
    default_renderer().zero_host(arrays)


def _flat(a, dtype):
    """1-D C-contiguous view of ``a`` with ``dtype``; copies only when the layout / dtype requires it (the pyx always
    copies through ``flatten()``, pyx:117-131, which is pure overhead on a 1M-triangle scene)."""
    if hasattr(a, "detach"):  # torch CPU tensor (Scene3DPytorch leaves colors / depths as tensors)
        a = a.detach().numpy()
    a = np.asarray(a)
    if a.dtype == np.bool_ and dtype == np.uint8 and a.flags["C_CONTIGUOUS"]:
        return a.view(np.uint8).reshape(-1)  # the pyx takes bool arrays as uint8 buffers (cast=True): no copy
    return np.ascontiguousarray(a, dtype=dtype).reshape(-1)


def _marshal(scene, nb_colors, with_grads):
    keep = {
        "faces": _flat(scene.faces, np.uint32),
        "faces_uv": _flat(scene.faces_uv, np.uint32),
        "depths": _flat(scene.depths, np.double),
        "uv": _flat(scene.uv, np.double),
        "ij": _flat(scene.ij, np.double),
        "shade": _flat(scene.shade, np.double),
        "colors": _flat(scene.colors, np.double),
        "edgeflags": _flat(scene.edgeflags, np.uint8),
        "textured": _flat(scene.textured, np.uint8),
        "shaded": _flat(scene.shaded, np.uint8),
        "texture": _flat(scene.texture, np.double),
    }
    h = _cabi.HostScene()
    for name, arr in keep.items():
        setattr(h, name, arr.ctypes.data)
    if scene.background_image is not None:
        keep["background_image"] = _flat(scene.background_image, np.double)
        h.background_image, h.background_color = keep["background_image"].ctypes.data, None
    else:
        keep["background_color"] = _flat(scene.background_color, np.double)
        h.background_image, h.background_color = None, keep["background_color"].ctypes.data
    h.nb_triangles = scene.faces.shape[0]
    h.nb_vertices = scene.depths.shape[0]
    h.nb_uv = scene.uv.shape[0]
    h.height, h.width, h.nb_colors = int(scene.height), int(scene.width), int(nb_colors)
    h.texture_height, h.texture_width = int(scene.texture.shape[0]), int(scene.texture.shape[1])
    h.clockwise = int(bool(scene.clockwise))
    h.backface_culling = int(bool(scene.backface_culling))
    h.strict_edge = int(bool(scene.strict_edge))
    h.perspective_correct = int(bool(scene.perspective_correct))
    h.integer_pixel_centers = int(bool(scene.integer_pixel_centers))
    if with_grads:
        # The pyx accumulates into flattened COPIES of scene.*_b and rebinds the attributes (pyx:297-312, 406-410).
        # Copying (and later re-faulting) tens of MB per call is pure overhead, so C-contiguous float64 arrays are
        # accumulated in place; anything else goes through a converted copy that is rebound, exactly like the pyx.
        for name in ("uv_b", "ij_b", "shade_b", "colors_b", "texture_b"):
            cur = getattr(scene, name)
            if isinstance(cur, np.ndarray) and cur.dtype == np.float64 and cur.flags["C_CONTIGUOUS"] and cur.flags["WRITEABLE"]:
                keep[name] = cur.reshape(-1)
            else:
                keep[name] = _flat(cur, np.double).copy()
            setattr(h, name, keep[name].ctypes.data)
    return h, keep


def _check_common(scene, image, z_buffer):
    assert image is not None
    assert z_buffer is not None
    height, width, nb_colors = image.shape[0], image.shape[1], image.shape[2]
    nb_triangles = scene.faces.shape[0]
    assert nb_triangles == scene.faces_uv.shape[0]
    nb_vertices = scene.depths.shape[0]
    # index ranges (`assert np.all(scene.faces < nb_vertices)`, pyx:76-77) are validated on the device by the call
    # itself (k_check_scene), see _call below
    assert scene.colors.ndim == 2
    assert scene.uv.ndim == 2
    assert scene.ij.ndim == 2
    assert scene.shade.ndim == 1
    assert scene.edgeflags.ndim == 2
    assert scene.textured.ndim == 1
    assert scene.shaded.ndim == 1
    assert scene.uv.shape[1] == 2
    assert scene.ij.shape[0] == nb_vertices
    assert scene.ij.shape[1] == 2
    assert scene.shade.shape[0] == nb_vertices
    assert scene.colors.shape[0] == nb_vertices
    assert scene.colors.shape[1] == nb_colors
    assert scene.edgeflags.shape[0] == nb_triangles
    assert scene.edgeflags.shape[1] == 3
    assert scene.textured.shape[0] == nb_triangles
    assert scene.shaded.shape[0] == nb_triangles
    if scene.background_image is not None:
        assert scene.background_image.ndim == 3
        assert scene.background_image.shape[0] == height
        assert scene.background_image.shape[1] == width
        assert scene.background_image.shape[2] == nb_colors
    else:
        assert scene.background_color.shape[0] == nb_colors
    if scene.texture.size > 0:
        assert scene.texture.ndim == 3
        assert scene.texture.shape[0] > 0
        assert scene.texture.shape[1] > 0
        assert scene.texture.shape[2] == nb_colors
    assert z_buffer.shape[0] == height
    assert z_buffer.shape[1] == width


def _require_f64(name, a, ndim):
    if a is None:
        return
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.ndim == ndim and a.flags["C_CONTIGUOUS"]):
        # the pyx declares np.ndarray[double, ndim, mode="c"]: Cython raises ValueError / TypeError on mismatch
        raise ValueError(f"Buffer dtype mismatch or wrong layout for '{name}': expected C-contiguous float64, ndim={ndim}")


def renderSceneCpp(scene, sigma, image, z_buffer, antialiase_error=0, obs=None, err_buffer=None, check_valid=1):
    _require_f64("image", image, 3)
    _require_f64("z_buffer", z_buffer, 2)
    _require_f64("obs", obs, 3)
    _require_f64("err_buffer", err_buffer, 2)
    if check_valid:
        _check_common(scene, image, z_buffer)
    nb_colors = image.shape[2]
    h, keep = _marshal(scene, nb_colors, with_grads=False)
    if antialiase_error:
        assert err_buffer.shape[0] == image.shape[0] and err_buffer.shape[1] == image.shape[1]
        assert obs.shape == image.shape
    _call(default_renderer().render_host, check_valid, h, image, z_buffer, sigma, bool(antialiase_error), obs,
          err_buffer)
    del keep


def renderSceneBCpp(scene, sigma, image, z_buffer, image_b=None, antialiase_error=0, obs=None, err_buffer=None,
                    err_buffer_b=None, check_valid=1):
    _require_f64("image", image, 3)
    _require_f64("z_buffer", z_buffer, 2)
    _require_f64("image_b", image_b, 3)
    _require_f64("obs", obs, 3)
    _require_f64("err_buffer", err_buffer, 2)
    _require_f64("err_buffer_b", err_buffer_b, 2)
    if check_valid:
        _check_common(scene, image, z_buffer)
        assert scene.uv_b.ndim == 2 and scene.ij_b.ndim == 2 and scene.shade_b.ndim == 1 and scene.colors_b.ndim == 2
        assert scene.uv_b.shape == scene.uv.shape and scene.ij_b.shape == scene.ij.shape
        assert scene.shade_b.shape == scene.shade.shape and scene.colors_b.shape == scene.colors.shape
        if scene.texture.size > 0:
            assert scene.texture_b.shape == scene.texture.shape
        if not antialiase_error:
            assert image_b is not None
            assert image_b.shape[0] == image.shape[0] and image_b.shape[1] == image.shape[1]
    nb_colors = image.shape[2]
    h, keep = _marshal(scene, nb_colors, with_grads=True)
    _call(default_renderer().render_b_host, check_valid, h, image, z_buffer, image_b, sigma, bool(antialiase_error),
          obs, err_buffer, err_buffer_b)
    # the pyx rebinds the gradient attributes to the arrays the C core accumulated into (pyx:406-410)
    scene.uv_b = keep["uv_b"].reshape(scene.uv_b.shape)
    scene.ij_b = keep["ij_b"].reshape(scene.ij_b.shape)
    scene.shade_b = keep["shade_b"].reshape(scene.shade_b.shape)
    scene.colors_b = keep["colors_b"].reshape(scene.colors_b.shape)
    scene.texture_b = keep["texture_b"].reshape(scene.texture_b.shape)
