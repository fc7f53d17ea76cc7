# cython: language_level=3
# distutils: language = c
"""Cython binding of libdeodr_b200.so with the two names of deodr/differentiable_renderer_cython.pyx (:50, :206).

What a maintainer of the reference would add next to the existing pyx: the same Python-facing functions, the extern
block pointing at include/deodr_b200.h instead of C++/DifferentiableRenderer.h.  The marshalling is table-driven (one
loop over the scene's array fields) and the argument checks are those of the ctypes shim
(deodr_b200.differentiable_renderer_cython._check_common: the pyx's asserts, pyx:61-114).

Build (tests/test_bindings.py does exactly this):
    cython bindings/differentiable_renderer_b200.pyx
    gcc -shared -fPIC $(python3-config --includes) -Iinclude bindings/differentiable_renderer_b200.c \
        -Ldeodr_b200 -ldeodr_b200 -Wl,-rpath,$PWD/deodr_b200 -o differentiable_renderer_b200$(python3-config --extension-suffix)
"""
import numpy as np
from libc.stdint cimport int32_t, uint8_t, uint32_t
from libc.string cimport memset


cdef extern from "deodr_b200.h":
    ctypedef struct DeodrWorkspace:
        pass
    ctypedef struct DeodrHostScene:
        const uint32_t *faces
        const uint32_t *faces_uv
        const double *depths
        const double *uv
        const double *ij
        const double *shade
        const double *colors
        const uint8_t *edgeflags
        const uint8_t *textured
        const uint8_t *shaded
        int32_t nb_triangles
        int32_t nb_vertices
        int32_t clockwise
        int32_t backface_culling
        int32_t nb_uv
        int32_t height
        int32_t width
        int32_t nb_colors
        const double *texture
        int32_t texture_height
        int32_t texture_width
        const double *background_image
        const double *background_color
        double *uv_b
        double *ij_b
        double *shade_b
        double *colors_b
        double *texture_b
        int32_t strict_edge
        int32_t perspective_correct
        int32_t integer_pixel_centers
    int deodr_b200_workspace_create(DeodrWorkspace **ws, int device)
    int deodr_b200_render_host(DeodrWorkspace *ws, const DeodrHostScene *scene, double *image, double *z_buffer,
                               double sigma, int antialiase_error, const double *obs, double *err_buffer)
    int deodr_b200_render_b_host(DeodrWorkspace *ws, const DeodrHostScene *scene, double *image, double *z_buffer,
                                 double *image_b, double sigma, int antialiase_error, const double *obs,
                                 double *err_buffer, double *err_buffer_b)
    const char *deodr_b200_last_error()


class DeodrB200Error(RuntimeError):
    pass


cdef DeodrWorkspace *_workspace = NULL


cdef DeodrWorkspace *workspace() except NULL:
    global _workspace
    if _workspace == NULL and deodr_b200_workspace_create(&_workspace, 0) != 0:
        _workspace = NULL
        raise DeodrB200Error(deodr_b200_last_error().decode())
    return _workspace


cdef size_t address(a):
    return <size_t>a.ctypes.data


def _flat(a, dtype):
    if hasattr(a, "detach"):  # torch CPU tensor
        a = a.detach().numpy()
    return np.ascontiguousarray(a, dtype=dtype).reshape(-1)


cdef dict fill(DeodrHostScene *s, scene, int nb_colors, bint with_grads):
    """Fills `s` from the duck-typed scene; returns the arrays that must outlive the call."""
    keep = {}
    for name, dtype in (("faces", np.uint32), ("faces_uv", np.uint32), ("edgeflags", np.uint8), ("textured", np.uint8),
                        ("shaded", np.uint8), ("depths", np.double), ("uv", np.double), ("ij", np.double),
                        ("shade", np.double), ("colors", np.double), ("texture", np.double)):
        keep[name] = _flat(getattr(scene, name), dtype)
    memset(s, 0, sizeof(DeodrHostScene))
    s.faces = <const uint32_t *>address(keep["faces"])
    s.faces_uv = <const uint32_t *>address(keep["faces_uv"])
    s.edgeflags = <const uint8_t *>address(keep["edgeflags"])
    s.textured = <const uint8_t *>address(keep["textured"])
    s.shaded = <const uint8_t *>address(keep["shaded"])
    s.depths = <const double *>address(keep["depths"])
    s.uv = <const double *>address(keep["uv"])
    s.ij = <const double *>address(keep["ij"])
    s.shade = <const double *>address(keep["shade"])
    s.colors = <const double *>address(keep["colors"])
    s.texture = <const double *>address(keep["texture"])
    if scene.background_image is not None:
        keep["background"] = _flat(scene.background_image, np.double)
        s.background_image = <const double *>address(keep["background"])
    else:
        keep["background"] = _flat(scene.background_color, np.double)
        s.background_color = <const double *>address(keep["background"])
    s.nb_triangles, s.nb_vertices, s.nb_uv = scene.faces.shape[0], scene.depths.shape[0], scene.uv.shape[0]
    s.height, s.width, s.nb_colors = scene.height, scene.width, nb_colors
    s.texture_height, s.texture_width = scene.texture.shape[0], scene.texture.shape[1]
    s.clockwise, s.backface_culling = bool(scene.clockwise), bool(scene.backface_culling)
    s.strict_edge, s.perspective_correct = bool(scene.strict_edge), bool(scene.perspective_correct)
    s.integer_pixel_centers = bool(scene.integer_pixel_centers)
    if with_grads:  # accumulated into flattened copies that are rebound afterwards, like pyx:297-312, 406-410
        for name in ("uv_b", "ij_b", "shade_b", "colors_b", "texture_b"):
            keep[name] = _flat(getattr(scene, name), np.double).copy()
        s.uv_b = <double *>address(keep["uv_b"])
        s.ij_b = <double *>address(keep["ij_b"])
        s.shade_b = <double *>address(keep["shade_b"])
        s.colors_b = <double *>address(keep["colors_b"])
        s.texture_b = <double *>address(keep["texture_b"])
    return keep


cdef check(int rc, bint check_valid):
    if rc != 0:
        message = deodr_b200_last_error().decode()
        if check_valid and "greater than scene.nb_" in message:  # the pyx asserts the index ranges (pyx:76-77)
            raise AssertionError(message)
        raise DeodrB200Error(message)


def renderSceneCpp(scene, double sigma, double[:, :, ::1] image, double[:, ::1] z_buffer, bint antialiase_error=0,
                   double[:, :, ::1] obs=None, double[:, ::1] err_buffer=None, bint check_valid=1):
    cdef DeodrHostScene s
    if check_valid:
        from deodr_b200.differentiable_renderer_cython import _check_common
        _check_common(scene, np.asarray(image), np.asarray(z_buffer))
    keep = fill(&s, scene, image.shape[2], False)
    check(deodr_b200_render_host(workspace(), &s, &image[0, 0, 0], &z_buffer[0, 0], sigma, antialiase_error,
                                 &obs[0, 0, 0] if obs is not None else NULL,
                                 &err_buffer[0, 0] if err_buffer is not None else NULL), check_valid)
    del keep


def renderSceneBCpp(scene, double sigma, double[:, :, ::1] image, double[:, ::1] z_buffer,
                    double[:, :, ::1] image_b=None, bint antialiase_error=0, double[:, :, ::1] obs=None,
                    double[:, ::1] err_buffer=None, double[:, ::1] err_buffer_b=None, bint check_valid=1):
    cdef DeodrHostScene s
    if check_valid:
        from deodr_b200.differentiable_renderer_cython import _check_common
        _check_common(scene, np.asarray(image), np.asarray(z_buffer))
        assert antialiase_error or image_b is not None
    keep = fill(&s, scene, image.shape[2], True)
    check(deodr_b200_render_b_host(workspace(), &s, &image[0, 0, 0], &z_buffer[0, 0],
                                   &image_b[0, 0, 0] if image_b is not None else NULL, sigma, antialiase_error,
                                   &obs[0, 0, 0] if obs is not None else NULL,
                                   &err_buffer[0, 0] if err_buffer is not None else NULL,
                                   &err_buffer_b[0, 0] if err_buffer_b is not None else NULL), check_valid)
    for name in ("uv_b", "ij_b", "shade_b", "colors_b", "texture_b"):
        setattr(scene, name, keep[name].reshape(np.shape(getattr(scene, name))))
