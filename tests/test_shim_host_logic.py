"""Host logic of the FFI drop-in (deodr_b200/differentiable_renderer_cython.py) without a GPU: the argument checks of
deodr/differentiable_renderer_cython.pyx:61-114, 219-293 fire BEFORE anything native is touched, and the marshalling
block (the pyx's flatten()/ascontiguousarray lines, pyx:117-171, 297-312) fills DeodrHostScene with the caller's own
buffers wherever the layout allows it."""
import ctypes as C

import numpy as np
import pytest

from deodr_b200 import _cabi
from deodr_b200 import differentiable_renderer_cython as shim
from deodr_b200.differentiable_renderer import Scene2D
from deodr_b200.scenes import soup_scene

FIELDS = ("faces", "faces_uv", "ij", "depths", "textured", "uv", "shade", "colors", "shaded", "edgeflags", "height",
          "width", "nb_colors", "texture", "background_image", "background_color", "clockwise", "backface_culling",
          "strict_edge", "perspective_correct", "integer_pixel_centers")


@pytest.fixture()
def scene(texture):
    np.random.seed(2)
    s = soup_scene(n_tri=12, width=96, height=88, texture=texture[::4, ::4].copy())
    s.background_image, s.background_color = None, np.array([0.3, 0.5, 0.7])
    for k in ("ij", "depths", "uv", "shade", "colors", "texture"):  # canonical layout: what a fitter holds
        setattr(s, k, np.ascontiguousarray(getattr(s, k), dtype=np.float64))
    for k in ("faces", "faces_uv"):
        setattr(s, k, np.ascontiguousarray(getattr(s, k), dtype=np.uint32))
    s2 = Scene2D(**{k: getattr(s, k) for k in FIELDS})
    s2.clear_gradients()
    return s2


@pytest.fixture()
def buffers(scene):
    return np.empty((scene.height, scene.width, scene.nb_colors)), np.empty((scene.height, scene.width))


class NeverCalled:
    """Stands in for the renderer: a check that lets a bad argument through would reach it."""

    def __getattr__(self, name):
        raise RuntimeError(f"native entry point {name} reached")


@pytest.fixture()
def no_native(monkeypatch):
    monkeypatch.setattr(shim, "default_renderer", lambda: NeverCalled())


def ptr(field):
    return C.cast(field, C.c_void_p).value


@pytest.mark.parametrize("breaker", [
    lambda s: setattr(s, "faces_uv", s.faces_uv[:-1]),                  # pyx:67
    lambda s: setattr(s, "colors", s.colors[:, 0]),                     # colors.ndim == 2
    lambda s: setattr(s, "uv", s.uv[:, :1]),                            # uv.shape[1] == 2
    lambda s: setattr(s, "ij", s.ij[:-1]),                              # ij.shape[0] == nb_vertices
    lambda s: setattr(s, "shade", s.shade[:-1]),
    lambda s: setattr(s, "colors", s.colors[:, :2]),                    # colors.shape[1] == nb_colors
    lambda s: setattr(s, "edgeflags", s.edgeflags[:, :2]),              # edgeflags.shape[1] == 3
    lambda s: setattr(s, "textured", s.textured[:-1]),
    lambda s: setattr(s, "shaded", s.shaded[:-1]),
    lambda s: setattr(s, "background_color", s.background_color[:2]),   # background_color.shape[0] == nb_colors
    lambda s: setattr(s, "texture", s.texture[:, :, :2]),               # texture.shape[2] == nb_colors
])
def test_forward_argument_checks_raise_assertion_error_before_any_native_call(scene, buffers, no_native, breaker):
    image, z = buffers
    breaker(scene)
    with pytest.raises(AssertionError):
        shim.renderSceneCpp(scene, 1.0, image, z)


def test_background_image_shape_is_checked(scene, buffers, no_native):
    image, z = buffers
    scene.background_color = None
    scene.background_image = np.zeros((scene.height, scene.width + 1, scene.nb_colors))
    with pytest.raises(AssertionError):
        shim.renderSceneCpp(scene, 1.0, image, z)


def test_zbuffer_shape_is_checked(scene, buffers, no_native):
    image, z = buffers
    with pytest.raises(AssertionError):
        shim.renderSceneCpp(scene, 1.0, image, z[:, :-1].copy())


@pytest.mark.parametrize("bad", ["float32", "fortran", "ndim", "list"])
def test_typed_buffer_arguments_are_refused_like_cython_does(scene, buffers, no_native, bad):
    image, z = buffers
    wrong = {"float32": image.astype(np.float32), "fortran": np.asfortranarray(image), "ndim": image[:, :, 0].copy(),
             "list": image.tolist()}[bad]
    with pytest.raises((ValueError, TypeError)):  # np.ndarray[double, ndim=3, mode="c"] (pyx:52)
        shim.renderSceneCpp(scene, 1.0, wrong, z)
    with pytest.raises((ValueError, TypeError)):
        shim.renderSceneBCpp(scene, 1.0, image, z, wrong)


def test_backward_argument_checks(scene, buffers, no_native):
    image, z = buffers
    image_b = np.zeros_like(image)
    with pytest.raises(AssertionError):  # image_b is required unless antialiase_error (pyx:291)
        shim.renderSceneBCpp(scene, 1.0, image, z, None)
    with pytest.raises(AssertionError):  # image_b.shape[:2] == image.shape[:2] (pyx:292-293)
        shim.renderSceneBCpp(scene, 1.0, image, z, image_b[:-1].copy())
    for name in ("uv_b", "ij_b", "shade_b", "colors_b", "texture_b"):
        keep = getattr(scene, name)
        setattr(scene, name, keep[:-1])
        with pytest.raises(AssertionError):
            shim.renderSceneBCpp(scene, 1.0, image, z, image_b)
        setattr(scene, name, keep)


def test_check_valid_0_skips_the_python_side_checks(scene, buffers, no_native):
    image, z = buffers
    scene.shade = scene.shade[:-1]
    with pytest.raises(RuntimeError, match="native entry point"):  # the checks are skipped: the call goes through
        shim.renderSceneCpp(scene, 1.0, image, z, check_valid=0)


def test_marshalling_points_at_the_callers_buffers(scene):
    h, keep = shim._marshal(scene, scene.nb_colors, with_grads=True)
    # C-contiguous arrays of the device-side dtype are passed as they are (the pyx copies each through flatten())
    for name in ("depths", "uv", "ij", "shade", "colors", "texture", "faces", "faces_uv"):
        assert ptr(getattr(h, name)) == np.asarray(getattr(scene, name)).ctypes.data, name
    for name in ("uv_b", "ij_b", "shade_b", "colors_b", "texture_b"):  # accumulated in place
        assert ptr(getattr(h, name)) == getattr(scene, name).ctypes.data, name
    assert (h.nb_triangles, h.nb_vertices, h.nb_uv) == (scene.faces.shape[0], scene.depths.shape[0], scene.uv.shape[0])
    assert (h.height, h.width, h.nb_colors) == (scene.height, scene.width, scene.nb_colors)
    assert (h.texture_height, h.texture_width) == scene.texture.shape[:2]
    assert (h.clockwise, h.backface_culling, h.strict_edge, h.perspective_correct, h.integer_pixel_centers) == tuple(
        int(bool(getattr(scene, k))) for k in ("clockwise", "backface_culling", "strict_edge", "perspective_correct",
                                               "integer_pixel_centers"))
    assert ptr(h.background_image) is None and ptr(h.background_color) == keep["background_color"].ctypes.data


def test_marshalling_converts_what_the_pyx_converts(scene):
    import torch

    faces64 = scene.faces.astype(np.int64)          # the pyx casts through ascontiguousarray(dtype=np.uint32)
    strided_ij = np.asfortranarray(scene.ij)        # not C-contiguous: flatten() copy
    bool_flags = scene.edgeflags.astype(bool)       # bool arrays are read as uint8 buffers, no copy
    torch_colors = torch.from_numpy(scene.colors)   # Scene3DPytorch leaves torch CPU tensors in scene_2d
    scene.faces, scene.ij, scene.edgeflags, scene.colors = faces64, strided_ij, bool_flags, torch_colors
    scene.ij_b = np.zeros(scene.ij.shape, dtype=np.float32)  # not float64: accumulate into a copy, rebind
    h, keep = shim._marshal(scene, scene.nb_colors, with_grads=True)
    assert keep["faces"].dtype == np.uint32 and np.array_equal(keep["faces"], faces64.reshape(-1))
    assert ptr(h.faces) == keep["faces"].ctypes.data != faces64.ctypes.data
    assert np.array_equal(keep["ij"], np.ascontiguousarray(strided_ij).reshape(-1))
    assert ptr(h.edgeflags) == bool_flags.ctypes.data and keep["edgeflags"].dtype == np.uint8
    assert ptr(h.colors) == torch_colors.numpy().ctypes.data
    assert keep["ij_b"].dtype == np.float64 and ptr(h.ij_b) == keep["ij_b"].ctypes.data != scene.ij_b.ctypes.data


def test_background_image_takes_precedence_like_the_pyx(scene):
    scene.background_image = np.full((scene.height, scene.width, scene.nb_colors), 0.25)
    h, keep = shim._marshal(scene, scene.nb_colors, with_grads=False)
    assert ptr(h.background_image) == scene.background_image.ctypes.data and ptr(h.background_color) is None
    assert "uv_b" not in keep and ptr(h.uv_b) is None  # forward: no gradient slots


def test_out_of_range_indices_surface_as_assertion_error_only_with_check_valid(scene, buffers, monkeypatch):
    image, z = buffers

    class Refusing:
        def render_host(self, *a):
            raise _cabi.DeodrB200Error(_cabi.EINVAL, "scene.faces value greater than scene.nb_vertices")

    monkeypatch.setattr(shim, "default_renderer", lambda: Refusing())
    with pytest.raises(AssertionError):              # pyx:76-77 `assert np.all(scene.faces < nb_vertices)`
        shim.renderSceneCpp(scene, 1.0, image, z, check_valid=1)
    with pytest.raises(_cabi.DeodrB200Error):        # the C core's own throw (DR.h:2703-2714), surfaced as an error
        shim.renderSceneCpp(scene, 1.0, image, z, check_valid=0)


# ---- Scene2D / renderScene / renderSceneB (deodr/differentiable_renderer.py:48-249, 599-734): error behaviour ---------


def test_scene2d_refuses_gradients_the_core_cannot_give(scene, no_native):
    """BaseException, like the reference (differentiable_renderer.py:630-637, 666-672), and before any rendering."""
    obs = np.zeros((scene.height, scene.width, scene.nb_colors))
    scene.perspective_correct = True
    with pytest.raises(BaseException, match="perspective_correct"):
        scene.render_compare_and_backward(obs)
    scene.store_backward = (1.0, obs.copy(), np.zeros((scene.height, scene.width)))
    with pytest.raises(BaseException, match="perspective_correct"):
        scene.render_backward(obs)
    scene.perspective_correct, scene.backface_culling = False, False
    with pytest.raises(BaseException, match="backface_culling"):
        scene.render_backward(obs)
    scene.store_backward = (1.0, obs, obs.copy(), np.zeros((scene.height, scene.width)), np.zeros((scene.height, scene.width)))
    with pytest.raises(BaseException, match="backface_culling"):
        scene.render_error_backward(np.ones((scene.height, scene.width)))


def test_render_scene_checks_of_the_python_layer(scene, buffers, no_native):
    from deodr_b200.differentiable_renderer import renderScene, renderSceneB

    image, z = buffers
    with pytest.raises(AssertionError):
        renderScene(scene, 1.0, None, z)
    with pytest.raises(AssertionError):
        renderScene(scene, 1.0, image, None)
    with pytest.raises(AssertionError, match="err_buffer"):  # differentiable_renderer.py:115-118
        renderScene(scene, 1.0, image, z, antialiase_error=True, obs=image.copy())
    with pytest.raises(AssertionError, match="obs"):
        renderScene(scene, 1.0, image, z, antialiase_error=True, err_buffer=z.copy())
    with pytest.raises(AssertionError):  # image_b is required without antialiase_error
        renderSceneB(scene, 1.0, image, z)
    with pytest.raises(AssertionError):  # ... and obs / err_buffer / err_buffer_b with it
        renderSceneB(scene, 1.0, image, z, antialiase_error=True, obs=image.copy(), err_buffer=z.copy())
    scene.colors_b = None
    with pytest.raises(AssertionError, match="colors_b"):
        renderSceneB(scene, 1.0, image, z, image_b=image.copy())
    scene.colors_b = np.zeros(scene.colors.shape)
    scene.faces = scene.faces.astype(np.int32)  # `assert scene.faces.dtype == np.uint32` (:75)
    with pytest.raises(AssertionError):
        renderScene(scene, 1.0, image, z)
    scene.faces = scene.faces.astype(np.uint32)
    scene.background_image = np.zeros((scene.height, scene.width, scene.nb_colors))  # both backgrounds: refused
    with pytest.raises(AssertionError, match="background"):
        renderScene(scene, 1.0, image, z)
    scene.background_image = scene.background_color = None                            # ... and neither
    with pytest.raises(AssertionError, match="background"):
        renderScene(scene, 1.0, image, z)


def test_clear_gradients_zeroes_in_place(scene):
    views = {}
    for name in ("uv_b", "ij_b", "shade_b", "colors_b", "texture_b"):
        getattr(scene, name)[...] = 1.5
        views[name] = getattr(scene, name)
    scene.clear_gradients()  # (small arrays: numpy fill; the copy-crew path needs the library's workspace, i.e. a GPU)
    for name, view in views.items():
        assert getattr(scene, name) is view and not view.any()
