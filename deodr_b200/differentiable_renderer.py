"""2.5D scene API with the surface of ``deodr.differentiable_renderer`` (hot-path part only).

Mirrors, name for name, the callers of the raster core in deodr/differentiable_renderer.py: ``Scene2DBase`` (:16-45),
``renderScene`` (:48-126), ``renderSceneB`` (:129-249) and ``Scene2D`` (:525-734), so that code written against the
reference keeps working when it imports these names from ``deodr_b200`` instead.  All rendering goes through
``deodr_b200.differentiable_renderer_cython`` -> ``libdeodr_b200.so`` (sm_100a); numpy in, numpy out.

Not mirrored here (outside the hot-path scope): ``Camera``, ``Scene3D`` and the mesh classes.
"""

from __future__ import annotations

import copy
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import differentiable_renderer_cython


@dataclass
class Scene2DBase:
    """The 2.5D scene structure consumed by the raster core (DifferentiableRenderer.h:56-90)."""

    faces: np.ndarray
    faces_uv: np.ndarray
    ij: np.ndarray
    depths: np.ndarray
    textured: np.ndarray
    uv: np.ndarray
    shade: np.ndarray
    colors: np.ndarray
    shaded: np.ndarray
    edgeflags: np.ndarray
    height: int
    width: int
    nb_colors: int
    texture: np.ndarray
    background_image: Optional[np.ndarray] = None
    background_color: Optional[np.ndarray] = None
    uv_b: Optional[np.ndarray] = None
    ij_b: Optional[np.ndarray] = None
    shade_b: Optional[np.ndarray] = None
    colors_b: Optional[np.ndarray] = None
    texture_b: Optional[np.ndarray] = None
    clockwise: bool = False
    backface_culling: bool = True
    strict_edge: bool = True
    perspective_correct: bool = False
    integer_pixel_centers: bool = True


def _check_backgrounds(scene) -> None:
    assert (scene.background_image is not None) != (scene.background_color is not None), (
        "You need to provide either background_image or background_color"
    )


def renderScene(scene, sigma, image, z_buffer, antialiase_error=False, obs=None, err_buffer=None,
                check_valid=True) -> None:
    """Forward render into the caller's ``image[H,W,C]`` / ``z_buffer[H,W]`` (float64, overwritten)."""
    if check_valid:
        assert image is not None
        assert z_buffer is not None
        assert scene.faces.dtype == np.uint32
        _check_backgrounds(scene)
        if antialiase_error:
            assert err_buffer is not None, "You need to provide err_buffer"
            assert obs is not None, "You need to provide obs"
    differentiable_renderer_cython.renderSceneCpp(scene, sigma, image, z_buffer, antialiase_error, obs, err_buffer,
                                                  check_valid)


def renderSceneB(scene, sigma, image, z_buffer, image_b=None, antialiase_error=False, obs=None, err_buffer=None,
                 err_buffer_b=None, check_valid=True) -> None:
    """Adjoint of ``renderScene``: accumulates into ``scene.ij_b / colors_b / uv_b / shade_b / texture_b``."""
    if check_valid:
        assert image is not None
        assert z_buffer is not None
        assert scene.faces.dtype == np.uint32
        _check_backgrounds(scene)
        for name in ("uv_b", "ij_b", "shade_b", "colors_b", "texture_b"):
            assert getattr(scene, name) is not None, f"scene.{name} must be allocated"
        if antialiase_error:
            assert err_buffer is not None and obs is not None and err_buffer_b is not None
        else:
            assert image_b is not None
    differentiable_renderer_cython.renderSceneBCpp(scene, sigma, image, z_buffer, image_b, antialiase_error, obs,
                                                   err_buffer, err_buffer_b, check_valid)


class Scene2D(Scene2DBase):
    """A set of 2D vertices with depths, and faces indexing them; see deodr/differentiable_renderer.py:525-598.

    Pixel-centre convention: with ``integer_pixel_centers`` (default) the centre of the upper-left pixel is (0, 0) and
    of the lower-right one (width - 1, height - 1); otherwise centres sit at half-integer coordinates.
    """

    def __init__(self, faces, faces_uv, ij, depths, textured, uv, shade, colors, shaded, edgeflags, height, width,
                 nb_colors, texture, background_image=None, background_color=None, clockwise=False,
                 backface_culling=False, strict_edge=True, perspective_correct=False, integer_pixel_centers=True):
        super().__init__(
            faces=faces, faces_uv=faces_uv, ij=ij, depths=depths, textured=textured, uv=uv, shade=shade,
            colors=colors, shaded=shaded, edgeflags=edgeflags, height=height, width=width, nb_colors=nb_colors,
            texture=texture, background_image=background_image, background_color=background_color,
            clockwise=clockwise, backface_culling=backface_culling, strict_edge=strict_edge,
            perspective_correct=perspective_correct, integer_pixel_centers=integer_pixel_centers,
        )
        self.uv_b = np.zeros(np.shape(self.uv))
        self.ij_b = np.zeros(np.shape(self.ij))
        self.shade_b = np.zeros(np.shape(self.shade))
        self.colors_b = np.zeros(np.shape(self.colors))
        self.texture_b = np.zeros(np.shape(self.texture))
        self.store_backward: Tuple = ()

    def clear_gradients(self) -> None:
        """In-place zero of the five gradient arrays (deodr/differentiable_renderer.py:599-610).  Large C-contiguous
        arrays are zeroed by the library's copy threads (30+ MB of `fill(0)` per step on a 1M-triangle scene)."""
        grads = []
        for name in ("uv_b", "ij_b", "shade_b", "colors_b", "texture_b"):
            grad = getattr(self, name)
            assert grad is not None
            grads.append(grad)
        big = [g for g in grads if isinstance(g, np.ndarray) and g.flags.c_contiguous and g.flags.writeable]
        if sum(g.nbytes for g in big) >= (4 << 20):
            from . import differentiable_renderer_cython as shim

            shim.zero_arrays(big)
            grads = [g for g in grads if not any(g is b for b in big)]
        for grad in grads:
            grad.fill(0)

    def _new_buffers(self) -> Tuple[np.ndarray, np.ndarray]:
        return np.zeros((self.height, self.width, self.nb_colors)), np.zeros((self.height, self.width))

    def render(self, sigma: float = 1) -> Tuple[np.ndarray, np.ndarray]:
        image, z_buffer = self._new_buffers()
        renderScene(self, sigma, image, z_buffer, False, None, None)
        self.store_backward = (sigma, image, z_buffer)
        return image, z_buffer

    def render_error(self, obs: np.ndarray, sigma: float = 1):
        image, z_buffer = self._new_buffers()
        err_buffer = np.empty((self.height, self.width))
        renderScene(self, sigma, image, z_buffer, True, obs, err_buffer)
        self.store_backward = (sigma, obs, image, z_buffer, err_buffer)
        return image, z_buffer, err_buffer

    def _require_differentiable(self) -> None:
        if self.perspective_correct:
            raise BaseException("perspective_correct not supported yet for gradient back propagation")
        if not self.backface_culling:
            raise BaseException(
                "use backface_culling=True if you use gradient backpropagation "
                "to get valid gradient through edge anti-aliasing."
            )

    def render_backward(self, image_b: np.ndarray, make_copies: bool = True) -> None:
        self._require_differentiable()
        sigma, image, z_buffer = self.store_backward
        # the reference un-blends `image` in place during the adjoint, hence its optional copy; our adjoint never
        # writes to image / image_b, the flag is kept for signature compatibility
        renderSceneB(self, sigma, image.copy() if make_copies else image, z_buffer, image_b, False, None, None, None)

    def render_error_backward(self, err_buffer_b: np.ndarray, make_copies: bool = True) -> None:
        self._require_differentiable()
        sigma, obs, image, z_buffer, err_buffer = self.store_backward
        renderSceneB(self, sigma, image, z_buffer, None, True, obs,
                     err_buffer.copy() if make_copies else err_buffer, err_buffer_b)

    def render_compare_and_backward(self, obs: np.ndarray, sigma: float = 1, antialiase_error: bool = False,
                                    mask: Optional[np.ndarray] = None, clear_gradients: bool = True,
                                    make_copies: bool = True):
        """L2 comparison with ``obs`` and back-propagation: returns ``(image, z_buffer, err_buffer, err)``."""
        if self.perspective_correct:
            raise BaseException("perspective_correct not supported yet for gradient back propagation")
        if mask is None:
            mask = np.ones((obs.shape[0], obs.shape[1]))
        if antialiase_error:
            image, z_buffer, err_buffer = self.render_error(obs, sigma)
        else:
            image, z_buffer = self.render(sigma)
        if clear_gradients:
            self.clear_gradients()
        if antialiase_error:
            err_buffer = err_buffer * mask
            err = float(np.sum(err_buffer))
            self.render_error_backward(copy.copy(mask), make_copies=make_copies)
        else:
            residual = (image - obs) * mask[:, :, None]
            err_buffer = residual**2
            err = float(np.sum(err_buffer))
            self.render_backward(2 * residual, make_copies=make_copies)
        return image, z_buffer, err_buffer, err
