"""Seeded synthetic 2.5D scenes for the parity tests and the benchmark.

Two families:

* ``soup_scene``   - S-soup(n, W, H, seed): a triangle soup drawn with the same sequence of ``np.random`` calls as
  the reference generator ``create_example_scene`` (deodr/examples/triangle_soup_fitting.py:18-97), so that
  ``np.random.seed(2)`` reproduces the scene whose SHA-256 hashes are pinned by the reference test
  tests/test_render_mesh.py:34-53.  The texture (the decoded ``deodr/data/trefle.jpg``) is passed in by the caller
  (tests load it from tests/golden/).
* ``torus_scene``  - S-mesh(n): a closed bumpy torus with ``2 n^2`` triangles seen through a pinhole camera, with
  silhouette edge flags computed as in ``TriMeshAdjacencies.edge_on_silhouette`` (deodr/triangulated_mesh.py:153-166).

Everything here is host-side numpy input preparation; nothing is on the rendering hot path.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np


@dataclass
class SceneArrays:
    """Plain container with the ``Scene2DBase`` fields (deodr/differentiable_renderer.py:16-45)."""

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
    extra: dict = field(default_factory=dict)

    def zero_gradients(self) -> None:
        self.uv_b = np.zeros(self.uv.shape)
        self.ij_b = np.zeros(self.ij.shape)
        self.shade_b = np.zeros(self.shade.shape)
        self.colors_b = np.zeros(self.colors.shape)
        self.texture_b = np.zeros(self.texture.shape)


def _det3(tri2x3: np.ndarray) -> float:
    return float(np.linalg.det(np.vstack((tri2x3, np.ones(3)))))


def confetti_scene(n_tri: int = 3000, width: int = 64, height: int = 48, size: float = 2.5, seed: int = 0,
                   edge_ratio: float = 0.05, texture: Optional[np.ndarray] = None, nb_colors: int = 3) -> SceneArrays:
    """Many SMALL overlapping triangles (a few pixels each, random depths): dozens of micro-triangle records per tile
    and more candidates per pixel than the z pass keeps in its per-pixel lists - the regime of the 1M-triangle
    headline scene compressed into a test-sized image.  Own RNG (does not touch the global stream)."""
    rng = np.random.default_rng(seed)
    centre = rng.random((n_tri, 1, 2)) * np.array([height, width], dtype=np.float64)
    tri = centre + rng.normal(size=(n_tri, 3, 2)) * size
    # consistent orientation (front-facing for clockwise=False), degenerate ones nudged
    det = (tri[:, 1, 0] - tri[:, 0, 0]) * (tri[:, 2, 1] - tri[:, 0, 1]) - (tri[:, 2, 0] - tri[:, 0, 0]) * (
        tri[:, 1, 1] - tri[:, 0, 1])
    flip = det > 0
    tri[flip] = tri[flip][:, ::-1]
    faces = np.arange(3 * n_tri, dtype=np.uint32).reshape(-1, 3)
    if texture is None:
        texture = np.zeros((2, 2, nb_colors))
    return SceneArrays(
        faces=faces,
        faces_uv=faces.copy(),
        ij=np.ascontiguousarray(tri.reshape(-1, 2)),
        depths=np.repeat(rng.random(n_tri) * 4 + 1, 3) + rng.random(3 * n_tri) * 0.05,
        textured=np.zeros(n_tri, dtype=bool),
        uv=np.zeros((3 * n_tri, 2)),
        shade=np.zeros(3 * n_tri),
        colors=rng.random((3 * n_tri, nb_colors)),
        shaded=np.zeros(n_tri, dtype=bool),
        edgeflags=rng.random((n_tri, 3)) < edge_ratio,
        height=height,
        width=width,
        nb_colors=nb_colors,
        texture=texture,
        background_image=None,
        background_color=np.linspace(0.1, 0.9, nb_colors),
        clockwise=False,
        backface_culling=True,
        perspective_correct=False,
    )


def soup_scene(
    n_tri: int = 30,
    width: int = 200,
    height: int = 200,
    clockwise: bool = False,
    textured_ratio: float = 0.5,
    texture: Optional[np.ndarray] = None,
    min_det: float = 1500.0,
) -> SceneArrays:
    """Triangle soup; consumes the global ``np.random`` stream exactly like the reference generator.

    ``texture`` is the float64 ``[Ht, Wt, 3]`` material in [0, 1].  ``min_det`` (1500 px^2 in the reference) is the
    rejection threshold on the triangle determinant; lower it for small images (the reference loop never terminates
    for sides below ~80 px).
    """
    assert texture is not None and texture.ndim == 3
    h_mat, w_mat = texture.shape[0], texture.shape[1]
    to_pixels = np.array([[height, 0], [0, width]], dtype=np.float64)
    to_texels = np.array([[h_mat - 1, 0], [0, w_mat - 1]], dtype=np.float64)

    def draw() -> np.ndarray:
        centre = np.random.rand(2, 1)
        jitter = np.random.rand(2, 3)
        return to_pixels.dot(centre.dot(np.ones((1, 3))) + 0.5 * (-0.5 + jitter))

    ij, depths, textured, uv, shade, colors, shaded = [], [], [], [], [], [], []
    for _ in range(n_tri):
        tri = draw()
        while abs(_det3(tri)) < min_det:
            tri = draw()
        if _det3(tri) > 0:
            tri = np.fliplr(tri)
        ij.append(tri.T)
        depths.append(np.random.rand(1) * np.ones((3, 1)))
        is_textured = bool(np.random.rand(1) > (1 - textured_ratio))
        textured.append(is_textured)
        if is_textured:
            uv.append(to_texels.dot(np.array([[0, 1, 0.2], [0, 0.2, 1]])).T + 1)
            shade.append(np.random.rand(3, 1))
            colors.append(np.zeros((3, 3)))
            shaded.append(True)
        else:
            uv.append(np.zeros((3, 2)))
            shade.append(np.zeros((3, 1)))
            colors.append(np.random.rand(3, 3))
            shaded.append(False)

    faces = np.arange(3 * n_tri).reshape(-1, 3).astype(np.uint32)
    if clockwise:
        faces = np.ascontiguousarray(np.fliplr(faces))
    return SceneArrays(
        faces=faces,
        faces_uv=faces.copy(),
        ij=np.vstack(ij),
        depths=np.vstack(depths).squeeze(),
        textured=np.array(textured, dtype=bool),
        uv=np.vstack(uv),
        shade=np.vstack(shade).squeeze(),
        colors=np.vstack(colors),
        shaded=np.array(shaded, dtype=bool),
        edgeflags=np.ones((n_tri, 3), dtype=bool),
        height=height,
        width=width,
        nb_colors=3,
        texture=texture,
        background_image=np.tile(np.array([0.3, 0.5, 0.7])[None, None, :], (height, width, 1)),
        background_color=None,
        clockwise=clockwise,
        backface_culling=True,
        perspective_correct=False,
    )


# ---------------------------------------------------------------------------------------------------------------
# S-mesh: closed bumpy torus


def torus_mesh(n: int, big_r: float = 1.0, small_r: float = 0.4, bump: float = 0.05):
    """``n x n`` quads -> ``V = n^2`` vertices, ``T = 2 n^2`` triangles (closed manifold)."""
    u = 2 * np.pi * np.arange(n) / n
    v = 2 * np.pi * np.arange(n) / n
    uu, vv = np.meshgrid(u, v, indexing="ij")
    r = small_r + bump * np.sin(7 * uu) * np.sin(5 * vv)
    x = (big_r + r * np.cos(vv)) * np.cos(uu)
    y = (big_r + r * np.cos(vv)) * np.sin(uu)
    z = r * np.sin(vv)
    vertices = np.stack((x, y, z), axis=-1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    i1, j1 = (i + 1) % n, (j + 1) % n
    a = (i * n + j).ravel()
    b = (i1 * n + j).ravel()
    c = (i1 * n + j1).ravel()
    d = (i * n + j1).ravel()
    faces = np.empty((2 * n * n, 3), dtype=np.uint32)
    faces[0::2] = np.stack((a, b, c), axis=1)
    faces[1::2] = np.stack((a, c, d), axis=1)
    uv01 = np.stack((uu / (2 * np.pi), vv / (2 * np.pi)), axis=-1).reshape(-1, 2)
    return vertices, faces, uv01


def vertex_normals(vertices: np.ndarray, faces: np.ndarray) -> np.ndarray:
    tri = vertices[faces]
    fn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    vn = np.zeros_like(vertices)
    for k in range(3):
        np.add.at(vn, faces[:, k], fn)
    return vn / np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-30)


def rotation_yx(angle_y: float, angle_x: float) -> np.ndarray:
    cy, sy, cx, sx = np.cos(angle_y), np.sin(angle_y), np.cos(angle_x), np.sin(angle_x)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return rx @ ry


def project_pinhole(vertices: np.ndarray, rot: np.ndarray, width: int, height: int, fov_deg: float = 60.0):
    """Pinhole camera looking at the centroid; returns ``ij[V,2]`` (col 0 = x/column, col 1 = y/row) and depths."""
    centre = vertices.mean(axis=0)
    radius = np.max(np.linalg.norm(vertices - centre, axis=1))
    focal = 0.5 * min(width, height) / np.tan(np.deg2rad(fov_deg) / 2)
    dist = 1.02 * radius / np.sin(np.deg2rad(fov_deg) / 2)
    cam = (vertices - centre) @ rot.T + np.array([0.0, 0.0, dist])
    ij = np.stack((focal * cam[:, 0] / cam[:, 2] + width / 2.0, focal * cam[:, 1] / cam[:, 2] + height / 2.0), axis=1)
    return np.ascontiguousarray(ij), np.ascontiguousarray(cam[:, 2]), cam


def face_visible_2d(ij: np.ndarray, faces: np.ndarray, clockwise: bool) -> np.ndarray:
    tri = ij[faces]
    u = tri[:, 1] - tri[:, 0]
    v = tri[:, 2] - tri[:, 0]
    cross = u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]
    return cross > 0 if clockwise else cross < 0


def silhouette_edgeflags(ij: np.ndarray, faces: np.ndarray, clockwise: bool) -> np.ndarray:
    """edgeflags[f, n] is True iff exactly one of the faces sharing edge n of face f is visible.

    Edge n of a face joins vertices (n, n+1 mod 3), i.e. the pairs the C core addresses as
    (1,0), (2,1), (0,2) (DifferentiableRenderer.h:2822); semantics of deodr/triangulated_mesh.py:153-166.
    """
    nb_faces = faces.shape[0]
    visible = face_visible_2d(ij, faces, clockwise)
    e = np.concatenate((faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]), axis=0).astype(np.int64)
    nb_v = int(faces.max()) + 1
    key = np.minimum(e[:, 0], e[:, 1]) * nb_v + np.maximum(e[:, 0], e[:, 1])
    _, inv = np.unique(key, return_inverse=True)
    face_of = np.tile(np.arange(nb_faces), 3)
    nb_visible = np.zeros(inv.max() + 1, dtype=np.int64)
    np.add.at(nb_visible, inv, visible[face_of].astype(np.int64))
    flags = (nb_visible[inv] == 1).reshape(3, nb_faces).T
    return np.ascontiguousarray(flags)


def torus_scene(
    n: int,
    width: int,
    height: int,
    view: int = 0,
    n_views: int = 1,
    textured: bool = False,
    nb_colors: int = 3,
    texture_size: int = 512,
    seed: int = 0,
) -> SceneArrays:
    """S-mesh(n) rendered from view ``view`` of ``n_views`` (rotation about y, then an x-tilt of 0.3 rad)."""
    rng = np.random.default_rng(seed)
    vertices, faces, uv01 = torus_mesh(n)
    base_colors = rng.random((vertices.shape[0], 3))
    texture = rng.random((texture_size, texture_size, 3)) if textured else np.zeros((2, 2, nb_colors))
    rot = rotation_yx(2 * np.pi * view / max(n_views, 1) + 0.4, 0.75)
    ij, depths, _ = project_pinhole(vertices, rot, width, height)
    normals = vertex_normals(vertices, faces) @ rot.T
    light = np.array([0.3, -0.3, 0.0])
    luminosity = np.maximum(-(normals @ light), 0.0) + 0.3
    # orientation: pick the winding flag for which outward-facing triangles are the visible ones
    cam_normals_face = np.cross(
        (vertices[faces[:, 1]] - vertices[faces[:, 0]]) @ rot.T, (vertices[faces[:, 2]] - vertices[faces[:, 0]]) @ rot.T
    )
    towards_camera = cam_normals_face[:, 2] < 0
    clockwise = bool(np.mean(face_visible_2d(ij, faces, True) == towards_camera) > 0.5)
    if nb_colors == 3:
        colors = base_colors * luminosity[:, None]
    else:
        colors = np.repeat(depths[:, None], nb_colors, axis=1) if nb_colors == 1 else rng.random((len(depths), nb_colors))
    nb_faces = faces.shape[0]
    scene = SceneArrays(
        faces=faces,
        faces_uv=faces.copy(),
        ij=ij,
        depths=depths,
        textured=np.full(nb_faces, textured, dtype=bool),
        uv=np.ascontiguousarray(uv01 * (texture_size - 1)) if textured else np.zeros((vertices.shape[0], 2)),
        shade=np.ascontiguousarray(luminosity),
        colors=np.ascontiguousarray(colors),
        shaded=np.full(nb_faces, textured, dtype=bool),
        edgeflags=silhouette_edgeflags(ij, faces, clockwise),
        height=height,
        width=width,
        nb_colors=nb_colors,
        texture=texture,
        background_image=None,
        background_color=np.full(nb_colors, 0.8),
        clockwise=clockwise,
        backface_culling=True,
        strict_edge=True,
        perspective_correct=False,
        integer_pixel_centers=True,
    )
    scene.extra["vertices"] = vertices
    return scene


def dense_image_b(image: np.ndarray, seed: int = 1) -> np.ndarray:
    """``image_b = 2 (image - obs)`` with ``obs`` uniform noise: a dense worst-case loss gradient."""
    obs = np.random.default_rng(seed).random(image.shape)
    return 2.0 * (image - obs)
