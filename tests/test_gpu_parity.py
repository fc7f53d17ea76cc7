"""GPU (-m gpu): parity of the sm_100a path, called through the C-ABI, against the CPU oracle and the golden vectors.

Stated tolerances (north star: z-buffer and face-id bit-exact; image and gradients within fp32 tolerance):
  z-buffer   bit-exact (np.array_equal)
  face id    exact (== rint of the reference's interpolated face-id channel)
  image      |err| <= 1e-6            (colours are interpolated in fp32; geometry in fp64)
  gradients  |err| <= 5e-5 * max|grad| + 1e-6   (fp32 atomics, non-deterministic summation order)
             and ||err||_2 <= 2e-5 * ||grad||_2                         (no vertex population is systematically off)
             and |err| <= 2e-3 * |grad| on every element above 1e-3 * max|grad|   (no single vertex is badly off)
"""
import hashlib
import os

import numpy as np
import pytest
from conftest import GOLDEN, SMALL_TAGS, load_small

from deodr_b200.scenes import confetti_scene, dense_image_b, soup_scene, torus_scene

pytestmark = pytest.mark.gpu

IMAGE_TOL = 1e-6
GRAD_RTOL = 5e-5

sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731


@pytest.fixture(scope="module")
def gpu(build_native):
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from deodr_b200.renderer import Renderer

    return Renderer(0)


def run_device(gpu, scene, sigma, image_b=None):
    import torch

    from deodr_b200.renderer import DeviceScene

    ds = DeviceScene(scene, "cuda:0")
    fwd = gpu.render(ds, sigma, face_id=True)
    out = {k: v.cpu().numpy() for k, v in fwd.items() if hasattr(v, "cpu")}
    if image_b is not None:
        grads = gpu.render_b(ds, sigma, fwd, torch.from_numpy(np.ascontiguousarray(image_b)).cuda())
        out.update({k: v.cpu().numpy() for k, v in grads.items()})
    return out


def assert_gradient_close(got, ref, name, rtol=GRAD_RTOL):
    """The three-part gradient criterion of the module docstring."""
    scale = np.abs(ref).max()
    err = np.abs(got - ref)
    assert err.max() <= rtol * scale + 1e-6, (name, "max-norm", err.max(), scale)
    norm = np.linalg.norm(ref.ravel())
    assert np.linalg.norm((got - ref).ravel()) <= 2e-5 * norm + 1e-6, (name, "L2", np.linalg.norm((got - ref).ravel()), norm)
    big = np.abs(ref) > 1e-3 * scale
    if big.any():
        rel = (err[big] / np.abs(ref[big])).max()
        assert rel <= 2e-3, (name, "per-element", rel)


def check(gpu, checker, scene, sigma):
    image, z = checker.render(scene, sigma)
    differentiable = scene.backface_culling and not scene.perspective_correct
    image_b = dense_image_b(image) if differentiable else None
    got = run_device(gpu, scene, sigma, image_b)
    assert np.array_equal(got["z_buffer"], z), "z-buffer not bit-exact"
    assert np.abs(got["image"] - image).max() <= IMAGE_TOL
    assert np.array_equal(got["face_id"] >= 0, np.isfinite(z))
    if differentiable:
        ref = checker.render_b(scene, sigma, image, z, image_b)
        for name in ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b"):
            if ref[name].size == 0:
                continue
            assert_gradient_close(got[name], ref[name], name)
    return got


@pytest.mark.parametrize("tag", SMALL_TAGS)
def test_small_golden(tag, gpu, checker):
    scene, d = load_small(tag)
    got = check(gpu, checker, scene, float(d["sigma"]))
    assert np.array_equal(got["z_buffer"], d["z"])
    assert np.abs(got["image"] - d["image"]).max() <= IMAGE_TOL
    if "ij_b" in d:
        for name in ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b"):
            tol = GRAD_RTOL * np.abs(d[name]).max() + 1e-6
            assert np.abs(got[name] - d[name]).max() <= tol, name


def test_pinned_reference_zbuffer_hash(gpu, texture):
    """The z-buffer SHA-256 pinned by the reference's tests/test_render_mesh.py:70-73 comes out of the GPU."""
    np.random.seed(2)
    scene = soup_scene(clockwise=True, texture=texture)
    got = run_device(gpu, scene, 1.0)
    assert sha(got["z_buffer"]) == "b6f87e03c60bd820efa09d0536495b25d5852f67ecbecd2622f8bf1910d6052a"
    g = np.load(os.path.join(GOLDEN, "soup_pinned.npz"))
    assert np.abs(got["image"][::4, ::4] - g["image_sub"]).max() <= IMAGE_TOL


@pytest.mark.parametrize("clockwise", [False, True])
def test_soup_all_flags(clockwise, gpu, checker, texture):
    np.random.seed(2)
    scene = soup_scene(clockwise=clockwise, texture=texture)
    for sigma in (1.0, 0.0, 3.3):
        check(gpu, checker, scene, sigma)
    scene.strict_edge = False
    scene.integer_pixel_centers = False
    check(gpu, checker, scene, 2.5)
    scene.perspective_correct = True
    check(gpu, checker, scene, 1.0)
    scene.perspective_correct = False
    scene.backface_culling = False
    check(gpu, checker, scene, 1.0)


def test_more_primitives_than_one_shared_memory_chunk(gpu, checker, texture):
    np.random.seed(7)
    scene = soup_scene(n_tri=150, width=40, height=36, texture=texture[::4, ::4].copy(), min_det=100)
    check(gpu, checker, scene, 1.0)
    np.random.seed(8)
    scene = soup_scene(n_tri=2000, width=256, height=256, texture=texture, min_det=300)
    check(gpu, checker, scene, 1.0)


def test_micro_triangle_pile(gpu, checker):
    """Several record chunks per tile and more candidates per pixel than the per-pixel lists of the z pass hold."""
    check(gpu, checker, confetti_scene(3000, 64, 48, size=2.5, seed=1), 1.0)
    check(gpu, checker, confetti_scene(1500, 50, 40, size=1.2, seed=2, edge_ratio=0.3), 0.7)
    check(gpu, checker, confetti_scene(20000, 300, 200, size=2.0, seed=3), 1.0)


def test_exact_z_ties(gpu, checker, texture):
    np.random.seed(4)
    scene = soup_scene(n_tri=6, width=48, height=48, textured_ratio=0.0, texture=texture[::4, ::4].copy(), min_det=100)
    n_v = scene.depths.shape[0]
    dup = lambda a: np.concatenate((a, a), axis=0)  # noqa: E731
    scene.faces = np.concatenate((scene.faces, scene.faces + n_v)).astype(np.uint32)
    scene.faces_uv = scene.faces.copy()
    for name in ("ij", "depths", "uv", "shade", "colors", "textured", "shaded", "edgeflags"):
        setattr(scene, name, dup(getattr(scene, name)))
    scene.colors[n_v:] *= 0.5
    got = check(gpu, checker, scene, 1.0)
    assert (got["owner"] <= -2).any(), "tie pixels should be recorded in the side table"


def test_face_id_matches_reference_deferred_channel(gpu, ref_oracle, texture):
    """Reference face ids only exist as an interpolated colour channel (Scene3D.render_deferred,
    deodr/differentiable_renderer.py:1098-1100, sigma = 0): rint(channel) must equal our int32 face id."""
    np.random.seed(3)
    scene = soup_scene(n_tri=60, width=160, height=120, textured_ratio=0.0, texture=texture, min_det=500)
    scene.nb_colors = 1
    scene.colors = np.repeat(np.arange(60, dtype=np.float64), 3)[:, None]
    scene.background_image = None
    scene.background_color = np.array([-1.0])
    scene.texture = np.zeros((2, 2, 1))
    channel, z = ref_oracle.render(scene, 0.0)
    got = run_device(gpu, scene, 0.0)
    assert np.array_equal(got["z_buffer"], z)
    assert np.array_equal(got["face_id"], np.rint(channel[:, :, 0]).astype(np.int32))


def test_meshes(gpu, checker):
    check(gpu, checker, torus_scene(24, 160, 120), 1.0)
    check(gpu, checker, torus_scene(40, 250, 200, textured=True, texture_size=64), 1.0)
    check(gpu, checker, torus_scene(100, 512, 512, nb_colors=1), 1.0)
    check(gpu, checker, torus_scene(64, 300, 200, nb_colors=4), 1.0)
    check(gpu, checker, torus_scene(40, 160, 120, nb_colors=2), 1.0)  # 2 channels run in the <4> instance
    check(gpu, checker, torus_scene(48, 200, 200, nb_colors=7), 0.7)


def test_config3_textured_mesh_1024(gpu, checker):
    """BASELINE.json configs[2]: 50k-triangle textured mesh, 1024x1024, bilinear UV + edge overdraw, fwd+bwd."""
    check(gpu, checker, torus_scene(158, 1024, 1024, textured=True, texture_size=512), 1.0)


def test_config5_full_size_1m_triangles_2048(gpu, checker):
    """BASELINE.json configs[4] at full size, against the oracle (a few seconds of CPU) and through size-independent
    properties: re-rendering is idempotent, a zero image_b gives zero gradients, the adjoint is linear in image_b."""
    import torch

    from deodr_b200.renderer import DeviceScene

    scene = torus_scene(708, 2048, 2048)
    assert scene.faces.shape[0] == 1002528
    got = check(gpu, checker, scene, 1.0)
    ds = DeviceScene(scene, "cuda:0")
    fwd = gpu.render(ds, 1.0)
    fwd2 = gpu.render(ds, 1.0)
    assert torch.equal(fwd["z_buffer"], fwd2["z_buffer"]) and torch.equal(fwd["image"], fwd2["image"])
    assert np.array_equal(fwd["z_buffer"].cpu().numpy(), got["z_buffer"])
    zero = gpu.render_b(ds, 1.0, fwd2, torch.zeros_like(fwd["image"]))
    assert all(float(v.abs().max()) == 0.0 for v in zero.values() if v.numel())
    a = torch.rand_like(fwd["image"])
    b = torch.rand_like(fwd["image"])
    ga, gb = gpu.render_b(ds, 1.0, fwd2, a), gpu.render_b(ds, 1.0, fwd2, b)
    gab = gpu.render_b(ds, 1.0, fwd2, 2 * a - 3 * b)
    for name in ("ij_b", "colors_b"):
        lin = 2 * ga[name] - 3 * gb[name]
        assert float((gab[name] - lin).abs().max()) <= 1e-4 * float(lin.abs().max()) + 1e-5


def test_empty_and_degenerate_inputs(gpu, checker, texture):
    np.random.seed(2)
    scene = soup_scene(n_tri=4, width=33, height=17, texture=texture[::4, ::4].copy(), min_det=20)
    # no triangles at all
    empty = soup_scene.__globals__["SceneArrays"](
        faces=np.zeros((0, 3), np.uint32), faces_uv=np.zeros((0, 3), np.uint32), ij=scene.ij, depths=scene.depths,
        textured=np.zeros(0, bool), uv=scene.uv, shade=scene.shade, colors=scene.colors, shaded=np.zeros(0, bool),
        edgeflags=np.zeros((0, 3), bool), height=17, width=33, nb_colors=3, texture=scene.texture,
        background_color=np.array([0.1, 0.2, 0.3]))
    got = run_device(gpu, empty, 1.0, np.ones((17, 33, 3)))
    assert np.isinf(got["z_buffer"]).all() and np.allclose(got["image"], [0.1, 0.2, 0.3])
    assert not got["ij_b"].any()
    # everything behind the camera / off screen / zero area
    scene.depths = -np.abs(scene.depths)
    check(gpu, checker, scene, 1.0)
    scene.depths = np.abs(scene.depths)
    scene.ij = scene.ij + 1000.0
    check(gpu, checker, scene, 1.0)
    scene.ij = np.zeros_like(scene.ij)
    scene.backface_culling = False
    check(gpu, checker, scene, 0.0)


def test_error_behaviour(gpu, texture):
    import torch

    from deodr_b200 import _cabi
    from deodr_b200.renderer import DeviceScene

    np.random.seed(2)
    scene = soup_scene(texture=texture)
    scene.backface_culling = False
    ds = DeviceScene(scene, "cuda:0")
    fwd = gpu.render(ds, 1.0)
    with pytest.raises(_cabi.DeodrB200Error, match="backface_culling"):
        gpu.render_b(ds, 1.0, fwd, torch.zeros_like(fwd["image"]))
    scene.backface_culling, scene.perspective_correct = True, True
    ds = DeviceScene(scene, "cuda:0")
    fwd = gpu.render(ds, 1.0)
    with pytest.raises(_cabi.DeodrB200Error, match="perspective_correct"):
        gpu.render_b(ds, 1.0, fwd, torch.zeros_like(fwd["image"]))
    scene.perspective_correct = False
    scene.faces = scene.faces.copy()
    scene.faces[0, 0] = 10**6
    with pytest.raises(_cabi.DeodrB200Error, match="faces"):
        gpu.check_scene(DeviceScene(scene, "cuda:0"))


def test_config1_soup_128(gpu, checker, texture):
    """BASELINE.json configs[0] at its literal size: 30-triangle soup, 128x128 RGB, fwd+bwd (the reference pins this
    scene at 200x200 - covered by test_soup_all_flags / the pinned hashes; no reference vector exists at 128, so the
    checker is the compiled reference itself)."""
    for seed, clockwise in ((2, False), (3, True)):
        np.random.seed(seed)
        check(gpu, checker, soup_scene(n_tri=30, width=128, height=128, clockwise=clockwise, texture=texture), 1.0)


def test_config2_hand_mesh_640x480(gpu, checker, ref_oracle):
    """BASELINE.json configs[1] on the REAL hand mesh: deodr/data/hand.obj (1048 faces) projected by the reference's own
    Camera and lit by its Scene3D (tests/golden/scene_ops.npz, made by make_scene_ops_golden.py from the reference's
    Python), 640x480 RGB Gouraud, fwd+bwd against the compiled reference core; face ids against the reference's
    deferred face-id channel (Scene3D.render_deferred, sigma = 0)."""
    from deodr_b200.scenes import SceneArrays

    g = np.load(os.path.join(GOLDEN, "scene_ops.npz"))
    faces, V = g["hand_faces"], g["hand_vertices"].shape[0]
    T = faces.shape[0]

    def scene_with(colors, nb_colors, background):
        return SceneArrays(
            faces=faces, faces_uv=np.zeros((T, 3), np.uint32), ij=g["c2_ij"], depths=g["c2_depths"],
            textured=np.zeros(T, bool), uv=np.zeros((1, 2)), shade=np.zeros(V), colors=colors, shaded=np.zeros(T, bool),
            edgeflags=g["c2_edgeflags"].astype(bool), height=480, width=640, nb_colors=nb_colors,
            texture=np.zeros((2, 2, nb_colors)), background_image=None, background_color=background, clockwise=False,
            backface_culling=True, strict_edge=True, perspective_correct=False, integer_pixel_centers=True)

    scene = scene_with(g["c2_colors"], 3, g["c2_background_color"])
    got = check(gpu, checker, scene, 1.0)
    assert np.isfinite(got["z_buffer"]).mean() > 0.05
    # face ids: the reference interpolates a per-vertex channel, so give every face its own three vertices
    soup_faces = np.arange(3 * T, dtype=np.uint32).reshape(T, 3)
    flat = faces.astype(np.int64).reshape(-1)
    deferred = SceneArrays(
        faces=soup_faces, faces_uv=np.zeros((T, 3), np.uint32), ij=g["c2_ij"][flat], depths=g["c2_depths"][flat],
        textured=np.zeros(T, bool), uv=np.zeros((1, 2)), shade=np.zeros(3 * T),
        colors=np.repeat(np.arange(T, dtype=np.float64), 3)[:, None], shaded=np.zeros(T, bool),
        edgeflags=np.zeros((T, 3), bool), height=480, width=640, nb_colors=1, texture=np.zeros((2, 2, 1)),
        background_image=None, background_color=np.array([-1.0]), clockwise=False, backface_culling=True,
        strict_edge=True, perspective_correct=False, integer_pixel_centers=True)
    channel, z = ref_oracle.render(deferred, 0.0)
    out = run_device(gpu, deferred, 0.0)
    assert np.array_equal(out["z_buffer"], z) and np.array_equal(z, got["z_buffer"])
    assert np.array_equal(out["face_id"], np.rint(channel[:, :, 0]).astype(np.int32))
    assert np.array_equal(got["face_id"], out["face_id"])  # the shared-vertex mesh owns the same faces


def test_face_ids_at_config3_size(gpu, ref_oracle):
    """Face ids against the reference's deferred face-id channel on the 50k-triangle mesh at 1024x1024 (sigma = 0)."""
    scene = torus_scene(158, 1024, 1024)
    T = scene.faces.shape[0]
    flat = scene.faces.astype(np.int64).reshape(-1)
    scene.ij, scene.depths, scene.shade = scene.ij[flat], scene.depths[flat], np.zeros(3 * T)
    scene.faces = np.arange(3 * T, dtype=np.uint32).reshape(T, 3)
    scene.faces_uv = np.zeros((T, 3), np.uint32)
    scene.uv = np.zeros((1, 2))
    scene.nb_colors, scene.colors = 1, np.repeat(np.arange(T, dtype=np.float64), 3)[:, None]
    scene.background_image, scene.background_color = None, np.array([-1.0])
    scene.texture = np.zeros((2, 2, 1))
    scene.edgeflags = np.zeros((T, 3), bool)
    channel, z = ref_oracle.render(scene, 0.0)
    out = run_device(gpu, scene, 0.0)
    assert np.array_equal(out["z_buffer"], z)
    assert np.array_equal(out["face_id"], np.rint(channel[:, :, 0]).astype(np.int32))


def test_record_parallel_small_adjoint(checker, texture, monkeypatch):
    """DEODR_B200_SMALL_ADJOINT=record (read when a workspace is created): k_small_rec_bwd - one thread per pre-masked
    record, the tile's owner / image_b blocks staged by TMA tile loads - gives the gradients of the default
    triangle-parallel kernel."""
    from deodr_b200.renderer import Renderer

    monkeypatch.setenv("DEODR_B200_SMALL_ADJOINT", "record")
    gpu = Renderer(0)
    monkeypatch.delenv("DEODR_B200_SMALL_ADJOINT")
    check(gpu, checker, confetti_scene(3000, 64, 48, size=2.5, seed=1), 1.0)
    check(gpu, checker, torus_scene(60, 300, 300, nb_colors=1), 1.0)
    check(gpu, checker, torus_scene(100, 512, 512), 1.0)           # TMA tile loads (C = 3, W % 4 == 0)
    check(gpu, checker, torus_scene(24, 150, 130), 1.0)             # cooperative loads (row pitch not 16-byte)
    check(gpu, checker, torus_scene(400, 512, 512, textured=True), 1.0)  # textured small triangles (T >= 262144)
