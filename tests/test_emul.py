"""CPU: the kernel phases (deodr_b200/csrc/phases.h) emulated on the host vs the oracle.

The same __host__ __device__ functions run inside the CUDA kernels; here each CTA is executed as per-phase loops over
the thread index (tests/emul/emul.cpp).  z-buffer: bit-exact.  Image: |err| <= 1e-6 (fp32 colours).  Gradients:
|err| <= 2e-5 * max|grad| + 1e-6 (fp32 accumulation).
"""
import numpy as np
import pytest
from canon import Emulator
from conftest import SMALL_TAGS, load_small

from deodr_b200.scenes import confetti_scene, dense_image_b, soup_scene, torus_scene

IMAGE_TOL = 1e-6
GRAD_RTOL = 2e-5


@pytest.fixture(scope="module")
def emulator():
    return Emulator()


def check(emulator, checker, scene, sigma, image_tol=IMAGE_TOL):
    image, z = checker.render(scene, sigma)
    fwd = emulator.render(scene, sigma)
    assert np.array_equal(fwd["z"], z), "z-buffer not bit-exact"
    assert np.abs(fwd["image"] - image).max() <= image_tol
    covered = np.isfinite(z)
    assert np.array_equal(fwd["face_id"] >= 0, covered)
    if scene.backface_culling and not scene.perspective_correct:
        image_b = dense_image_b(image)
        ref = checker.render_b(scene, sigma, image, z, image_b)
        got = emulator.render_b(scene, sigma, fwd, image_b)
        for name in ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b"):
            if ref[name].size == 0:
                continue
            tol = GRAD_RTOL * np.abs(ref[name]).max() + 1e-6
            assert np.abs(got[name] - ref[name]).max() <= tol, name
    return fwd


@pytest.mark.parametrize("tag", SMALL_TAGS)
def test_emulated_kernels_small_golden(tag, emulator, checker):
    scene, d = load_small(tag)
    fwd = check(emulator, checker, scene, float(d["sigma"]))
    assert np.array_equal(fwd["z"], d["z"])


@pytest.mark.parametrize("clockwise", [False, True])
def test_emulated_kernels_soup(clockwise, emulator, checker, texture):
    np.random.seed(2)
    scene = soup_scene(clockwise=clockwise, texture=texture)
    check(emulator, checker, scene, 1.0)
    scene.strict_edge = False
    scene.integer_pixel_centers = False
    check(emulator, checker, scene, 2.5)
    scene.backface_culling = False
    check(emulator, checker, scene, 1.0)


def test_emulated_kernels_many_edges_per_tile(emulator, checker, texture):
    """More than EDGE_CHUNK (64) silhouette edges and more than TRI_CHUNK (128) triangles in one tile."""
    np.random.seed(7)
    scene = soup_scene(n_tri=150, width=40, height=36, texture=texture[::4, ::4].copy(), min_det=100)
    fwd = check(emulator, checker, scene, 1.0)
    assert fwd["edges"] == 450


def test_emulated_kernels_micro_triangle_pile(emulator, checker):
    """Thousands of few-pixel triangles on a small image: several 128-record chunks per tile, pixels with more
    candidates than the per-pixel lists hold (scan fall-back), small triangles carrying silhouette edges."""
    check(emulator, checker, confetti_scene(3000, 64, 48, size=2.5, seed=1), 1.0)
    check(emulator, checker, confetti_scene(1500, 50, 40, size=1.2, seed=2, edge_ratio=0.3), 0.7)


@pytest.mark.parametrize("flags", [
    dict(perspective_correct=True), dict(strict_edge=False), dict(integer_pixel_centers=False),
    dict(backface_culling=False), dict(strict_edge=False, integer_pixel_centers=False, perspective_correct=True)])
def test_emulated_kernels_micro_triangle_pile_flags(flags, emulator, checker):
    """The same regime under every rasterisation flag of the reference (the adjoint runs where it is defined).
    A pixel of the pile is overdrawn by dozens of stacked silhouette edges: every fp32 blend adds half an ulp, hence the
    looser image tolerance (the z-buffer stays bit-exact)."""
    scene = confetti_scene(1200, 48, 40, size=2.0, seed=5, edge_ratio=0.1)
    for name, value in flags.items():
        setattr(scene, name, value)
    check(emulator, checker, scene, 1.0, image_tol=5e-6)


def test_emulated_kernels_mesh(emulator, checker):
    check(emulator, checker, torus_scene(24, 160, 120), 1.0)
    emulator.set_record_rows(0)  # what the device picks for scenes of a few thousand triangles
    try:
        check(emulator, checker, torus_scene(24, 160, 120), 1.0)
        check(emulator, checker, torus_scene(30, 100, 90, textured=True, texture_size=32), 1.0)
    finally:
        emulator.set_record_rows(16)
    check(emulator, checker, torus_scene(30, 100, 90, textured=True, texture_size=32), 1.0)
    check(emulator, checker, torus_scene(16, 70, 50, nb_colors=1), 1.0)


def test_record_parallel_form_of_the_small_adjoint(emulator, checker, texture):
    """k_small_rec_bwd (DEODR_B200_SMALL_ADJOINT=record): one thread per pre-masked record of a tile instead of one per
    small triangle; same gradients."""
    emulator.set_small_records(True)
    try:
        check(emulator, checker, confetti_scene(3000, 64, 48, size=2.5, seed=1), 1.0)
        check(emulator, checker, torus_scene(24, 160, 120), 1.0)
        check(emulator, checker, torus_scene(30, 100, 90, textured=True, texture_size=32), 1.0)
        np.random.seed(2)
        check(emulator, checker, soup_scene(texture=texture), 1.0)
        scene = confetti_scene(1200, 48, 40, size=2.0, seed=5, edge_ratio=0.1)
        scene.strict_edge, scene.integer_pixel_centers = False, False
        check(emulator, checker, scene, 1.0, image_tol=5e-6)
    finally:
        emulator.set_small_records(False)


def test_textured_triangles_through_the_pixel_parallel_adjoint(emulator, checker, texture):
    """TriBins::small_textured = 0 (what the device picks below a few hundred thousand triangles): textured triangles
    keep their forward records but are owned without SMALL_FLAG, so their adjoint is the pixel-parallel kernel's."""
    emulator.set_small_textured(False)
    try:
        check(emulator, checker, torus_scene(30, 100, 90, textured=True, texture_size=32), 1.0)
        np.random.seed(2)
        check(emulator, checker, soup_scene(texture=texture), 1.0)  # textured and interpolated triangles mixed
    finally:
        emulator.set_small_textured(True)


def test_exact_z_ties_are_split_like_the_reference(emulator, checker, texture):
    """Duplicated interpolated triangles tie exactly in z: forward keeps the lowest index (strict '<', DR.h:961),
    the adjoint credits the highest (DR.h:1024 in reverse order)."""
    np.random.seed(4)
    scene = soup_scene(n_tri=6, width=48, height=48, textured_ratio=0.0, texture=texture[::4, ::4].copy(), min_det=100)
    dup = lambda a: np.concatenate((a, a), axis=0)  # noqa: E731
    n_v = scene.depths.shape[0]
    scene.faces = np.concatenate((scene.faces, scene.faces + n_v)).astype(np.uint32)
    scene.faces_uv = scene.faces.copy()
    for name in ("ij", "depths", "uv", "shade", "colors"):
        setattr(scene, name, dup(getattr(scene, name)))
    scene.colors[n_v:] *= 0.5  # distinct colours: shows which copy is drawn
    for name in ("textured", "shaded", "edgeflags"):
        setattr(scene, name, dup(getattr(scene, name)))
    fwd = check(emulator, checker, scene, 1.0)
    assert fwd["ties"] > 0


def test_emulated_antialiase_error_mode(emulator, checker, texture):
    """Row f3 ahead of its kernels: the error-mode phases (edges overdraw the squared residual against `obs`) emulated on
    the CPU against the oracle - forward image / z-buffer / err_buffer, adjoint with the reference's defect #2 kept
    (compat) - on soups with textured and interpolated triangles, a mesh and a micro-triangle pile."""
    rng = np.random.default_rng(3)
    np.random.seed(2)
    scenes = [(soup_scene(clockwise=False, texture=texture), 1.0),
              (soup_scene(n_tri=150, width=40, height=36, texture=texture[::4, ::4].copy(), min_det=100), 1.5),
              (torus_scene(24, 160, 120), 1.0), (torus_scene(20, 90, 80, textured=True, texture_size=32), 2.0),
              (confetti_scene(800, 48, 40, size=2.0, seed=4, edge_ratio=0.2), 1.0)]
    for scene, sigma in scenes:
        if scene.textured.any():
            scene.uv = scene.uv * 0.9973 + 0.0131  # keep the texture coordinates off the texel grid
        obs = rng.random((scene.height, scene.width, scene.nb_colors)).astype(np.float32).astype(np.float64)
        image, z, err = checker.render(scene, sigma, antialiase_error=True, obs=obs)
        fwd = emulator.render_error(scene, sigma, obs)
        assert np.array_equal(fwd["z"], z)
        assert np.abs(fwd["image"] - image).max() <= IMAGE_TOL
        assert np.abs(fwd["err"] - err).max() <= 2e-6 * max(1.0, err.max())
        err_b = rng.random((scene.height, scene.width)) * 2 - 1
        ref = checker.render_b(scene, sigma, image, z, None, antialiase_error=True, obs=obs, err_buffer=err,
                               err_buffer_b=err_b)
        got = emulator.render_error_b(scene, sigma, fwd, err_b, compat=True)
        for name in ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b"):
            if ref[name].size == 0:
                continue
            tol = 5e-5 * np.abs(ref[name]).max() + 2e-6
            assert np.abs(got[name] - ref[name]).max() <= tol, (name, np.abs(got[name] - ref[name]).max(), tol)


def test_plan_of_an_earlier_pass_serves_a_drifted_scene(emulator, checker):
    """The forward pass never waits for the host: its per-tile lists live in segments reserved by the PLAN of an earlier
    pass over a scene of the same shape (count pass + slack).  A drifted scene (vertices moved like one optimiser
    step) must either fit - and then be bit-exact like any other pass - or raise the verdict word, never corrupt."""
    base = torus_scene(20, 128, 96)
    emulator.build_plan(base, 1.0)
    rng = np.random.default_rng(0)
    fitted = 0
    for step in (0.05, 0.3, 1.0):
        import copy

        moved = copy.copy(base)
        moved.ij = base.ij + rng.normal(scale=step, size=base.ij.shape)
        moved.edgeflags = base.edgeflags  # topology and flags as the plan saw them
        verdict, out = emulator.render_planned(moved, 1.0)
        if verdict == 0:
            fitted += 1
            image, z = checker.render(moved, 1.0)
            assert np.array_equal(out["z"], z)
            assert np.abs(out["image"] - image).max() <= IMAGE_TOL * max(1.0, np.abs(image).max())
    assert fitted >= 1  # the smallest drift fits into the slack
    # a very different scene on the same plan: overflow is REPORTED (non-zero verdict), the host then re-plans
    crowded = confetti_scene(4000, 128, 96, size=3.0, seed=3)
    emulator.build_plan(torus_scene(4, 128, 96), 1.0)
    verdict, _ = emulator.render_planned(crowded, 1.0)
    assert verdict != 0


@pytest.mark.parametrize("sigma", [1e-3, 0.05, 9.0, 45.0])
def test_emulated_kernels_extreme_sigma(sigma, emulator, checker, texture):
    """Edge bands from a thousandth of a pixel to wider than the scene (every tile holds every silhouette edge: many
    edge chunks per tile, band boxes clipped on all sides): same answers as the reference."""
    np.random.seed(2)
    check(emulator, checker, soup_scene(clockwise=False, texture=texture), sigma)
    check(emulator, checker, torus_scene(24, 160, 120), sigma)
    check(emulator, checker, confetti_scene(800, 96, 80, size=3.0, seed=3, edge_ratio=0.5), sigma)


def far_vertex_scene(seed, far):
    """A confetti scene with 40 of its 600 vertices pushed up to `far` pixels away (scripts/probe_far_vertices.py)."""
    rng = np.random.default_rng(seed)
    scene = confetti_scene(200, 64, 48, size=float(rng.choice([5.0, 40.0])), seed=seed, edge_ratio=0.3)
    idx = rng.choice(scene.ij.shape[0], size=40, replace=False)
    scene.ij[idx] += rng.choice([-1, 1], size=(40, 2)) * far * rng.random((40, 2))
    # (a row that wraps to exactly 32767 starts the reference's `short` row counter at -32768 and makes it write at
    # negative pixel indices, DR.h:925-960: the checker runs in this process, keep clear of it)
    rows = np.floor(scene.ij[:, 1])
    hit = (np.abs(rows) < 2.0 ** 31) & ((rows.astype(np.int64) & 0xFFFF) == 32767)
    scene.ij[hit, 1] += 1.0
    return scene


@pytest.mark.parametrize("far", [1e3, 3.2e4])
def test_vertices_far_outside_the_image_within_the_short_range(far, emulator, checker):
    """Up to the reference's own limit (loop counters and bounds are `short`: +-32767 px) long triangles that cross the
    image from far outside are rasterised bit for bit like the reference, gradients included."""
    for seed in range(6):
        check(emulator, checker, far_vertex_scene(seed, far), 1.0)


@pytest.mark.parametrize("far", [4e4, 7e4, 1e6, 3e9, 1e15])
def test_exact_short_wrap_build_follows_the_reference_beyond_the_short_range(far, checker):
    """Beyond +-32767 px the reference's `(short)` casts wrap (x86) and a triangle with one far vertex is still drawn in
    part.  The build with DEODR_EXACT_SHORT_WRAP=1 (rmath.h; not the default) reproduces
    that bit for bit: z-buffer identical for every magnitude up to 1e15 px (image within tolerance up to 3e9 px)."""
    exact = Emulator(exact_short_wrap=True)
    for seed in range(8):
        scene = far_vertex_scene(seed, far)
        image, z = checker.render(scene, 1.0)
        fwd = exact.render(scene, 1.0)
        assert np.array_equal(fwd["z"], z), (far, seed)
        # colours: measured 1.3e-7 up to 3e9 px; beyond, the interpolation of a triangle with a vertex 1e12 px away
        # is ill-conditioned in any arithmetic (3e-5 at 1e12, 3e-2 at 1e15 between fp32 and fp64 attributes)
        if far <= 3e9:
            assert np.abs(fwd["image"] - image).max() <= IMAGE_TOL


def test_exact_short_wrap_build_changes_nothing_within_the_short_range(checker, texture):
    exact = Emulator(exact_short_wrap=True)
    np.random.seed(2)
    check(exact, checker, soup_scene(clockwise=False, texture=texture), 1.0)
    check(exact, checker, torus_scene(24, 160, 120), 1.0)
    check(exact, checker, torus_scene(30, 100, 90, textured=True, texture_size=32), 2.5)
    check(exact, checker, confetti_scene(3000, 64, 48, size=2.5, seed=1), 1.0)
    for seed in range(4):
        check(exact, checker, far_vertex_scene(seed, 3.2e4), 1.0)


@pytest.mark.parametrize("workload", ["c3", "c5"])
def test_emulated_kernels_at_the_full_benchmark_sizes(workload, emulator, checker):
    """BASELINE.json configs[2] (50k textured triangles, 1024^2) and configs[4] (1M triangles, 2048^2) through the CPU
    emulation of the kernel phases: z-buffer bit-exact, image within 1e-6, every gradient within the three-part
    criterion of the GPU parity tests (tests/test_gpu_parity.py::assert_gradient_close)."""
    import sys

    from conftest import ROOT

    sys.path.insert(0, ROOT)
    import bench

    scene = bench.build_scene(workload, view=0, n_views=1)
    image, z = checker.render(scene, 1.0)
    fwd = emulator.render(scene, 1.0)
    assert np.array_equal(fwd["z"], z)
    assert np.abs(fwd["image"] - image).max() <= IMAGE_TOL
    image_b = dense_image_b(image)
    ref, got = checker.render_b(scene, 1.0, image, z, image_b), emulator.render_b(scene, 1.0, fwd, image_b)
    for name in ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b"):
        if ref[name].size == 0 or not np.abs(ref[name]).max() > 0:
            continue
        err, scale = np.abs(got[name] - ref[name]), np.abs(ref[name]).max()
        assert err.max() <= 5e-5 * scale + 1e-6, name
        assert np.linalg.norm(err.ravel()) <= 2e-5 * np.linalg.norm(ref[name].ravel()) + 1e-6, name
        big = np.abs(ref[name]) > 1e-3 * scale
        assert (err[big] / np.abs(ref[name][big])).max() <= 2e-3, name


def test_emulated_kernels_on_views_of_config4(emulator, checker):
    """BASELINE.json configs[3]: views of the 200k-triangle mesh at 512^2, an RGB and a depth (C = 1) render each -
    eight of the 64 views here (all 64 run on the GPU, tests/test_gpu_views.py)."""
    import sys

    from conftest import ROOT

    sys.path.insert(0, ROOT)
    import bench

    for view in range(0, 64, 8):
        rgb = bench.build_scene("c4", view=view, n_views=64)
        depth = bench.build_scene("c4", view=view, n_views=64)
        depth.nb_colors, depth.colors = 1, np.ascontiguousarray(depth.depths[:, None])
        depth.background_color, depth.texture = np.array([float(depth.depths.max())]), np.zeros((2, 2, 1))
        for scene in (rgb, depth):
            image, z = checker.render(scene, 1.0)
            fwd = emulator.render(scene, 1.0)
            assert np.array_equal(fwd["z"], z), view
            assert np.abs(fwd["image"] - image).max() <= IMAGE_TOL * max(1.0, np.abs(image).max())
            image_b = dense_image_b(image)
            ref, got = checker.render_b(scene, 1.0, image, z, image_b), emulator.render_b(scene, 1.0, fwd, image_b)
            for name in ("ij_b", "colors_b"):
                assert np.abs(got[name] - ref[name]).max() <= GRAD_RTOL * np.abs(ref[name]).max() + 1e-6, (view, name)


def test_emulated_kernels_on_configs_1_and_2(emulator, checker, ref_oracle, texture):
    """BASELINE.json configs[0] at its literal size (30-triangle soup, 128 x 128) and configs[1] on the REAL hand mesh
    (deodr/data/hand.obj projected and lit by the reference's own Camera / Scene3D: tests/golden/scene_ops.npz), 640 x 480;
    face ids against the reference's deferred face-id channel."""
    import os

    from conftest import GOLDEN

    from deodr_b200.scenes import SceneArrays

    for seed, clockwise in ((2, False), (3, True)):
        np.random.seed(seed)
        check(emulator, checker, soup_scene(n_tri=30, width=128, height=128, clockwise=clockwise, texture=texture), 1.0)
    g = np.load(os.path.join(GOLDEN, "scene_ops.npz"))
    faces, V = g["hand_faces"], g["hand_vertices"].shape[0]
    T = faces.shape[0]
    hand = SceneArrays(
        faces=faces, faces_uv=np.zeros((T, 3), np.uint32), ij=g["c2_ij"], depths=g["c2_depths"],
        textured=np.zeros(T, bool), uv=np.zeros((1, 2)), shade=np.zeros(V), colors=g["c2_colors"],
        shaded=np.zeros(T, bool), edgeflags=g["c2_edgeflags"].astype(bool), height=480, width=640, nb_colors=3,
        texture=np.zeros((2, 2, 3)), background_image=None, background_color=g["c2_background_color"], clockwise=False,
        backface_culling=True, strict_edge=True, perspective_correct=False, integer_pixel_centers=True)
    fwd = check(emulator, checker, hand, 1.0)
    assert np.isfinite(fwd["z"]).mean() > 0.05
    flat = faces.astype(np.int64).reshape(-1)  # every face its own three vertices: a per-vertex channel = the face id
    deferred = SceneArrays(
        faces=np.arange(3 * T, dtype=np.uint32).reshape(T, 3), faces_uv=np.zeros((T, 3), np.uint32),
        ij=g["c2_ij"][flat], depths=g["c2_depths"][flat], textured=np.zeros(T, bool), uv=np.zeros((1, 2)),
        shade=np.zeros(3 * T), colors=np.repeat(np.arange(T, dtype=np.float64), 3)[:, None], shaded=np.zeros(T, bool),
        edgeflags=np.zeros((T, 3), bool), height=480, width=640, nb_colors=1, texture=np.zeros((2, 2, 1)),
        background_image=None, background_color=np.array([-1.0]), clockwise=False, backface_culling=True,
        strict_edge=True, perspective_correct=False, integer_pixel_centers=True)
    channel, z = ref_oracle.render(deferred, 0.0)
    assert np.array_equal(fwd["z"], z)
    assert np.array_equal(fwd["face_id"], np.rint(channel[:, :, 0]).astype(np.int32))


def test_emulated_antialiase_error_mode_at_config3_size(emulator, checker):
    """antialiase_error = True (row f3) on the 50k-triangle textured mesh at 1024^2: err_buffer and the (bug-compatible)
    adjoint of the emulated phases against the compiled reference."""
    import sys

    from conftest import ROOT

    sys.path.insert(0, ROOT)
    import bench

    scene = bench.build_scene("c3", view=0, n_views=1)
    rng = np.random.default_rng(0)
    obs = rng.random((scene.height, scene.width, scene.nb_colors)).astype(np.float32).astype(np.float64)
    err_b = rng.random((scene.height, scene.width)) * 2 - 1
    image, z, err = checker.render(scene, 1.0, antialiase_error=True, obs=obs)
    fwd = emulator.render_error(scene, 1.0, obs)
    assert np.array_equal(fwd["z"], z)
    assert np.abs(fwd["err"] - err).max() <= 2e-6 * max(1.0, err.max())
    ref = checker.render_b(scene, 1.0, image, z, None, antialiase_error=True, obs=obs, err_buffer=err, err_buffer_b=err_b)
    got = emulator.render_error_b(scene, 1.0, fwd, err_b, compat=True)
    for name in ("ij_b", "uv_b", "shade_b", "texture_b"):
        assert np.abs(got[name] - ref[name]).max() <= GRAD_RTOL * np.abs(ref[name]).max() + 1e-6, name
