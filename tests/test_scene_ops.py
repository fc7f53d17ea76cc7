"""Rows f1 / f2: the steps either side of the raster (deodr_b200/csrc/scene_ops.cu, deodr_b200/mesh_ops.py) against
fixtures generated from the REFERENCE's own Python (tests/golden/make_scene_ops_golden.py -> tests/golden/scene_ops.npz).

Tolerances: fp64 outputs 1e-12 relative (the reference is numpy fp64; only the summation order differs), fp32 outputs
(colours / luminosity: the rasteriser's attribute type) 1e-6, silhouette flags bit-exact.
"""
import os

import numpy as np
import pytest
from conftest import GOLDEN

from deodr_b200.mesh_ops import topology_arrays


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "scene_ops.npz"))


def _flags_from_topology(topo, ij, clockwise):
    """numpy walk of the adjacency arrays exactly as k_face_visible / k_edge_flags do it."""
    faces = topo["faces"].view(np.uint32).astype(np.int64)
    tri = ij[faces]
    u, v = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    cross = u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]
    visible = (cross > 0 if clockwise else cross < 0).astype(np.int64)
    per_edge = np.add.reduceat(visible[topo["edge_face_index"]], topo["edge_face_offset"][:-1])
    return (per_edge[topo["faces_edges"]] == 1).astype(np.uint8)


def test_topology_arrays_reproduce_the_reference_silhouette_flags(gold):
    for mesh in ("hand", "torus"):
        topo = topology_arrays(gold[f"{mesh}_faces"], gold[f"{mesh}_vertices"].shape[0])
        assert topo["edge_face_offset"][-1] == 3 * topo["nb_faces"] == topo["vertex_face_offset"][-1]
        for ij, flags in zip(gold[f"{mesh}_views_ij"], gold[f"{mesh}_views_edgeflags"]):
            assert np.array_equal(_flags_from_topology(topo, ij, False), flags)
            assert flags.any()


def test_reference_camera_adjoint_is_wrong_for_a_general_rotation(gold):
    """Defect #3: the reference multiplies by R^T where the adjoint of world_to_camera needs R."""
    a, b = gold["cam_plain_points_b_reference"], gold["cam_plain_points_b_adjoint"]
    assert np.abs(a - b).max() > 1e-3 * np.abs(b).max()


# ---------------------------------------------------------------------------------------------------------- GPU


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["plain", "dist"])
def test_project_points_and_backward(tag, gold, build_native):
    import torch

    from deodr_b200.mesh_ops import CameraParams, project_points, project_points_backward

    cam = CameraParams(gold[f"cam_{tag}_extrinsic"], gold[f"cam_{tag}_intrinsic"],
                       gold[f"cam_{tag}_distortion"] if tag == "dist" else None)
    pts = torch.from_numpy(gold["hand_vertices"]).cuda()
    ij, depths = project_points(pts, cam)
    assert _rel(ij.cpu().numpy(), gold[f"cam_{tag}_ij"]) < 1e-12
    assert _rel(depths.cpu().numpy(), gold[f"cam_{tag}_depths"]) < 1e-12
    ij_b = torch.from_numpy(gold[f"cam_{tag}_ij_b"]).cuda()
    d_b = torch.from_numpy(gold[f"cam_{tag}_depths_b"]).cuda()
    ref = project_points_backward(pts, cam, ij_b, d_b, reference_transpose=True).cpu().numpy()
    assert _rel(ref, gold[f"cam_{tag}_points_b_reference"]) < 1e-12
    adj = project_points_backward(pts, cam, ij_b, d_b).cpu().numpy()
    assert _rel(adj, gold[f"cam_{tag}_points_b_adjoint"]) < 1e-11
    # accumulation semantics
    twice = project_points_backward(pts, cam, ij_b, d_b, out=torch.from_numpy(adj.copy()).cuda()).cpu().numpy()
    assert _rel(twice, 2 * adj) < 1e-12


@pytest.mark.gpu
def test_vertex_luminosity_and_backward(gold, build_native):
    import torch

    from deodr_b200.mesh_ops import vertex_luminosity, vertex_luminosity_backward

    normals = torch.from_numpy(gold["lum_normals"]).cuda()
    vcol = torch.from_numpy(gold["lum_vertex_colors"]).cuda()
    light, ambient = gold["lum_light"], float(gold["lum_ambient"])
    lum, colors = vertex_luminosity(normals, vcol, light, ambient)
    assert np.abs(lum.cpu().numpy() - gold["lum_luminosity"]).max() < 1e-6
    assert np.abs(colors.cpu().numpy() - gold["lum_colors"]).max() < 1e-6
    g = vertex_luminosity_backward(normals, vcol, light, ambient, colors_b=torch.from_numpy(gold["lum_colors_b"]).cuda())
    assert _rel(g["vertex_colors_b"].cpu().numpy(), gold["lum_vertex_colors_b"]) < 1e-12
    assert _rel(g["normals_b"].cpu().numpy(), gold["lum_normals_b"]) < 1e-12
    assert _rel(g["light_b"].cpu().numpy(), gold["lum_light_b"]) < 1e-11
    lum_only, none = vertex_luminosity(normals, None, None, 0.25)
    assert none is None and np.all(lum_only.cpu().numpy() == np.float32(0.25))


@pytest.mark.gpu
@pytest.mark.parametrize("mesh", ["hand", "torus"])
def test_silhouette_flags_bit_exact_and_normals(mesh, gold, build_native):
    import torch

    from deodr_b200.mesh_ops import MeshTopology

    topo = MeshTopology(gold[f"{mesh}_faces"], gold[f"{mesh}_vertices"].shape[0])
    for ij, flags in zip(gold[f"{mesh}_views_ij"], gold[f"{mesh}_views_edgeflags"]):
        got = topo.edge_on_silhouette(torch.from_numpy(ij).cuda()).cpu().numpy()
        assert np.array_equal(got, flags)
    fn, vn = topo.vertex_normals(torch.from_numpy(gold[f"{mesh}_vertices"]).cuda())
    assert _rel(fn.cpu().numpy(), gold[f"{mesh}_face_normals"]) < 1e-12
    assert _rel(vn.cpu().numpy(), gold[f"{mesh}_vertex_normals"]) < 1e-12
    if mesh == "hand":
        vb = topo.vertex_normals_backward(torch.from_numpy(gold["hand_vertices"]).cuda(),
                                          torch.from_numpy(gold["hand_vertex_normals_b"]).cuda()).cpu().numpy()
        assert _rel(vb, gold["hand_vertices_b"]) < 1e-11


@pytest.mark.gpu
def test_silhouette_flags_at_benchmark_sizes(build_native):
    """c3 / c5 meshes: the device flags equal the numpy restatement used to build the synthetic scenes (itself checked
    against the reference on the fixtures above), under a view rotation."""
    import torch

    from deodr_b200.mesh_ops import MeshTopology
    from deodr_b200.scenes import torus_scene

    for n, size in ((158, 1024), (708, 2048)):
        scene = torus_scene(n, size, size, view=1, n_views=5)
        topo = MeshTopology(scene.faces, scene.depths.shape[0], clockwise=scene.clockwise)
        got = topo.edge_on_silhouette(torch.from_numpy(scene.ij).cuda()).cpu().numpy()
        assert np.array_equal(got.astype(bool), np.asarray(scene.edgeflags, dtype=bool))


@pytest.mark.gpu
def test_device_resident_mesh_view_matches_the_numpy_chain(gold, build_native):
    """DeviceMeshView: vertices -> image -> vertices_b with every step on the device, against the same chain assembled
    from the fixtures' numpy pieces and the CPU oracle for the raster step."""
    import torch

    from deodr_b200.mesh_ops import CameraParams, DeviceMeshView
    from deodr_b200.scenes import SceneArrays, dense_image_b
    from oracle.oracle import Oracle, available

    faces, vertices = gold["hand_faces"], gold["hand_vertices"]
    cam = CameraParams(gold["view_extrinsic"], gold["view_intrinsic"])  # the camera of the configs[1] fixture
    light, ambient = gold["lum_light"], float(gold["lum_ambient"])
    view = DeviceMeshView(faces, vertices.shape[0], cam, 480, 640, light_directional=light, ambient=ambient,
                          background_color=(0.3, 0.5, 0.7))
    vcol = gold["lum_vertex_colors"]
    image = view.render(torch.from_numpy(vertices).cuda(), torch.from_numpy(vcol).cuda())
    # the same Scene2D from the reference-generated pieces
    topo = topology_arrays(faces, vertices.shape[0])
    ij, depths = gold["c2_ij"], gold["c2_depths"]
    T, V = faces.shape[0], vertices.shape[0]
    scene = SceneArrays(
        faces=faces, faces_uv=np.zeros((T, 3), np.uint32), ij=ij, depths=depths, textured=np.zeros(T, bool),
        uv=np.zeros((1, 2)), shade=np.zeros(V), colors=gold["lum_colors"], shaded=np.zeros(T, bool),
        edgeflags=_flags_from_topology(topo, ij, False).astype(bool), height=480, width=640, nb_colors=3,
        texture=np.zeros((2, 2, 3)), background_image=None, background_color=np.array([0.3, 0.5, 0.7]),
        clockwise=False, backface_culling=True, strict_edge=True, perspective_correct=False, integer_pixel_centers=True)
    oracle = Oracle("reference" if available("reference") else "port")
    image_ref, z_ref = oracle.render(scene, 1.0)
    assert np.isfinite(z_ref).mean() > 0.05  # the hand is in the picture
    assert np.abs(image.cpu().numpy() - image_ref).max() < 2e-6
    image_b = dense_image_b(image_ref)
    g_ref = oracle.render_b(scene, 1.0, image_ref, z_ref, image_b)
    got = view.backward(torch.from_numpy(image_b).cuda())
    # chain of the reference formulas in numpy: colours -> luminosity -> normals ; ij -> camera (true adjoint)
    lum = gold["lum_luminosity"]
    directional = lum - ambient
    lum_b = np.sum(vcol * g_ref["colors_b"], axis=1)
    assert np.abs(got["vertex_colors_b"].cpu().numpy() - g_ref["colors_b"] * lum[:, None]).max() < 1e-4 * np.abs(g_ref["colors_b"]).max()
    light_b = np.concatenate((-np.sum((lum_b * (directional > 0))[:, None] * gold["lum_normals"], axis=0), [lum_b.sum()]))
    assert np.abs(got["light_b"].cpu().numpy() - light_b).max() < 1e-4 * np.abs(light_b).max()
    # vertices_b = camera adjoint of ij_b + normals adjoint of the luminosity chain, both checked piecewise above
    from deodr_b200.mesh_ops import MeshTopology, project_points_backward

    normals_b = -((lum_b * (directional > 0))[:, None]) * light
    expect = MeshTopology(faces, V).vertex_normals_backward(torch.from_numpy(vertices).cuda(), torch.from_numpy(normals_b).cuda())
    project_points_backward(torch.from_numpy(vertices).cuda(), cam, torch.from_numpy(g_ref["ij_b"]).cuda(), out=expect)
    expect = expect.cpu().numpy()
    assert np.abs(got["vertices_b"].cpu().numpy() - expect).max() < 1e-4 * np.abs(expect).max()
