"""Generates tests/golden/scene_ops.npz from the REFERENCE's own Python (imported from /root/reference in the build
container): inputs and expected outputs of the steps either side of the raster -
Camera.project_points(+_backward) with and without distortion, the Scene3D luminosity chain (+ backward),
TriMeshAdjacencies.edge_on_silhouette / normals (+ backward) on deodr/data/hand.obj (under several view rotations) and on
a small torus, plus the Scene2D of configs[1] (hand mesh, 640x480 RGB Gouraud) as produced by the reference's Scene3D.

    PYTHONPATH=tests/dropin/stubs python tests/golden/make_scene_ops_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "..", "dropin", "stubs"), ROOT]

# the reference package without its compiled extension: the functions used here never call it
pkg = types.ModuleType("deodr")
pkg.__path__ = ["/root/reference/deodr"]
sys.modules["deodr"] = pkg
stub = types.ModuleType("deodr.differentiable_renderer_cython")
stub.renderSceneCpp = stub.renderSceneBCpp = None
sys.modules["deodr.differentiable_renderer_cython"] = stub
import deodr.differentiable_renderer as dr  # noqa: E402
from deodr.obj import read_obj  # noqa: E402
from deodr.triangulated_mesh import ColoredTriMesh, TriMeshAdjacencies  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

from deodr_b200.scenes import torus_mesh  # noqa: E402

out = {}
rng = np.random.default_rng(7)

# ---- f1: camera
faces_h, vertices_h = read_obj("/root/reference/deodr/data/hand.obj")
faces_h = np.asarray(faces_h)
out["hand_faces"] = faces_h.astype(np.uint32)
out["hand_vertices"] = vertices_h
rot = Rotation.from_rotvec([0.3, -0.5, 0.9]).as_matrix()
for tag, dist in (("plain", None), ("dist", np.array([0.1, -0.05, 0.01, 0.02, 0.003]))):
    cam = dr.default_camera(640, 480, 60, vertices_h, rot, distortion=dist)
    sb = {}
    ij, depths = cam.project_points(vertices_h, store_backward=sb)
    ij_b = rng.normal(size=ij.shape).astype(np.float32).astype(np.float64)
    depths_b = rng.normal(size=depths.shape)
    out[f"cam_{tag}_extrinsic"], out[f"cam_{tag}_intrinsic"] = cam.extrinsic, cam.intrinsic
    if dist is not None:
        out[f"cam_{tag}_distortion"] = dist
    out[f"cam_{tag}_ij"], out[f"cam_{tag}_depths"] = ij, depths
    out[f"cam_{tag}_ij_b"], out[f"cam_{tag}_depths_b"] = ij_b, depths_b
    out[f"cam_{tag}_points_b_reference"] = cam.project_points_backward(ij_b, sb, depths_b=depths_b)
    # the true adjoint: the reference's last line multiplies by R^T instead of R (defect #3): undo, redo
    R = cam.extrinsic[:3, :3]
    p_camera_b = out[f"cam_{tag}_points_b_reference"].dot(np.linalg.inv(R.T))
    out[f"cam_{tag}_points_b_adjoint"] = p_camera_b.dot(R)

# ---- f1: luminosity chain through the reference's Scene3D
mesh = ColoredTriMesh(faces_h, vertices_h, clockwise=False, nb_colors=3)
mesh.set_vertices_colors(rng.random((vertices_h.shape[0], 3)))
scene = dr.Scene3D(sigma=1)
light = np.array([0.3, -0.3, 0.0]) * 1.3
scene.set_light(light_directional=light, light_ambient=0.3)
scene.set_mesh(mesh)
scene.store_backward_current = {}
colors = scene._compute_vertices_colors_with_illumination()
colors_b = rng.normal(size=colors.shape).astype(np.float32).astype(np.float64)
scene._compute_vertices_colors_with_illumination_backward(colors_b)
out["lum_light"], out["lum_ambient"] = light, np.array(0.3)
out["lum_normals"], out["lum_vertex_colors"] = mesh.vertex_normals, mesh.vertices_colors
out["lum_colors"] = colors
out["lum_luminosity"] = scene.compute_vertices_luminosity()
out["lum_colors_b"] = colors_b
out["lum_vertex_colors_b"] = mesh.vertices_colors_b
out["lum_normals_b"] = scene.vertex_normals_b
out["lum_light_b"] = np.concatenate((scene.light_directional_b, [scene.light_ambient_b]))

# ---- f2: normals (+ backward) and silhouette flags, hand mesh under view rotations + a small closed torus
out["hand_face_normals"], out["hand_vertex_normals"] = mesh.face_normals, mesh.vertex_normals
vn_b = rng.normal(size=mesh.vertex_normals.shape)
mesh.compute_vertex_normals_backward(vn_b)
out["hand_vertex_normals_b"], out["hand_vertices_b"] = vn_b, mesh._vertices_b.copy()
adj = TriMeshAdjacencies(faces_h, clockwise=False, nb_vertices=vertices_h.shape[0])
flags, ijs = [], []
for k in range(6):
    r = Rotation.from_euler("yx", [2 * np.pi * k / 6, 0.3]).as_matrix()
    cam = dr.default_camera(640, 480, 60, vertices_h, r)
    ij = cam.project_points(vertices_h, return_depths=False)
    ijs.append(ij)
    flags.append(np.asarray(adj.edge_on_silhouette(ij)))
out["hand_views_ij"], out["hand_views_edgeflags"] = np.stack(ijs), np.stack(flags).astype(np.uint8)
tv, tf, _ = torus_mesh(24)
out["torus_faces"], out["torus_vertices"] = tf.astype(np.uint32), tv
tadj = TriMeshAdjacencies(tf, clockwise=False, nb_vertices=tv.shape[0])
tflags, tijs = [], []
for k in range(4):
    r = Rotation.from_euler("yx", [2 * np.pi * k / 4, 0.3]).as_matrix()
    cam = dr.default_camera(256, 192, 60, tv, r)
    ij = cam.project_points(tv, return_depths=False)
    tijs.append(ij)
    tflags.append(np.asarray(tadj.edge_on_silhouette(ij)))
out["torus_views_ij"], out["torus_views_edgeflags"] = np.stack(tijs), np.stack(tflags).astype(np.uint8)
tmesh = ColoredTriMesh(tf, tv, clockwise=False, nb_colors=3)
out["torus_face_normals"], out["torus_vertex_normals"] = tmesh.face_normals, tmesh.vertex_normals

# ---- configs[1]: the Scene2D the reference's Scene3D hands to the raster core for the hand mesh at 640x480
# (camera of deodr/examples/render_mesh.py:26-29: half a turn about x - PerspectiveCamera composes `rot` in a way
# that only frames the mesh for symmetric rotation matrices)
cam = dr.default_camera(640, 480, 60, vertices_h, Rotation.from_euler("xyz", [180, 0, 0], degrees=True).as_matrix())
out["view_extrinsic"], out["view_intrinsic"] = cam.extrinsic, cam.intrinsic
scene.set_background_color([0.3, 0.5, 0.7])
scene.camera = cam
ij, depths = cam.project_points(vertices_h)
out["c2_ij"], out["c2_depths"] = ij, depths
out["c2_colors"] = colors
out["c2_edgeflags"] = np.asarray(adj.edge_on_silhouette(ij)).astype(np.uint8)
out["c2_background_color"] = np.array([0.3, 0.5, 0.7])
inside = (ij[:, 0] > 0) & (ij[:, 0] < 640) & (ij[:, 1] > 0) & (ij[:, 1] < 480)
assert inside.mean() > 0.9, inside.mean()

np.savez_compressed(os.path.join(HERE, "scene_ops.npz"), **out)
print({k: v.shape for k, v in out.items()})
