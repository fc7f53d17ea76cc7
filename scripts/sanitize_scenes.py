"""Development aid: a few small scenes through every device path (run under compute-sanitizer):

    compute-sanitizer --tool memcheck   python scripts/sanitize_scenes.py
    compute-sanitizer --tool racecheck  python scripts/sanitize_scenes.py
    compute-sanitizer --tool synccheck  python scripts/sanitize_scenes.py

Covers: forward + adjoint (interpolated, textured, 1 / 3 channels, perspective), a batch of views on several lanes, the
forward in two calls, the record-parallel adjoint, the
antialiase_error mode, the G-buffer outputs, a re-plan (lists outgrowing the plan), image widths that do and do not qualify
for the TMA tile store / loads, and the f1 / f2 scene ops."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deodr_b200.mesh_ops import CameraParams, MeshTopology, project_points, project_points_backward, vertex_luminosity  # noqa: E402
from deodr_b200.renderer import DeviceScene, Renderer  # noqa: E402
from deodr_b200.scenes import confetti_scene, dense_image_b, soup_scene, torus_scene  # noqa: E402

tex = np.load(os.path.join(ROOT, "tests/golden/trefle_texture_u8.npy")).astype(np.float64) / 255
r = Renderer(0)
np.random.seed(2)
persp = soup_scene(clockwise=False, texture=tex)
persp.perspective_correct, persp.strict_edge = True, False
scenes = [soup_scene(clockwise=True, texture=tex), torus_scene(24, 150, 130, nb_colors=3),  # 150: no TMA (pitch % 16)
          torus_scene(40, 200, 160, textured=True, nb_colors=3), torus_scene(60, 300, 300, nb_colors=1),
          confetti_scene(3000, 64, 48, size=2.5, seed=1), persp]
for sc in scenes:
    ds = DeviceScene(sc, "cuda:0")
    fwd = r.render(ds, 1.0, face_id=True, barycentric=True)
    torch.cuda.synchronize()
    if sc.backface_culling and not sc.perspective_correct:
        ib = torch.from_numpy(dense_image_b(fwd["image"].cpu().numpy().astype(np.float64))).cuda()
        g = r.render_b(ds, 1.0, fwd, ib)
        torch.cuda.synchronize()
        print("ok fwd+bwd", sc.faces.shape[0], float(g["ij_b"].abs().sum()))
# batch of views on two lanes, shared gradients
views = [DeviceScene(torus_scene(30, 128, 96, view=v, n_views=3), "cuda:0") for v in range(3)]
outs = r.render_views(views, 1.0)
shared = views[0].zero_grads()
grads = [dict(shared, ij_b=torch.zeros_like(shared["ij_b"])) for _ in views]
r.render_b_views(views, 1.0, outs, [torch.ones_like(o["image"]) for o in outs], grads)
torch.cuda.synchronize()
print("ok views", float(shared["colors_b"].abs().sum()))
# the forward in two calls (binning | rest), then the adjoint
half = r.render_views(views, 1.0, part="geometry")
half = r.render_views(views, 1.0, out=half, part="resume")
r.render_b_views(views, 1.0, half, [torch.ones_like(o["image"]) for o in half])
torch.cuda.synchronize()
print("ok two-call forward")
# record-parallel adjoint of the small triangles (read from the environment when a workspace is created)
os.environ["DEODR_B200_SMALL_ADJOINT"] = "record"
r2 = Renderer(0)
del os.environ["DEODR_B200_SMALL_ADJOINT"]
for sc in (confetti_scene(3000, 64, 48, size=2.5, seed=1), torus_scene(24, 150, 130, nb_colors=3), torus_scene(60, 300, 300, nb_colors=1)):
    ds = DeviceScene(sc, "cuda:0")
    fwd = r2.render(ds, 1.0)
    r2.render_b(ds, 1.0, fwd, torch.ones_like(fwd["image"]))
torch.cuda.synchronize()
print("ok record-parallel adjoint")
# antialiase_error mode
sc = torus_scene(24, 160, 120)
ds = DeviceScene(sc, "cuda:0")
obs = torch.rand((120, 160, 3), device="cuda")
fwd = r.render(ds, 1.0, obs=obs)
g = r.render_b(ds, 1.0, fwd, err_buffer_b=torch.rand((120, 160), device="cuda"))
torch.cuda.synchronize()
print("ok error mode", float(fwd["err_buffer"].sum()), float(g["colors_b"].abs().sum()))
# re-plan: same shape, far larger lists
r.render(DeviceScene(confetti_scene(2000, 128, 96, size=1.2, seed=1, edge_ratio=0.05), "cuda:0"), 1.0)
r.render(DeviceScene(confetti_scene(2000, 128, 96, size=14.0, seed=2, edge_ratio=0.6), "cuda:0"), 1.0)
torch.cuda.synchronize()
print("ok re-plan")
# f1 / f2
g = np.load(os.path.join(ROOT, "tests/golden/scene_ops.npz"))
cam = CameraParams(g["cam_dist_extrinsic"], g["cam_dist_intrinsic"], g["cam_dist_distortion"])
pts = torch.from_numpy(g["hand_vertices"]).cuda()
ij, depths = project_points(pts, cam)
project_points_backward(pts, cam, torch.ones_like(ij, dtype=torch.float32), torch.ones_like(depths))
topo = MeshTopology(g["hand_faces"], pts.shape[0])
flags = topo.edge_on_silhouette(ij)
fn, vn = topo.vertex_normals(pts)
topo.vertex_normals_backward(pts, torch.ones_like(vn))
vertex_luminosity(vn, torch.from_numpy(g["lum_vertex_colors"]).cuda(), g["lum_light"], 0.3)
torch.cuda.synchronize()
print("ok scene ops", int(flags.sum()))
