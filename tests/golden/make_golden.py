"""Generates the committed golden vectors from the REAL reference (run in the build container only).

Needs oracle/_ref (the unmodified reference core compiled by oracle/Makefile from /root/reference) and
/root/reference/deodr/data/trefle.jpg.  Outputs (all small):

* trefle_texture_u8.npy         - decoded texture of the reference soup scene (input fixture)
* soup_pinned.npz               - S-soup(30, 200x200, seed 2, clockwise=True): SHA-256 of image / z-buffer (the values
                                  pinned by the reference's tests/test_render_mesh.py:66-74), a 4x sub-sampled image,
                                  and every gradient of renderScene_B for image_b = 2 (image - obs)
* soup_fitting.json             - first two image hashes + first two losses of the four runs of
                                  tests/test_triangle_soup_fitting.py (antialiase_error=False runs only are consumed)
* small_*.npz                   - full inputs/outputs of a few 48x40 scenes covering the flag combinations

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from deodr_b200.scenes import dense_image_b, soup_scene, torus_scene  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
ref = Oracle("reference")
fix = Oracle("reference", texfix=True)

tex_u8 = np.asarray(Image.open("/root/reference/deodr/data/trefle.jpg"))
np.save(os.path.join(HERE, "trefle_texture_u8.npy"), tex_u8)
tex = tex_u8.astype(np.float64) / 255

# ---- pinned soup scene
np.random.seed(2)
scene = soup_scene(clockwise=True, texture=tex)
image, z = ref.render(scene, 1.0)
assert sha(image).startswith("4de52cc3e902f92f") and sha(z).startswith("b6f87e03c60bd820"), "reference pins broken"
image_b = dense_image_b(image)
g = ref.render_b(scene, 1.0, image, z, image_b)
gf = fix.render_b(scene, 1.0, image, z, image_b)
np.savez_compressed(
    os.path.join(HERE, "soup_pinned.npz"), image_sha=sha(image), z_sha=sha(z), image_sub=image[::4, ::4],
    z_sub=z[::4, ::4], ij_b=g["ij_b"], colors_b=g["colors_b"], uv_b=g["uv_b"], shade_b=g["shade_b"],
    texture_b_last_writer=g["texture_b"].astype(np.float32), texture_b_summed=gf["texture_b"].astype(np.float32),
)

# ---- soup fitting runs (reference tests/test_triangle_soup_fitting.py; loop of examples/triangle_soup_fitting.py:150)
def fitting_run(clockwise, n_iter=3):
    np.random.seed(2)
    gt = soup_scene(clockwise=clockwise, texture=tex)
    target, _ = ref.render(gt, 1.0)
    n_vertices = len(gt.depths)
    ij0 = gt.ij + np.random.randn(n_vertices, 2) * 10
    uv0 = np.minimum(np.maximum(gt.uv + np.random.randn(n_vertices, 2) * 0, 0), np.array(gt.texture.shape[:2]) - 1)
    _ = np.random.randn(n_vertices, 3)  # colours displacement draw (magnitude 0)
    gt.ij, gt.uv = ij0, uv0
    speed = np.zeros((n_vertices, 2))
    losses, hashes, ijs = [], [], []
    for _ in range(n_iter):
        ijs.append(gt.ij.copy())
        image, z = ref.render(gt, 1.0)
        diff = image - target
        losses.append(float(np.sum(diff**2)))
        hashes.append(sha(image))
        grads = ref.render_b(gt, 1.0, image, z, 2 * diff)
        speed = 0.80 * speed - grads["ij_b"] * 0.01
        gt.ij = gt.ij + speed
    return {"losses": losses, "hashes": hashes, "target_sha": sha(target), "ij_sha": [sha(a) for a in ijs]}

runs = {"ccw": fitting_run(False), "cw": fitting_run(True)}
assert runs["ccw"]["hashes"][0].startswith("38b6f695") and runs["ccw"]["hashes"][1].startswith("0434ea72")
assert runs["cw"]["hashes"][0].startswith("eb9f335a") and runs["cw"]["hashes"][1].startswith("6b4cc11e")
json.dump(runs, open(os.path.join(HERE, "soup_fitting.json"), "w"), indent=1)

# ---- small full-array fixtures
def small(tag, scene, sigma):
    image, z = ref.render(scene, sigma)
    out = dict(image=image, z=z, sigma=sigma)
    if scene.backface_culling and not scene.perspective_correct:
        image_b = dense_image_b(image)
        g = fix.render_b(scene, sigma, image, z, image_b)
        out.update(ij_b=g["ij_b"], colors_b=g["colors_b"], uv_b=g["uv_b"], shade_b=g["shade_b"],
                   texture_b=g["texture_b"].astype(np.float32))
    for k in ("faces", "faces_uv", "ij", "depths", "textured", "uv", "shade", "colors", "shaded", "edgeflags"):
        out["in_" + k] = getattr(scene, k)
    out["in_texture"] = scene.texture
    if scene.background_image is not None:
        out["in_background_image"] = scene.background_image
    else:
        out["in_background_color"] = scene.background_color
    out["flags"] = np.array([scene.height, scene.width, scene.nb_colors, scene.clockwise, scene.backface_culling,
                             scene.strict_edge, scene.perspective_correct, scene.integer_pixel_centers], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, f"small_{tag}.npz"), **out)

small_tex = tex[::6, ::6].copy()
np.random.seed(5)
s = soup_scene(n_tri=12, width=48, height=40, clockwise=False, texture=small_tex, min_det=60)
small("soup_s1", s, 1.0)
s.strict_edge, s.integer_pixel_centers = False, False
small("soup_nonstrict_halfpix_s2", s, 2.0)
s.perspective_correct = True
small("soup_persp", s, 1.0)
small("torus", torus_scene(10, 48, 40), 1.0)
small("torus_tex", torus_scene(10, 48, 40, textured=True, texture_size=16), 1.5)
print("golden vectors written to", HERE)
