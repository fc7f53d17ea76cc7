"""Development aid (CPU only): unusual inputs through the CPU emulation of the kernel phases (tests/emul, the same
__host__ __device__ code the CUDA kernels run) against the compiled reference (oracle/_ref).  INTEGRATION.md section 5
quotes what this prints.

    python scripts/probe_edge_cases.py [exact]      # exact: the DEODR_EXACT_SHORT_WRAP=1 build of the emulation

Sections: non-finite / degenerate geometry and attributes; textures down to one texel, texel-grid uv, every
textured / shaded flag combination, 2 .. 16 channels; images at the maximum sides (32767) and extreme aspect ratios."""
import copy
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from canon import Emulator  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

from deodr_b200.scenes import confetti_scene, dense_image_b, soup_scene, torus_scene  # noqa: E402

warnings.filterwarnings("ignore")
emu, ref = Emulator(exact_short_wrap="exact" in sys.argv[1:]), Oracle("reference", texfix=True)
tex = np.load(os.path.join(ROOT, "tests/golden/trefle_texture_u8.npy")).astype(np.float64) / 255
GRADS = ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b")


def compare(name, scene, sigma=1.0):
    image, z = ref.render(scene, sigma)
    fwd = emu.render(scene, sigma)
    same = (fwd["z"] == z) | (np.isnan(fwd["z"]) & np.isnan(z))
    with np.errstate(all="ignore"):
        d = np.abs(fwd["image"] - image)
        worst = float(np.nanmax(d)) if np.isfinite(d).any() else 0.0
    msg = (f"{name}: z-buffer {'identical' if same.all() else f'DIFFERS at {int((~same).sum())} px'}, image {worst:.1e}, "
           f"NaN pattern {'same' if np.array_equal(np.isnan(fwd['image']), np.isnan(image)) else 'DIFFERS'}")
    if scene.backface_culling and not scene.perspective_correct:
        image_b = dense_image_b(np.nan_to_num(image))
        r, g = ref.render_b(scene, sigma, image, z, image_b), emu.render_b(scene, sigma, fwd, image_b)
        rel = 0.0
        for k in GRADS:
            fin = np.isfinite(r[k]) & np.isfinite(g[k])
            if fin.any() and np.abs(r[k][fin]).max() > 0:
                rel = max(rel, float(np.abs(r[k][fin] - g[k][fin]).max() / np.abs(r[k][fin]).max()))
            if not np.array_equal(np.isfinite(r[k]), np.isfinite(g[k])):
                msg += f" [{k}: {int((~np.isfinite(r[k])).sum())} / {int((~np.isfinite(g[k])).sum())} non-finite entries]"
        msg += f", gradients {rel:.1e}"
    print(msg, flush=True)


rng = np.random.default_rng(0)


def base():
    return confetti_scene(300, 64, 48, size=6.0, seed=3, edge_ratio=0.4)


print("-- non-finite and degenerate inputs")
s = base(); s.ij[rng.choice(900, 10, replace=False)] = np.nan; compare("NaN in 10 vertex positions", s)
s = base(); s.ij[rng.choice(900, 10, replace=False)] = np.inf; compare("inf in 10 vertex positions", s)
s = base(); s.depths[rng.choice(900, 10, replace=False)] = np.nan; compare("NaN in 10 depths (the reference's depth sort is undefined)", s)
s = base(); s.depths[rng.choice(900, 30, replace=False)] *= -1; compare("30 negative depths", s)
s = base(); s.depths[rng.choice(900, 30, replace=False)] = 0.0; compare("30 zero depths", s)
s = base(); s.depths *= 1e-300; compare("depths x 1e-300", s)
s = base(); s.depths *= 1e300; compare("depths x 1e300", s)
s = base(); s.colors[rng.choice(900, 10, replace=False)] = np.nan; compare("NaN in 10 colours", s)
s = base(); s.colors *= 1e6; compare("colours x 1e6 (image error is absolute)", s)
s = base(); s.faces[::7, 1] = s.faces[::7, 0]; compare("faces with a repeated vertex", s)
s = base(); s.ij[s.faces[::5, 2]] = s.ij[s.faces[::5, 1]]; compare("coincident vertices (zero-area triangles)", s)
s = base(); s.ij = np.round(s.ij); compare("integer vertex positions (pixel centres ON silhouette edges: T = 0)", s)
s = base(); s.edgeflags[:] = 1; compare("every edge flagged as silhouette", s)
s = base(); s.backface_culling = False; compare("no culling", s)
s = base(); s.perspective_correct = True; s.depths[::9] = 1e-12; compare("perspective_correct with depths of 1e-12", s)
for far in (3.2e4, 7e4, 1e9):
    s = base(); idx = rng.choice(900, 40, replace=False)
    s.ij[idx] += rng.choice([-1, 1], size=(40, 2)) * far * rng.random((40, 2))
    rows = np.floor(s.ij[:, 1])  # (row 32767 after the wrap: the reference then writes at negative pixel indices)
    s.ij[(np.abs(rows) < 2.0 ** 31) & ((rows.astype(np.int64) & 0xFFFF) == 32767), 1] += 1.0
    compare(f"40 vertices up to {far:g} px away", s)

print("-- textures, flags, channels")
for th, tw in ((2, 2), (3, 5), (1, 7), (9, 1), (1, 1)):
    np.random.seed(2)
    s = soup_scene(clockwise=False, texture=rng.random((th, tw, 3)))
    s.uv = rng.random(s.uv.shape) * np.array([max(tw - 1, 1), max(th - 1, 1)]) * 1.2 - 0.1
    compare(f"texture {th} x {tw}" + (" (texel -1 is read with weight 0: undefined in the reference too)" if min(th, tw) == 1 else ""), s)
np.random.seed(2); s = soup_scene(clockwise=False, texture=tex[::4, ::4].copy()); s.uv = np.round(s.uv); compare("uv on the texel grid", s)
np.random.seed(2); s = soup_scene(clockwise=False, texture=tex[::4, ::4].copy()); s.uv = s.uv * 50 - 500; compare("uv far outside the texture", s)
for t, sh in ((1, 0), (0, 1), (1, 1), (0, 0)):
    np.random.seed(2); s = soup_scene(clockwise=False, texture=tex[::4, ::4].copy()); s.textured[:] = t; s.shaded[:] = sh
    compare(f"textured = {t}, shaded = {sh} on every triangle", s)
for C in (2, 5, 8, 13, 16):
    s = confetti_scene(400, 64, 48, size=5.0, seed=4, edge_ratio=0.4, nb_colors=C); compare(f"{C} channels", s)
s = confetti_scene(400, 64, 48, size=5.0, seed=4, edge_ratio=0.4, nb_colors=16)
s.background_image, s.background_color = rng.random((48, 64, 16)), None; compare("16 channels over a background image", s)

print("-- image sides")
for W, H in ((1, 1), (1, 97), (130, 1), (32767, 3), (3, 32767), (20000, 17), (32767, 40)):
    s = copy.copy(base()); s.width, s.height = W, H
    s.ij = s.ij * np.array([W / 64.0, H / 48.0])
    s.ij[:, 0] = np.clip(s.ij[:, 0] + rng.normal(size=900) * 3, -32000, 32766.5)  # (inside the `short` range)
    s.ij[:, 1] = np.clip(s.ij[:, 1] + rng.normal(size=900) * 3, -32000, 32766.5)
    compare(f"{W} x {H} image", s)
for sigma in (1e-3, 45.0):
    compare(f"sigma = {sigma:g}", torus_scene(24, 160, 120), sigma)
