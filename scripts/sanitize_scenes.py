"""Development aid: a few small scenes through the device and host paths (run under compute-sanitizer)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deodr_b200.renderer import DeviceScene, Renderer  # noqa: E402
from deodr_b200.scenes import dense_image_b, soup_scene, torus_scene  # noqa: E402

tex = np.load(os.path.join(ROOT, "tests/golden/trefle_texture_u8.npy")).astype(np.float64) / 255
r = Renderer(0)
np.random.seed(2)
scenes = [soup_scene(clockwise=True, texture=tex), torus_scene(24, 150, 130, nb_colors=3),
          torus_scene(40, 200, 160, textured=True, nb_colors=3), torus_scene(60, 300, 300, nb_colors=1)]
for sc in scenes:
    ds = DeviceScene(sc, "cuda:0")
    fwd = r.render(ds, 1.0, face_id=True)
    torch.cuda.synchronize()
    ib = torch.from_numpy(dense_image_b(fwd["image"].cpu().numpy().astype(np.float64))).cuda()
    g = r.render_b(ds, 1.0, fwd, ib)
    torch.cuda.synchronize()
    print("ok", sc.faces.shape[0], float(g["ij_b"].abs().sum()))
