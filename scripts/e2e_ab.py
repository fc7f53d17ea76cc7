"""Development aid: median wall-clock of the three reference-shaped host calls on a workload (A/B of host-path settings
through environment variables; one process per arm)."""
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deodr_b200 import differentiable_renderer_cython as shim  # noqa: E402
from deodr_b200.differentiable_renderer import Scene2D  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c5"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
scene = bench.build_scene(wl, 0, 1)
s2 = Scene2D(**{k: getattr(scene, k) for k in (
    "faces", "faces_uv", "ij", "depths", "textured", "uv", "shade", "colors", "shaded", "edgeflags", "height", "width",
    "nb_colors", "texture", "background_image", "background_color", "clockwise", "backface_culling", "strict_edge",
    "perspective_correct", "integer_pixel_centers")})
H, W, C = scene.height, scene.width, scene.nb_colors
image, z = np.empty((H, W, C)), np.empty((H, W))
image_b = np.random.default_rng(1).random((H, W, C)) * 2 - 1
rows = []
for it in range(iters + 3):
    t0 = time.perf_counter()
    s2.clear_gradients()
    t1 = time.perf_counter()
    shim.renderSceneCpp(s2, 1.0, image, z)
    t2 = time.perf_counter()
    shim.renderSceneBCpp(s2, 1.0, image, z, image_b)
    t3 = time.perf_counter()
    if it >= 3:
        rows.append((t1 - t0, t2 - t1, t3 - t2, t3 - t0))
med = [1e3 * statistics.median(r[i] for r in rows) for i in range(4)]
print(f"{os.environ.get('ARM', 'default'):>24}: clear {med[0]:.2f}  fwd {med[1]:.2f}  bwd {med[2]:.2f}  step {med[3]:.2f} ms")
