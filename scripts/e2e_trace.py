"""Development aid: wall-clock breakdown of the reference-shaped host path on the c5 workload."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deodr_b200 import differentiable_renderer_cython as shim  # noqa: E402
from deodr_b200.differentiable_renderer import Scene2D  # noqa: E402

scene = bench.build_scene(sys.argv[1] if len(sys.argv) > 1 else "c5", 0, 1)
s2 = Scene2D(**{k: getattr(scene, k) for k in (
    "faces", "faces_uv", "ij", "depths", "textured", "uv", "shade", "colors", "shaded", "edgeflags", "height", "width",
    "nb_colors", "texture", "background_image", "background_color", "clockwise", "backface_culling", "strict_edge",
    "perspective_correct", "integer_pixel_centers")})
H, W, C = scene.height, scene.width, scene.nb_colors
image, z = np.empty((H, W, C)), np.empty((H, W))
image_b = np.random.default_rng(1).random((H, W, C)) * 2 - 1
for it in range(4):
    if it == 3:
        os.environ["DEODR_B200_TRACE"] = "1"
    t0 = time.perf_counter()
    s2.clear_gradients()
    t1 = time.perf_counter()
    shim.renderSceneCpp(s2, 1.0, image, z)
    t2 = time.perf_counter()
    shim.renderSceneBCpp(s2, 1.0, image, z, image_b)
    t3 = time.perf_counter()
    print(f"iter {it}: clear {1e3*(t1-t0):.2f} ms  renderSceneCpp {1e3*(t2-t1):.2f} ms  renderSceneBCpp {1e3*(t3-t2):.2f} ms")
