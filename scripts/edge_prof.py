"""Development aid: per-phase cycle totals of k_edge_fwd from a -DDEODR_PROFILE_EDGE build (DEODR_B200_LIB=...)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deodr_b200 import _cabi  # noqa: E402
from deodr_b200.renderer import DeviceScene, Renderer  # noqa: E402
from deodr_b200.scenes import torus_scene  # noqa: E402

lib = ctypes.CDLL(_cabi.LIB_PATH)
scene = torus_scene(708, 2048, 2048, view=0, n_views=1, textured=False, nb_colors=3)
r = Renderer(0)
ds = DeviceScene(scene, "cuda:0")
out = (ctypes.c_ulonglong * 16)()
for it in range(3):
    fwd = r.render(ds, 1.0)
    torch.cuda.synchronize()
    lib.deodr_b200_debug_prof(out, 1)
v = np.array(list(out), dtype=np.float64)
n = v[8]
names = ["loads", "setup+bar", "spans+bar", "blend", "bar", "store"]
print("CTAs", n, "total cycles/CTA", v[:6].sum() / n)
for k, nm in enumerate(names):
    print(f"  {nm:10s} {v[k] / n:9.0f} cycles/CTA")
