"""Development aid (CPU only): vertices beyond the `short` range of the reference's loop counters (DR.h:676-711, 925).

The reference casts row / column bounds to `short`; on x86 that wraps modulo 65536, so a triangle with one vertex more
than 32767 pixels away is drawn in part (typically the half between its two near vertices) or at wrapped positions -
behaviour the C restatement (oracle/deodr_oracle.c) reproduces bit for bit.  The kernel phases (emulated here, same
source as the CUDA kernels) apply the same wrap to every bound but bin a triangle by the rows [first, last] of the union
of its halves, which is empty once the first bound has wrapped past the last: such triangles are culled.  This script
measures how often that shows: INTEGRATION.md section 5 quotes it.

    python scripts/probe_far_vertices.py            # the default build
    python scripts/probe_far_vertices.py exact      # built with DEODR_EXACT_SHORT_WRAP=1 (rmath.h): no difference left
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from canon import Emulator  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

from deodr_b200.scenes import confetti_scene  # noqa: E402

emu, ref, port = Emulator(exact_short_wrap="exact" in sys.argv[1:]), Oracle("reference"), Oracle("port")
rng = np.random.default_rng(1)
for far in (1e3, 1e4, 3.2e4, 4e4, 7e4, 1e6, 3e9, 1e15):
    n, differs, port_differs, pixels = 20, 0, 0, 0
    for trial in range(n):
        scene = confetti_scene(200, 64, 48, size=float(rng.choice([5.0, 40.0])), seed=trial, edge_ratio=0.3)
        idx = rng.choice(scene.ij.shape[0], size=40, replace=False)  # 40 of 600 vertices pushed far away
        scene.ij[idx] += rng.choice([-1, 1], size=(40, 2)) * far * rng.random((40, 2))
        rows = np.floor(scene.ij[:, 1])  # (row 32767 after the wrap: the reference then writes at negative pixel indices)
        scene.ij[(np.abs(rows) < 2.0 ** 31) & ((rows.astype(np.int64) & 0xFFFF) == 32767), 1] += 1.0
        _, z = ref.render(scene, 1.0)
        _, zp = port.render(scene, 1.0)
        got = emu.render(scene, 1.0)["z"]
        differs += not np.array_equal(got, z)
        port_differs += not np.array_equal(zp, z)
        pixels += int((got != z).sum())
    print(f"vertices up to {far:g} px away: kernel phases differ from the reference in {differs}/{n} scenes "
          f"({pixels} of {n * 64 * 48} pixels), C restatement in {port_differs}/{n}")
