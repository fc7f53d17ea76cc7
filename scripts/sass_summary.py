"""SASS evidence of the TMA / bulk-copy paths: mnemonic counts per kernel of the objects linked into libdeodr_b200.so.

    python scripts/sass_summary.py > profiles/<tag>_sass_tma.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = re.compile(r"\b(UBLKCP[A-Z0-9_.]*|UTMASTG[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|UTMAREDG[A-Z0-9_.]*|SYNCS\.ARRIVE\.TRANS64[A-Z0-9_.]*|"
                  r"SYNCS\.PHASECHK\.TRANS64[A-Z0-9_.]*|UTC[A-Z]*MMA[A-Z0-9_.]*|HMMA[A-Z0-9_.]*)")
print("# SASS evidence of the TMA paths: mnemonic counts per kernel (cuobjdump -sass of the objects linked into\n"
      "# deodr_b200/libdeodr_b200.so).  UBLKCP.S.G = cp.async.bulk (per-tile record lists, k_tile_z); UTMASTG.2D =\n"
      "# cp.async.bulk.tensor store (image tile, fused k_tile_z); UTMALDG.2D = cp.async.bulk.tensor loads (image_b /\n"
      "# owner / z tiles, k_raster_bwd); SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK.TRANS64.TRYWAIT = mbarrier expect_tx / wait.\n"
      "# No tensor-core mnemonics (UTC*MMA / HMMA) anywhere: there is no contraction on this path.\n")
for obj in ("kernels.o", "kernels_bwd.o", "scene_ops.o"):
    path = os.path.join(ROOT, "build", obj)
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            name = name.replace("void ", "")
            per.setdefault(name, collections.Counter())
            continue
        if name:
            for hit in WANT.findall(line):
                per[name][hit] += 1
    print(f"## build/{obj}")
    any_hit = False
    for k, c in per.items():
        if c:
            any_hit = True
            print(f"{k:<44}  " + "  ".join(f"{m} x{n}" for m, n in sorted(c.items())))
    if not any_hit:
        print("(none)")
    print()
