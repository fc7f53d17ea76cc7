#!/usr/bin/env python
"""Compiles one .cu of deodr_b200/csrc for sm_100a with `-Xptxas -v` and prints one line per kernel:
registers, stack frame, spill stores / loads, shared memory.  (Development aid; also how profiles/*_ptxas.txt is made.)"""
import re
import subprocess
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "deodr_b200/csrc/kernels.cu"
extra = sys.argv[2:]
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
       "-Xptxas", "-v", "-c", "-o", "/dev/null", src] + extra
out = subprocess.run(cmd, capture_output=True, text=True)
text = out.stderr
if out.returncode:
    print(text)
    sys.exit(out.returncode)
names = []
cur = None
rows = []
for line in text.splitlines():
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        frame = tuple(int(x) for x in m.groups())
        continue
    m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", line)
    if m and cur:
        rows.append((cur, int(m.group(1)), frame, m.group(3) or "0"))
        cur = None
dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"{'kernel':70s} regs stack spill_st spill_ld smem")
for (mangled, regs, frame, smem), d in zip(rows, dem):
    short = re.sub(r"\(.*", "", d).replace("void ", "")
    print(f"{short:70s} {regs:4d} {frame[0]:5d} {frame[1]:8d} {frame[2]:8d} {smem}")
