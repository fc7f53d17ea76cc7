"""Development aid: per-source-line instruction counts of one kernel from an ncu report.

The ncu CSV source page is SASS-only; nvdisasm -g gives the file:line of every SASS instruction of the same cubin.
Both list the kernel's instructions in address order, so they are joined by position.

    python scripts/ncu_lines.py <report.ncu-rep> <cubin> <kernel-regex> <mangled-kernel-name-substring>
"""
import csv
import os
import collections
import re
import subprocess
import sys

rep, cubin, kregex, mangled = sys.argv[1:5]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kregex}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
i_exec, i_src = hdr.index("Instructions Executed"), hdr.index("Source")
i_samp = hdr.index("# Samples")
sass = []
for r in rows[start + 1:]:
    if len(r) != len(hdr):
        break  # next kernel instance
    sass.append((r[i_src].strip(), float(r[i_exec] or 0), float(r[i_samp] or 0)))
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
beg = next(i for i, l in enumerate(dis) if ".text." in l and mangled in l and ".section" in l)
lines, cur = [], ("?", 0)
for l in dis[beg + 1:]:
    if l.startswith("\t.section") or l.startswith("//-----"):
        if lines:
            break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print(f"{len(sass)} profiled instructions, {len(lines)} disassembled")
n = min(len(sass), len(lines))
agg, samp = collections.Counter(), collections.Counter()
for (src, ex, sm), loc in zip(sass[:n], lines[:n]):
    agg[loc] += ex
    samp[loc] += sm
tot, tots = sum(agg.values()), sum(samp.values())
print(f"total warp-instructions {tot:.3g}, samples {tots:.0f}")
by_file = collections.Counter()
for (f, ln), v in agg.items():
    by_file[f] += v
print({k: f"{100 * v / tot:.1f}%" for k, v in by_file.most_common()})
order = samp.most_common(40) if len(sys.argv) > 5 else agg.most_common(40)
for (f, ln), _ in order:
    v = agg[(f, ln)]
    text = ""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deodr_b200", "csrc", f)
    rev = os.environ.get("SRC_REV")  # report taken from an older commit: show that commit's source text
    if rev:
        src_lines = subprocess.run(["git", "show", f"{rev}:deodr_b200/csrc/{f}"], capture_output=True, text=True).stdout.splitlines()
        text = src_lines[ln - 1].strip()[:100] if 0 < ln <= len(src_lines) else ""
    elif os.path.exists(path):
        src_lines = open(path).read().splitlines()
        text = src_lines[ln - 1].strip()[:100] if 0 < ln <= len(src_lines) else ""
    print(f"{f}:{ln:<5d} inst {100 * v / tot:5.1f}%  stall {100 * samp[(f, ln)] / max(tots, 1):5.1f}%  | {text}")
