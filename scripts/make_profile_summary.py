"""Turns the ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

    python scripts/make_profile_summary.py <tag> <launches.csv> <full.ncu-rep> [bench.json]
"""
import collections
import csv
import json
import subprocess
import sys

tag, launches, report = sys.argv[1:4]
bench = sys.argv[4] if len(sys.argv) > 4 else None
out = [f"# ncu summary {tag}\n"]

rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]
i_name, i_val, i_grid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[1:]:  # one line per (kernel, grid): two launches of one kernel with different grids are different work
    agg.setdefault(r[i_name].split("(")[0][:48] + " grid " + r[i_grid].split(",")[0].strip("( "), []).append(float(r[i_val].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
out.append(f"## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, {len(rows) - 1} launches, "
           f"cold-cache / serialised: compare SHARES)\n")
out.append("| kernel | launches | mean us | share |\n|---|---|---|---|")
for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"| `{n}` | {len(v)} | {sum(v) / len(v) / 1000:.1f} | {100 * sum(v) / tot:.1f}% |")

raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
idx = {k: i for i, k in enumerate(h)}
want = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM rd MB"),
        ("dram__bytes_write.sum", "DRAM wr MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp inst"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %")]
stalls = [k for k in h if "smsp__average_warps_issue_stalled" in k and "per_issue_active" in k]
out.append(f"\n## `ncu --set full --clock-control none` (one launch per kernel; {report.split('/')[-1]})\n")
out.append("| kernel | " + " | ".join(w[1] for w in want) + " | top stalls |\n|---|" + "---|" * (len(want) + 1))
seen = set()
for r in rr[2:]:
    name = r[idx["Kernel Name"]].split("(")[0][:40] + " grid " + r[idx["Grid Size"]].split(",")[0].strip("( ")
    if name in seen:
        continue
    seen.add(name)
    vals = []
    for k, _ in want:
        v = r[idx[k]] if k in idx else ""
        try:
            v = f"{float(v):.4g}"
        except ValueError:
            pass
        vals.append(v)
    st = sorted(((float(r[idx[k]]), k) for k in stalls if r[idx[k]] not in ("", "n/a")), reverse=True)[:3]
    st = ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.1f}"
                   for v, k in st)
    out.append(f"| `{name}` | " + " | ".join(vals) + f" | {st} |")
# per-launch DRAM traffic of every raster kernel -> profiles/ncu_traffic.json (bench.py's roofline.traffic)
phase_of = {"k_tile_z": "tile_z", "k_shade": "shade", "k_edge_fwd": "edge_fwd", "k_small_tri_bwd": "small_tri_bwd",
            "k_interior_bwd": "interior_bwd", "k_raster_bwd": "edge_bwd", "k_bin": "bin", "k_bin_edges": "edge_bin",
            "k_sort_tile_edges": "edge_tile_sort", "k_finalize_edges": "edge_finalize"}
workload = tag.split("_")[1] if "_" in tag else "c5"
traffic = {}
for r in rr[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    if name in phase_of and phase_of[name] not in traffic:
        def mb(k):
            v, unit = float(r[idx[k]].replace(",", "")), rr[1][idx[k]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        traffic[phase_of[name]] = int(mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"))
try:
    all_traffic = json.load(open("profiles/ncu_traffic.json"))
except Exception:
    all_traffic = {}
traffic["step"] = sum(traffic.values())  # every kernel of one forward + adjoint (memsets aside)
all_traffic[workload] = traffic
all_traffic["source"] = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture {report.split('/')[-1]} "
                         f"(summary profiles/{tag}.md, scripts/make_profile_summary.py)")
json.dump(all_traffic, open("profiles/ncu_traffic.json", "w"), indent=1, sort_keys=True)
if bench:
    d = json.loads(open(bench).read().strip().splitlines()[-1])
    out.append("\n## bench.py line of the same build\n\n```json\n" + json.dumps(d, indent=1) + "\n```")
open(f"profiles/{tag}.md", "w").write("\n".join(out) + "\n")
print("wrote", f"profiles/{tag}.md")
