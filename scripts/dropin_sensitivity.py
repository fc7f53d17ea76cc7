"""Development aid (build container, CPU only): how far does an fp32-sized perturbation drive the reference's own example
fits apart?  Runs the REFERENCE (its own Cython extension in a scratch copy, see tests/golden/make_dropin_golden.py) with
every image / gradient it returns multiplied by 1 + eps * N(0, 1) (tests/dropin/runner.py, impl ref_noise) and compares
the energy trajectories with the unperturbed ones of tests/golden/dropin_reference.json.

    DEODR_STAGED_REFERENCE=/tmp/dref python scripts/dropin_sensitivity.py [seeds=8] [eps=1e-7]

The tolerances of tests/test_dropin_reference.py on the hand fits (chaotic 50-step momentum descents) come from this."""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "dropin", "runner.py")
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "dropin_reference.json")))
seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eps = sys.argv[2] if len(sys.argv) > 2 else "1e-7"


def run(job):
    mode, lib, seed = job
    env = dict(os.environ, DEODR_NOISE_SEED=str(seed), DEODR_NOISE_EPS=eps)
    args = [mode, lib[2], lib[7], "50"] if mode == "soup" else [mode, lib, "50"]  # soup: lib = "cw<0|1>_err<0|1>"
    out = subprocess.run([sys.executable, RUNNER, "ref_noise", *args], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    key = "losses" if mode == "soup" else "energies"
    got = np.array(json.loads(line[len("RESULT "):])[key])
    ref = np.array(golden[mode][lib][key])
    rel = np.abs(got - ref) / ref
    return mode, lib, seed, float(rel[:4].max()), float(rel.max()), int(rel.argmax()), float(rel[-1])


FITS = (("hand_depth", "none"), ("hand_depth", "pytorch"), ("hand_rgb", "none"), ("soup", "cw0_err0"), ("soup", "cw0_err1"),
        ("soup", "cw1_err0"), ("soup", "cw1_err1"))
if os.environ.get("FITS"):  # e.g. FITS=soup
    FITS = tuple(f for f in FITS if f[0] in os.environ["FITS"].split(","))
jobs = [(m, l, s) for m, l in FITS for s in range(1, seeds + 1)]
with ThreadPoolExecutor(max_workers=int(os.environ.get("JOBS", "3"))) as pool:
    results = list(pool.map(run, jobs))
for mode, lib in FITS:
    rows = [r for r in results if r[0] == mode and r[1] == lib]
    worst = np.array([r[4] for r in rows])
    print(f"{mode}/{lib}: eps {eps}, {len(rows)} seeds: first-4 max {max(r[3] for r in rows):.1e}; trajectory max: "
          f"median {np.median(worst):.2e}, largest {worst.max():.2e}; final-energy deviation largest {max(r[6] for r in rows):.2e}")
    print("   per seed:", " ".join(f"{w:.1e}@{r[5]}" for w, r in zip(worst, rows)))
