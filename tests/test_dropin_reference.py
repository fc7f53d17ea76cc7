"""GPU (-m gpu): drop-in acceptance against the REAL reference package.

baseline/_ref/ holds a verbatim, git-ignored copy of /root/reference/{deodr,tests} (scripts/stage_reference.py, run by
__graft_entry__.build() in the build container; it travels to the GPU box with the tree).  tests/dropin/runner.py binds
``deodr_b200.differentiable_renderer_cython`` as ``deodr.differentiable_renderer_cython`` - the one-line swap of
INTEGRATION.md - and then runs the reference's OWN, unmodified test files and example fitters.

* convention tests + the pinned soup render: the reference's pytest files pass as they are;
* fitting loops (losses the reference pins with `==` / 1e-5 on a float64 CPU path): the unmodified example code runs
  on the sm_100a path and its loss / energy trajectory is compared with the one the reference's own extension produced
  in the build container (tests/golden/dropin_reference.json, made by tests/golden/make_dropin_golden.py).
  Stated tolerances: the soup fits stay within 2e-3 over the whole 50-iteration trajectory and within 2e-5 over the
  first 4 iterations (measured on the B200: 1e-8 throughout).  The hand fits are chaotic 50-step momentum descents:
  the REFERENCE ITSELF, with every image / gradient it returns perturbed by a relative 1e-7 (fp32-sized;
  scripts/dropin_sensitivity.py, 16 seeds per fit), leaves its own trajectory by up to 3.5e-2 (median 2.5e-2 for the
  depth fits, 1.5e-2 for the RGB fit), its final energy by up to 3.5e-2, and - when a pixel of the depth image changes
  owner - the first four iterations by up to 3e-5.  Hence: first 4 iterations within 1e-4 (measured on the B200: 9e-9),
  whole trajectory and final energy within 1e-1 = 3 x the largest excursion of the perturbed reference (measured on the
  B200: 1.3e-2 / 1.5e-2 / 2.0e-2), and the fit must end as low as the reference's.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

STAGED = os.path.join(ROOT, "baseline", "_ref")
RUNNER = os.path.join(ROOT, "tests", "dropin", "runner.py")


@pytest.fixture(scope="module")
def staged(build_native):
    if not os.path.isdir(os.path.join(STAGED, "deodr")):
        pytest.fail("baseline/_ref/deodr is missing: run scripts/stage_reference.py where /root/reference exists")
    return STAGED


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(GOLDEN, "dropin_reference.json")))


def _run(*args, timeout=900):
    out = subprocess.run([sys.executable, RUNNER, "b200", *[str(a) for a in args]], capture_output=True, text=True,
                         timeout=timeout)
    return out


def _result(out, tag=None):
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    scratch = os.path.join(ROOT, "gpurun_out")
    if tag and os.path.isdir(scratch):  # keep the trajectory next to the other artefacts of a GPU session
        json.dump(res, open(os.path.join(scratch, f"dropin_{tag}.json"), "w"))
    return res


def test_reference_convention_tests_pass_unmodified(staged):
    tests = os.path.join(staged, "tests")
    out = _run("pytest", os.path.join(tests, "test_pixel_center_coordinates.py"),
               os.path.join(tests, "test_texture_coordinates.py"),
               os.path.join(tests, "test_render_mesh.py") + "::test_render_mesh_triangle_soup")
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "3 passed" in out.stdout


@pytest.mark.parametrize("clockwise", [0, 1])
@pytest.mark.parametrize("antialiase_error", [0, 1])
def test_reference_soup_fitting_example(clockwise, antialiase_error, staged, golden):
    """deodr/examples/triangle_soup_fitting.py:run - the body of the reference's tests/test_triangle_soup_fitting.py
    (runs 1-4: both windings, with and without antialiase_error)."""
    ref = golden["soup"][f"cw{clockwise}_err{antialiase_error}"]["losses"]
    got = _result(_run("soup", clockwise, antialiase_error, 50), f"soup_cw{clockwise}_err{antialiase_error}")["losses"]
    assert len(got) == len(ref) == 50
    rel = np.abs(np.array(got) - np.array(ref)) / np.array(ref)
    assert rel[:4].max() <= 2e-5, rel[:4]
    assert rel.max() <= 2e-3, (rel.argmax(), rel.max())
    assert got[-1] < 0.5 * got[0]  # and it fits


@pytest.mark.parametrize("library", ["none", "pytorch"])
def test_reference_depth_hand_fitting_example(library, staged, golden):
    """deodr/examples/depth_image_hand_fitting.py:run with MeshDepthFitter (numpy) and its PyTorch twin - the body of
    the reference's tests/test_depth_image_hand_fitting.py."""
    ref = np.array(golden["hand_depth"][library]["energies"])
    got = np.array(_result(_run("hand_depth", library, 50), f"hand_depth_{library}")["energies"])
    rel = np.abs(got - ref) / ref
    assert rel[:4].max() <= 1e-4, rel[:4]
    # The fit is a 50-step momentum descent over a piecewise-smooth energy: the reference's own test lists final
    # energies that differ by 4e-5 between fp64 platforms (251.3271 / 251.3165), and a 1e-7 perturbation of the
    # reference's own gradients moves its trajectory by up to 3.5e-2 (module docstring).  The trajectory must stay a
    # descent of the same quality.
    assert rel.max() <= 1e-1, (rel.argmax(), rel.max())
    assert abs(got[49] - 251.327) / 251.327 <= 1e-1  # the value the reference's test pins (1e-5 on its own path)
    assert got[49] < 0.15 * got[0]


def test_reference_rgb_hand_fitting_example(staged, golden):
    ref = np.array(golden["hand_rgb"]["none"]["energies"])
    got = np.array(_result(_run("hand_rgb", "none", 50), "hand_rgb_none")["energies"])
    rel = np.abs(got - ref) / ref
    assert rel[:4].max() <= 1e-4, rel[:4]
    assert rel.max() <= 1e-1, (rel.argmax(), rel.max())
    assert got[49] < 0.6 * got[0]  # the reference's fit ends at 0.54 of its first energy
