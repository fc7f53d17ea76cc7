"""TEST INFRASTRUCTURE: runs the REFERENCE package's own tests / examples with deodr_b200 swapped in.

    python tests/dropin/runner.py <impl> pytest <pytest args ...>     # the reference's test files, unmodified
    python tests/dropin/runner.py <impl> soup <clockwise 0|1> <antialiase_error 0|1> <iterations>
    python tests/dropin/runner.py <impl> hand_depth <none|pytorch> <iterations>
    python tests/dropin/runner.py <impl> hand_rgb <none|pytorch> <iterations>

<impl> = b200: ``sys.modules['deodr.differentiable_renderer_cython']`` is bound to
``deodr_b200.differentiable_renderer_cython`` BEFORE the reference package is imported - the one-line swap
INTEGRATION.md describes; nothing else of the staged package (baseline/_ref/deodr: a verbatim, git-ignored copy of
/root/reference/deodr made by scripts/stage_reference.py) is touched.  <impl> = ref: no swap; needs the reference's own
Cython extension next to the package (only used in the build container to validate this runner).
<impl> = ref_noise (build container only, scripts/dropin_sensitivity.py): the reference's own extension with every image
and gradient it returns multiplied element-wise by 1 + DEODR_NOISE_EPS * N(0, 1) (seed DEODR_NOISE_SEED) - a stand-in
for fp32 rounding, used to MEASURE how far a perturbation of that size drives the example fits apart.
Results of the example modes are printed as one JSON line prefixed with RESULT.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
STAGED = os.environ.get("DEODR_STAGED_REFERENCE", os.path.join(ROOT, "baseline", "_ref"))


def main():
    impl, mode, args = sys.argv[1], sys.argv[2], sys.argv[3:]
    sys.path[:0] = [os.path.join(HERE, "stubs"), STAGED, ROOT]
    import cv2

    cv2.waitKey = lambda *a, **k: -1
    cv2.imshow = lambda *a, **k: None
    if impl == "b200":
        import deodr_b200.differentiable_renderer_cython as shim

        sys.modules["deodr.differentiable_renderer_cython"] = shim
    import deodr  # the staged reference package

    assert os.path.abspath(deodr.__file__).startswith(os.path.abspath(STAGED)), deodr.__file__
    if impl == "b200":
        from deodr import differentiable_renderer_cython as bound

        assert bound.__name__ == "deodr_b200.differentiable_renderer_cython", bound.__name__
    if impl == "ref_noise":
        import numpy as np
        from deodr import differentiable_renderer_cython as ffi

        rng = np.random.default_rng(int(os.environ.get("DEODR_NOISE_SEED", "0")))
        eps = float(os.environ.get("DEODR_NOISE_EPS", "1e-7"))
        fwd, bwd = ffi.renderSceneCpp, ffi.renderSceneBCpp

        def noisy(a):
            a *= 1.0 + eps * rng.standard_normal(a.shape)

        def render(scene, sigma, image, z_buffer, *a, **k):
            fwd(scene, sigma, image, z_buffer, *a, **k)
            noisy(image)

        def render_b(scene, *a, **k):
            bwd(scene, *a, **k)
            for name in ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b"):
                noisy(getattr(scene, name))

        ffi.renderSceneCpp, ffi.renderSceneBCpp = render, render_b
    if mode == "pytest":
        import pytest

        sys.exit(pytest.main(["-q", "-x", "-p", "no:cacheprovider", "--rootdir", STAGED] + args))
    if mode == "soup":
        from deodr.examples.triangle_soup_fitting import run

        losses, hashes = run(nb_max_iter=int(args[2]), display=False, clockwise=bool(int(args[0])),
                             antialiase_error=bool(int(args[1])))
        print("RESULT " + json.dumps({"losses": losses, "hashes": hashes}))
        return
    if mode in ("hand_depth", "hand_rgb"):
        if mode == "hand_depth":
            from deodr.examples.depth_image_hand_fitting import run
        else:
            from deodr.examples.rgb_image_hand_fitting import run
        energies = run(dl_library=args[0], plot_curves=False, display=False, save_images=False, max_iter=int(args[1]))
        print("RESULT " + json.dumps({"energies": [float(e) for e in energies]}))
        return
    raise SystemExit(f"unknown mode {mode}")


if __name__ == "__main__":
    main()
