"""The Cython binding a maintainer of the reference would add (bindings/differentiable_renderer_b200.pyx, INTEGRATION.md
section 2): cythonised, compiled against include/deodr_b200.h and linked with libdeodr_b200.so here; CPU: it exports the
two names of the reference's pyx and fails like it / fails loudly; -m gpu: its results against the oracle."""
import importlib.util
import os
import subprocess
import sys
import sysconfig

import numpy as np
import pytest
from conftest import ROOT

from deodr_b200.differentiable_renderer import Scene2D
from deodr_b200.scenes import dense_image_b, soup_scene

FIELDS = ("faces", "faces_uv", "ij", "depths", "textured", "uv", "shade", "colors", "shaded", "edgeflags", "height",
          "width", "nb_colors", "texture", "background_image", "background_color", "clockwise", "backface_culling",
          "strict_edge", "perspective_correct", "integer_pixel_centers")


@pytest.fixture(scope="module")
def binding(build_native, tmp_path_factory):
    work = tmp_path_factory.mktemp("cython_binding")
    src = os.path.join(ROOT, "bindings", "differentiable_renderer_b200.pyx")
    c_file = str(work / "differentiable_renderer_b200.c")
    subprocess.run([sys.executable, "-m", "cython", "-3", src, "-o", c_file], check=True, capture_output=True)
    so = str(work / ("differentiable_renderer_b200" + sysconfig.get_config_var("EXT_SUFFIX")))
    libdir = os.path.join(ROOT, "deodr_b200")
    subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-I", sysconfig.get_paths()["include"], "-I",
                    os.path.join(ROOT, "include"), c_file, "-L", libdir, "-ldeodr_b200", f"-Wl,-rpath,{libdir}",
                    "-o", so], check=True, capture_output=True)
    spec = importlib.util.spec_from_file_location("differentiable_renderer_b200", so)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


@pytest.fixture()
def scene(texture):
    np.random.seed(2)
    s = soup_scene(clockwise=False, texture=texture)
    s2 = Scene2D(**{k: getattr(s, k) for k in FIELDS})
    s2.clear_gradients()
    return s2


def test_binding_has_the_reference_names_and_its_failure_modes(binding, scene):
    import torch

    assert callable(binding.renderSceneCpp) and callable(binding.renderSceneBCpp)
    image, z = np.empty((scene.height, scene.width, 3)), np.empty((scene.height, scene.width))
    with pytest.raises(ValueError):  # np.ndarray[double, ndim=3, mode="c"] of the pyx (:52): typed buffer
        binding.renderSceneCpp(scene, 1.0, image.astype(np.float32), z)
    keep = scene.shade
    scene.shade = scene.shade[:-1]
    with pytest.raises(AssertionError):  # the pyx's shape asserts (:61-114) before anything native
        binding.renderSceneCpp(scene, 1.0, image, z)
    scene.shade = keep
    if not torch.cuda.is_available():
        with pytest.raises(binding.DeodrB200Error, match="no CPU fallback"):
            binding.renderSceneCpp(scene, 1.0, image, z)


@pytest.mark.gpu
def test_binding_renders_and_differentiates_like_the_reference(binding, scene, checker):
    image_ref, z_ref = checker.render(scene, 1.0)
    image, z = np.empty_like(image_ref), np.empty_like(z_ref)
    binding.renderSceneCpp(scene, 1.0, image, z)
    assert np.array_equal(z, z_ref) and np.abs(image - image_ref).max() <= 1e-6
    image_b = dense_image_b(image_ref)
    ref = checker.render_b(scene, 1.0, image_ref, z_ref, image_b)
    binding.renderSceneBCpp(scene, 1.0, image, z, image_b)
    for name in ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b"):
        got = getattr(scene, name)
        assert got.shape == ref[name].shape
        assert np.abs(got - ref[name]).max() <= 5e-5 * np.abs(ref[name]).max() + 1e-6, name
