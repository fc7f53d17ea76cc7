"""GPU (-m gpu): the reference-facing Python surface (Scene2D / renderSceneCpp / autograd Functions) on the CUDA path.
The first two tests restate the reference's own convention tests with the same inputs and assertions."""
import numpy as np
import pytest

from deodr_b200.scenes import soup_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def need_gpu(build_native):
    import torch

    assert torch.cuda.is_available()


def test_upper_left_pixel_center_coordinates():
    """Same scene and assertion as the reference tests/test_pixel_center_coordinates.py:8-103."""
    from deodr_b200.differentiable_renderer import Scene2D

    height, width, eps = 4, 3, 0.001
    corners = [(0, 0), (width - 1, 0), (0, height - 1), (width - 1, height - 1)]
    for integer_pixel_centers in (False, True):
        shift = 0.0 if integer_pixel_centers else 0.5
        for cx, cy in corners:
            ij = np.array([[-eps, -eps], [-eps, eps], [eps, -eps]]) + np.array((cx + shift, cy + shift))
            scene = Scene2D(
                ij=ij, faces=np.array([[0, 2, 1]], dtype=np.uint32), faces_uv=np.array([[0, 2, 1]], dtype=np.uint32),
                uv=np.zeros((3, 2), dtype=bool), texture=np.ones((2, 2, 1)), height=height, width=width, nb_colors=1,
                background_image=None, background_color=np.array([0]), depths=np.array([1, 1, 1]),
                textured=np.array([0], dtype=bool), shade=np.array([1, 1, 1]), colors=np.array([[1], [1], [1]]),
                shaded=np.array([0], dtype=bool), edgeflags=np.zeros((1, 3), dtype=bool), strict_edge=False,
                perspective_correct=True, clockwise=True, integer_pixel_centers=integer_pixel_centers)
            image, _ = scene.render(sigma=0)
            expected = np.zeros((height, width, 1))
            expected[cy, cx, 0] = 1
            assert np.allclose(expected, image)


def test_texture_coordinates():
    """Same scene and assertions as the reference tests/test_texture_coordinates.py:8-71."""
    from deodr_b200.differentiable_renderer import Scene2D

    texture = np.array([[[1, 0, 0], [0, 1, 0]], [[0, 0, 1], [1, 1, 1]]], dtype=np.float64)
    for clockwise in (False, True):
        order = [0, 2, 1] if clockwise else [0, 1, 2]
        scene = Scene2D(
            ij=np.array([[1, 1], [1, 15], [15, 1]]), faces=np.array([order], dtype=np.uint32),
            faces_uv=np.array([order], dtype=np.uint32), uv=np.array([[0, 0], [1, 0], [0, 1]]), texture=texture,
            height=40, width=60, nb_colors=3, background_image=None, background_color=np.array([0, 0, 0]),
            depths=np.array([1, 1, 1]), textured=np.array([1], dtype=bool), shade=np.array([1, 1, 1]),
            colors=np.eye(3), shaded=np.array([1], dtype=bool), edgeflags=np.zeros((1, 3), dtype=bool),
            strict_edge=False, perspective_correct=True, clockwise=clockwise)
        image, _ = scene.render(sigma=0)
        assert np.allclose(image[0, :, :], [0, 0, 0])
        assert np.allclose(image[:, 0, :], [0, 0, 0])
        assert np.allclose(image[1, 1, :], [1, 0, 0])
        assert np.allclose(image[15, 1, :], [0, 1, 0])
        assert np.allclose(image[1, 15, :], [0, 0, 1])


def _scene2d(arrays, **over):
    from deodr_b200.differentiable_renderer import Scene2D

    keys = ("faces", "faces_uv", "ij", "depths", "textured", "uv", "shade", "colors", "shaded", "edgeflags", "height",
            "width", "nb_colors", "texture", "background_image", "background_color", "clockwise", "backface_culling",
            "strict_edge", "perspective_correct", "integer_pixel_centers")
    kw = {k: getattr(arrays, k) for k in keys}
    kw.update(over)
    return Scene2D(**kw)


def test_scene2d_render_compare_and_backward_matches_reference(checker, texture):
    """numpy path of the reference call stack (B): Scene2D.render_compare_and_backward."""
    np.random.seed(2)
    arrays = soup_scene(clockwise=True, texture=texture)
    scene = _scene2d(arrays)
    obs = np.random.default_rng(1).random((200, 200, 3))
    image, z, err_buffer, err = scene.render_compare_and_backward(obs, sigma=1.0)
    image_ref, z_ref = checker.render(arrays, 1.0)
    assert image.dtype == np.float64 and np.array_equal(z, z_ref) and np.abs(image - image_ref).max() < 1e-6
    ref = checker.render_b(arrays, 1.0, image_ref, z_ref, 2 * (image_ref - obs))
    assert abs(err - float(np.sum((image_ref - obs) ** 2))) < 1e-3
    for name in ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b"):
        got = getattr(scene, name)
        assert got.dtype == np.float64
        assert np.abs(got - ref[name]).max() <= 5e-5 * np.abs(ref[name]).max() + 1e-6, name
    # gradients accumulate across calls unless cleared (reference semantics: += into scene.*_b)
    before = scene.ij_b.copy()
    scene.render_compare_and_backward(obs, sigma=1.0, clear_gradients=False)
    assert np.allclose(scene.ij_b, 2 * before, rtol=1e-4, atol=1e-5)
    # backward requires culling, like the reference (deodr/differentiable_renderer.py:666-672)
    scene.backface_culling = False
    with pytest.raises(BaseException, match="backface_culling"):
        scene.render_backward(np.zeros_like(image))


def test_soup_fitting_loop_descends(checker, texture):
    """The fitting loop of the reference example (examples/triangle_soup_fitting.py:150-173) on the GPU path:
    identical iteration-0 loss, and losses staying within 1e-6 relative of the reference's for the first steps."""
    np.random.seed(2)
    gt = soup_scene(clockwise=False, texture=texture)
    target, _ = checker.render(gt, 1.0)
    n = len(gt.depths)
    start = gt.ij + np.random.randn(n, 2) * 10
    uv = np.minimum(np.maximum(gt.uv, 0), np.array(gt.texture.shape[:2]) - 1)
    gpu_scene = _scene2d(gt, ij=start.copy(), uv=uv)
    cpu = soup_scene.__globals__["SceneArrays"](**{k: getattr(gpu_scene, k) for k in (
        "faces", "faces_uv", "ij", "depths", "textured", "uv", "shade", "colors", "shaded", "edgeflags", "height",
        "width", "nb_colors", "texture", "background_image", "background_color", "clockwise", "backface_culling")})
    speed_g, speed_c = np.zeros((n, 2)), np.zeros((n, 2))
    for it in range(4):
        _, _, _, loss_g = gpu_scene.render_compare_and_backward(sigma=1, obs=target)
        image, z = checker.render(cpu, 1.0)
        loss_c = float(np.sum((image - target) ** 2))
        assert abs(loss_g - loss_c) <= 2e-5 * loss_c, (it, loss_g, loss_c)
        grads = checker.render_b(cpu, 1.0, image, z, 2 * (image - target))
        speed_g = 0.8 * speed_g - gpu_scene.ij_b * 0.01
        speed_c = 0.8 * speed_c - grads["ij_b"] * 0.01
        gpu_scene.ij = gpu_scene.ij + speed_g
        cpu.ij = cpu.ij + speed_c


def test_ffi_shim_semantics(texture):
    from deodr_b200 import differentiable_renderer_cython as shim

    np.random.seed(2)
    scene = _scene2d(soup_scene(texture=texture))
    image = np.full((200, 200, 3), 7.0)
    z = np.full((200, 200), 7.0)
    shim.renderSceneCpp(scene, 1.0, image, z)
    assert not (image == 7.0).all() and np.isinf(z).any()
    with pytest.raises(AssertionError):
        shim.renderSceneCpp(scene, 1.0, np.zeros((100, 200, 3)), z)
    with pytest.raises(ValueError):
        shim.renderSceneCpp(scene, 1.0, image.astype(np.float32), z)
    scene.faces = scene.faces.copy()
    scene.faces[0, 0] = 10**6
    with pytest.raises(AssertionError):
        shim.renderSceneCpp(scene, 1.0, image, z)
    with pytest.raises(Exception, match="faces"):
        shim.renderSceneCpp(scene, 1.0, image, z, check_valid=0)


def test_autograd_functions(checker, texture):
    import torch

    from deodr_b200.pytorch import CudaDifferentiableRender2D, TorchDifferentiableRender2D
    from deodr_b200.renderer import DeviceScene

    np.random.seed(2)
    arrays = soup_scene(clockwise=True, textured_ratio=0.0, texture=texture)
    image_ref, z_ref = checker.render(arrays, 1.0)
    obs = np.random.default_rng(1).random(image_ref.shape)
    ref = checker.render_b(arrays, 1.0, image_ref, z_ref, 2 * (image_ref - obs))

    # reference-contract Function: CPU float64 tensors, scene object exposing .scene_2d (Scene3DPytorch duck type)
    class Holder:
        pass

    holder = Holder()
    holder.scene_2d = _scene2d(arrays)
    ij = torch.tensor(arrays.ij, requires_grad=True)
    colors = torch.tensor(arrays.colors, requires_grad=True)
    image = TorchDifferentiableRender2D(ij, colors, holder)
    loss = torch.sum((image - torch.tensor(obs)) ** 2)
    loss.backward()
    assert np.abs(ij.grad.numpy() - ref["ij_b"]).max() <= 5e-5 * np.abs(ref["ij_b"]).max() + 1e-6
    assert np.abs(colors.grad.numpy() - ref["colors_b"]).max() <= 5e-5 * np.abs(ref["colors_b"]).max() + 1e-6

    # zero-copy CUDA Function
    ds = DeviceScene(arrays, "cuda:0")
    ij_c = torch.tensor(arrays.ij, device="cuda", requires_grad=True)
    col_c = torch.tensor(arrays.colors, device="cuda", dtype=torch.float32, requires_grad=True)
    image_c = CudaDifferentiableRender2D(ij_c, col_c, ds, 1.0)
    loss_c = torch.sum((image_c - torch.tensor(obs, device="cuda", dtype=torch.float32)) ** 2)
    loss_c.backward()
    assert abs(float(loss_c) - float(loss)) <= 1e-4 * float(loss)
    assert np.abs(ij_c.grad.cpu().numpy() - ref["ij_b"]).max() <= 5e-5 * np.abs(ref["ij_b"]).max() + 1e-5
    assert np.abs(col_c.grad.cpu().numpy() - ref["colors_b"]).max() <= 5e-5 * np.abs(ref["colors_b"]).max() + 1e-5
