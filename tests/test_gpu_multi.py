"""GPU, >= 2 devices (skipped on a one-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
the product multi-GPU module on the device it is written for.  `ViewShardedBackward` (deodr_b200/distributed.py) with the
sm_100a renderer as `render_view`, one process per GPU over NCCL, against the sequential `+=` accumulation of the same
views by the compiled reference on the CPU (deodr/mesh_fitter.py:511-549).  The single-process CPU twin of this test
(gloo, oracle as renderer) is tests/test_distributed.py."""
import os
import socket

import numpy as np
import pytest
import torch

N_VIEWS = 5  # odd on purpose: the ranks own different numbers of views
SIZE = dict(n=40, width=192, height=160)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _scene(view):
    from deodr_b200.scenes import torus_scene

    return torus_scene(SIZE["n"], SIZE["width"], SIZE["height"], view=view, n_views=N_VIEWS)


def _image_b(view):
    rng = np.random.default_rng(100 + view)
    return rng.random((SIZE["height"], SIZE["width"], 3)) * 2 - 1


def _worker(rank, world, port, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist

    from deodr_b200.distributed import ViewShardedBackward, views_of_rank
    from deodr_b200.renderer import DeviceScene, Renderer

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    renderer = Renderer(rank)

    def render_view(view):  # forward + adjoint of one view on THIS rank's GPU, through the C-ABI
        ds = DeviceScene(_scene(view), dev)
        fwd = renderer.render(ds, 1.0)
        return renderer.render_b(ds, 1.0, fwd, torch.from_numpy(_image_b(view)).to(dev))

    template = DeviceScene(_scene(0), dev).zero_grads()
    sharded = ViewShardedBackward(N_VIEWS, render_view, shared_like={k: template[k] for k in
                                                                    ("colors_b", "uv_b", "shade_b", "texture_b")})
    total, per_view = sharded.step()
    torch.cuda.synchronize()
    assert sorted(per_view) == views_of_rank(N_VIEWS, rank, world)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{k: v.cpu().numpy() for k, v in total.items()},
             **{f"ij_b_{v}": g["ij_b"].cpu().numpy() for v, g in per_view.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_view_sharded_backward_over_nccl_matches_sequential_reference(tmp_path, checker):
    world = min(torch.cuda.device_count(), 2) if torch.cuda.is_available() else 0
    if world < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # sequential accumulation by the reference (CPU), view after view
    want = None
    want_ij = {}
    for view in range(N_VIEWS):
        scene = _scene(view)
        image, z = checker.render(scene, 1.0)
        g = checker.render_b(scene, 1.0, image, z, _image_b(view))
        want_ij[view] = g["ij_b"]
        if want is None:
            want = {k: g[k].astype(np.float64).copy() for k in ("colors_b", "uv_b", "shade_b", "texture_b")}
        else:
            for k in want:
                want[k] += g[k]
    ranks = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    for k, ref in want.items():
        if ref.size == 0:
            continue
        tol = 5e-5 * np.abs(ref).max() + 1e-6
        for r in range(world):  # every rank holds the sum over ALL views after the one all-reduce
            assert np.abs(ranks[r][k] - ref).max() <= tol, (k, r)
    seen = set()
    for r in range(world):
        for name in ranks[r].files:
            if name.startswith("ij_b_"):
                view = int(name[5:])
                seen.add(view)
                ref = want_ij[view]
                assert np.abs(ranks[r][name] - ref).max() <= 5e-5 * np.abs(ref).max() + 1e-6, name
    assert seen == set(range(N_VIEWS))
