"""CPU, world_size 2 over gloo: view sharding + the single flat all-reduce of the shared gradients reproduce the
sequential `+=` accumulation of the reference multi-frame fitter (deodr/mesh_fitter.py:511-549)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deodr_b200.distributed import ViewShardedBackward, allreduce_flat, views_of_rank

N_VIEWS = 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _view_grads(view):
    """Gradients of one view computed by the CPU oracle (stand-in renderer for this host-logic test)."""
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from deodr_b200.scenes import dense_image_b, torus_scene
    from oracle.oracle import Oracle

    oracle = Oracle("port")
    scene = torus_scene(12, 64, 48, view=view, n_views=N_VIEWS)
    image, z = oracle.render(scene, 1.0)
    g = oracle.render_b(scene, 1.0, image, z, dense_image_b(image, seed=view))
    return {k: torch.from_numpy(g[k]) for k in ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b")}


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded = ViewShardedBackward(N_VIEWS, _view_grads)
    total, per_view = sharded.step()
    assert sorted(per_view) == views_of_rank(N_VIEWS, rank, world)
    a, b = torch.full((3,), float(rank + 1)), torch.full((2, 2), float(10 * (rank + 1)), dtype=torch.float32)
    a = a.to(torch.float32)
    allreduce_flat([a, b])
    assert torch.all(a == 3.0) and torch.all(b == 30.0)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{k: v.numpy() for k, v in total.items()},
             **{f"ij_b_{v}": g["ij_b"].numpy() for v, g in per_view.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_view_partition():
    assert views_of_rank(5, 0, 2) == [0, 2, 4] and views_of_rank(5, 1, 2) == [1, 3]
    assert sorted(sum((views_of_rank(64, r, 8) for r in range(8)), [])) == list(range(64))
    assert views_of_rank(1, 3, 8) == []


def test_sharded_backward_equals_sequential_accumulation(tmp_path, build_native):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    serial = [_view_grads(v) for v in range(N_VIEWS)]
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for name in ("colors_b", "shade_b", "uv_b", "texture_b"):
        expect = sum(g[name] for g in serial).numpy()
        for r in ranks:
            assert np.allclose(r[name], expect, rtol=1e-12, atol=1e-12), name
    for v in range(N_VIEWS):
        owner = ranks[v % world]
        assert np.array_equal(owner[f"ij_b_{v}"], serial[v]["ij_b"].numpy())


def _worker_fewer_views(rank, world, port, out_dir):
    """world = 3, ONE view: ranks 1 and 2 have nothing to render and must still join the collective with zeros."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def render_view(view):
        calls.append(view)
        return {"ij_b": torch.ones(4, 2), "colors_b": torch.full((4, 3), 2.0), "uv_b": torch.full((2, 2), 3.0)}

    like = {"colors_b": torch.empty(4, 3), "uv_b": torch.empty(2, 2), "shade_b": torch.empty(4), "texture_b": torch.empty(0)}
    try:
        ViewShardedBackward(1, render_view)  # no templates: refused instead of hanging later
        raise AssertionError("expected ValueError")
    except ValueError:
        pass
    total, per_view = ViewShardedBackward(1, render_view, shared_like=like).step()
    assert calls == ([0] if rank == 0 else [])
    assert torch.all(total["colors_b"] == 2.0) and torch.all(total["uv_b"] == 3.0) and torch.all(total["shade_b"] == 0.0)
    dist.barrier()
    dist.destroy_process_group()


def test_ranks_without_views_contribute_zeros(tmp_path):
    world = 3
    mp.spawn(_worker_fewer_views, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
