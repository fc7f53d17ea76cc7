"""Multi-GPU: shard the batch-of-views axis, one NCCL all-reduce of the shared-parameter gradients per step.

The hot path shards only where the reference loops: ``MeshRGBFitterWithPoseMultiFrame`` renders its frames one after
the other and ``+=``-accumulates the gradients of the parameters the frames share (deodr/mesh_fitter.py:511-549).
Views are independent (own ``ij`` / ``depths`` / ``edgeflags`` / framebuffer / ``image_b``), so view ``i`` goes to rank
``i mod world`` with no data-path collective; the only exchange is the sum of the shared gradients (vertex colours,
``uv``, ``shade``, texture), folded into ONE flat buffer and ONE ``all_reduce`` (latency-bound, a few MB at most).
Per-view gradients (``ij_b``) stay on the rank that owns the view.

One process per GPU (``torchrun``); backend ``nccl`` on GPUs (NVLink 5 / NVSwitch), ``gloo`` in the CPU tests.
"""

from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

SHARED_GRADS = ("colors_b", "uv_b", "shade_b", "texture_b")


def views_of_rank(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin partition of the view axis: view i -> rank i mod world."""
    return list(range(rank, n_views, world))


def allreduce_flat(tensors: Sequence[torch.Tensor], group=None, async_op: bool = False):
    """Sum ``tensors`` across ranks with a single collective: pack -> all_reduce -> unpack (in place).

    A caller that already keeps its shared gradients in ONE contiguous buffer (``bench.py`` does: the gradient slots are
    views of it) passes that buffer alone: no packing, and ``async_op=True`` returns the collective's work handle."""
    tensors = [t for t in tensors if t is not None and t.numel() > 0]
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    dtype = tensors[0].dtype
    assert all(t.dtype == dtype and t.device == tensors[0].device for t in tensors)
    if len(tensors) == 1 and tensors[0].is_contiguous():
        return dist.all_reduce(tensors[0], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    offset = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[offset:offset + n].view_as(t))
        offset += n


class ViewShardedBackward:
    """Runs ``render_view(view_index) -> dict of gradient tensors`` for the views of this rank, accumulates the shared
    gradients over the local views (the ``+=`` of deodr/mesh_fitter.py:518-527) and all-reduces them once.

    Every rank must issue the SAME collective.  The list of tensors that go into it therefore comes from
    ``shared_like`` - rank-independent templates ``{name: tensor}`` (the shared parameters themselves: the mesh is
    replicated) - and a rank without views, or whose ``render_view`` leaves a shared gradient out, contributes zeros of
    the template's shape.  Without templates the names are taken from what ``render_view`` returned, which is only
    rank-independent when every rank has at least one view: the constructor refuses ``n_views < world`` in that case.
    """

    def __init__(self, n_views: int, render_view: Callable[[int], Dict[str, torch.Tensor]],
                 shared: Iterable[str] = SHARED_GRADS, group=None,
                 shared_like: Optional[Dict[str, torch.Tensor]] = None):
        self.n_views = n_views
        self.render_view = render_view
        self.shared = tuple(shared)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local_views = views_of_rank(n_views, self.rank, self.world)
        self.shared_like = dict(shared_like) if shared_like is not None else None
        if self.shared_like is None and n_views < self.world:
            raise ValueError(
                f"{n_views} views on {self.world} ranks: some ranks have no view, so the tensors of the all-reduce must "
                "be described by `shared_like` (rank-independent templates of the shared gradients)")

    def step(self):
        """-> (shared gradients summed over ALL views of ALL ranks, {view: per-view gradients of the local views})."""
        total: Dict[str, torch.Tensor] = {}
        per_view: Dict[int, Dict[str, torch.Tensor]] = {}
        for view in self.local_views:
            grads = self.render_view(view)
            per_view[view] = {k: v for k, v in grads.items() if k not in self.shared}
            for name in self.shared:
                if name not in grads or grads[name] is None:
                    continue
                if name in total:
                    total[name] += grads[name]
                else:
                    total[name] = grads[name].clone()
        if self.shared_like is not None:
            # rank-independent list: zeros where this rank has nothing to add
            names = [n for n in self.shared if n in self.shared_like and self.shared_like[n] is not None]
            for n in names:
                like = self.shared_like[n]
                if n not in total:
                    total[n] = torch.zeros_like(like)
                elif total[n].shape != like.shape:
                    raise ValueError(f"shared gradient {n}: shape {tuple(total[n].shape)} != template {tuple(like.shape)}")
        else:
            names = [n for n in self.shared if n in total]
        if self.world > 1:
            allreduce_flat([total[n] for n in names], self.group)
        return total, per_view
