"""Camera-sharded multi-view training step.

The only data-parallel axis of the path is "independent cameras": every stage is per-camera and only the
per-Gaussian *gradients* couple views.  Gaussians are replicated on every GPU, rank r renders views
{r, r+G, ...} of the step's batch, and ONE ``all_reduce(sum)`` over a single flat fp32 gradient buffer per step
makes the replicas agree (NCCL over NVLink on the B200 box, gloo in the CPU tests).  Densification statistics
(per-Gaussian |grad2d| sums, visibility counts, max radii) are per-camera as well and need the same treatment
or the replicas diverge when they densify (last hard part).

Nothing here touches kernels: it is host-side plumbing over ``torch.distributed``.
"""

from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Views rendered by `rank`: round-robin so every rank gets ceil or floor(n_views / world) of them."""
    return list(range(rank, n_views, world))


class FlatGradBucket:
    """One contiguous fp32 buffer holding the gradients of all Gaussian parameters, in a fixed order.

    ``pack()`` copies ``p.grad`` of every parameter into the buffer (missing gradients count as zero),
    ``all_reduce()`` sums it across ranks, ``unpack()`` points every ``p.grad`` at its slice of the buffer
    (views, no copy).  Payload for 3 M Gaussians: 168 MB at sh0 (14 floats), 708 MB at sh3 (59 floats).
    """

    def __init__(self, params: Dict[str, Tensor]):
        self.names = list(params.keys())
        self.shapes = [tuple(params[k].shape) for k in self.names]
        self.sizes = [int(params[k].numel()) for k in self.names]
        self.offsets = [0]
        for s in self.sizes:
            self.offsets.append(self.offsets[-1] + s)
        first = params[self.names[0]]
        self.buffer = torch.zeros(self.offsets[-1], dtype=torch.float32, device=first.device)

    def slices(self) -> Dict[str, Tensor]:
        return {k: self.buffer[o:o + n].view(shape)
                for k, o, n, shape in zip(self.names, self.offsets, self.sizes, self.shapes)}

    @torch.no_grad()
    def pack(self, params: Dict[str, Tensor]):
        views = self.slices()
        for k in self.names:
            g = params[k].grad
            if g is None:
                views[k].zero_()
            elif g.data_ptr() != views[k].data_ptr():
                views[k].copy_(g)
        return self.buffer

    def all_reduce(self, group=None, async_op: bool = False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None

    @torch.no_grad()
    def unpack(self, params: Dict[str, Tensor]):
        views = self.slices()
        for k in self.names:
            params[k].grad = views[k]


@torch.no_grad()
def sync_strategy_state(state: Dict, group=None):
    """Make the densification statistics identical on all ranks: sums for grad2d / count, max for radii."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for key, op in (("grad2d", dist.ReduceOp.SUM), ("count", dist.ReduceOp.SUM), ("radii", dist.ReduceOp.MAX)):
        t = state.get(key)
        if isinstance(t, Tensor):
            dist.all_reduce(t, op=op, group=group)


def global_batch_loss_scale(n_views: int) -> float:
    """Each rank averages the loss over the whole batch so that the summed gradient equals the gradient of the
    mean loss over all views of the step."""
    return 1.0 / float(n_views)
