"""world_size-2 gloo test of the camera-sharded step's host logic: view sharding,
the flat gradient bucket and its all-reduce, and the densification-statistics sync.  The kernels themselves
need a GPU; here each rank produces a synthetic per-view gradient so the arithmetic of the exchange is
checked exactly."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _view_grad(view: int, shapes):
    g = torch.Generator().manual_seed(1000 + view)
    return {k: torch.randn(s, generator=g) for k, s in shapes.items()}


def _worker(rank, world, port, n_views, out):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root / "collab-splats_b200"))
    from radegs_b200.multiview import FlatGradBucket, shard_views, sync_strategy_state
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = {"means": (7, 3), "quats": (7, 4), "scales": (7, 3), "opacities": (7,), "features_dc": (7, 3),
              "features_rest": (7, 15, 3)}
    params = {k: torch.zeros(s, requires_grad=True) for k, s in shapes.items()}
    bucket = FlatGradBucket(params)
    mine = shard_views(n_views, rank, world)
    for k in shapes:
        params[k].grad = sum((_view_grad(v, shapes)[k] for v in mine), torch.zeros(shapes[k]))
    params["opacities"].grad = None if rank == 1 else params["opacities"].grad   # a missing grad counts as zero
    bucket.pack(params)
    bucket.all_reduce()
    bucket.unpack(params)
    state = {"grad2d": torch.full((7,), float(rank + 1)), "count": torch.ones(7), "radii": torch.full((7,), float(rank)),
             "scene_scale": 1.0}
    sync_strategy_state(state)
    if rank == 0:
        torch.save({"grads": {k: params[k].grad.clone() for k in shapes}, "state": state, "views": mine}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2(tmp_path):
    n_views, world = 5, 2
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_views, out), nprocs=world, join=True)
    res = torch.load(out)
    shapes = {k: tuple(v.shape) for k, v in res["grads"].items()}
    assert res["views"] == [0, 2, 4]
    for k in shapes:
        views = range(n_views) if k != "opacities" else [0, 2, 4]    # rank 1 had no opacity gradient
        want = sum((_view_grad(v, shapes)[k] for v in views), torch.zeros(shapes[k]))
        assert torch.allclose(res["grads"][k], want, atol=1e-6), k
    assert torch.equal(res["state"]["grad2d"], torch.full((7,), 3.0))
    assert torch.equal(res["state"]["count"], torch.full((7,), 2.0))
    assert torch.equal(res["state"]["radii"], torch.full((7,), 1.0))


def test_shard_views_partition():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "collab-splats_b200"))
    from radegs_b200.multiview import shard_views
    for n_views in (1, 8, 13):
        for world in (1, 2, 4, 8):
            parts = [shard_views(n_views, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n_views))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
