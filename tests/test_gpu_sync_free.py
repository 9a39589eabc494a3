"""Sync-free intersections and the CUDA-graph step (gsplat.cuda._wrapper.SYNC_FREE): the number of intersections stays on
the device (rs_isect_emit_ordered_bounded / rs_sort_pairs_dev / rs_offset_encode_dev, n_isects_dev of rs_rasterize_*),
buffers are sized for a learned capacity, an overflow raises a device flag.  Everything must equal the default path,
which reads the count back every call (as the reference does, gsplat/cuda/_wrapper.py isect_tiles)."""

import pytest
import torch

from radegs_b200 import scenes
from tests.util import small_scene

pytestmark = pytest.mark.gpu


@pytest.fixture
def sync_free():
    from gsplat.cuda import _wrapper as W
    old = (W.SYNC_FREE, W.ISECT_HEADROOM, W.ISECT_AUTO_GROW)
    W._ISECT_CAPACITY.clear()
    W.SYNC_FREE = True
    yield W
    W.SYNC_FREE, W.ISECT_HEADROOM, W.ISECT_AUTO_GROW = old
    W._ISECT_CAPACITY.clear()


def _call(params, vm, Ks, W_, H_, views_mode="RGB+ED"):
    from gsplat.rendering import rasterization
    return rasterization(*params, vm, Ks, W_, H_, sh_degree=3, packed=False, render_mode=views_mode,
                         rasterize_mode="antialiased", return_depth_normal=True)


def _leaves(gs, dev):
    return [p.detach().to(dev).requires_grad_(True) for p in scenes.activate(gs, 3)]


def _loss(out):
    return sum((o * torch.cos(0.01 * torch.arange(o.numel(), device=o.device, dtype=torch.float32)).view(o.shape)).sum()
               for o in out[:5])


@pytest.mark.parametrize("views,w,h,n", [(1, 160, 96, 4000), (3, 250, 130, 5000)])
def test_sync_free_equals_default(cuda_dev, sync_free, views, w, h, n):
    cfg, gs, vm, Ks = small_scene(n=n, w=w, h=h, views=views)
    vm, Ks = vm.to(cuda_dev), Ks.to(cuda_dev)
    sync_free.SYNC_FREE = False
    p0 = _leaves(gs, cuda_dev)
    ref = _call(p0, vm, Ks, w, h)
    _loss(ref).backward()
    sync_free.SYNC_FREE = True
    first = _call(_leaves(gs, cuda_dev), vm, Ks, w, h)              # learns the capacity (still synchronising)
    assert isinstance(first[5]["n_isects"], int) and first[5]["isect_overflow"] is None
    p1 = _leaves(gs, cuda_dev)
    got = _call(p1, vm, Ks, w, h)                                   # sync-free from here on
    meta, rmeta = got[5], ref[5]
    M = rmeta["flatten_ids"].numel()
    assert torch.is_tensor(meta["n_isects"]) and int(meta["n_isects"]) == M
    assert meta["flatten_ids"].numel() > M                          # capacity-sized
    assert int(meta["isect_overflow"]) == 0 and not sync_free.isect_overflowed()
    assert torch.equal(meta["isect_ids"][:M], rmeta["isect_ids"])
    assert torch.equal(meta["flatten_ids"][:M], rmeta["flatten_ids"])
    assert torch.equal(meta["isect_offsets"], rmeta["isect_offsets"])
    assert torch.equal(meta["tiles_per_gauss"], rmeta["tiles_per_gauss"])
    for a, b in zip(got[:5], ref[:5]):                              # same lists, same kernels: bit-identical images
        assert torch.equal(a, b)
    _loss(got).backward()
    for a, b in zip(p1, p0):                                        # float atomics commit in a different order
        scale = float(b.grad.abs().max()) + 1e-30
        assert float((a.grad - b.grad).abs().max()) <= 2e-5 * scale


def test_sync_free_overflow_is_flagged_not_fatal(cuda_dev, sync_free):
    cfg, gs, vm, Ks = small_scene(n=4000, w=160, h=96)
    vm, Ks = vm.to(cuda_dev), Ks.to(cuda_dev)
    first = _call(_leaves(gs, cuda_dev), vm, Ks, 160, 96)
    (entry,) = sync_free._ISECT_CAPACITY.values()
    cap = entry[0] = 2048 * (first[5]["n_isects"] // 4096)          # pretend a far too small capacity was learned
    assert cap >= 2048
    del first
    sync_free.ISECT_AUTO_GROW = False                               # (the automatic growth has its own test below)
    p = _leaves(gs, cuda_dev)
    out = _call(p, vm, Ks, 160, 96)
    _loss(out).backward()                                           # truncated lists: wrong image, but no fault
    torch.cuda.synchronize()
    assert int(out[5]["n_isects"]) > cap and out[5]["flatten_ids"].numel() == cap
    assert int(out[5]["isect_overflow"]) == 1
    assert sync_free.isect_overflowed() and not sync_free._ISECT_CAPACITY   # forgotten: the next call re-learns
    assert all(torch.isfinite(q.grad).all() for q in p)
    # every list entry a tile sees is a real one (the first `cap` of the sorted order are NOT guaranteed, the emitted
    # subset is): offsets are monotone and within the capacity
    offs = out[5]["isect_offsets"].flatten()
    assert int(offs.max()) <= cap and bool((offs[1:] >= offs[:-1]).all())
    _call(_leaves(gs, cuda_dev), vm, Ks, 160, 96)
    again = _call(_leaves(gs, cuda_dev), vm, Ks, 160, 96)
    assert int(again[5]["isect_overflow"]) == 0


def test_sync_free_capacity_follows_the_count_one_step_late(cuda_dev, sync_free):
    """Views with more intersections than the first one: the emit kernel mirrors the count into pinned host memory, the
    next render reads it (no synchronisation) and grows the buffers; a render that overflowed is reported then."""
    import warnings
    cfg, gs, vm, Ks = small_scene(n=4000, w=160, h=96)
    vm, Ks = vm.to(cuda_dev), Ks.to(cuda_dev)

    def render(boost):
        g2 = dict(gs)
        g2["log_scales"] = gs["log_scales"] + boost
        return _call(_leaves(g2, cuda_dev), vm, Ks, 160, 96)

    m0 = render(0.0)[5]["n_isects"]                                 # learns the capacity (synchronising call)
    (entry,) = sync_free._ISECT_CAPACITY.values()
    cap0 = entry[0] = (int(m0 * 1.1) + 2047) // 2048 * 2048         # (small scene: drop the constant part of the headroom)
    out = render(0.02)                                              # a few more intersections: fits, mirrored
    torch.cuda.synchronize()
    m1 = int(entry[2][0])
    assert m1 == int(out[5]["n_isects"]) and m0 < m1 <= cap0 and int(out[5]["isect_overflow"]) == 0
    big = render(1.5)                                               # far more than the headroom: truncated ...
    torch.cuda.synchronize()
    assert int(entry[2][0]) > cap0 and int(big[5]["isect_overflow"]) == 1
    assert entry[0] >= cap0                                         # (may already have grown for m1)
    cap0 = entry[0]
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        again = render(1.5)                                         # ... reported and repaired on the next call
    assert any("truncated" in str(w.message) for w in caught)
    assert entry[0] > cap0 and again[5]["flatten_ids"].numel() == entry[0]
    sync_free.SYNC_FREE = False
    ref = render(1.5)
    M = ref[5]["flatten_ids"].numel()
    assert torch.equal(again[5]["flatten_ids"][:M], ref[5]["flatten_ids"])
    for a, b in zip(again[:5], ref[:5]):
        assert torch.equal(a, b)


def test_sync_free_as_a_call_argument(cuda_dev):
    from gsplat.cuda import _wrapper as W
    from gsplat.rendering import rasterization
    assert not W.SYNC_FREE
    W._ISECT_CAPACITY.clear()
    cfg, gs, vm, Ks = small_scene(n=3000, w=128, h=80)
    vm, Ks = vm.to(cuda_dev), Ks.to(cuda_dev)
    kw = dict(sh_degree=3, packed=False, render_mode="RGB+ED", rasterize_mode="antialiased", return_depth_normal=True)
    ref = rasterization(*_leaves(gs, cuda_dev), vm, Ks, 128, 80, **kw)
    assert not W._ISECT_CAPACITY                                    # the default call leaves no state behind
    rasterization(*_leaves(gs, cuda_dev), vm, Ks, 128, 80, sync_free=True, **kw)
    got = rasterization(*_leaves(gs, cuda_dev), vm, Ks, 128, 80, sync_free=True, **kw)
    try:
        assert torch.is_tensor(got[5]["n_isects"]) and int(got[5]["n_isects"]) == ref[5]["n_isects"]
        for a, b in zip(got[:5], ref[:5]):
            assert torch.equal(a, b)
    finally:
        W._ISECT_CAPACITY.clear()


def test_sync_free_nothing_visible(cuda_dev, sync_free):
    cfg, gs, vm, Ks = small_scene(n=500, w=64, h=48)
    gs["means"] = gs["means"] + 100.0                              # everything behind / outside
    vm, Ks = vm.to(cuda_dev), Ks.to(cuda_dev)
    for _ in range(2):
        p = _leaves(gs, cuda_dev)
        out = _call(p, vm, Ks, 64, 48)
    assert int(out[5]["n_isects"]) == 0 and float(out[1].abs().max()) == 0.0
    _loss(out).backward()
    assert float(p[0].grad.abs().max()) == 0.0


def test_training_step_captured_in_a_cuda_graph(cuda_dev, sync_free):
    """forward + fused loss + backward of the training step, captured once and replayed: no device->host read inside."""
    from radegs_b200.losses import fused_rade_loss
    w, h = 160, 96
    cfg, gs, vm, Ks = small_scene(n=4000, w=w, h=h)
    vm, Ks = vm.to(cuda_dev), Ks.to(cuda_dev)
    fx, fy = float(Ks[0, 0, 0]), float(Ks[0, 1, 1])
    gt = torch.randint(0, 256, (h, w, 3), device=cuda_dev, dtype=torch.uint8)
    p = _leaves(gs, cuda_dev)

    def step():
        out = _call(p, vm, Ks, w, h)
        loss, _ = fused_rade_loss(out[0].view(h, w, -1), out[1].view(h, w), out[2].view(h, w), out[3].view(h, w),
                                  out[4].view(h, w, 3), gt, fx, fy)
        loss.backward()
        return loss

    side = torch.cuda.Stream(cuda_dev)
    side.wait_stream(torch.cuda.current_stream(cuda_dev))
    with torch.cuda.stream(side):
        for _ in range(3):                                          # learns the capacity, warms the allocator
            for q in p:
                q.grad = None
            eager_loss = step().detach()                            # (a live graph would pin its AccumulateGrad nodes
                                                                    #  to this stream and invalidate the capture)
    torch.cuda.current_stream(cuda_dev).wait_stream(side)
    torch.cuda.synchronize()
    eager = [q.grad.clone() for q in p]
    for q in p:
        q.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        loss = step().detach()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    assert not sync_free.isect_overflowed(reset=False)
    assert abs(float(loss) - float(eager_loss)) <= 1e-6 * abs(float(eager_loss))
    for a, b in zip(p, eager):
        scale = float(b.abs().max()) + 1e-30
        assert float((a.grad - b).abs().max()) <= 2e-5 * scale
    # new parameter values flow through the replay (the graph reads the parameter storage, not a snapshot)
    with torch.no_grad():
        p[0].add_(0.01)
    graph.replay()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(eager_loss)) > 0.0
