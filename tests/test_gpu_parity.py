"""GPU parity tests: the sm_100a kernels (through the gsplat-compatible API, i.e. through the C ABI) against
the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): integer artefacts (radii, tile lists, sorted keys, offsets) bit-exact;
images and gradients within max-abs 1e-4 + rel 1e-3 (images) / 2e-3 of the gradient scale (gradients, whose
atomics reorder fp32 sums).  Pixels the oracle flags as *fragile* (a discrete decision within ~1e-5 relative
of its threshold, where a 1-ulp difference in exp() legitimately flips it) are excluded from image checks and
their count is bounded.
"""

import math

import numpy as np
import pytest
import torch

from oracle import rade_oracle as O
from radegs_b200 import scenes
from tests.util import ATOL, RTOL, close_report, grad_close_report, small_scene

pytestmark = pytest.mark.gpu


def _gpu(ts, dev, grad=False):
    out = []
    for t in ts:
        t = t.detach().to(dev)
        if grad and t.is_floating_point():
            t.requires_grad_(True)
        out.append(t)
    return out


# ------------------------------------------------------------------------------------------------ projection
@pytest.mark.parametrize("views,spread", [(1, 1.0), (3, 1.6)])
def test_projection_forward(cuda_dev, views, spread):
    from gsplat.cuda._wrapper import fully_fused_projection
    cfg, gs, vm, Ks = small_scene(n=20000, w=256, h=192, views=views, spread=spread)
    means, quats, scales, _, _ = scenes.activate(gs, 3)
    ref = O.fully_fused_projection(means, quats, scales, vm, Ks, cfg.width, cfg.height, calc_compensations=True)
    m, q, s, v, k = _gpu((means, quats, scales, vm, Ks), cuda_dev)
    got = fully_fused_projection(m, None, q, s, v, k, cfg.width, cfg.height, calc_compensations=True)
    names = ["radii", "means2d", "depths", "conics", "compensations", "ray_ts", "ray_planes", "normals"]
    n_valid = int((ref[0] > 0).all(-1).sum())
    assert 0 < n_valid < views * cfg.n_gaussians or spread == 1.0
    # exact-order section: bit-exact
    for i in range(5):
        assert torch.equal(got[i].cpu(), ref[i]), f"{names[i]} not bit-exact: " \
            f"{int((got[i].cpu() != ref[i]).sum())} elements differ"
    # RaDe terms: ordinary arithmetic, tolerance
    for i in range(5, 8):
        ok, msg = close_report(names[i], got[i], ref[i], atol=1e-6, rtol=1e-4)
        assert ok, msg


def test_projection_prefilter_call(cuda_dev):
    """The exact call of RadegsModel._prefilter_voxel (rade_gs_model.py:373-397)."""
    from gsplat.cuda._wrapper import fully_fused_projection
    cfg, gs, vm, Ks = small_scene(n=5000, spread=2.5)
    means, quats, scales, _, _ = scenes.activate(gs, 3)
    m, q, s, v, k = _gpu((means, quats, scales, vm, Ks), cuda_dev)
    proj = fully_fused_projection(m, None, q, s, v, k, int(cfg.width), int(cfg.height), eps2d=0.3, packed=False,
                                  near_plane=0.01, far_plane=1e10, radius_clip=0.0, sparse_grad=False,
                                  calc_compensations=False)
    radii, means2d, depths, conics, compensations, ray_ts, ray_planes, normals = proj
    assert compensations is None and radii.shape == (1, 5000, 2) and radii.dtype == torch.int32
    mask = torch.sum(radii, dim=-1).squeeze() > 0
    ref = O.fully_fused_projection(means, quats, scales, vm, Ks, cfg.width, cfg.height)
    assert torch.equal(mask.cpu(), (ref[0].sum(-1).squeeze() > 0))
    assert 0 < int(mask.sum()) < 5000


@pytest.mark.parametrize("views", [1, 2])
def test_projection_backward(cuda_dev, views):
    from gsplat.cuda._wrapper import fully_fused_projection
    cfg, gs, vm, Ks = small_scene(n=4000, views=views, spread=1.3)
    means, quats, scales, _, _ = scenes.activate(gs, 3)
    cpu = [t.detach().clone().requires_grad_(True) for t in (means, quats, scales)]
    vm_c = vm.clone().requires_grad_(True)
    ref = O.fully_fused_projection(*cpu, vm_c, Ks, cfg.width, cfg.height, calc_compensations=True)
    g = torch.Generator().manual_seed(3)
    ws = [torch.randn(r.shape, generator=g) for r in ref[1:]]
    sum((r * w).sum() for r, w in zip(ref[1:], ws)).backward()
    gpu = _gpu((means, quats, scales), cuda_dev, grad=True)
    vm_g = vm.to(cuda_dev).requires_grad_(True)
    got = fully_fused_projection(gpu[0], None, gpu[1], gpu[2], vm_g, Ks.to(cuda_dev), cfg.width, cfg.height,
                                 calc_compensations=True)
    sum((r * w.to(cuda_dev)).sum() for r, w in zip(got[1:], ws)).backward()
    for name, a, b in zip(("means", "quats", "scales"), gpu, cpu):
        ok, msg = grad_close_report("v_" + name, a.grad, b.grad, rel=1e-3)
        assert ok, msg
    ok, msg = grad_close_report("v_viewmats", vm_g.grad[:, :3, :], vm_c.grad[:, :3, :], rel=2e-3)
    assert ok, msg


# ------------------------------------------------------------------------------------------------ SH
@pytest.mark.parametrize("degree", [0, 1, 2, 3])
@pytest.mark.parametrize("shared", [False, True])
def test_spherical_harmonics(cuda_dev, degree, shared):
    from gsplat.cuda._wrapper import spherical_harmonics
    g = torch.Generator().manual_seed(degree)
    C, N, K = 2, 3000, 16
    dirs = torch.randn(C, N, 3, generator=g)
    coeffs = torch.randn(*((N,) if shared else (C, N)), K, 3, generator=g)
    masks = torch.rand(C, N, generator=g) > 0.2
    w = torch.randn(C, N, 3, generator=g)
    dc, cc = dirs.clone().requires_grad_(True), coeffs.clone().requires_grad_(True)
    ref = O.spherical_harmonics(degree, dc, cc[None].expand(C, N, K, 3) if shared else cc, masks)
    (ref * w).sum().backward()
    dg, cg = dirs.to(cuda_dev).requires_grad_(True), coeffs.to(cuda_dev).requires_grad_(True)
    got = spherical_harmonics(degree, dg, cg, masks=masks.to(cuda_dev))
    (got * w.to(cuda_dev)).sum().backward()
    ok, msg = close_report("sh colors", got, ref, atol=1e-5, rtol=1e-4)
    assert ok, msg
    ok, msg = grad_close_report("v_coeffs", cg.grad, cc.grad, rel=1e-4)
    assert ok, msg
    ok, msg = grad_close_report("v_dirs", dg.grad, dc.grad, rel=1e-3)
    assert ok, msg


# ------------------------------------------------------------------------------------------------ scan / sort / isect
@pytest.mark.parametrize("n", [1, 7, 2048, 2049, 100_003, 3_000_000])
def test_cumsum(cuda_dev, n):
    from radegs_b200 import backend as be
    lib = be.load()
    g = torch.Generator().manual_seed(n)
    x = torch.randint(0, 50, (n,), generator=g, dtype=torch.int32)
    xd = x.to(cuda_dev)
    out = torch.empty(n, device=cuda_dev, dtype=torch.int64)
    tb = lib.rs_cumsum_temp_bytes(n)
    temp = torch.empty(tb, device=cuda_dev, dtype=torch.uint8)
    be.check(lib.rs_cumsum_i32_i64(be.ptr(xd), be.ptr(out), n, be.ptr(temp), tb, be.stream_ptr(cuda_dev)), "cumsum")
    assert torch.equal(out.cpu(), torch.cumsum(x.long(), 0))


@pytest.mark.parametrize("m,end_bit", [(1, 46), (33, 46), (4096, 46), (4097, 40), (50_000, 46), (1_000_003, 46),
                                        (5_000_000, 47), (200_000, 64), (200_000, 13)])
def test_radix_sort_pairs(cuda_dev, m, end_bit):
    """Stable ascending sort on key bits [0,end_bit): identical to a stable argsort of the masked keys."""
    from radegs_b200 import backend as be
    lib = be.load()
    g = torch.Generator().manual_seed(m + end_bit)
    # many duplicate keys (few distinct depths) so stability matters
    hi = torch.randint(0, 1 << 13, (m,), generator=g, dtype=torch.int64)
    lo = torch.randint(0, 1 << 20, (m,), generator=g, dtype=torch.int64) * 3001 % (1 << 32)
    keys = ((hi << 32) | lo) if end_bit > 32 else lo
    if end_bit == 64:
        keys = keys | (torch.randint(0, 2, (m,), generator=g, dtype=torch.int64) << 62)
    vals = torch.arange(m, dtype=torch.int32)
    mask = (1 << end_bit) - 1 if end_bit < 64 else -1
    order = np.argsort((keys.numpy().astype(np.uint64) & np.uint64(mask & 0xFFFFFFFFFFFFFFFF)), kind="stable")
    ka, va = keys.to(cuda_dev), vals.to(cuda_dev)
    kb, vb = torch.empty_like(ka), torch.empty_like(va)
    tb = lib.rs_sort_pairs_temp_bytes(m, 0, end_bit)
    temp = torch.empty(tb, device=cuda_dev, dtype=torch.uint8)
    where = be.check(lib.rs_sort_pairs(be.ptr(ka), be.ptr(va), be.ptr(kb), be.ptr(vb), m, 0, end_bit, be.ptr(temp),
                                       tb, be.stream_ptr(cuda_dev)), "sort")
    ko, vo = (kb, vb) if where == 0 else (ka, va)
    assert torch.equal(vo.cpu(), vals[torch.from_numpy(order)]), "values are not in stable sorted order"
    assert torch.equal(ko.cpu(), keys[torch.from_numpy(order)])


@pytest.mark.parametrize("m", [1, 31, 2048, 2049, 100_003, 1_000_000])
def test_argsort_u32(cuda_dev, m):
    """rs_argsort_u32 (depth keys of the presorted intersection path) = a stable argsort, ties included."""
    from radegs_b200 import backend as be
    lib = be.load()
    g = torch.Generator().manual_seed(m)
    depth = torch.rand(m, generator=g) * 5.0 + 0.01
    depth[torch.rand(m, generator=g) < 0.3] = 2.5            # many equal keys: stability matters
    keys = depth.view(torch.int32)
    order = np.argsort(keys.numpy().astype(np.uint32), kind="stable")
    ka = keys.to(cuda_dev).clone()
    kb, va, vb = torch.empty_like(ka), torch.empty_like(ka), torch.empty_like(ka)
    tb = lib.rs_sort_pairs_temp_bytes(m, 0, 32)
    temp = torch.empty(tb, device=cuda_dev, dtype=torch.uint8)
    where = be.check(lib.rs_argsort_u32(be.ptr(ka), be.ptr(va), be.ptr(kb), be.ptr(vb), m, 0, 32, be.ptr(temp), tb,
                                        be.stream_ptr(cuda_dev)), "argsort")
    ko, vo = (kb, vb) if where == 0 else (ka, va)
    assert torch.equal(vo.cpu().long(), torch.from_numpy(order)), "not the stable argsort"
    assert torch.equal(ko.cpu(), keys[torch.from_numpy(order)])


@pytest.fixture
def isect_method(request):
    import gsplat.cuda._wrapper as wr
    default = wr.ISECT_SORT_METHOD
    wr.ISECT_SORT_METHOD = request.param
    yield request.param
    wr.ISECT_SORT_METHOD = default


@pytest.mark.parametrize("isect_method", ["presort", "radix"], indirect=True)
@pytest.mark.parametrize("views,w,h", [(1, 160, 96), (2, 250, 130), (5, 64, 64)])
def test_isect_bit_exact(cuda_dev, isect_method, views, w, h):
    """Stage-wise (oracle consumes the GPU's own floats) and end to end (oracle's own projection)."""
    from gsplat.cuda._wrapper import fully_fused_projection, isect_offset_encode, isect_tiles
    cfg, gs, vm, Ks = small_scene(n=6000, w=w, h=h, views=views, spread=1.4)
    means, quats, scales, _, _ = scenes.activate(gs, 3)
    m, q, s, v, k = _gpu((means, quats, scales, vm, Ks), cuda_dev)
    radii, means2d, depths = fully_fused_projection(m, None, q, s, v, k, w, h)[:3]
    tw, th = math.ceil(w / 16), math.ceil(h / 16)
    tiles, ids, flat = isect_tiles(means2d, radii, depths, 16, tw, th)
    offs = isect_offset_encode(ids, views, tw, th)
    # stage-wise
    r_tiles, r_ids, r_flat = O.isect_tiles(means2d.cpu(), radii.cpu(), depths.cpu(), 16, tw, th)
    r_offs = O.isect_offset_encode(r_ids, views, tw, th)
    assert torch.equal(tiles.cpu(), r_tiles)
    assert torch.equal(ids.cpu(), r_ids), "sorted tile|depth keys differ"
    assert torch.equal(flat.cpu(), r_flat), "flatten ids differ (stability / emission order)"
    assert torch.equal(offs.cpu(), r_offs)
    assert ids.numel() > 1000
    # unsorted emission order too
    _, u_ids, u_flat = isect_tiles(means2d, radii, depths, 16, tw, th, sort=False)
    _, ru_ids, ru_flat = O.isect_tiles(means2d.cpu(), radii.cpu(), depths.cpu(), 16, tw, th, sort=False)
    assert torch.equal(u_ids.cpu(), ru_ids) and torch.equal(u_flat.cpu(), ru_flat)
    # end to end: the oracle's own projection gives the same lists
    o_radii, o_m2, o_depths = O.fully_fused_projection(means, quats, scales, vm, Ks, w, h)[:3]
    e_tiles, e_ids, e_flat = O.isect_tiles(o_m2, o_radii, o_depths, 16, tw, th)
    assert torch.equal(ids.cpu(), e_ids) and torch.equal(flat.cpu(), e_flat) and torch.equal(tiles.cpu(), e_tiles)


@pytest.mark.parametrize("isect_method", ["presort", "radix"], indirect=True)
def test_isect_equal_depths_stable(cuda_dev, isect_method):
    """Many Gaussians share a depth (quantised to 1/8): the order inside a tile is then decided by stability alone
    (flatten id, tile row, tile column), which the depth-presorted path must reproduce bit for bit."""
    from gsplat.cuda._wrapper import isect_tiles
    views, w, h = 3, 200, 120
    cfg, gs, vm, Ks = small_scene(n=8000, w=w, h=h, views=views, spread=1.3)
    means, quats, scales, _, _ = scenes.activate(gs, 3)
    radii, m2, depths = O.fully_fused_projection(means, quats, scales, vm, Ks, w, h)[:3]
    depths = (depths * 8).round() / 8
    tw, th = math.ceil(w / 16), math.ceil(h / 16)
    r_tiles, r_ids, r_flat = O.isect_tiles(m2, radii, depths, 16, tw, th)
    tiles, ids, flat = isect_tiles(m2.to(cuda_dev), radii.to(cuda_dev), depths.to(cuda_dev), 16, tw, th)
    assert int((r_ids[1:] == r_ids[:-1]).sum()) > 1000, "the scene must contain equal keys"
    assert torch.equal(tiles.cpu(), r_tiles) and torch.equal(ids.cpu(), r_ids) and torch.equal(flat.cpu(), r_flat)


# ------------------------------------------------------------------------------------------------ compositing
def _raster_inputs(cfg, gs, vm, Ks, D, seed=5, antialiased=True):
    means, quats, scales, opac, _ = scenes.activate(gs, 3)
    radii, m2, depths, conics, comps, ray_ts, ray_planes, normals = O.fully_fused_projection(
        means, quats, scales, vm, Ks, cfg.width, cfg.height, calc_compensations=True)
    C, N = depths.shape
    g = torch.Generator().manual_seed(seed)
    colors = torch.rand(C, N, D, generator=g)
    o = opac[None].expand(C, N) * (comps if antialiased else 1.0)
    tw, th = math.ceil(cfg.width / 16), math.ceil(cfg.height / 16)
    _, ids, flat = O.isect_tiles(m2, radii, depths, 16, tw, th)
    offs = O.isect_offset_encode(ids, C, tw, th)
    return dict(means2d=m2, conics=conics, colors=colors, opacities=o.contiguous(), ray_ts=ray_ts,
                ray_planes=ray_planes, normals=normals), offs, flat


@pytest.fixture
def raster_variant(request):
    """Selects the compositing variant for <= 4 channels (0: 8x4 pixels per warp, 1: 8x8, two pixels per lane -- the
    default --, 2: 8x8 with the backward's per-Gaussian reduction on the tensor cores, 3: 8x8 with the staging schemes swapped (forward mbarrier ring, backward CTA barrier); +10: with the bbox footprint test
    instead of the exact ellipse-vs-rectangle one) through the per-call flags for the duration of a test."""
    from radegs_b200 import backend as be
    from gsplat.cuda import _wrapper as W
    v = request.param % 10
    flags = {0: be.RS_RASTER_ONE_PIXEL, 1: 0, 2: be.RS_RASTER_BWD_MMA, 3: be.RS_RASTER_FWD_RING | be.RS_RASTER_BWD_BARRIER}[v]
    if request.param >= 10:
        flags |= be.RS_RASTER_CULL_BBOX
    old = W.RASTER_FLAGS
    W.RASTER_FLAGS = flags
    yield v
    W.RASTER_FLAGS = old


@pytest.mark.parametrize("raster_variant", [0, 1, 2, 3, 10, 11, 13], indirect=True)
@pytest.mark.parametrize("D,views,w,h,bg,n", [(3, 1, 160, 96, False, 2500), (4, 2, 100, 70, True, 2500),
                                               (17, 1, 96, 64, False, 2500), (67, 1, 64, 48, True, 2500),
                                               (8, 1, 64, 64, False, 2500),
                                               # > 256 Gaussians per tile: several staged batches per tile
                                               (3, 1, 64, 48, False, 9000), (36, 1, 48, 32, False, 4000),
                                               # image not a multiple of the tile / warp block
                                               (2, 1, 75, 53, True, 3000)])
def test_rasterize_to_pixels_fwd_bwd(cuda_dev, raster_variant, D, views, w, h, bg, n):
    from gsplat.cuda._wrapper import rasterize_to_pixels
    if raster_variant == 1 and D > 4:
        pytest.skip("the two-pixels-per-lane kernels cover <= 4 channels; wider rows use the generic kernels")
    cfg, gs, vm, Ks = small_scene(n=n, w=w, h=h, views=views)
    inp, offs, flat = _raster_inputs(cfg, gs, vm, Ks, D)
    g = torch.Generator().manual_seed(9)
    backgrounds = torch.rand(views, D, generator=g) if bg else None
    order = ["means2d", "conics", "colors", "opacities", "ray_ts", "ray_planes", "normals"]
    cpu = {k: inp[k].detach().clone().requires_grad_(True) for k in order}
    ref = O.rasterize_to_pixels(cpu["means2d"], cpu["conics"], cpu["colors"], cpu["opacities"], cpu["ray_ts"],
                                cpu["ray_planes"], cpu["normals"], Ks, w, h, 16, offs, flat, backgrounds=backgrounds,
                                return_aux=True)
    aux = ref[5]
    gpu = {k: inp[k].detach().to(cuda_dev).requires_grad_(True) for k in order}
    got = rasterize_to_pixels(gpu["means2d"], gpu["conics"], gpu["colors"], gpu["opacities"], w, h, 16,
                              offs.to(cuda_dev), flat.to(cuda_dev),
                              backgrounds=None if backgrounds is None else backgrounds.to(cuda_dev),
                              ray_ts=gpu["ray_ts"], ray_planes=gpu["ray_planes"], normals=gpu["normals"],
                              Ks=Ks.to(cuda_dev), return_ids=True)
    ok_px = ~aux["fragile"]
    # near-threshold decisions scale with the number of pairs visited (~1e-5 relative window per decision)
    assert int(aux["fragile"].sum()) <= max(4, aux["fragile"].numel() // 500, aux["n_tested"] // 100_000)
    names = ["colors", "alphas", "expected_depths", "median_depths", "normals"]
    for i, nm in enumerate(names):
        ok, msg = close_report(nm, got[i], ref[i], mask=ok_px)
        assert ok, msg
    same_last = (got[5].cpu() == aux["last_ids"]) | aux["fragile"]
    assert bool(same_last.all()), f"last_ids differ at {int((~same_last).sum())} robust pixels"
    same_med = (got[6].cpu() == aux["median_ids"]) | aux["fragile"]
    assert bool(same_med.all()), f"median_ids differ at {int((~same_med).sum())} robust pixels"
    if n > 3000:
        per_tile = (offs.flatten()[1:] - offs.flatten()[:-1]).max().item()
        assert per_tile > 600, per_tile
    assert float(ref[1].mean()) > 0.2, "scene does not cover the image enough to be a meaningful test"
    # backward: random cotangents, zeroed at fragile pixels so both sides differentiate the same branch
    ws = [torch.randn(r.shape, generator=g) * ok_px[..., None] for r in ref[:5]]
    sum((r * x).sum() for r, x in zip(ref[:5], ws)).backward()
    sum((r * x.to(cuda_dev)).sum() for r, x in zip(got[:5], ws)).backward()
    for k in order:
        ok, msg = grad_close_report("v_" + k, gpu[k].grad, cpu[k].grad)
        assert ok, msg


@pytest.mark.parametrize("raster_variant,D", [(0, 5), (0, 3), (1, 3)], indirect=["raster_variant"])
def test_rasterize_absgrad_and_shared_colors(cuda_dev, raster_variant, D):
    """colors [N,D] shared by all cameras (no expansion) and the absgrad side channel."""
    from gsplat.cuda._wrapper import rasterize_to_pixels
    cfg, gs, vm, Ks = small_scene(n=2000, w=96, h=64, views=2)
    inp, offs, flat = _raster_inputs(cfg, gs, vm, Ks, D)
    shared = inp["colors"][0].contiguous()
    cpu_col = shared.clone().requires_grad_(True)
    cpu_m2 = inp["means2d"].clone().requires_grad_(True)
    ref = O.rasterize_to_pixels(cpu_m2, inp["conics"], cpu_col[None].expand(2, -1, -1), inp["opacities"],
                                inp["ray_ts"], inp["ray_planes"], inp["normals"], Ks, 96, 64, 16, offs, flat,
                                return_aux=True)
    d = {k: v.to(cuda_dev) for k, v in inp.items()}
    col = shared.to(cuda_dev).requires_grad_(True)
    m2 = d["means2d"].clone().requires_grad_(True)
    got = rasterize_to_pixels(m2, d["conics"], col, d["opacities"], 96, 64, 16, offs.to(cuda_dev), flat.to(cuda_dev),
                              absgrad=True, ray_ts=d["ray_ts"], ray_planes=d["ray_planes"], normals=d["normals"],
                              Ks=Ks.to(cuda_dev))
    ok_px = ~ref[5]["fragile"]
    ok, msg = close_report("colors", got[0], ref[0], mask=ok_px)
    assert ok, msg
    w = torch.randn(ref[0].shape, generator=torch.Generator().manual_seed(1)) * ok_px[..., None]
    (ref[0] * w).sum().backward()
    (got[0] * w.to(cuda_dev)).sum().backward()
    ok, msg = grad_close_report("v_colors(shared)", col.grad, cpu_col.grad)
    assert ok, msg
    ok, msg = grad_close_report("v_means2d", m2.grad, cpu_m2.grad)
    assert ok, msg
    assert hasattr(m2, "absgrad") and m2.absgrad.shape == m2.shape
    assert bool((m2.absgrad + 1e-7 >= m2.grad.abs()).all()), "sum of |g| must dominate |sum of g|"


# ------------------------------------------------------------------------------------------------ end to end
@pytest.mark.parametrize("sh_degree,mode,raster_mode,views", [(3, "RGB+ED", "antialiased", 1), (None, "RGB", "classic", 2),
                                                               (1, "RGB", "antialiased", 1), (None, "ED", "classic", 1)])
def test_rasterization_end_to_end(cuda_dev, sh_degree, mode, raster_mode, views):
    """The reference's call (rade_gs_model.py:439-465), forward + backward incl. the depth-normal loss."""
    from gsplat.rendering import rasterization
    cfg, gs, vm, Ks = small_scene(n=3000, w=128, h=80, views=views, sh_degree=sh_degree if sh_degree else 3)
    params = scenes.activate(gs, sh_degree)
    W, H = cfg.width, cfg.height
    cpu = [p.detach().clone().requires_grad_(True) for p in params]
    ref = O.rasterization(*cpu, vm, Ks, W, H, sh_degree=sh_degree, render_mode=mode, rasterize_mode=raster_mode,
                          return_depth_normal=True, return_aux=True)
    gpu = _gpu(params, cuda_dev, grad=True)
    got = rasterization(means=gpu[0], quats=gpu[1], scales=gpu[2], opacities=gpu[3], colors=gpu[4],
                        viewmats=vm.to(cuda_dev), Ks=Ks.to(cuda_dev), width=W, height=H, packed=False,
                        near_plane=0.01, far_plane=1e10, render_mode=mode, sh_degree=sh_degree, sparse_grad=False,
                        absgrad=False, rasterize_mode=raster_mode, return_depth_normal=True)
    meta, rmeta = got[5], ref[5]
    # integer artefacts: bit-exact end to end
    for key in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert torch.equal(meta[key].cpu(), rmeta[key]), f"meta[{key}] differs"
    ok_px = ~rmeta["fragile"]
    for i, nm in enumerate(["render", "alpha", "expected_depths", "median_depths", "expected_normals"]):
        ok, msg = close_report(nm, got[i], ref[i], mask=ok_px)
        assert ok, msg
    # loss = L1-ish on colour + depth-normal consistency (rade_gs_model.py:202-219, 292-307) on camera 0
    def loss_of(out, K0):
        rc, ra, de, dm, nr = out[:5]
        keep = ok_px.to(rc.device)
        l_dn, _ = O.depth_normal_loss(K0, W, H, de[0, ..., 0] * keep[0], dm[0, ..., 0] * keep[0], nr[0] * keep[0, ..., None])
        return (rc * keep[..., None]).abs().mean() + 0.1 * (ra * keep[..., None]).mean() + l_dn
    loss_of(ref, Ks[0]).backward()
    loss_of(got, Ks[0].to(cuda_dev)).backward()
    for nm, a, b in zip(("means", "quats", "scales", "opacities", "colors"), gpu, cpu):
        ok, msg = grad_close_report("v_" + nm, a.grad, b.grad, rel=3e-3)
        assert ok, msg
    # meta contract
    for key in ("means2d", "radii", "depths", "conics", "opacities", "width", "height", "n_cameras", "tile_size",
                "tile_width", "tile_height", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets",
                "camera_ids", "gaussian_ids"):
        assert key in meta
    assert meta["means2d"].shape == (views, cfg.n_gaussians, 2) and meta["radii"].dtype == torch.int32


def test_rasterization_retain_grad_absgrad(cuda_dev):
    """strategy.step_pre_backward does info['means2d'].retain_grad() (rade_gs_model.py:191-198)."""
    from gsplat.rendering import rasterization
    from gsplat.strategy import DefaultStrategy
    cfg, gs, vm, Ks = small_scene(n=1500, w=96, h=64)
    params = _gpu(scenes.activate(gs, None), cuda_dev, grad=True)
    strat = DefaultStrategy(absgrad=True)
    out = rasterization(*params, vm.to(cuda_dev), Ks.to(cuda_dev), 96, 64, packed=False, absgrad=strat.absgrad,
                        return_depth_normal=True)
    info = out[5]
    strat.step_pre_backward({}, {}, {}, 0, info)
    (out[0].mean() + out[3].mean()).backward()
    assert info["means2d"].grad is not None and info["means2d"].absgrad is not None
    assert float(info["means2d"].absgrad.sum()) > 0


def test_unsupported_options_raise(cuda_dev):
    from gsplat.rendering import rasterization
    cfg, gs, vm, Ks = small_scene(n=100, w=32, h=32)
    p = _gpu(scenes.activate(gs, None), cuda_dev)
    with pytest.raises(NotImplementedError):
        rasterization(*p, vm.to(cuda_dev), Ks.to(cuda_dev), 32, 32)  # packed defaults to True upstream
    with pytest.raises(NotImplementedError):
        rasterization(*p, vm.to(cuda_dev), Ks.to(cuda_dev), 32, 32, packed=False, sparse_grad=True)
    with pytest.raises(RuntimeError):
        rasterization(*[t.cpu() for t in p], vm, Ks, 32, 32, packed=False)  # no CPU path


@pytest.mark.parametrize("case", ["no_gaussians_visible", "single_gaussian", "odd_image", "wide_channels_chunked"])
def test_edge_cases(cuda_dev, case):
    from gsplat.rendering import rasterization
    if case == "no_gaussians_visible":
        cfg, gs, vm, Ks = small_scene(n=500, w=64, h=48)
        gs["means"] = gs["means"] + 100.0  # everything behind / outside
        p = _gpu(scenes.activate(gs, None), cuda_dev, grad=True)
        out = rasterization(*p, vm.to(cuda_dev), Ks.to(cuda_dev), 64, 48, packed=False, return_depth_normal=True,
                            backgrounds=torch.full((1, 3), 0.25, device=cuda_dev))
        assert out[5]["isect_ids"].numel() == 0
        assert torch.allclose(out[0], torch.full_like(out[0], 0.25)) and float(out[1].abs().max()) == 0.0
        out[0].sum().backward()
        assert float(p[0].grad.abs().max()) == 0.0
    elif case == "single_gaussian":
        # analytic known answer: one isotropic Gaussian on the optical axis
        dev = cuda_dev
        means = torch.tensor([[0.0, 0.0, 0.0]], device=dev)
        quats = torch.tensor([[1.0, 0.0, 0.0, 0.0]], device=dev)
        scales = torch.full((1, 3), 0.2, device=dev)
        opac = torch.tensor([0.9], device=dev)
        cols = torch.tensor([[0.2, 0.5, 0.8]], device=dev)
        vm = torch.eye(4, device=dev)[None].clone()
        vm[0, 2, 3] = 4.0
        Ks = torch.tensor([[[100.0, 0, 32.0], [0, 100.0, 32.0], [0, 0, 1]]], device=dev)
        rc, ra, de, dm, nr, meta = rasterization(means, quats, scales, opac, cols, vm, Ks, 64, 64, packed=False,
                                                 return_depth_normal=True)
        c = (31, 31)  # pixel centre (31.5,31.5) is 0.5 px from the principal point
        a = float(ra[0, c[0], c[1], 0])
        assert 0.85 < a < 0.9001
        assert abs(float(dm[0, c[0], c[1], 0]) - 4.0) < 2e-3
        assert abs(float(de[0, c[0], c[1], 0]) / a - 4.0) < 2e-3
        n = nr[0, c[0], c[1]] / a
        assert float(n[2]) < -0.999 and float(n[:2].abs().max()) < 2e-2
    elif case == "odd_image":
        cfg, gs, vm, Ks = small_scene(n=1500, w=75, h=37)
        params = scenes.activate(gs, None)
        ref = O.rasterization(*params, vm, Ks, 75, 37, return_depth_normal=True, return_aux=True)
        got = rasterization(*_gpu(params, cuda_dev), vm.to(cuda_dev), Ks.to(cuda_dev), 75, 37, packed=False,
                            return_depth_normal=True)
        for i in range(5):
            ok, msg = close_report(f"out{i}", got[i], ref[i], mask=~ref[5]["fragile"])
            assert ok, msg
    else:
        cfg, gs, vm, Ks = small_scene(n=800, w=48, h=32, sh_degree=None, n_features=97)
        params = scenes.activate(gs, None)
        assert params[4].shape[-1] == 100
        cpu = [p.detach().clone().requires_grad_(True) for p in params]
        ref = O.rasterization(*cpu, vm, Ks, 48, 32, return_depth_normal=True, return_aux=True)
        gpu = _gpu(params, cuda_dev, grad=True)
        got = rasterization(*gpu, vm.to(cuda_dev), Ks.to(cuda_dev), 48, 32, packed=False, return_depth_normal=True)
        keep = ~ref[5]["fragile"]
        ok, msg = close_report("colors100", got[0], ref[0], mask=keep)
        assert ok, msg
        (ref[0] * keep[..., None]).sum().backward()
        (got[0] * keep[..., None].to(cuda_dev)).sum().backward()
        ok, msg = grad_close_report("v_colors100", gpu[4].grad, cpu[4].grad)
        assert ok, msg


def test_full_size_properties(cuda_dev):
    """BASELINE config 2 size (1M Gaussians, 1080p): size-independent invariants instead of the oracle."""
    from gsplat.rendering import rasterization
    cfg = scenes.BASELINE_CONFIGS[2]
    gs, vm, Ks = scenes.make_scene(cfg)
    means, quats, scales, opac, _ = _gpu(scenes.activate(gs, 3), cuda_dev)
    ones = torch.ones(cfg.n_gaussians, 3, device=cuda_dev)
    rc, ra, de, dm, nr, meta = rasterization(means, quats, scales, opac, ones, vm.to(cuda_dev), Ks.to(cuda_dev),
                                             cfg.width, cfg.height, packed=False, render_mode="RGB+ED",
                                             rasterize_mode="antialiased", return_depth_normal=True)
    ids = meta["isect_ids"]
    assert ids.numel() > 1_000_000
    assert bool((ids[1:] >= ids[:-1]).all()), "keys are not sorted"
    offs = meta["isect_offsets"].flatten()
    assert bool((offs[1:] >= offs[:-1]).all()) and int(offs[0]) == 0 and int(offs[-1]) <= ids.numel()
    assert int(meta["tiles_per_gauss"].sum()) == ids.numel()
    # rendering constant colour 1 telescopes to alpha: sum_i vis_i = 1 - T_final
    assert float((rc[..., :3] - ra).abs().max()) < 2e-5
    assert float(ra.min()) >= 0.0 and float(ra.max()) <= 1.0
    assert float(nr.norm(dim=-1).max()) <= 1.0 + 1e-4
    inside = ra[..., 0] > 0.99
    assert int(inside.sum()) > 10000
    zd = rc[..., 3][inside]   # alpha-normalised expected z depth: must lie inside the scene's depth range
    assert float(zd.min()) > 1.0 and float(zd.max()) < 5.5
    assert float(dm[..., 0][inside].min()) > 1.0 and float(dm[..., 0][inside].max()) < 5.5
    # every flatten id is a visible Gaussian
    assert bool((meta["radii"].reshape(-1, 2)[meta["flatten_ids"].long()] > 0).all())


@pytest.mark.parametrize("cfg_id,win", [(2, (31, 37, 57, 63)), (3, (14, 20, 27, 33))])
def test_full_size_window_matches_oracle(cuda_dev, cfg_id, win):
    """BASELINE configs 2 and 3 at FULL size (1 M Gaussians, 1920x1080, sh3; 500 k Gaussians with 3 + 64 feature
    channels, 960x540; RGB+ED antialiased) against the oracle: the
    oracle projects, intersects and sorts everything (so radii, tile counts, the 6.9 M sorted 64-bit keys, flatten ids
    and the 8160 tile offsets are compared bit for bit) and composites a 6x6-tile window in the image centre, where
    images and -- for a loss over that window -- the gradients of all five inputs are compared."""
    from gsplat.rendering import rasterization
    cfg = scenes.BASELINE_CONFIGS[cfg_id]
    gs, vm, Ks = scenes.make_scene(cfg, n_views=1)
    sh = 3 if cfg_id == 2 else None
    params = scenes.activate(gs, sh)
    W, H = cfg.width, cfg.height
    y0, y1, x0, x1 = win[0] * 16, win[1] * 16, win[2] * 16, win[3] * 16    # win = tile rows / tile columns
    cpu = [t.detach().clone().requires_grad_(True) for t in params]
    ref = O.rasterization(*cpu, vm, Ks, W, H, sh_degree=sh, render_mode="RGB+ED", rasterize_mode="antialiased",
                          return_depth_normal=True, return_aux=True, tile_window=win)
    gpu = _gpu(params, cuda_dev, grad=True)
    got = rasterization(*gpu, vm.to(cuda_dev), Ks.to(cuda_dev), W, H, packed=False, sh_degree=sh, render_mode="RGB+ED",
                        rasterize_mode="antialiased", return_depth_normal=True)
    meta, rmeta = got[5], ref[5]
    assert rmeta["isect_ids"].numel() > 1_000_000 and got[0].shape[-1] == (4 if cfg_id == 2 else 68)
    for key in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert torch.equal(meta[key].cpu(), rmeta[key]), f"meta[{key}] differs at full size"
    keep = ~rmeta["fragile"][:, y0:y1, x0:x1]
    assert float(ref[1][0, y0:y1, x0:x1].detach().mean()) > 0.5 and int((~keep).sum()) < keep.numel() // 20
    for i, nm in enumerate(["render", "alpha", "expected_depths", "median_depths", "expected_normals"]):
        ok, msg = close_report(nm, got[i][:, y0:y1, x0:x1], ref[i][:, y0:y1, x0:x1], mask=keep)
        assert ok, msg
    g = torch.Generator().manual_seed(4)
    ws = [torch.randn(1, y1 - y0, x1 - x0, t.shape[-1], generator=g) * keep[..., None] for t in ref[:5]]
    sum((t[:, y0:y1, x0:x1] * w).sum() for t, w in zip(ref[:5], ws)).backward()
    sum((t[:, y0:y1, x0:x1] * w.to(cuda_dev)).sum() for t, w in zip(got[:5], ws)).backward()
    for nm, a, b in zip(("means", "quats", "scales", "opacities", "sh_coeffs"), gpu, cpu):
        ok, msg = grad_close_report("v_" + nm, a.grad, b.grad, rel=3e-3)
        assert ok, msg
        assert float(b.grad.abs().max()) > 0


FULL_VIEW_CASES = {
    # BASELINE config -> (views rendered, what it adds).  Config 4 is a 2-view slice of its 8-view batch (camera bits
    # in the sort keys at 3 M Gaussians), config 5 is forward only (the meshing sweep renders under no_grad).
    2: dict(views=1, sh=3, backward=True), 3: dict(views=1, sh=None, backward=True),
    4: dict(views=2, sh=3, backward=True), 5: dict(views=1, sh=3, backward=False),
}


@pytest.mark.parametrize("cfg_id", [2, 3, 4, 5])
def test_full_view_matches_c_oracle(cuda_dev, cfg_id):
    """Every BASELINE GPU config at FULL size against the oracle on the COMPLETE view(s): the oracle's projection, SH,
    intersection and sort run in PyTorch on the CPU, its compositing forward/backward in C (oracle/raster_oracle.c,
    pinned to the PyTorch restatement by tests/test_c_oracle.py).  Integer artefacts bit for bit; every pixel of every
    output within max-abs 1e-4 + rel 1e-3 outside the oracle's fragile pixels; the gradients of all five inputs, for a
    random cotangent over the whole view, per element against the fp64 oracle and against the fp32 oracle's own error
    (tests/util.py:grad_parity_report)."""
    from gsplat.rendering import rasterization
    from tests.util import grad_parity_report
    case = FULL_VIEW_CASES[cfg_id]
    cfg = scenes.BASELINE_CONFIGS[cfg_id]
    gs, vm, Ks = scenes.make_scene(cfg, n_views=case["views"])
    sh = case["sh"]
    params = scenes.activate(gs, sh)
    W, H, C = cfg.width, cfg.height, case["views"]
    kw = dict(sh_degree=sh, render_mode="RGB+ED", rasterize_mode="antialiased", return_depth_normal=True)
    cpu = [t.detach().clone().requires_grad_(case["backward"]) for t in params]
    ref = O.rasterization(*cpu, vm, Ks, W, H, return_aux=True, compositor="c", **kw)
    gpu = _gpu(params, cuda_dev, grad=case["backward"])
    with torch.set_grad_enabled(case["backward"]):
        got = rasterization(*gpu, vm.to(cuda_dev), Ks.to(cuda_dev), W, H, packed=False, **kw)
    meta, rmeta = got[5], ref[5]
    assert rmeta["isect_ids"].numel() > 1_500_000
    if C > 1:
        assert int(rmeta["isect_ids"].max() >> 45) >= 1, "camera bits are not exercised"
    for key in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert torch.equal(meta[key].cpu(), rmeta[key]), f"meta[{key}] differs at full size"
    keep = ~rmeta["fragile"]
    assert float(ref[1].detach().mean()) > 0.5 and int((~keep).sum()) < keep.numel() // 20
    for i, nm in enumerate(["render", "alpha", "expected_depths", "median_depths", "expected_normals"]):
        ok, msg = close_report(nm, got[i], ref[i], mask=keep)
        assert ok, msg
    if not case["backward"]:
        return
    # The gradients are compared with an fp64 evaluation, which only makes sense on pixels whose discrete decisions
    # (alpha >= 1/255, T <= 1e-4, median crossing) fp32 and fp64 take alike: wider margins than for the images.
    with torch.no_grad():
        wide = O.rasterization(*[t.detach() for t in cpu], vm, Ks, W, H, return_aux=True, compositor="c",
                               discrete_from=rmeta, fragile_scale=10.0, **kw)[5]["fragile"]
    keep64 = keep & ~wide
    assert int((~keep64).sum()) < keep64.numel() // 5
    g = torch.Generator().manual_seed(4)
    ws = [torch.randn(t.shape, generator=g) * keep64[..., None] for t in ref[:5]]
    sum((t * w).sum() for t, w in zip(ref[:5], ws)).backward()
    sum((t * w.to(cuda_dev)).sum() for t, w in zip(got[:5], ws)).backward()
    d64 = [t.detach().double().clone().requires_grad_(True) for t in params]
    # fp64 ground truth along the fp32 run's discrete decisions (radii, tile lists): the render is discontinuous in them
    r64 = O.rasterization(*d64, vm.double(), Ks.double(), W, H, compositor="c", discrete_from=rmeta, **kw)
    sum((t * w.double()).sum() for t, w in zip(r64[:5], ws)).backward()
    report = []
    all_ok = True
    for nm, a, b32, b64 in zip(("means", "quats", "scales", "opacities", "colors"), gpu, cpu, d64):
        ok, msg = grad_parity_report("v_" + nm, a.grad, b64.grad, b32.grad, outlier_frac=2e-6)
        report.append(msg)
        all_ok &= ok
        assert float(b64.grad.abs().max()) > 0
    print("\n".join(report))
    assert all_ok, "\n".join(report)


# ------------------------------------------------------------------------------------------------ fused loss (8f row f1)
@pytest.mark.parametrize("use_dn,with_bg,D", [(True, False, 4), (True, True, 3), (False, False, 3), (True, False, 12)])
def test_fused_loss_matches_reference_glue(cuda_dev, use_dn, with_bg, D):
    """csrc/loss.cu against the reference's own post-render arithmetic (camera_utils.py:176-279,
    rade_gs_model.py:202-219,292-307) restated in torch (radegs_b200.losses / the oracle)."""
    from radegs_b200.losses import fused_rade_loss
    g = torch.Generator().manual_seed(2)
    H, W = 70, 93
    fx, fy = 110.0, 95.0
    K = torch.tensor([[fx, 0, W / 2], [0, fy, H / 2], [0, 0, 1.0]])
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    base = 3.0 + 0.01 * xx + 0.02 * yy + 0.3 * torch.sin(xx / 7.0) * torch.cos(yy / 5.0)
    leaves = dict(render=torch.rand(H, W, D, generator=g) * 1.4 - 0.2, alpha=torch.rand(H, W, generator=g),
                  de=base + 0.05 * torch.rand(H, W, generator=g), dm=base + 0.05 * torch.rand(H, W, generator=g),
                  nrm=torch.nn.functional.normalize(torch.randn(H, W, 3, generator=g), dim=-1) * 0.9)
    gt = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8)
    bg = torch.rand(3, generator=g) if with_bg else None
    cpu = {k: v.clone().requires_grad_(True) for k, v in leaves.items()}
    rgb = cpu["render"][..., :3] + ((1 - cpu["alpha"])[..., None] * bg if with_bg else 0.0)
    ref = (torch.clamp(rgb, 0, 1) - gt.float() / 255.0).abs().mean()
    if use_dn:
        ref = ref + O.depth_normal_loss(K, W, H, cpu["de"], cpu["dm"], cpu["nrm"])[0]
    (ref * 3.0).backward()
    gpu = {k: v.clone().to(cuda_dev).requires_grad_(True) for k, v in leaves.items()}
    loss, terms = fused_rade_loss(gpu["render"], gpu["alpha"], gpu["de"], gpu["dm"], gpu["nrm"], gt.to(cuda_dev), fx, fy,
                                  background=None if bg is None else bg.to(cuda_dev), use_depth_normal=use_dn)
    (loss * 3.0).backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item())), (loss.item(), ref.item())
    assert abs(float(terms[:3].sum()) - loss.item()) < 1e-6
    for k in leaves:
        ok, msg = grad_close_report("v_" + k, gpu[k].grad, cpu[k].grad, rel=1e-3, floor=1e-9)
        assert ok, msg


@pytest.mark.parametrize("degree,with_depth,views", [(0, False, 1), (2, True, 2), (3, True, 1), (3, False, 3)])
def test_sh_colors_fused(cuda_dev, degree, with_depth, views):
    """csrc/colors.cu against the chain it replaces inside rasterization():
    inverse(viewmats) -> dirs -> spherical_harmonics(masks) -> +0.5 -> clamp_min(0) -> cat(depth)."""
    from gsplat.cuda._wrapper import sh_colors
    cfg, gs, vm, Ks = small_scene(n=5000, views=views, spread=1.8)
    means, quats, scales, _, sh = scenes.activate(gs, 3)
    sh = sh * 6.0          # large coefficients so that the clamp at 0 is active for some Gaussians
    radii, _, depths = O.fully_fused_projection(means, quats, scales, vm, Ks, cfg.width, cfg.height)[:3]
    mc, sc = means.clone().requires_grad_(True), sh.clone().requires_grad_(True)
    dc = depths.clone().requires_grad_(True)
    campos = torch.linalg.inv(vm)[:, :3, 3]
    ref = O.spherical_harmonics(degree, mc[None] - campos[:, None], sc[None].expand(views, -1, -1, -1),
                                masks=(radii > 0).all(-1))
    ref = torch.clamp_min(ref + 0.5, 0.0)
    ref = torch.cat([ref, dc[..., None] if with_depth else torch.zeros(views, cfg.n_gaussians, 1)], dim=-1)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(4))
    (ref * w).sum().backward()
    mg, sg = means.to(cuda_dev).requires_grad_(True), sh.to(cuda_dev).requires_grad_(True)
    dg = depths.to(cuda_dev).requires_grad_(True)
    got = sh_colors(degree, mg, sg, vm.to(cuda_dev), radii.to(cuda_dev), dg if with_depth else None)
    (got * w.to(cuda_dev)).sum().backward()
    assert 0.02 < float((ref[..., :3] == 0).float().mean()) < 0.9
    ok, msg = close_report("colors4", got, ref, atol=2e-5, rtol=1e-4)
    assert ok, msg
    ok, msg = grad_close_report("v_coeffs", sg.grad, sc.grad, rel=1e-4)
    assert ok, msg
    ok, msg = grad_close_report("v_means", mg.grad, mc.grad, rel=2e-3)
    assert ok, msg
    if with_depth:
        ok, msg = grad_close_report("v_depths", dg.grad, dc.grad, rel=1e-6)
        assert ok, msg


@pytest.mark.parametrize("degree,cams", [(3, [1, 2, 1]), (1, [2, 2]), (0, [1]), (3, [1] * 8)])
def test_sh_gradient_split_over_virtual_ranks(cuda_dev, degree, cams):
    """Camera-sharded form of the SH backward (csrc/colors.cu: rs_sh_colors_bwd_local on every 'rank', then
    rs_sh_coeffs_gather over all regions) against the oracle's autograd over ALL cameras at once: the coefficient
    gradient rebuilt from the published colour gradients equals the sum the all-reduce would have produced."""
    import ctypes as ct
    from radegs_b200 import backend as be
    lib = be.load()
    views = sum(cams)
    cfg, gs, vm, Ks = small_scene(n=5000, views=views, spread=1.8)
    means, quats, scales, _, sh = scenes.activate(gs, 3)
    sh = sh * 6.0
    radii, _, depths = O.fully_fused_projection(means, quats, scales, vm, Ks, cfg.width, cfg.height)[:3]
    N, K = means.shape[0], sh.shape[1]
    mc, sc = means.clone().requires_grad_(True), sh.clone().requires_grad_(True)
    campos = torch.linalg.inv(vm)[:, :3, 3]
    ref = O.spherical_harmonics(degree, mc[None] - campos[:, None], sc[None].expand(views, -1, -1, -1),
                                masks=(radii > 0).all(-1))
    ref = torch.clamp_min(ref + 0.5, 0.0)
    w = torch.randn(views, N, 4, generator=torch.Generator().manual_seed(4))
    (ref * w[..., :3]).sum().backward()
    md, sd = means.to(cuda_dev), sh.to(cuda_dev).contiguous()
    st = be.stream_ptr(cuda_dev)
    regions, v_means_sum, c0 = [], torch.zeros(N, 3, device=cuda_dev), 0
    for C in cams:
        region = torch.zeros(lib.rs_sh_region_bytes(C, N), device=cuda_dev, dtype=torch.uint8)
        v_means = torch.empty(N, 3, device=cuda_dev)
        v_depths = torch.empty(C, N, device=cuda_dev)
        vm_g, radii_g = vm[c0:c0 + C].to(cuda_dev).contiguous(), radii[c0:c0 + C].to(cuda_dev).contiguous()
        w_g = w[c0:c0 + C].to(cuda_dev).contiguous()       # named: the pointers must outlive the launch
        be.check(lib.rs_sh_colors_bwd_local(degree, K, C, N, be.ptr(md), be.ptr(sd), be.ptr(vm_g), be.ptr(radii_g),
                                            be.ptr(w_g), 1, be.ptr(region), be.ptr(v_means), be.ptr(v_depths), st),
                 "local")
        assert torch.equal(v_depths.cpu(), w[c0:c0 + C, :, 3])
        v_means_sum += v_means
        regions.append(region)
        c0 += C
    v_coeffs = torch.full((N, K, 3), float("nan"), device=cuda_dev)
    ptrs = (ct.c_void_p * len(cams))(*[r.data_ptr() for r in regions])
    be.check(lib.rs_sh_coeffs_gather(degree, K, N, be.ptr(md), ptrs, (ct.c_int * len(cams))(*cams), len(cams),
                                     be.ptr(v_coeffs), st), "gather")
    ok, msg = grad_close_report("v_coeffs (gathered)", v_coeffs, sc.grad, rel=1e-4)
    assert ok, msg
    ok, msg = grad_close_report("v_means (sum of the local parts)", v_means_sum, mc.grad, rel=2e-3)
    assert ok, msg
    nb = (degree + 1) ** 2
    assert bool((v_coeffs[:, nb:] == 0).all()), "bands above the active degree get exact zeros"


@pytest.mark.parametrize("mode", ["push", "p2p", "allgather"])
def test_sh_grad_exchange_single_process(cuda_dev, mode):
    """ShGradExchange with one rank (signal to self / local gather) through the full rasterization() backward:
    same gradients as the plain path, on two consecutive steps (alternating regions)."""
    from gsplat.rendering import rasterization
    from radegs_b200.multiview import ShGradExchange
    cfg, gs, vm, Ks = small_scene(n=3000, w=96, h=64, views=2)
    params = scenes.activate(gs, 3)

    def run(exchange):
        leaves = [p.detach().to(cuda_dev).requires_grad_(True) for p in params]
        out = rasterization(*leaves, vm.to(cuda_dev), Ks.to(cuda_dev), cfg.width, cfg.height, sh_degree=3, packed=False,
                            render_mode="RGB+ED", rasterize_mode="antialiased", return_depth_normal=True)
        g = torch.Generator().manual_seed(0)
        loss = sum((o * torch.randn(o.shape, generator=g).to(cuda_dev)).sum() for o in out[:5])
        if exchange is None:
            loss.backward()
        else:
            exchange.begin_step()
            with exchange:
                loss.backward()
            assert leaves[4].grad is None
            leaves[4].grad = exchange.finish()
        return [l.grad.clone() for l in leaves]

    ref = run(None)
    ex = ShGradExchange(3000, 2, cuda_dev, mode=mode)
    try:
        for _ in range(3):
            got = run(ex)
            for a, b, nm in zip(got, ref, ["means", "quats", "scales", "opacities", "sh"]):
                ok, msg = grad_close_report("v_" + nm, a, b, rel=1e-5)
                assert ok, msg
        ex.check()
    finally:
        ex.close()


@pytest.mark.parametrize("cfg_id", [2, 3])
def test_full_size_adjoint_identity(cuda_dev, cfg_id):
    """BASELINE config 2 (1M, 1080p, D=3) and config 3 (500k, 960x540, 3+64 feature channels) at full size.
    The render is exactly linear in the colours, so forward and backward must satisfy the adjoint identity
    <render(c), w> = <c, d<render,w>/dc> and additivity render(a+b) = render(a)+render(b) -- size-independent
    checks of the wide-channel forward AND backward where the oracle would take hours."""
    from gsplat.rendering import rasterization
    cfg = scenes.BASELINE_CONFIGS[cfg_id]
    gs, vm, Ks = scenes.make_scene(cfg, n_views=1)
    means, quats, scales, opac, _ = _gpu(scenes.activate(gs, cfg.sh_degree)[:5], cuda_dev)
    D = 3 + cfg.n_features
    g = torch.Generator(device=cuda_dev).manual_seed(5)
    ca = torch.rand(cfg.n_gaussians, D, device=cuda_dev, generator=g).requires_grad_(True)
    cb = torch.rand(cfg.n_gaussians, D, device=cuda_dev, generator=g)
    vmd, Kd = vm.to(cuda_dev), Ks.to(cuda_dev)

    def render(c):
        return rasterization(means, quats, scales, opac, c, vmd, Kd, cfg.width, cfg.height, packed=False,
                             rasterize_mode="antialiased", return_depth_normal=True)[0]

    ra = render(ca)
    w = torch.randn(ra.shape, device=cuda_dev, generator=g)
    lhs = (ra.double() * w.double()).sum()
    (ra * w).sum().backward()
    rhs = (ca.detach().double() * ca.grad.double()).sum()
    assert abs(lhs.item() - rhs.item()) <= 2e-4 * max(abs(lhs.item()), float((ra.abs() * w.abs()).sum()) * 1e-3), \
        (lhs.item(), rhs.item())
    with torch.no_grad():
        rb, rab = render(cb), render(ca.detach() + cb)
    assert float((rab - ra.detach() - rb).abs().max()) < 5e-5 * D ** 0.5
    assert float(ra.detach().abs().max()) > 0.5


def test_multi_camera_batch_equals_single_camera_calls(cuda_dev):
    """C cameras in one call (flatten ids c*N+n, camera bits in the sort keys) must reproduce C single-camera
    calls: forward images bit-identical, gradients equal to the sum over cameras (config 4's batch axis)."""
    from gsplat.rendering import rasterization
    cfg = scenes.SceneConfig("mc", 200_000, 640, 360, 3, 3, 0, 77)
    gs, vm, Ks = scenes.make_scene(cfg)
    gs["log_scales"] = gs["log_scales"] + 0.5
    params = scenes.activate(gs, 3)
    vmd, Kd = vm.to(cuda_dev), Ks.to(cuda_dev)
    kw = dict(width=cfg.width, height=cfg.height, packed=False, sh_degree=3, render_mode="RGB+ED",
              rasterize_mode="antialiased", return_depth_normal=True)
    g = torch.Generator(device=cuda_dev).manual_seed(1)
    batch = _gpu(params, cuda_dev, grad=True)
    out = rasterization(*batch, viewmats=vmd, Ks=Kd, **kw)
    ws = [torch.randn(t.shape, device=cuda_dev, generator=g) for t in out[:5]]
    sum((t * w).sum() for t, w in zip(out[:5], ws)).backward()
    single = _gpu(params, cuda_dev, grad=True)
    for c in range(3):
        o = rasterization(*single, viewmats=vmd[c:c + 1], Ks=Kd[c:c + 1], **kw)
        for k in range(5):
            assert torch.equal(o[k][0], out[k][c]), f"camera {c} output {k} differs between batch and single call"
        sum((t * w[c:c + 1]).sum() for t, w in zip(o[:5], ws)).backward()
    for nm, a, b in zip(("means", "quats", "scales", "opacities", "sh"), batch, single):
        ok, msg = grad_close_report("v_" + nm, a.grad, b.grad, rel=1e-3)
        assert ok, msg


@pytest.mark.parametrize("views,w,h,n,boost", [(1, 160, 96, 6000, 1.2), (3, 250, 130, 5000, 1.2), (1, 64, 48, 20000, 2.0),
                                               (2, 16, 16, 300, 1.0), (1, 1000, 700, 3000, 0.5)])
def test_compact_isect_pipeline_matches_radix_path(cuda_dev, views, w, h, n, boost):
    """The opt-in "compact" intersection pipeline (32-bit camera|tile keys through the radix passes) against the
    default path and the oracle: bit-exact keys, flatten ids, offsets and tile counts, including equal (tile, depth)
    keys whose order is the tie rule."""
    from gsplat.cuda._wrapper import (fully_fused_projection, isect_offset_encode, isect_tiles,
                                      isect_tiles_and_offsets)
    cfg, gs, vm, Ks = small_scene(n=n, w=w, h=h, views=views, spread=1.3, scale_boost=boost)
    means, quats, scales, _, _ = scenes.activate(gs, 3)
    # duplicate some Gaussians so that equal (tile, depth) keys exist and the tie order matters
    means = torch.cat([means, means[:200]]); quats = torch.cat([quats, quats[:200]]); scales = torch.cat([scales, scales[:200]])
    m, q, s, v, k = _gpu((means, quats, scales, vm, Ks), cuda_dev)
    radii, means2d, depths = fully_fused_projection(m, None, q, s, v, k, w, h)[:3]
    tw, th = math.ceil(w / 16), math.ceil(h / 16)
    t1, i1, f1 = isect_tiles(means2d, radii, depths, 16, tw, th)
    o1 = isect_offset_encode(i1, views, tw, th)
    assert i1.numel() > 0
    t3, i3, f3, o3 = isect_tiles_and_offsets(means2d, radii, depths, 16, tw, th, method="compact")
    assert torch.equal(t1, t3) and torch.equal(o1, o3) and torch.equal(i1, i3) and torch.equal(f1, f3)
    r_t, r_i, r_f = O.isect_tiles(means2d.cpu(), radii.cpu(), depths.cpu(), 16, tw, th)
    assert torch.equal(i3.cpu(), r_i) and torch.equal(f3.cpu(), r_f)
    assert torch.equal(o3.cpu(), O.isect_offset_encode(r_i, views, tw, th))
    assert int((i1[1:] == i1[:-1]).sum()) > 20, "the scene must contain equal keys"


def test_compact_isect_pipeline_nothing_visible(cuda_dev):
    from gsplat.cuda._wrapper import isect_tiles_and_offsets
    N = 1000
    means2d = torch.rand(1, N, 2, device=cuda_dev) * 100
    radii = torch.zeros(1, N, 2, device=cuda_dev, dtype=torch.int32)
    depths = torch.rand(1, N, device=cuda_dev)
    t, i, f, o = isect_tiles_and_offsets(means2d, radii, depths, 16, 7, 7, method="compact")
    assert i.numel() == 0 and f.numel() == 0 and int(t.sum()) == 0 and int(o.abs().sum()) == 0 and o.shape == (1, 7, 7)


def test_forward_two_pixels_per_lane_variant(cuda_dev):
    """The two-pixels-per-lane forward (8x8 pixels per warp, the default) against the oracle and against the
    one-pixel-per-lane kernel (RS_RASTER_ONE_PIXEL), incl. multi-batch tiles, image edges and a background."""
    from gsplat.cuda import _wrapper as W
    from gsplat.cuda._wrapper import rasterize_to_pixels
    from radegs_b200 import backend as be
    for (n, w, h, views) in ((2500, 150, 90, 2), (9000, 64, 48, 1)):
        cfg, gs, vm, Ks = small_scene(n=n, w=w, h=h, views=views)
        inp, offs, flat = _raster_inputs(cfg, gs, vm, Ks, 4)
        bg = torch.rand(views, 4, generator=torch.Generator().manual_seed(3))
        ref = O.rasterize_to_pixels(inp["means2d"], inp["conics"], inp["colors"], inp["opacities"], inp["ray_ts"],
                                    inp["ray_planes"], inp["normals"], Ks, w, h, 16, offs, flat, backgrounds=bg,
                                    return_aux=True)
        d = {k: v.to(cuda_dev) for k, v in inp.items()}
        outs = []
        try:
            for flags in (be.RS_RASTER_ONE_PIXEL, 0):
                W.RASTER_FLAGS = flags
                outs.append(rasterize_to_pixels(d["means2d"], d["conics"], d["colors"], d["opacities"], w, h, 16,
                                                offs.to(cuda_dev), flat.to(cuda_dev), backgrounds=bg.to(cuda_dev),
                                                ray_ts=d["ray_ts"], ray_planes=d["ray_planes"], normals=d["normals"],
                                                Ks=Ks.to(cuda_dev), return_ids=True))
        finally:
            W.RASTER_FLAGS = 0
        keep = ~ref[5]["fragile"]
        for i, nm in enumerate(["colors", "alphas", "expected_depths", "median_depths", "normals"]):
            ok, msg = close_report(nm, outs[1][i], ref[i], mask=keep)
            assert ok, msg
            ok, msg = close_report(nm + " (variant 1 vs 0)", outs[1][i], outs[0][i], mask=keep)
            assert ok, msg
        assert bool(((outs[1][5].cpu() == ref[5]["last_ids"]) | ref[5]["fragile"]).all())
        assert bool(((outs[1][6].cpu() == ref[5]["median_ids"]) | ref[5]["fragile"]).all())


# ------------------------------------------------------------------------------------------------ row f4
@pytest.mark.parametrize("absgrad,views,track_radii", [(False, 1, True), (True, 3, True), (False, 2, False)])
def test_densify_stats_fused_matches_host_sequence(cuda_dev, absgrad, views, track_radii):
    """DefaultStrategy._update_state on CUDA tensors (csrc/stats.cu) against the same method on host copies, i.e.
    against the published gsplat sequence, after two accumulation steps."""
    from gsplat.strategy import DefaultStrategy
    N, W, H = 5000, 320, 200
    g = torch.Generator().manual_seed(11)
    strat = DefaultStrategy(absgrad=absgrad, refine_scale2d_stop_iter=1000 if track_radii else 0)
    st_cpu, st_gpu = strat.initialize_state(), strat.initialize_state()
    params = {"means": torch.zeros(N, 3)}
    for step in range(2):
        radii = torch.randint(-1, 40, (views, N, 2), generator=g, dtype=torch.int32).clamp_min(0)
        radii[torch.rand(views, N, generator=g) < 0.3] = 0
        grad = torch.randn(views, N, 2, generator=g) * 1e-3

        class M:   # stands in for meta["means2d"] after the backward
            pass
        infos = []
        for dev in ("cpu", cuda_dev):
            m = M()
            m.grad = grad.to(dev)
            m.absgrad = grad.abs().to(dev)
            infos.append({"width": W, "height": H, "n_cameras": views, "radii": radii.to(dev), "means2d": m})
        strat._update_state(params, st_cpu, infos[0])
        strat._update_state({"means": params["means"].to(cuda_dev)}, st_gpu, infos[1])
        # running maximum screen radius over ALL cameras that see the Gaussian.  (Upstream's
        # `state["radii"][ids] = maximum(state["radii"][ids], r)` keeps an arbitrary camera's value when a Gaussian is
        # visible in several cameras of one call -- duplicate indices in an index_put; with the reference's C == 1
        # the two agree, and the kernel implements the maximum that line is meant to compute.)
        vis = (radii > 0).all(-1)
        r_step = torch.where(vis, radii.max(-1).values.float() / max(W, H), torch.zeros(())).max(dim=0).values
        r_ref = r_step if step == 0 else torch.maximum(r_ref, r_step)
    for key in ["grad2d", "count"]:
        a, b = st_gpu[key].cpu(), st_cpu[key]
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-9), (key, float((a - b).abs().max()))
    if track_radii:
        assert torch.allclose(st_gpu["radii"].cpu(), r_ref, rtol=1e-6), float((st_gpu["radii"].cpu() - r_ref).abs().max())
        if views == 1:
            assert torch.allclose(st_gpu["radii"].cpu(), st_cpu["radii"], rtol=1e-6)
    assert float(st_cpu["count"].max()) == 2 * views


def test_project_gaussians_on_device(cuda_dev):
    """radegs_b200.meta_utils.project_gaussians against collab_splats/utils/utils.py:13-40 restated with torch."""
    from radegs_b200.meta_utils import project_gaussians
    N, W, H = 20000, 333, 201
    g = torch.Generator().manual_seed(5)
    means2d = (torch.rand(1, N, 2, generator=g) * 1.4 - 0.2) * torch.tensor([W, H])
    means2d[0, :50] = torch.arange(50)[:, None] + 0.5          # exact .5 ties: torch.round is half-to-even
    radii = torch.randint(0, 4, (1, N, 2), generator=g, dtype=torch.int32)
    depths = torch.rand(1, N, generator=g)
    meta = {"width": W, "height": H, "radii": radii.to(cuda_dev), "means2d": means2d.to(cuda_dev),
            "depths": depths.to(cuda_dev)}
    got = project_gaussians(meta)
    r = radii.squeeze()
    valid = (r > 1.0).sum(dim=1) > 0
    xy = torch.round(means2d).squeeze().long()
    flat = torch.clamp(xy[:, 0], 0, W - 1) + torch.clamp(xy[:, 1], 0, H - 1) * W
    assert torch.equal(got["proj_flattened"].cpu(), flat)
    assert torch.equal(got["valid_mask"].cpu(), valid)
    assert torch.equal(got["gaussian_ids"].cpu(), valid.nonzero(as_tuple=False).squeeze())
    assert torch.equal(got["proj_depths"].cpu(), depths.squeeze())
    assert all(v.is_cuda for v in got.values())
    assert all(not v.is_cuda for v in project_gaussians(meta, to_cpu=True).values())
