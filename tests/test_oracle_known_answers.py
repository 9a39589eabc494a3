"""Analytic known-answer tests that pin the CPU oracle ("known-answer tests we can still
derive analytically").  The reference holds no golden vectors for this path (tests/test_models.py:45-63 only
construct the models), so the oracle is pinned by geometry whose answer is known in closed form, by
invariants of the algorithm, by finite differences, and by the reference's own depth->normal code
(collab_splats/utils/camera_utils.py:176-279) restated in oracle.depth_double_to_normal.
"""

import math

import numpy as np
import pytest
import torch

from oracle import rade_oracle as O
from radegs_b200 import scenes


def _cam(W=64, H=64, f=100.0, z=4.0, dtype=torch.float64):
    vm = torch.eye(4, dtype=dtype)[None].clone()
    vm[0, 2, 3] = z
    Ks = torch.tensor([[[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]]], dtype=dtype)
    return vm, Ks


def test_single_isotropic_gaussian_on_axis():
    """(i) one isotropic Gaussian facing the camera: depth = z, normal = (0,0,-1), alpha = o*exp(-sigma)."""
    dt = torch.float64
    W = H = 64
    vm, Ks = _cam(W, H)
    means = torch.zeros(1, 3, dtype=dt)
    quats = torch.tensor([[1.0, 0, 0, 0]], dtype=dt)
    s = 0.2
    scales = torch.full((1, 3), s, dtype=dt)
    opac = torch.tensor([0.9], dtype=dt)
    cols = torch.tensor([[0.2, 0.5, 0.8]], dtype=dt)
    radii, m2, depths, conics, comps, ray_ts, ray_planes, normals = O.fully_fused_projection(
        means, quats, scales, vm, Ks, W, H, calc_compensations=True)
    sig2 = (100.0 * s / 4.0) ** 2
    assert torch.allclose(m2[0, 0], torch.tensor([32.0, 32.0], dtype=dt))
    assert abs(depths[0, 0].item() - 4.0) < 1e-12 and abs(ray_ts[0, 0].item() - 4.0) < 1e-12
    assert torch.allclose(conics[0, 0], torch.tensor([1 / (sig2 + 0.3), 0.0, 1 / (sig2 + 0.3)], dtype=dt))
    assert radii[0, 0].tolist() == [math.ceil(3.33 * math.sqrt(sig2 + 0.3))] * 2
    assert abs(comps[0, 0].item() - math.sqrt(sig2 * sig2 / (sig2 + 0.3) ** 2)) < 1e-12
    assert torch.allclose(normals[0, 0], torch.tensor([0.0, 0.0, -1.0], dtype=dt), atol=1e-12)
    assert torch.allclose(ray_planes[0, 0], torch.zeros(2, dtype=dt), atol=1e-12)
    rc, ra, de, dm, nr, meta = O.rasterization(means, quats, scales, opac, cols, vm, Ks, W, H,
                                               return_depth_normal=True)
    # pixel (31,31): centre (31.5,31.5), offset (0.5,0.5) from the mean
    sigma = 0.5 * (0.25 + 0.25) / (sig2 + 0.3)
    a = 0.9 * math.exp(-sigma)
    assert abs(ra[0, 31, 31, 0].item() - a) < 1e-12
    assert torch.allclose(rc[0, 31, 31], a * cols[0])
    ln = math.sqrt((0.5 / 100) ** 2 * 2 + 1)
    assert abs(de[0, 31, 31, 0].item() - a * 4.0 / ln) < 1e-9        # raw sum(vis*t)/ln (Q1)
    assert abs(dm[0, 31, 31, 0].item() - 4.0 / ln) < 1e-9            # T: 1 -> 0.1 crosses 0.5
    assert torch.allclose(nr[0, 31, 31], a * torch.tensor([0.0, 0.0, -1.0], dtype=dt), atol=1e-12)
    # far corner: alpha < 1/255 -> nothing
    assert ra[0, 0, 0, 0].item() == 0.0 and dm[0, 0, 0, 0].item() == 0.0


@pytest.mark.parametrize("tilt_deg", [0.0, 25.0, -40.0])
def test_tilted_planar_sheet_depth_and_normal(tilt_deg):
    """(ii) a dense sheet of flat Gaussians on a tilted plane: rendered normal = plane normal, median and
    alpha-normalised expected depth = plane depth, and the reference's depth->normal agrees."""
    dt = torch.float64
    W, H, f, z0 = 96, 64, 120.0, 4.0
    vm, Ks = _cam(W, H, f, z0)
    th = math.radians(tilt_deg)
    # plane through (0,0,0) (camera z = z0) spanned by e1 = (cos th, 0, sin th), e2 = (0,1,0)
    e1 = torch.tensor([math.cos(th), 0.0, math.sin(th)], dtype=dt)
    e2 = torch.tensor([0.0, 1.0, 0.0], dtype=dt)
    n_plane = torch.linalg.cross(e1, e2)          # (-sin th, 0, cos th)
    if n_plane[2] > 0:
        n_plane = -n_plane                        # facing the camera (camera looks along +z)
    gu, gv = torch.meshgrid(torch.linspace(-3.4, 3.4, 140, dtype=dt), torch.linspace(-1.6, 1.6, 66, dtype=dt),
                            indexing="ij")
    means = gu.reshape(-1, 1) * e1 + gv.reshape(-1, 1) * e2
    N = means.shape[0]
    # rotation taking local x->e1, y->e2, z->normal  (rotation about y by -th): quaternion
    q = torch.tensor([math.cos(-th / 2), 0.0, math.sin(-th / 2), 0.0], dtype=dt)
    quats = q[None].repeat(N, 1)
    scales = torch.tensor([[0.06, 0.06, 0.0005]], dtype=dt).repeat(N, 1)
    opac = torch.full((N,), 0.95, dtype=dt)
    cols = torch.full((N, 3), 0.5, dtype=dt)
    rc, ra, de, dm, nr, meta = O.rasterization(means, quats, scales, opac, cols, vm, Ks, W, H,
                                               return_depth_normal=True)
    inner = (slice(8, H - 8), slice(8, W - 8))
    a = ra[0][inner][..., 0]
    assert a.min().item() > 0.99
    # analytic plane depth per pixel: ray r=((x-cx)/f,(y-cy)/f,1), point = z r with n.(z r - p0) = 0, p0=(0,0,z0)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt) + 0.5, torch.arange(W, dtype=dt) + 0.5, indexing="ij")
    r = torch.stack([(xs - W / 2) / f, (ys - H / 2) / f, torch.ones_like(xs)], dim=-1)
    z_plane = (n_plane[2] * z0) / (r @ n_plane)
    nrm = nr[0][inner] / a[..., None]
    assert (nrm - n_plane).abs().max().item() < 2e-2
    assert ((de[0, ..., 0] / ra[0, ..., 0])[inner] - z_plane[inner]).abs().max().item() < 2e-2
    assert (dm[0, ..., 0][inner] - z_plane[inner]).abs().max().item() < 3e-2
    # the reference's depth -> normal stencil on these maps reproduces the rendered normal
    n_d = O.depth_double_to_normal(Ks[0], W, H, de[0, ..., 0] / ra[0, ..., 0], dm[0, ..., 0])
    assert (n_d[0][inner] - n_plane).abs().max().item() < 3e-2
    assert (n_d[1][inner] - n_plane).abs().max().item() < 1e-1   # median depth is piecewise: noisier
    _, err = O.depth_normal_loss(Ks[0], W, H, de[0, ..., 0] / ra[0, ..., 0], dm[0, ..., 0], nrm_full(nr, ra))
    assert err[:, 8:H - 8, 8:W - 8].abs().max().item() < 1e-2


def nrm_full(nr, ra):
    return nr[0] / ra[0].clamp(min=1e-6)


def test_transmittance_telescopes_and_outputs_are_bounded():
    """(iii) sum_i vis_i + T_final = 1: rendering a constant colour 1 gives exactly alpha."""
    cfg = scenes.SceneConfig("t", 3000, 96, 64, 1, None, 0, 3)
    gs, vm, Ks = scenes.make_scene(cfg, dtype=torch.float64)
    gs["log_scales"] += 1.2
    means, quats, scales, opac, _ = scenes.activate(gs, None)
    ones = torch.ones(cfg.n_gaussians, 3, dtype=torch.float64)
    rc, ra, de, dm, nr, meta = O.rasterization(means, quats, scales, opac, ones, vm, Ks, 96, 64,
                                               return_depth_normal=True)
    assert (rc - ra).abs().max().item() < 1e-12
    assert ra.min().item() >= 0 and ra.max().item() <= 1 - 1e-4 + 1e-12
    assert nr.norm(dim=-1).max().item() <= 1 + 1e-9


def test_isect_invariants():
    """(v) keys sorted, offsets monotone, counts add up, every listed tile is inside the Gaussian's bbox."""
    cfg = scenes.SceneConfig("t", 5000, 200, 120, 3, None, 0, 5)
    gs, vm, Ks = scenes.make_scene(cfg)
    gs["log_scales"] += 1.0
    means, quats, scales, _, _ = scenes.activate(gs, None)
    radii, m2, depths = O.fully_fused_projection(means * 1.5, quats, scales, vm, Ks, 200, 120)[:3]
    tw, th = math.ceil(200 / 16), math.ceil(120 / 16)
    tiles, ids, flat = O.isect_tiles(m2, radii, depths, 16, tw, th)
    offs = O.isect_offset_encode(ids, 3, tw, th)
    M = ids.numel()
    assert M == int(tiles.sum()) and M > 5000
    assert bool((ids[1:] >= ids[:-1]).all())
    o = offs.flatten()
    assert bool((o[1:] >= o[:-1]).all()) and o[0] == 0 and o[-1] <= M
    tile_bits = (tw * th).bit_length()
    cam = (ids >> (32 + tile_bits)).long()
    tile = ((ids >> 32) & ((1 << tile_bits) - 1)).long()
    assert bool((cam == flat.long() // 5000).all())
    ty, tx = tile // tw, tile % tw
    mm = m2.reshape(-1, 2)[flat.long()]
    rr = radii.reshape(-1, 2)[flat.long()].float()
    assert bool(((tx + 1) * 16 > mm[:, 0] - rr[:, 0]).all()) and bool((tx * 16 < mm[:, 0] + rr[:, 0]).all())
    assert bool(((ty + 1) * 16 > mm[:, 1] - rr[:, 1]).all()) and bool((ty * 16 < mm[:, 1] + rr[:, 1]).all())
    dbits = torch.from_numpy((ids.numpy() & 0xFFFFFFFF).astype(np.uint32).view(np.float32))
    assert torch.equal(dbits, depths.reshape(-1)[flat.long()])
    # offsets index the first key of each (camera, tile)
    key_hi = ids >> 32
    for t in (0, 7, tw * th + 3, 3 * tw * th - 1):
        s, e = int(o[t]), int(o[t + 1]) if t + 1 < o.numel() else M
        q = ((t // (tw * th)) << tile_bits) | (t % (tw * th))
        assert bool((key_hi[s:e] == q).all())
        assert s == 0 or key_hi[s - 1] < q


def test_gradients_match_finite_differences_fp64():
    """(iv) oracle autograd vs central finite differences on a tiny scene (robust pixels only)."""
    torch.manual_seed(0)
    dt = torch.float64
    cfg = scenes.SceneConfig("t", 40, 32, 32, 1, None, 0, 9)
    gs, vm, Ks = scenes.make_scene(cfg, dtype=dt)
    gs["log_scales"] += 2.2
    means, quats, scales, opac, cols = [t.clone() for t in scenes.activate(gs, None)]
    w = [torch.randn(1, 32, 32, k, dtype=dt) for k in (3, 1, 1, 3)]   # colours, alpha, expected depth, normals

    def f(m, q, s, o, c):
        rc, ra, de, dm, nr, meta = O.rasterization(m, q, s, o, c, vm, Ks, 32, 32, return_depth_normal=True)
        return (rc * w[0]).sum() + (ra * w[1]).sum() + (de * w[2]).sum() + (nr * w[3]).sum()

    leaves = [t.clone().requires_grad_(True) for t in (means, quats, scales, opac, cols)]
    f(*leaves).backward()
    rng = np.random.default_rng(0)
    for li, leaf in enumerate(leaves):
        flat = leaf.detach().reshape(-1)
        for idx in rng.choice(flat.numel(), size=6, replace=False):
            h = 1e-6 * max(1.0, abs(flat[idx].item()))
            args_p = [t.detach().clone() for t in leaves]
            args_m = [t.detach().clone() for t in leaves]
            args_p[li].reshape(-1)[idx] += h
            args_m[li].reshape(-1)[idx] -= h
            fd = (f(*args_p) - f(*args_m)).item() / (2 * h)
            an = leaf.grad.reshape(-1)[idx].item()
            assert abs(fd - an) <= 2e-4 * max(1.0, abs(an), abs(fd)), (li, idx, fd, an)


def test_render_modes_and_antialiasing():
    cfg = scenes.SceneConfig("t", 1500, 64, 48, 2, 2, 0, 4)
    gs, vm, Ks = scenes.make_scene(cfg)
    gs["log_scales"] += 1.3
    p = scenes.activate(gs, 2)
    rgb = O.rasterization(*p, vm, Ks, 64, 48, sh_degree=2, render_mode="RGB")
    rgbd = O.rasterization(*p, vm, Ks, 64, 48, sh_degree=2, render_mode="RGB+D")
    rgbed = O.rasterization(*p, vm, Ks, 64, 48, sh_degree=2, render_mode="RGB+ED")
    assert rgb[0].shape == (2, 48, 64, 3) and rgbd[0].shape == (2, 48, 64, 4)
    assert torch.equal(rgb[0], rgbd[0][..., :3])
    m = rgbd[1][..., 0] > 0
    assert torch.allclose(rgbed[0][..., 3][m], (rgbd[0][..., 3] / rgbd[1][..., 0])[m])
    aa = O.rasterization(*p, vm, Ks, 64, 48, sh_degree=2, rasterize_mode="antialiased")
    assert float(aa[1].mean()) < float(rgb[1].mean())      # compensation only ever lowers opacity
    assert torch.equal(aa[2]["isect_ids"], rgb[2]["isect_ids"])


@pytest.mark.parametrize("compositor", ["torch", "c"])
@pytest.mark.parametrize("front_first_in_memory", [True, False])
def test_two_gaussian_stack_closed_form(compositor, front_first_in_memory):
    """Two isotropic Gaussians on the optical axis at depths 3 and 5: the centre pixel is the textbook front-to-back
    composite  C = a1 c1 + (1 - a1) a2 c2,  alpha = 1 - (1 - a1)(1 - a2),  expected depth = (a1 t1 + (1 - a1) a2 t2) / |ray|,
    expected normal = -(a1 + (1 - a1) a2) z,  and the median depth is that of the Gaussian under which the transmittance
    first falls below 1/2 -- whatever the order of the two in memory (sorted by depth), for both compositors."""
    dt = torch.float64
    W = H = 32
    f = 80.0
    vm = torch.eye(4, dtype=dt)[None].clone()                       # camera at the origin looking down +z
    Ks = torch.tensor([[[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]]], dtype=dt)
    z = [3.0, 5.0]
    s = [0.15, 0.4]
    o = [0.4, 0.8]
    c = [[0.9, 0.1, 0.2], [0.1, 0.7, 0.3]]
    order = [0, 1] if front_first_in_memory else [1, 0]
    means = torch.tensor([[0.0, 0.0, z[i]] for i in order], dtype=dt)
    quats = torch.tensor([[1.0, 0, 0, 0]] * 2, dtype=dt)
    scales = torch.tensor([[s[i]] * 3 for i in order], dtype=dt)
    opac = torch.tensor([o[i] for i in order], dtype=dt)
    cols = torch.tensor([c[i] for i in order], dtype=dt)
    kw = dict(compositor=compositor) if compositor == "c" else {}
    rc, ra, de, dm, nr, meta = O.rasterization(means, quats, scales, opac, cols, vm, Ks, W, H, return_depth_normal=True,
                                               **kw)
    # pixel (15,15): centre (15.5,15.5), offset (0.5,0.5) px from both projected means (16,16)
    a = []
    for i in range(2):
        sig2 = (f * s[i] / z[i]) ** 2 + 0.3                         # projected variance + the 0.3 px^2 blur
        a.append(min(0.99, o[i] * math.exp(-0.5 * (0.25 + 0.25) / sig2)))
    a1, a2 = a
    ln = math.sqrt(2 * (0.5 / f) ** 2 + 1)                          # |ray| through the pixel centre, z = 1
    # ray distance t of a fronto-parallel isotropic Gaussian at the pixel: the ray-space plane through its centre
    vis = [a1, (1 - a1) * a2]
    assert abs(ra[0, 15, 15, 0].item() - (1 - (1 - a1) * (1 - a2))) < 1e-12
    want_c = vis[0] * torch.tensor(c[0], dtype=dt) + vis[1] * torch.tensor(c[1], dtype=dt)
    assert torch.allclose(rc[0, 15, 15], want_c, atol=1e-12)
    assert torch.allclose(nr[0, 15, 15], torch.tensor([0.0, 0.0, -(vis[0] + vis[1])], dtype=dt), atol=1e-9)
    assert abs(de[0, 15, 15, 0].item() - (vis[0] * z[0] + vis[1] * z[1]) / ln) < 1e-6
    # T after the front Gaussian is 1 - a1 = 0.6+ > 1/2, after the back one (1 - a1)(1 - a2) < 1/2: median = the back one
    assert (1 - a1) > 0.5 > (1 - a1) * (1 - a2)
    assert abs(dm[0, 15, 15, 0].item() - z[1] / ln) < 1e-6
    # the tile lists are depth-ordered: front Gaussian first, whatever the memory order
    ids = meta["flatten_ids"]
    first = int(ids[0])
    assert first == order.index(0)


def _real_sh_reference(l, m, x, y, z):
    """Real spherical harmonics from the textbook Cartesian forms, written independently of the oracle (normalisation
    constants from the factorial formula N_lm = sqrt((2l+1)/(4 pi) (l-|m|)!/(l+|m|)!))."""
    def N(l_, m_):
        return math.sqrt((2 * l_ + 1) / (4 * math.pi) * math.factorial(l_ - abs(m_)) / math.factorial(l_ + abs(m_)))
    r2 = math.sqrt(2.0)
    table = {
        (0, 0): N(0, 0),
        (1, -1): r2 * N(1, 1) * y, (1, 0): N(1, 0) * z, (1, 1): r2 * N(1, 1) * x,
        (2, -2): r2 * N(2, 2) * 3 * (2 * x * y), (2, -1): r2 * N(2, 1) * 3 * y * z,
        (2, 0): N(2, 0) * 0.5 * (3 * z * z - 1), (2, 1): r2 * N(2, 1) * 3 * x * z,
        (2, 2): r2 * N(2, 2) * 3 * (x * x - y * y),
        (3, -3): r2 * N(3, 3) * 15 * y * (3 * x * x - y * y), (3, -2): r2 * N(3, 2) * 15 * (2 * x * y) * z,
        (3, -1): r2 * N(3, 1) * 1.5 * y * (5 * z * z - 1), (3, 0): N(3, 0) * 0.5 * z * (5 * z * z - 3),
        (3, 1): r2 * N(3, 1) * 1.5 * x * (5 * z * z - 1), (3, 2): r2 * N(3, 2) * 15 * (x * x - y * y) * z,
        (3, 3): r2 * N(3, 3) * 15 * x * (x * x - 3 * y * y),
    }
    return table[(l, m)]


def test_sh_basis_is_the_real_spherical_harmonics_up_to_the_3dgs_sign_convention():
    """The oracle's basis (gsplat ordering k = l^2 + l + m) against textbook real SH evaluated independently: equal up to
    the Condon-Shortley-free sign (-1)^m that the 3DGS code base uses for odd m (Y_1^{+-1}, Y_2^{+-1}, Y_3^{+-1,+-3})."""
    g = torch.Generator().manual_seed(5)
    dirs = torch.nn.functional.normalize(torch.randn(7, 3, generator=g, dtype=torch.float64), dim=-1)
    for l in range(4):
        for m in range(-l, l + 1):
            k = l * l + l + m
            coeffs = torch.zeros(7, 16, 3, dtype=torch.float64)
            coeffs[:, k, :] = 1.0
            got = O.spherical_harmonics(3, dirs, coeffs)[:, 0]
            want = torch.tensor([_real_sh_reference(l, m, *d.tolist()) for d in dirs], dtype=torch.float64)
            sign = -1.0 if (m % 2 != 0) else 1.0
            assert torch.allclose(got, sign * want, atol=1e-12), (l, m, got[:3], want[:3])
