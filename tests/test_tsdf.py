"""TSDF fusion of rendered frames (reference loop collab_splats/utils/mesh.py:1562-1632).

CPU: the numpy oracle against closed-form answers (a fronto-parallel and a tilted plane: the fused TSDF of a voxel
is its signed distance along the viewing ray / sdf_trunc, the zero crossing is on the plane, weights count frames).
GPU: ``radegs_b200.tsdf.ScalableTSDFVolume`` (csrc/tsdf.cu through the C ABI) against the oracle, bit for bit, unit
by unit, over several frames, with and without colour, strides 1 and 4, plus the host-side error behaviour."""

import numpy as np
import pytest
import torch

from oracle import tsdf_oracle as to


def _look_at(eye, target=(0.0, 0.0, 0.0)):
    """world->camera 4x4, OpenCV axes (x right, y down, z forward)."""
    eye, target = np.asarray(eye, np.float64), np.asarray(target, np.float64)
    z = target - eye
    z /= np.linalg.norm(z)
    up = np.array([0.0, 0.0, 1.0]) if abs(z[2]) < 0.9 else np.array([0.0, 1.0, 0.0])
    x = np.cross(z, up)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = -R @ eye
    return E


def _plane_depth(E, fx, fy, cx, cy, W, H, n, d0):
    """z-depth map of the world plane n.x = d0 seen by camera E (pixel (j,i) samples the ray through (j,i), which
    is Open3D's convention: x = (j - cx) z / fx)."""
    R, t = E[:3, :3], E[:3, 3]
    jj, ii = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    rays = np.stack([(jj - cx) / fx, (ii - cy) / fy, np.ones_like(jj)], -1)         # camera frame, z = 1
    n_c = R @ np.asarray(n, np.float64)
    d_c = d0 + n_c @ t                                                             # plane in camera frame: n_c.p = d_c
    z = d_c / (rays @ n_c)
    return np.where(z > 0, z, 0).astype(np.float32)


def _sphere_frame(E, fx, fy, cx, cy, W, H, radius=0.5):
    """z-depth + colour of a sphere at the origin."""
    R, t = E[:3, :3], E[:3, 3]
    jj, ii = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    rays = np.stack([(jj - cx) / fx, (ii - cy) / fy, np.ones_like(jj)], -1)
    c = t                                                                          # sphere centre in camera frame
    a = (rays * rays).sum(-1)
    b = -2 * rays @ c
    cc = c @ c - radius ** 2
    disc = b * b - 4 * a * cc
    z = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), 0.0)
    depth = np.where(z > 0, z, 0).astype(np.float32)
    pw = (np.linalg.inv(E) @ np.concatenate([rays * z[..., None], np.ones_like(z)[..., None]], -1)[..., None])[..., :3, 0]
    col = np.clip((pw / radius * 0.5 + 0.5) * 255, 0, 255).astype(np.uint8)
    col[depth == 0] = 0
    return depth, col


CAM = dict(W=160, H=120, fx=150.0, fy=150.0, cx=79.5, cy=59.5)


# ----------------------------------------------------------------------------- oracle known answers (CPU)
def test_oracle_fronto_parallel_plane():
    W, H, fx, fy, cx, cy = (CAM[k] for k in ("W", "H", "fx", "fy", "cx", "cy"))
    E = np.eye(4)                                                   # camera at origin looking down +z
    depth = np.full((H, W), 1.0, np.float32)                         # plane z = 1
    vol = to.TsdfOracleVolume(voxel_length=0.01, sdf_trunc=0.04, with_color=False, depth_sampling_stride=4)
    touched = vol.integrate(depth, None, fx, fy, cx, cy, E, depth_trunc=3.0)
    assert len(touched) > 0
    # every touched unit is within sdf_trunc (+ one unit) of the plane
    ul = 0.16
    for (ux, uy, uz) in touched:
        assert (uz + 1) * ul >= 1.0 - 0.04 - 1e-6 and uz * ul <= 1.0 + 0.04 + 1e-6
    n_checked = 0
    for key, unit in vol.units.items():
        idx = np.arange(16)
        vx, vy, vz = np.meshgrid(idx, idx, idx, indexing="ij")
        x = (key[0] * 16 + vx + 0.5) * 0.01
        y = (key[1] * 16 + vy + 0.5) * 0.01
        z = (key[2] * 16 + vz + 0.5) * 0.01
        uf, vf = x * fx / z + cx + 0.5, y * fy / z + cy + 0.5
        u, v = np.floor(uf), np.floor(vf)
        robust = (np.abs(uf - np.round(uf)) > 1e-3) & (np.abs(vf - np.round(vf)) > 1e-3)   # pixel choice is fp32-safe
        inside = (u >= 1) & (u < W - 1) & (v >= 1) & (v < H - 1) & robust
        xn, yn = (u - cx) / fx, (v - cy) / fy
        sdf = (1.0 - z) * np.sqrt(xn * xn + yn * yn + 1)
        seen = inside & (sdf > -0.04 + 1e-6)
        hidden = inside & (sdf < -0.04 - 1e-6)
        assert np.all(unit["weight"][seen] == 1) and np.all(unit["weight"][hidden] == 0)
        np.testing.assert_allclose(unit["tsdf"][seen], np.minimum(1, sdf[seen] / 0.04), atol=2e-5)
        n_checked += int(seen.sum())
    assert n_checked > 10000


def test_oracle_tilted_plane_two_frames_zero_crossing():
    W, H, fx, fy, cx, cy = (CAM[k] for k in ("W", "H", "fx", "fy", "cx", "cy"))
    n = np.array([0.3, -0.2, 1.0])
    n /= np.linalg.norm(n)
    vol = to.TsdfOracleVolume(voxel_length=0.02, sdf_trunc=0.08, with_color=False, depth_sampling_stride=2)
    for eye in [(0.2, 0.1, -2.0), (-0.3, 0.2, -1.8)]:
        E = _look_at(eye)
        vol.integrate(_plane_depth(E, fx, fy, cx, cy, W, H, n, 0.0), None, fx, fy, cx, cy, E, depth_trunc=5.0)
    # voxels seen by both frames with |tsdf| < 0.5: sign(tsdf) = side of the plane (camera side positive)
    agree = total = 0
    for key, unit in vol.units.items():
        idx = np.arange(16)
        vx, vy, vz = np.meshgrid(idx, idx, idx, indexing="ij")
        p = np.stack([(key[0] * 16 + vx + 0.5), (key[1] * 16 + vy + 0.5), (key[2] * 16 + vz + 0.5)], -1) * 0.02
        dist = -(p @ n)                                             # cameras sit at n.p < 0
        # the depth is sampled at the nearest pixel: up to half a pixel footprint (2/150/2) x plane slope (0.4) off
        m = (unit["weight"] == 2) & (np.abs(unit["tsdf"]) < 0.9) & (np.abs(dist) > 0.01)
        agree += int((np.sign(unit["tsdf"][m]) == np.sign(dist[m])).sum())
        total += int(m.sum())
        # along-ray distance >= perpendicular distance, and within 1/cos(60 deg) of it
        m &= np.abs(dist) > 0.03
        r = unit["tsdf"][m] * 0.08 / dist[m]
        assert np.all(r > 0.85) and np.all(r < 2.0), (r.min(), r.max())
    assert total > 5000 and agree == total


def test_oracle_depth_trunc_and_holes():
    W, H, fx, fy, cx, cy = (CAM[k] for k in ("W", "H", "fx", "fy", "cx", "cy"))
    depth = np.full((H, W), 1.0, np.float32)
    depth[:, : W // 2] = 0.0                                         # left half: no measurement
    depth[: H // 2, W // 2:] = 4.0                                   # top right: beyond depth_trunc
    vol = to.TsdfOracleVolume(0.01, 0.04, with_color=False)
    vol.integrate(depth, None, fx, fy, cx, cy, np.eye(4), depth_trunc=3.0)
    for key, unit in vol.units.items():
        assert key[0] >= -1 and key[1] >= -1, key                    # only the bottom-right quadrant (x>0, y>0)
        assert key[2] * 0.16 < 1.2                                   # nothing near z = 4
    empty = to.TsdfOracleVolume(0.01, 0.04, with_color=False)
    assert empty.integrate(np.zeros((H, W), np.float32), None, fx, fy, cx, cy, np.eye(4), 3.0) == []


# ----------------------------------------------------------------------------- CUDA vs oracle (GPU)
def _compare(vol_gpu, vol_cpu, with_color):
    u = {k: v.cpu().numpy() for k, v in vol_gpu.units().items()}
    got = {tuple(int(c) for c in xyz): i for i, xyz in enumerate(u["xyz"])}
    assert set(got) == set(vol_cpu.units), (len(got), len(vol_cpu.units))
    for key, unit in vol_cpu.units.items():
        i = got[key]
        assert np.array_equal(u["weight"][i], unit["weight"]), key
        assert np.array_equal(u["tsdf"][i].view(np.uint32), unit["tsdf"].view(np.uint32)), key
        if with_color:
            assert np.array_equal(u["rgb"][i].view(np.uint32), unit["rgb"].view(np.uint32)), key


@pytest.mark.gpu
@pytest.mark.parametrize("stride,with_color", [(4, True), (1, False), (3, True)])
def test_tsdf_integrate_matches_oracle_bit_exact(cuda_dev, stride, with_color):
    from radegs_b200 import tsdf
    W, H, fx, fy, cx, cy = (CAM[k] for k in ("W", "H", "fx", "fy", "cx", "cy"))
    vol = tsdf.ScalableTSDFVolume(0.02, 0.08, tsdf.TSDFVolumeColorType.RGB8 if with_color
                                  else tsdf.TSDFVolumeColorType.NoColor, depth_sampling_stride=stride,
                                  max_units=4096, device=cuda_dev)
    ref = to.TsdfOracleVolume(0.02, 0.08, with_color=with_color, depth_sampling_stride=stride)
    intr = tsdf.PinholeCameraIntrinsic(W, H, fx, fy, cx, cy)
    eyes = [(1.6, 0.2, 0.3), (-0.4, 1.5, 0.5), (0.1, -0.3, 1.7), (-1.2, -1.0, -0.6)]
    for f, eye in enumerate(eyes):
        E = _look_at(eye)
        depth, col = _sphere_frame(E, fx, fy, cx, cy, W, H)
        ref.integrate(depth, col if with_color else None, fx, fy, cx, cy, E, depth_trunc=3.0)
        vol.integrate(torch.from_numpy(depth).to(cuda_dev)[..., None],
                      torch.from_numpy(col).to(cuda_dev) if with_color else None, intr, E, depth_trunc=3.0)
        # the frame's touched list holds every unit the oracle touched, once
        c = vol.counters.cpu()
        assert int(c[1]) == len(ref.last_touched) and int(c[2]) == 0
        slots = vol.vals[vol.touched[: int(c[1])].long()].long()
        xyz = vol.unit_xyz[slots].cpu().numpy()
        assert sorted(map(tuple, xyz.tolist())) == ref.last_touched
    _compare(vol, ref, with_color)
    pc = vol.extract_voxel_point_cloud()
    r = pc["points"].norm(dim=1)
    assert pc["points"].shape[0] > 1000 and float((r - 0.5).abs().max()) < 0.08 + 0.02 * 1.8


@pytest.mark.gpu
def test_tsdf_float_colour_and_depth_trunc(cuda_dev):
    from radegs_b200 import tsdf
    W, H, fx, fy, cx, cy = (CAM[k] for k in ("W", "H", "fx", "fy", "cx", "cy"))
    E = _look_at((0.0, 0.0, -2.0))
    depth, col = _sphere_frame(E, fx, fy, cx, cy, W, H)
    depth[depth > 1.7] = 5.0                                          # beyond depth_trunc: dropped
    vol = tsdf.ScalableTSDFVolume(0.02, 0.08, max_units=2048, device=cuda_dev)
    ref = to.TsdfOracleVolume(0.02, 0.08)
    colf = col.astype(np.float32) * np.float32(0.5)
    ref.integrate(depth, colf, fx, fy, cx, cy, E, depth_trunc=3.0)
    vol.integrate(torch.from_numpy(depth).to(cuda_dev), torch.from_numpy(colf).to(cuda_dev),
                  tsdf.PinholeCameraIntrinsic(W, H, fx, fy, cx, cy), torch.from_numpy(E), depth_trunc=3.0)
    _compare(vol, ref, True)


@pytest.mark.gpu
def test_tsdf_overflow_and_bad_arguments_are_loud(cuda_dev):
    from radegs_b200 import tsdf
    W, H, fx, fy, cx, cy = (CAM[k] for k in ("W", "H", "fx", "fy", "cx", "cy"))
    E = _look_at((0.0, 0.0, -2.0))
    depth, col = _sphere_frame(E, fx, fy, cx, cy, W, H)
    intr = tsdf.PinholeCameraIntrinsic(W, H, fx, fy, cx, cy)
    small = tsdf.ScalableTSDFVolume(0.02, 0.08, max_units=8, device=cuda_dev)
    small.integrate(torch.from_numpy(depth).to(cuda_dev), torch.from_numpy(col).to(cuda_dev), intr, E)
    with pytest.raises(RuntimeError, match="overflow"):
        small.n_units()
    vol = tsdf.ScalableTSDFVolume(0.02, 0.08, max_units=64, device=cuda_dev)
    with pytest.raises(RuntimeError, match="no CPU path"):
        vol.integrate(torch.from_numpy(depth), torch.from_numpy(col), intr, E)
    with pytest.raises(ValueError):
        vol.integrate(torch.from_numpy(depth).to(cuda_dev), None, intr, E)
    with pytest.raises(ValueError):
        vol.integrate(torch.from_numpy(depth).to(cuda_dev)[:, :-1], torch.from_numpy(col).to(cuda_dev), intr, E)
    with pytest.raises(NotImplementedError):
        tsdf.ScalableTSDFVolume(0.02, 0.08, volume_unit_resolution=8, device=cuda_dev)


def test_tsdf_volume_refuses_cpu_device():
    from radegs_b200 import tsdf
    with pytest.raises(RuntimeError, match="no CPU path"):
        tsdf.ScalableTSDFVolume(0.02, 0.08, device="cpu")


@pytest.mark.gpu
def test_fuse_render_sweep_equals_per_frame_host_loop(cuda_dev):
    """The device sweep (render -> rgb/depth outputs -> integrate, nothing leaves HBM) fuses the same volume as the
    reference's loop shape (mesh.py:1571-1632: render, copy the frame to the host, integrate on the CPU), the CPU
    side played by the oracle on the downloaded frames."""
    from gsplat.rendering import rasterization
    from radegs_b200 import scenes, tsdf
    from tests.util import small_scene
    cfg, gs, vm, Ks = small_scene(n=4000, w=160, h=96, views=3, sh_degree=3, seed=5)
    params = [t.to(cuda_dev) for t in scenes.activate(gs, 3)]
    vmd, Kd = vm.to(cuda_dev), Ks.to(cuda_dev)
    vol = tsdf.ScalableTSDFVolume(0.02, 0.06, max_units=8192, device=cuda_dev)
    assert tsdf.fuse_render_sweep(vol, params, vmd, Kd, cfg.width, cfg.height, sh_degree=3, depth_trunc=6.0) == 3
    ref = to.TsdfOracleVolume(0.02, 0.06)
    with torch.no_grad():
        for v in range(3):
            render, alpha, exp_d, _, _, _ = rasterization(*params, vmd[v:v + 1], Kd[v:v + 1], cfg.width, cfg.height,
                                                          packed=False, sh_degree=3, render_mode="RGB+ED",
                                                          rasterize_mode="antialiased", return_depth_normal=True)
            rgb = torch.clamp(render[0, ..., :3], 0.0, 1.0)
            d = torch.where(alpha[0] > 0, exp_d[0], exp_d[0].max())
            K = Ks[v].double().numpy()
            ref.integrate(d.squeeze(-1).cpu().numpy(), np.asarray(rgb.cpu().numpy() * 255, order="C", dtype=np.uint8),
                          K[0, 0], K[1, 1], K[0, 2], K[1, 2], vm[v].double().numpy(), depth_trunc=6.0)
    assert len(ref.units) > 20
    _compare(vol, ref, True)
    # several cameras per rasterization() call: same frames, same volume
    vol3 = tsdf.ScalableTSDFVolume(0.02, 0.06, max_units=8192, device=cuda_dev)
    tsdf.fuse_render_sweep(vol3, params, vmd, Kd, cfg.width, cfg.height, sh_degree=3, depth_trunc=6.0, views_per_launch=2)
    _compare(vol3, ref, True)
