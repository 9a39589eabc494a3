"""Device-side versions of the small `meta` consumers of the reference."""

from __future__ import annotations

from typing import Dict

import torch


def project_gaussians(meta: dict, to_cpu: bool = False) -> Dict[str, torch.Tensor]:
    """collab_splats/utils/utils.py:13-40 (``project_gaussians``) without the four device->host copies: the flat
    pixel index of every projected centre, its depth, the "radius > 1 px" visibility mask and the ids of the
    visible Gaussians, for the single camera of ``meta`` (``rasterization()``'s meta dict, C == 1).

    Same keys and values as the reference; tensors stay on the device unless ``to_cpu`` (the reference's
    behaviour) is requested.  One kernel (csrc/stats.cu) replaces round / long / clamp x2 / mul / add / compare /
    sum / compare."""
    from radegs_b200 import backend as be
    lib = be.load()
    W, H = int(meta["width"]), int(meta["height"])
    radii, means2d = meta["radii"], meta["means2d"]
    if not means2d.is_cuda:
        raise RuntimeError("project_gaussians: meta tensors must be CUDA tensors (there is no CPU path)")
    if radii.dim() == 3:
        assert radii.shape[0] == 1, "project_gaussians handles one camera (the reference squeezes the camera axis)"
        radii, means2d = radii[0], means2d[0]
    N = radii.shape[0]
    dev = means2d.device
    r32 = radii.detach().to(torch.int32).contiguous()
    m2 = means2d.detach().to(torch.float32).contiguous()
    flat = torch.empty(N, device=dev, dtype=torch.int64)
    valid = torch.empty(N, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        be.check(lib.rs_project_lookup(be.ptr(m2), be.ptr(r32), N, W, H, be.ptr(flat), be.ptr(valid),
                                       be.stream_ptr(dev)), "rs_project_lookup")
    valid = valid.bool()
    out = {"proj_flattened": flat, "proj_depths": meta["depths"].squeeze().detach(), "valid_mask": valid,
           "gaussian_ids": valid.nonzero(as_tuple=False).squeeze()}
    return {k: v.cpu() for k, v in out.items()} if to_cpu else out
