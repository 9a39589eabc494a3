"""Generates tests/golden/intree_glue.npz by running the REFERENCE's own in-tree glue on seeded inputs:

* ``depth_double_to_normal`` -- collab_splats/utils/camera_utils.py:176-279 (the stencil of the depth-normal
  consistency loss), with the module's ``nerfstudio`` import replaced by a minimal ``Cameras``
  stand-in (camera_to_worlds / width / height / get_intrinsics_matrices -- all the function reads) and
  ``Tensor.cuda()`` made a no-op (this container has no GPU; the arithmetic is unchanged);
* ``project_gaussians``      -- collab_splats/utils/utils.py:13-40.

Only runs in the build container (the GPU box has no /root/reference).
Re-generate with:  python tests/golden/make_intree_golden.py
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/collab_splats/utils")
OUT = Path(__file__).resolve().parent / "intree_glue.npz"


class Cameras:                                       # stand-in for nerfstudio.cameras.cameras.Cameras
    def __init__(self, c2w, K, W, H):
        self.camera_to_worlds, self._K = c2w, K
        self.width, self.height = torch.tensor([[W]]), torch.tensor([[H]])

    def get_intrinsics_matrices(self):
        return self._K


for name in ("nerfstudio", "nerfstudio.cameras", "nerfstudio.cameras.cameras"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["nerfstudio.cameras.cameras"].Cameras = Cameras
torch.Tensor.cuda = lambda self, *a, **k: self


def load(fname):
    spec = importlib.util.spec_from_file_location("ref_" + fname, REF / (fname + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


cu, uu = load("camera_utils"), load("utils")
g = torch.Generator().manual_seed(77)
W, H = 96, 64
fx, fy = 0.9 * W, 1.1 * W
K = torch.tensor([[[fx, 0.0, W / 2.0], [0.0, fy, H / 2.0], [0.0, 0.0, 1.0]]])
c2w = torch.eye(4)[None, :3, :]
cam = Cameras(c2w, K, W, H)
yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
d1 = 2.0 + 0.01 * xx - 0.02 * yy + 0.05 * torch.rand(H, W, generator=g)
d2 = 2.5 + 0.3 * torch.sin(xx / 9.0) * torch.cos(yy / 7.0) + 0.02 * torch.rand(H, W, generator=g)
normals = cu.depth_double_to_normal(cam, d1[None, ..., None], d2[None, ..., None])        # [2,H,W,3]

N = 500
meta = {"width": W, "height": H,
        "radii": (torch.rand(1, N, 2, generator=g) * 4).to(torch.int32),
        "means2d": torch.rand(1, N, 2, generator=g) * torch.tensor([W + 20.0, H + 20.0]) - 10.0,
        "depths": torch.rand(1, N, generator=g) * 5}
meta["means2d"][0, :8] = torch.tensor([[0.5, 1.5], [2.5, 3.5], [-0.5, 0.49], [W - 0.5, H - 0.5], [W, H], [1e9, -1e9],
                                       [10.5, 11.5], [12.5, 0.0]])
pg = uu.project_gaussians(meta)

np.savez_compressed(OUT, K=K.numpy(), W=W, H=H, d1=d1.numpy(), d2=d2.numpy(), normals=normals.numpy(),
                    radii=meta["radii"].numpy(), means2d=meta["means2d"].numpy(), depths=meta["depths"].numpy(),
                    **{"pg_" + k: v.numpy() for k, v in pg.items()})
print("wrote", OUT, OUT.stat().st_size, "bytes")
