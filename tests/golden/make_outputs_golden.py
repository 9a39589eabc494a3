"""Generates tests/golden/rade_outputs.npz by EXECUTING the reference's own post-render glue, the part of
``RadegsModel.get_outputs`` that follows the rasterization call (collab_splats/models/rade_gs_model.py:200-271,
SURVEY.md row a14): the source lines are read from /root/reference at generation time, wrapped in a function and run on
seeded inputs with stand-ins for what they touch of nerfstudio (``self.config``, ``self.step``,
``self._get_background_color()``, ``camera.rescale_output_resolution``); ``depth_double_to_normal`` is the reference's
own (collab_splats/utils/camera_utils.py:176-279), loaded as in make_intree_golden.py.  Nothing is copied into the repo.

Only runs in the build container (the GPU box has no /root/reference).
Re-generate with:  python tests/golden/make_outputs_golden.py
"""
import importlib.util
import sys
import textwrap
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/collab_splats")
OUT = Path(__file__).resolve().parent / "rade_outputs.npz"


class Cameras:                                       # stand-in for nerfstudio.cameras.cameras.Cameras
    def __init__(self, c2w, K, W, H):
        self.camera_to_worlds, self._K = c2w, K
        self.width, self.height = torch.tensor([[W]]), torch.tensor([[H]])
        self.metadata = None

    def get_intrinsics_matrices(self):
        return self._K

    def rescale_output_resolution(self, *_):
        pass


for name in ("nerfstudio", "nerfstudio.cameras", "nerfstudio.cameras.cameras"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["nerfstudio.cameras.cameras"].Cameras = Cameras
torch.Tensor.cuda = lambda self, *a, **k: self
spec = importlib.util.spec_from_file_location("ref_camera_utils", REF / "utils" / "camera_utils.py")
cu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cu)

# ---- the reference's lines, from the depth-normal block to the returned dict
lines = (REF / "models" / "rade_gs_model.py").read_text().splitlines()
start = next(i for i, l in enumerate(lines) if "# Calculate depth_middepth_normal" in l)
end = next(i for i in range(start, len(lines)) if '"background": background,' in lines[i]) + 1   # + the closing brace
body = textwrap.dedent("\n".join(lines[start:end + 1]))
src = ("def glue(self, camera, render, alpha, expected_depths, median_depths, expected_normals, render_mode, "
       "camera_scale_fac, H, W):\n" + textwrap.indent(body, "    "))
ns = {"torch": torch, "depth_double_to_normal": cu.depth_double_to_normal}
exec(compile(src, "rade_gs_model.py[get_outputs glue]", "exec"), ns)


class Self:
    training = True
    step = 100

    class config:
        use_depth_normal_loss = True
        regularization_from_iter = 0
        use_bilateral_grid = False

    def __init__(self, bg):
        self._bg = bg

    def _get_background_color(self):
        return self._bg


g = torch.Generator().manual_seed(99)
W, H = 80, 56
fx, fy = 0.9 * W, 1.05 * W
K = torch.tensor([[[fx, 0.0, W / 2.0], [0.0, fy, H / 2.0], [0.0, 0.0, 1.0]]])
cam = Cameras(torch.eye(4)[None, :3, :], K, W, H)
yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
alpha = torch.rand(1, H, W, 1, generator=g)
alpha[:, :9, :17] = 0.0                                            # pixels no Gaussian reached: the masked fills
alpha[:, 30:34, 40:60] = 0.0
render = torch.cat([torch.rand(1, H, W, 3, generator=g) * 1.5 - 0.25,      # some values leave [0,1]: clamp
                    (3.0 + torch.rand(1, H, W, 1, generator=g))], dim=-1)
exp_d = (2.0 + 0.01 * xx - 0.015 * yy + 0.05 * torch.rand(H, W, generator=g))[None, ..., None]
med_d = (2.3 + 0.2 * torch.sin(xx / 8.0) * torch.cos(yy / 6.0) + 0.02 * torch.rand(H, W, generator=g))[None, ..., None]
nrm = torch.nn.functional.normalize(torch.randn(1, H, W, 3, generator=g), dim=-1) * 0.9
bg = torch.tensor([0.1, 0.6, 0.3])
out = {}
for mode in ("RGB+ED", "RGB"):
    o = ns["glue"](Self(bg), cam, render if mode == "RGB+ED" else render[..., :3], alpha, exp_d, med_d, nrm[0], mode,
                   1.0, H, W)
    for k, v in o.items():
        if v is not None:
            out[f"{mode}:{k}"] = v.detach().numpy()

np.savez_compressed(OUT, K=K.numpy(), W=W, H=H, alpha=alpha.numpy(), render=render.numpy(), exp_d=exp_d.numpy(),
                    med_d=med_d.numpy(), normals=nrm.numpy(), background=bg.numpy(),
                    **{k.replace(":", "__").replace("+", "p"): v for k, v in out.items()})
print("wrote", OUT, OUT.stat().st_size, "bytes;", sorted(out))
