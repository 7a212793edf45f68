#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/get_rays_a.npz by running the reference's OWN datasets/ray_utils.py
(get_ray_directions + get_rays, imported unchanged from /root/reference; `kornia.create_meshgrid` - absent here - is stubbed with
its documented pixel-grid semantics) the way NeRFSystem.forward uses them (train_nerf.py:167-182: directions[pix_idxs],
poses[img_idxs]).

Run in the build container only (needs /root/reference):  python oracle/gen_golden_rays.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("NCN_REFERENCE_ROOT", "/root/reference")

if __name__ == "__main__":
    kornia = types.ModuleType("kornia")

    def create_meshgrid(height, width, normalized_coordinates=True, device="cpu"):
        assert not normalized_coordinates
        ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32, device=device),
                                torch.arange(width, dtype=torch.float32, device=device), indexing="ij")
        return torch.stack([xs, ys], -1)[None]                    # (1, H, W, 2) = (x, y), kornia's layout

    kornia.create_meshgrid = create_meshgrid
    sys.modules["kornia"] = kornia
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_ray_utils", os.path.join(REF, "datasets", "ray_utils.py"))
    ru = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ru)                                    # the reference file, unmodified
    torch.manual_seed(0)
    H, W, P, n = 48, 64, 6, 777
    K = torch.tensor([[55.4, 0.0, 32.0], [0.0, 55.4, 24.0], [0.0, 0.0, 1.0]])
    directions = ru.get_ray_directions(H, W, K)                    # (H*W, 3)
    directions = directions / torch.norm(directions, dim=-1, keepdim=True)      # unit-norm dirs (cam_model.py:192-194)
    q, _ = torch.linalg.qr(torch.randn(P, 3, 3))
    poses = torch.cat([q, torch.randn(P, 3, 1)], -1)               # (P, 3, 4)
    img = torch.randint(0, P, (n,)); pix = torch.randint(0, H * W, (n,))
    rays_o, rays_d = ru.get_rays(directions[pix], poses[img])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "get_rays_a.npz"), directions=directions.numpy(), poses=poses.numpy(),
                        img_idx=img.numpy(), pix_idx=pix.numpy(), rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), K=K.numpy(), H=H, W=W)
    print("get_rays_a", tuple(rays_o.shape), tuple(rays_d.shape), float(rays_d.norm(dim=-1).mean()))
