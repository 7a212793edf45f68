#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/normals_image_a.npz by running the reference's OWN
datasets/hypersim_src/utils.py::_extract_normals_from_depth_batch (imported unchanged from /root/reference, CPU tensors; the
module's unavailable imports h5py / imgviz / pytorch3d are stubbed exactly as in gen_golden_loss.py).

Run in the build container only (needs /root/reference):  python oracle/gen_golden_normals_image.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import gen_golden_loss as ggl  # noqa: E402

if __name__ == "__main__":
    ggl._install_stubs()
    sys.path.insert(0, ggl.REF)
    from datasets.hypersim_src.utils import _extract_normals_from_depth_batch as ref_fn      # the reference function, unmodified
    torch.manual_seed(0)
    B, H, W = 3, 24, 32
    fx = 30.0
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    dirs = torch.stack([(xs - W / 2 + 0.5) / fx, (ys - H / 2 + 0.5) / fx, torch.ones_like(xs)], -1).reshape(-1, 3)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)                      # unit-norm camera directions (cam_model.py:192-194)
    # depth of a tilted plane + a step edge + noise, with the invalid codes the reference handles
    depth = 1.5 + 0.01 * xs[None] - 0.02 * ys[None] + 0.003 * torch.randn(B, H, W)
    depth[:, :, W // 2:] += 0.4
    depth[0, 5, 7] = 0.0; depth[1, 10, 3] = float("nan"); depth[2, 12, 20] = float("inf"); depth[0, 0, 0] = 0.0
    depth[1, 6, 9] = 0.0; depth[1, 6, 10] = 1.0      # a valid pixel whose neighbour is invalid (the reference does not mask it)
    q, _ = torch.linalg.qr(torch.randn(B, 3, 3))
    poses = torch.eye(4).repeat(B, 1, 1)
    poses[:, :3, :3] = q
    poses[:, :3, 3] = torch.randn(B, 3)
    out = ref_fn(depth, dirs, poses)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "normals_image_a.npz"), depth=depth.numpy(), dirs=dirs.numpy(),
                        poses=poses.numpy(), normals=out.numpy())
    print("normals_image_a", tuple(out.shape), "zero rows", int((out.abs().sum(-1) == 0).sum()), "nan rows", int(torch.isnan(out).any(-1).sum()))
