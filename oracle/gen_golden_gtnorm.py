#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/gt_normals_a.npz: the reference's OWN losses.py (unchanged, CPU, stubs of
gen_golden_loss.py) with the optional GT-normal supervision of the depth-derived normals switched on (loss_norm_depth_L1_w /
loss_norm_depth_dot_w > 0, load_norm_gt; losses.py:387-409) on a triangle batch of the synthetic room; some GT normals are the
(0,0,0) "no label" code.

Run in the build container only (needs /root/reference):  python oracle/gen_golden_gtnorm.py"""
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
    import importlib.util
    spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "normal-clustering-nerf_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
    import losses as ref_losses          # the reference file, unmodified
    hp = dict(loss_opacity_w=1e-3, loss_norm_depth_L1_w=0.02, loss_norm_depth_dot_w=0.03, load_norm_gt=True, ray_sampling_strategy="all_images_triang",
              random_tr_poses=False, pred_norm_nn=False, pred_norm_depth=True)
    n_rays, seed = 3 * 256, 9
    torch.manual_seed(seed)
    b = synth.random_batch(n_rays, seed=seed)
    rays_d = torch.from_numpy(b["rays_d"]); rays_o = torch.from_numpy(b["rays_o"])
    t_wall = torch.where(rays_d > 0, (0.4 - rays_o) / rays_d, (-0.4 - rays_o) / rays_d).min(-1)[0]
    depth = (t_wall + 0.002 * torch.randn(n_rays)).clamp_min(0.02).requires_grad_(True)
    normals_gt = torch.nn.functional.normalize(torch.randn(n_rays, 3), dim=-1)
    normals_gt[::5] = 0.0                                      # no label
    pred = {"rgb": torch.rand(n_rays, 3, requires_grad=True), "depth": depth, "opacity": torch.rand(n_rays).clamp(0.05, 0.99), "rays_o": rays_d,
            "rays_d": rays_d, "deltas": torch.zeros(1), "ts": torch.zeros(1), "rays_a": torch.zeros(1, 3, dtype=torch.int64)}
    target = {"rgb": torch.rand(n_rays, 3), "normals": normals_gt}
    loss_d = ref_losses.NeRFMTLoss(hp)(pred, target, global_step=3000)
    loss_d["total"].backward()
    print({k: float(v) for k, v in loss_d.items()})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gt_normals_a.npz"), rays_d=b["rays_d"], depth=depth.detach().numpy(),
                        normals_gt=normals_gt.numpy(), w_l1=hp["loss_norm_depth_L1_w"], w_dot=hp["loss_norm_depth_dot_w"],
                        loss_l1=float(loss_d["norm_D_L1"]), loss_dot=float(loss_d["norm_D_dot"]), grad_depth=depth.grad.numpy())
