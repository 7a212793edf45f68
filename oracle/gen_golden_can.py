#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/cluster_can_a.npz: the reference's OWN losses.py (unchanged, CPU, stubs and
k-means stand-in of gen_golden_loss.py) with the optional canonical-axis terms switched on (loss_norm_D_C_can_dot_w /
_can_L1_w > 0, losses.py:480-502) on an axis-aligned synthetic room, where the cluster means do fall within 3*tres of the axes.

Run in the build container only (needs /root/reference):  python oracle/gen_golden_can.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import gen_golden_loss as ggl  # noqa: E402
from oracle import cluster_loss as cl  # noqa: E402

if __name__ == "__main__":
    ggl._install_stubs()
    sys.path.insert(0, ggl.REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "normal-clustering-nerf_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
    import losses as ref_losses          # the reference file, unmodified
    hp = dict(loss_opacity_w=1e-3, loss_norm_can_tres=0.01, loss_norm_D_C_ort_dot_w=2e-3, loss_norm_D_C_centr_dot_w=2e-3,
              loss_norm_D_C_centr_L1_w=2e-3, loss_norm_D_C_can_dot_w=3e-3, loss_norm_D_C_can_L1_w=1e-3, loss_norm_can_start=500,
              loss_norm_can_grow=2500, loss_norm_can_end=-1, ray_sampling_strategy="all_images_triang_patch", random_tr_poses=False,
              pred_norm_nn=False, pred_norm_depth=True)
    n_rays, seed = 1024, 5
    torch.manual_seed(seed)
    b = synth.patch_batch(n_rays, seed=seed)
    rays_d = torch.from_numpy(b["rays_d"]); rays_o = torch.from_numpy(b["rays_o"])
    t_wall = torch.where(rays_d > 0, (0.4 - rays_o) / rays_d, (-0.4 - rays_o) / rays_d).min(-1)[0]
    depth = (t_wall + 0.0005 * torch.randn(n_rays)).clamp_min(0.02).requires_grad_(True)
    opacity = torch.rand(n_rays).clamp(0.05, 0.99)
    rgb = torch.rand(n_rays, 3, requires_grad=True)
    tri = b["tri"][:, :49] % 64
    # NOTE rays_o := rays_d is what rendering.py:227 hands the loss; with it the "points" are rays_d * (1 + depth), so the room's walls
    # are no longer planes - the canonical terms still see near-axis clusters on this batch (printed below)
    pred = {"rgb": rgb, "depth": depth, "opacity": opacity, "rays_o": rays_d, "rays_d": rays_d, "deltas": torch.zeros(1), "ts": torch.zeros(1),
            "rays_a": torch.zeros(1, 3, dtype=torch.int64)}
    target = {"rgb": torch.rand(n_rays, 3), "patch_area": 64, "x1_offsets_local": torch.from_numpy(tri[0]),
              "x2_offsets_local": torch.from_numpy(tri[1]), "x3_offsets_local": torch.from_numpy(tri[2])}
    loss_fn = ref_losses.NeRFMTLoss(hp)
    cap = {}
    orig = ref_losses._normals_clustering

    def wrapped(*a, **k):
        r = orig(*a, **k)
        cap["labels"] = r[0].numpy().copy()
        return r

    ref_losses._normals_clustering = wrapped
    loss_d = loss_fn(pred, target, global_step=3000)
    ref_losses._normals_clustering = orig
    loss_d["total"].backward()
    normals = ref_losses._extract_normals_from_ray_batch(rays_d, rays_d, depth.detach(),
                                                         {k: torch.from_numpy(b["tri"][i]) for i, k in enumerate(("x1", "x2", "x3"))})
    valid = cl.valid_rows(normals).numpy()
    print({k: float(v) for k, v in loss_d.items()})
    assert "norm_D_C_can_dot" in loss_d, "no cluster mean fell within 3*tres of a canonical axis - pick another batch"
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cluster_can_a.npz"), rays_d=b["rays_d"], depth=depth.detach().numpy(),
                        tri=b["tri"], valid=valid, labels=cap["labels"], step=3000, tres=hp["loss_norm_can_tres"],
                        w_can_dot=loss_fn.w_sched(hp["loss_norm_D_C_can_dot_w"], 3000), w_can_l1=loss_fn.w_sched(hp["loss_norm_D_C_can_L1_w"], 3000),
                        w_clu=loss_fn.w_sched(2e-3, 3000), loss_can_dot=float(loss_d["norm_D_C_can_dot"]), loss_can_l1=float(loss_d["norm_D_C_can_L1"]),
                        loss_ort=float(loss_d["norm_D_C_ort_dot"]), loss_dot=float(loss_d["norm_D_C_centr_dot"]), loss_l1=float(loss_d["norm_D_C_centr_L1"]),
                        grad_depth=depth.grad.numpy())
