#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/sem_loss_a.npz by running the reference's OWN losses.py::NeRFMTLoss
(imported unchanged from /root/reference, CPU tensors, stubs of gen_golden_loss.py) with the semantic head enabled
(--pred_sem, loss_sem_w > 0, load_sem_gt): the 'sem' term (losses.py:240-242, 569-573), the rgb and opacity terms next to it,
and the gradient that reaches the rendered logits.  Two cases: NYU40-style labels (0 = void) and an all-void batch (the
reference drops the NaN term).

Run in the build container only (needs /root/reference):  python oracle/gen_golden_sem.py"""
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
    import contextlib
    import io
    import losses as ref_losses          # the reference file, unmodified
    hp = dict(loss_opacity_w=1e-3, loss_sem_w=4e-2, pred_sem=True, pred_norm_nn=True, pred_norm_depth=False, load_sem_gt=True,
              load_sem_WF_gt=False, ray_sampling_strategy="all_images", random_tr_poses=False)
    out = {}
    for case, (R, n_cls, all_void) in {"a": (1024, 13, False), "b": (1024, 3, False), "void": (256, 5, True)}.items():
        torch.manual_seed(7 + R + n_cls)
        sem = (torch.randn(R, n_cls) * 2.0).requires_grad_(True)
        rgb = torch.rand(R, 3, requires_grad=True)
        opacity = torch.rand(R).clamp(0.05, 0.99)
        labels = torch.zeros(R, dtype=torch.int64) if all_void else torch.randint(0, n_cls + 1, (R,))
        pred = {"rgb": rgb, "depth": torch.rand(R), "opacity": opacity, "sem": sem, "norm_nn": torch.randn(R, 3), "rays_o": torch.randn(R, 3),
                "rays_d": torch.randn(R, 3), "deltas": torch.zeros(1), "ts": torch.zeros(1), "rays_a": torch.zeros(1, 3, dtype=torch.int64)}
        target = {"rgb": torch.rand(R, 3), "semantics": labels}
        loss_fn = ref_losses.NeRFMTLoss(hp)
        with contextlib.redirect_stdout(io.StringIO()):      # the reference prints when it drops a NaN term
            loss_d = loss_fn(pred, target, global_step=3000)
        loss_d["total"].backward()
        out.update({f"{case}_sem": sem.detach().numpy(), f"{case}_labels": labels.numpy(), f"{case}_rgb": rgb.detach().numpy(),
                    f"{case}_opacity": opacity.numpy(), f"{case}_target_rgb": target["rgb"].numpy(),
                    f"{case}_loss_sem": float(loss_d["sem"]), f"{case}_loss_rgb": float(loss_d["rgb"]), f"{case}_loss_opacity": float(loss_d["opacity"]),
                    f"{case}_loss_total": float(loss_d["total"]),
                    f"{case}_grad_sem": (sem.grad if sem.grad is not None else torch.zeros_like(sem)).numpy(), f"{case}_grad_rgb": rgb.grad.numpy()})
        print(case, {k: float(v) for k, v in loss_d.items()}, "labelled", int((labels > 0).sum()))
    # depth supervision + RegNeRF-style depth smoothness (losses.py:371-385, 411-417) on a triangle batch: off in the shipped
    # experiments, carried by the module path
    hp2 = dict(loss_opacity_w=1e-3, loss_depth_w=0.1, loss_reg_depth_w=0.05, loss_norm_can_start=500, pred_sem=False, pred_norm_nn=False,
               pred_norm_depth=False, ray_sampling_strategy="all_images_triang", random_tr_poses=False)
    torch.manual_seed(3)
    R = 3 * 300
    depth = (1.0 + torch.rand(R)).requires_grad_(True)
    rgb = torch.rand(R, 3, requires_grad=True)
    opacity = torch.rand(R).clamp(0.05, 0.99)
    d_gt = 1.0 + torch.rand(R); d_gt[::7] = 0.0                       # 0 = no label
    pred = {"rgb": rgb, "depth": depth, "opacity": opacity, "rays_o": torch.randn(R, 3), "rays_d": torch.randn(R, 3), "deltas": torch.zeros(1),
            "ts": torch.zeros(1), "rays_a": torch.zeros(1, 3, dtype=torch.int64)}
    target = {"rgb": torch.rand(R, 3), "depth": d_gt}
    loss_d = ref_losses.NeRFMTLoss(hp2)(pred, target, global_step=3000)
    loss_d["total"].backward()
    out.update(depth_depth=depth.detach().numpy(), depth_rgb=rgb.detach().numpy(), depth_opacity=opacity.numpy(), depth_target_rgb=target["rgb"].numpy(),
               depth_target_depth=d_gt.numpy(), depth_loss_depth=float(loss_d["depth"]), depth_loss_reg=float(loss_d["reg_depth"]),
               depth_loss_total=float(loss_d["total"]), depth_grad_depth=depth.grad.numpy(), depth_w=hp2["loss_depth_w"], reg_w=hp2["loss_reg_depth_w"])
    print("depth", {k: float(v) for k, v in loss_d.items()})
    out["sem_w"] = hp["loss_sem_w"]; out["opacity_w"] = hp["loss_opacity_w"]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sem_loss_a.npz"), **out)
