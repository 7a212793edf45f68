#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/batches_a.npz by running the reference's OWN datasets/base.py (imported
unchanged from /root/reference, by file path so that datasets/__init__.py and its HDF5 / PNG loaders are not pulled in):

  * BaseDataset.__getitem__ (index half, base.py:94-173) for every ray-sampling strategy, with and without --random_tr_poses and
    --triang_max_expand, on a seeded numpy stream  -> img_idxs / pix_idxs / rnd_img_idxs
  * generate_random_poses (base.py:235-263), with and without the focus-point jitter, on a seeded numpy stream

Pins ncn_b200.batches.{sample_batch_indices, generate_random_poses} (CPU test) and the index arithmetic of the device sampler.

Run in the build container only (needs /root/reference):  python oracle/gen_golden_batches.py"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("NCN_REFERENCE_ROOT", "/root/reference")

if __name__ == "__main__":
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_base", os.path.join(REF, "datasets", "base.py"))
    base = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(base)                                  # the reference file, unmodified
    H, W, P, Q, B, PATCH = 24, 32, 5, 37, 768, 8
    out = dict(H=H, W=W, P=P, Q=Q, B=B, PATCH=PATCH)
    q, _ = torch.linalg.qr(torch.randn(P, 3, 3, generator=torch.Generator().manual_seed(0)))
    poses = torch.cat([q, 0.2 * torch.randn(P, 3, 1, generator=torch.Generator().manual_seed(1))], -1)      # (P,3,4) float32
    n_case = 0
    for strategy in ("all_images", "same_image", "all_images_triang", "same_image_triang", "all_images_triang_patch",
                     "same_image_triang_patch"):
        for rtp in (False, True):
            if rtp and "triang" not in strategy:
                continue                                            # asserted by the reference (base.py:98-102)
            for expand in ((0, 3) if strategy.endswith("triang") else (0,)):
                ds = base.BaseDataset(root_dir=None, split="train")
                ds.ray_sampling_strategy = strategy
                ds.random_tr_poses = rtp
                ds.batch_size = B
                ds.img_wh = (W, H)
                ds.poses = poses
                ds.rays = torch.zeros(P, H * W, 3)
                ds.random_poses = torch.zeros(Q, 3, 4)
                if strategy.endswith("_patch"):
                    ds._triang_patche_images_metadata(H, W, PATCH)
                elif strategy.endswith("triang"):
                    ds._triang_images_metadata(H, W, expand)
                seed = 100 + n_case
                np.random.seed(seed)
                s = ds[0]
                key = f"case{n_case}"
                out[key + "_meta"] = np.array([strategy, str(int(rtp)), str(expand), str(seed)])
                out[key + "_img"] = np.asarray(s["img_idxs"]).astype(np.int64).reshape(-1)
                out[key + "_pix"] = np.asarray(s["pix_idxs"]).astype(np.int64).reshape(-1)
                if rtp:
                    out[key + "_rnd"] = np.asarray(s["rnd_img_idxs"]).astype(np.int64).reshape(-1)
                print(key, strategy, "rtp" if rtp else "", expand, out[key + "_pix"].shape)
                n_case += 1
    out["n_cases"] = n_case
    # generated poses
    lo = poses[:, :3, 3].min(0)[0]; hi = poses[:, :3, 3].max(0)[0]
    for name, jitter in (("plain", False), ("jitter", True)):
        np.random.seed(7)
        rp, avg = base.generate_random_poses(poses, lo, hi, 64, random_pose_focusptjitter=jitter)
        out["rp_" + name] = rp.numpy()
        out["rp_avg"] = np.asarray(avg)
    out["poses"] = poses.numpy(); out["xyz_min"] = lo.numpy(); out["xyz_max"] = hi.numpy()
    out["focus_pt"] = np.asarray(base.focus_pt_fn(poses.numpy()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "batches_a.npz"), **out)
    print("batches_a:", n_case, "index cases; random poses", out["rp_plain"].shape, out["rp_plain"].dtype)
