#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/cluster_loss_*.npz by running the reference's OWN
losses.py (imported unchanged from /root/reference) on CPU tensors.

The reference file imports vren / faiss / h5py / imgviz / pytorch3d, none of which exist here; they are stubbed
with inert modules, and `faiss.Kmeans` is bound to oracle.cluster_loss.spherical_kmeans (the stand-in of SURVEY.md
Appendix C - faiss is unpinned, so the k-means ENGINE is not what the fixtures pin; everything downstream of
its (assign, centroids) output is: cluster selection, merging, opposite labelling, the three loss terms and the
gradient that reaches the rendered depth).

Run in the build container only (needs /root/reference):  python oracle/gen_golden_loss.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("NCN_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import cluster_loss as cl  # noqa: E402

CAPTURE = {}


def _install_stubs():
    for name in ("vren", "h5py", "imgviz", "pytorch3d", "pytorch3d.transforms"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorch3d.transforms"].matrix_to_euler_angles = lambda *a, **k: None
    sys.modules["imgviz"].label_colormap = lambda *a, **k: None
    sys.modules["imgviz"].depth2rgb = lambda *a, **k: None
    sys.modules["pytorch3d"].transforms = sys.modules["pytorch3d.transforms"]
    faiss = types.ModuleType("faiss")

    class _Index:
        def __init__(self, o):
            self.o = o

        def search(self, x, k):
            sim = x @ self.o.centroids.T
            i = sim.argmax(1)
            return sim[np.arange(len(x)), i][:, None], i[:, None].astype(np.int64)

    class Kmeans:
        def __init__(self, d, k, niter=25, gpu=False, spherical=False, verbose=False):
            self.k, self.niter = k, niter
            self.index = _Index(self)

        def train(self, x):
            self.centroids, a = cl.spherical_kmeans(x, self.k, self.niter)
            CAPTURE["centroids"], CAPTURE["assign"], CAPTURE["kmeans_in"] = self.centroids.copy(), a.copy(), x.copy()

    faiss.Kmeans = Kmeans
    contrib = types.ModuleType("faiss.contrib")
    tu = types.ModuleType("faiss.contrib.torch_utils")
    sys.modules.update({"faiss": faiss, "faiss.contrib": contrib, "faiss.contrib.torch_utils": tu})
    faiss.contrib = contrib; contrib.torch_utils = tu


if __name__ == "__main__":
    _install_stubs()
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "normal-clustering-nerf_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
    import losses as ref_losses          # the reference file, unmodified

    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    hp = dict(loss_opacity_w=1e-3, loss_norm_can_tres=0.01, loss_norm_D_C_ort_dot_w=2e-3, loss_norm_D_C_centr_dot_w=2e-3,
              loss_norm_D_C_centr_L1_w=2e-3, loss_norm_can_start=500, loss_norm_can_grow=2500, loss_norm_can_end=-1,
              ray_sampling_strategy="all_images_triang_patch", random_tr_poses=False, pred_norm_nn=False,
              pred_norm_depth=True)
    # case c = BASELINE.json config 1 (8192 rays -> 6272 normals): above 256*K = 5120 points the k-means trains on a sub-sample and
    # assigns ALL points afterwards (faiss max_points_per_centroid) - the branch the small cases never reach
    cases = {"a": (1024, 0), "b": (2048, 3), "c": (8192, 5)}
    only = [a for a in sys.argv[1:] if a in cases]
    for case, (n_rays, seed) in cases.items():
        if only and case not in only:
            continue
        torch.manual_seed(seed)
        b = synth.patch_batch(n_rays, seed=seed)
        rays_d = torch.from_numpy(b["rays_d"])
        rays_o = torch.from_numpy(b["rays_o"])
        # depth of a synthetic room: distance to the walls of [-0.4,0.4]^3 along the ray + noise
        t_wall = torch.where(rays_d > 0, (0.4 - rays_o) / rays_d, (-0.4 - rays_o) / rays_d).min(-1)[0]
        depth = (t_wall + 0.002 * torch.randn(n_rays)).clamp_min(0.02).requires_grad_(True)
        depth.data[::97] = 0.0
        opacity = torch.rand(n_rays).clamp(0.05, 0.99)
        rgb = torch.rand(n_rays, 3, requires_grad=True)
        tri = b["tri"][:, :49] % 64
        pred = {"rgb": rgb, "depth": depth, "opacity": opacity, "rays_o": rays_d, "rays_d": rays_d,   # rays_o := rays_d (rendering.py:227)
                "deltas": torch.zeros(1), "ts": torch.zeros(1), "rays_a": torch.zeros(1, 3, dtype=torch.int64)}
        target = {"rgb": torch.rand(n_rays, 3), "patch_area": 64,
                  "x1_offsets_local": torch.from_numpy(tri[0]), "x2_offsets_local": torch.from_numpy(tri[1]),
                  "x3_offsets_local": torch.from_numpy(tri[2])}
        loss_fn = ref_losses.NeRFMTLoss(hp)
        step = 3000
        # capture the reference's own cluster labels by wrapping its function
        orig = ref_losses._normals_clustering
        def wrapped(*a, **k):
            r = orig(*a, **k)
            CAPTURE["labels"], CAPTURE["centrs_new"] = r[0].numpy().copy(), r[2].numpy().copy()
            return r
        ref_losses._normals_clustering = wrapped
        loss_d = loss_fn(pred, target, global_step=step)
        ref_losses._normals_clustering = orig
        loss_d["total"].backward()
        normals = ref_losses._extract_normals_from_ray_batch(rays_d, rays_d, depth.detach(),
                                                             {k: torch.from_numpy(b["tri"][i]) for i, k in enumerate(("x1", "x2", "x3"))})
        valid = cl.valid_rows(normals).numpy()
        w = loss_fn.w_sched(2e-3, step)
        np.savez_compressed(
            os.path.join(out_dir, f"cluster_loss_{case}.npz"),
            rays_d=b["rays_d"], depth=depth.detach().numpy(), opacity=opacity.numpy(), rgb=rgb.detach().numpy(),
            target_rgb=target["rgb"].numpy(), tri=b["tri"], tri_local=tri, normals=normals.numpy(), valid=valid,
            kmeans_centroids=CAPTURE["centroids"], kmeans_assign=CAPTURE["assign"], labels=CAPTURE["labels"],
            centrs_new=CAPTURE["centrs_new"], step=step, w_sched=w,
            loss_rgb=float(loss_d["rgb"]), loss_opacity=float(loss_d["opacity"]),
            loss_ort=float(loss_d["norm_D_C_ort_dot"]), loss_dot=float(loss_d["norm_D_C_centr_dot"]),
            loss_l1=float(loss_d["norm_D_C_centr_L1"]), loss_total=float(loss_d["total"]),
            grad_depth=depth.grad.numpy(), grad_rgb=rgb.grad.numpy())
        print(case, {k: float(v) for k, v in loss_d.items()}, "labels", np.bincount(CAPTURE["labels"] + 3, minlength=7),
              "valid", valid.sum(), "/", len(valid))
