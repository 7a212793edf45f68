"""CPU: the torch restatement in oracle/cluster_loss.py against golden vectors produced by the reference's own
losses.py (oracle/gen_golden_loss.py) - pins selection / merge / opposite labelling, the three cluster terms and
the gradient that reaches the rendered depth."""
import glob
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "cluster_loss_*.npz")))


def _load(path):
    return {k: v for k, v in np.load(path).items()}


@pytest.mark.parametrize("path", CASES)
def test_oracle_matches_reference_golden(path):
    from oracle import cluster_loss as cl
    g = _load(path)
    rays_d = torch.from_numpy(g["rays_d"])
    depth = torch.from_numpy(g["depth"]).requires_grad_(True)
    x123 = {k: torch.from_numpy(g["tri"][i]) for i, k in enumerate(("x1", "x2", "x3"))}
    normals = cl.normals_from_rays(rays_d, rays_d, depth, x123)          # rays_o := rays_d (rendering.py:227)
    torch.testing.assert_close(normals.detach(), torch.from_numpy(g["normals"]), rtol=1e-6, atol=1e-7)
    valid = cl.valid_rows(normals.detach())
    assert np.array_equal(valid.numpy(), g["valid"])
    labels_v, sel = cl.select_clusters(torch.from_numpy(g["kmeans_assign"]), torch.from_numpy(g["kmeans_centroids"]),
                                       1.0 - 0.01)
    assert np.array_equal(labels_v.numpy(), g["labels"])
    torch.testing.assert_close(torch.from_numpy(g["kmeans_centroids"])[list(sel)], torch.from_numpy(g["centrs_new"]))
    ort, dot, l1 = cl.cluster_terms(normals[valid], labels_v)
    w = float(g["w_sched"])
    assert abs(w * float(ort) - float(g["loss_ort"])) <= 1e-6 * max(1, abs(float(g["loss_ort"])))
    assert abs(w * float(dot) - float(g["loss_dot"])) <= 2e-6
    assert abs(w * float(l1) - float(g["loss_l1"])) <= 2e-6
    (w * (ort + dot + l1)).backward()
    torch.testing.assert_close(depth.grad, torch.from_numpy(g["grad_depth"]), rtol=1e-4, atol=1e-7)


def test_kmeans_standin_recovers_planted_frame():
    """oracle k-means + selection on config-1 normals recovers the planted Manhattan frame within 1 degree"""
    from oracle import cluster_loss as cl
    import importlib.util
    spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "normal-clustering-nerf_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
    x, q = synth.manhattan_normals(8192, seed=0)
    xt = torch.from_numpy(x)
    valid = cl.valid_rows(xt)
    cent, assign = cl.spherical_kmeans(x[valid.numpy()], 20, 20)
    labels, sel = cl.select_clusters(torch.from_numpy(assign), torch.from_numpy(cent), 0.99)
    axes = cent[list(sel)]
    cos = np.abs(axes @ q)            # each selected centroid aligns with one planted axis
    assert (cos.max(1) > np.cos(np.deg2rad(1.5))).all(), cos
    assert sorted(cos.argmax(1).tolist()) == [0, 1, 2]


def test_normals_image_oracle_matches_reference_golden():
    """oracle.cluster_loss.normals_from_depth_image vs the reference's own _extract_normals_from_depth_batch
    (tests/golden/normals_image_a.npz, generator oracle/gen_golden_normals_image.py): border, invalid depth codes, NaN
    propagation from invalid neighbours, rotation to the world frame."""
    import os
    import numpy as np
    import torch
    from oracle import cluster_loss as cl
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "normals_image_a.npz"))
    out = cl.normals_from_depth_image(torch.from_numpy(g["depth"]), torch.from_numpy(g["dirs"]), torch.from_numpy(g["poses"])).numpy()
    ref = g["normals"]
    fin = np.isfinite(ref).all(-1)
    assert (~fin).sum() == 4 and not np.isfinite(out[~fin]).all(-1).any()
    np.testing.assert_allclose(out[fin], ref[fin], rtol=0, atol=2e-6)
    zero = (ref[fin] == 0).all(-1)
    assert zero.sum() > 300 and (out[fin][zero] == 0).all()


def test_rotation_from_centroids_oracle():
    """train_nerf.py:505-517 restated: a noisy, sign-flipped, permuted copy of a rotation's columns is mapped back onto it"""
    import numpy as np
    import torch
    from oracle import cluster_loss as cl
    from scipy.spatial.transform import Rotation
    R = torch.from_numpy(Rotation.from_euler("ZYX", [25.0, -10.0, 5.0], degrees=True).as_matrix())
    rows = torch.stack([-R[:, 2], R[:, 0], -R[:, 1]]) + 0.01 * torch.randn(3, 3, dtype=torch.float64)     # centroids as rows
    rot = cl.rotation_from_centroids(rows.float(), R)
    assert abs(float(torch.det(rot)) - 1.0) < 1e-9
    ang = np.rad2deg(np.arccos(np.clip((np.trace((rot.T @ R).numpy()) - 1) / 2, -1, 1)))
    assert ang < 1.5, ang


def test_semantic_loss_module_path_matches_reference_golden():
    """ncn_b200.losses.NeRFMTLoss with the semantic term (CPU tensors: the clustering weights are 0, so no kernel is involved)
    against the reference's own losses.py run on the same inputs (tests/golden/sem_loss_a.npz, oracle/gen_golden_sem.py):
    'sem', 'rgb', 'opacity', 'total' and the gradients reaching the rendered logits / colours; all-void batch -> term dropped."""
    import os
    import numpy as np
    import torch
    import ncn_b200  # noqa: F401
    from ncn_b200.losses import NeRFMTLoss
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sem_loss_a.npz"))
    hp = dict(loss_opacity_w=float(g["opacity_w"]), loss_sem_w=float(g["sem_w"]), pred_sem=True, pred_norm_nn=True, pred_norm_depth=False,
              ray_sampling_strategy="all_images")
    for case in ("a", "b", "void"):
        sem = torch.from_numpy(g[f"{case}_sem"]).requires_grad_(True)
        rgb = torch.from_numpy(g[f"{case}_rgb"]).requires_grad_(True)
        pred = {"rgb": rgb, "depth": torch.zeros(len(rgb)), "opacity": torch.from_numpy(g[f"{case}_opacity"]), "sem": sem}
        target = {"rgb": torch.from_numpy(g[f"{case}_target_rgb"]), "semantics": torch.from_numpy(g[f"{case}_labels"])}
        loss_d = NeRFMTLoss(hp)(pred, target, global_step=3000)
        loss_d["total"].backward()
        for k in ("sem", "rgb", "opacity", "total"):
            np.testing.assert_allclose(float(loss_d[k]), float(g[f"{case}_loss_{k}"]), rtol=1e-6, atol=1e-9)
        gs = sem.grad if sem.grad is not None else torch.zeros_like(sem)
        assert torch.isfinite(gs).all()                       # all-void batch: zero gradient, not 0 * NaN
        np.testing.assert_allclose(gs.numpy(), g[f"{case}_grad_sem"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(rgb.grad.numpy(), g[f"{case}_grad_rgb"], rtol=1e-5, atol=1e-9)


def test_depth_and_reg_depth_terms_match_reference_golden():
    """module-path NeRFMTLoss: depth supervision (labels 0 = none) and the RegNeRF-style depth smoothness on a triangle batch
    (losses.py:371-385, 411-417 - the reference does NOT multiply the latter by its weight; reproduced) against the reference's
    own losses.py (tests/golden/sem_loss_a.npz, case 'depth')"""
    import os
    import numpy as np
    import torch
    import ncn_b200  # noqa: F401
    from ncn_b200.losses import NeRFMTLoss
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sem_loss_a.npz"))
    hp = dict(loss_opacity_w=float(g["opacity_w"]), loss_depth_w=float(g["depth_w"]), loss_reg_depth_w=float(g["reg_w"]), loss_norm_can_start=500,
              pred_norm_depth=False, ray_sampling_strategy="all_images_triang")
    depth = torch.from_numpy(g["depth_depth"]).requires_grad_(True)
    pred = {"rgb": torch.from_numpy(g["depth_rgb"]), "depth": depth, "opacity": torch.from_numpy(g["depth_opacity"])}
    target = {"rgb": torch.from_numpy(g["depth_target_rgb"]), "depth": torch.from_numpy(g["depth_target_depth"])}
    loss_d = NeRFMTLoss(hp)(pred, target, global_step=3000)
    loss_d["total"].backward()
    np.testing.assert_allclose(float(loss_d["depth"]), float(g["depth_loss_depth"]), rtol=1e-6)
    np.testing.assert_allclose(float(loss_d["reg_depth"]), float(g["depth_loss_reg"]), rtol=1e-6)
    np.testing.assert_allclose(float(loss_d["total"]), float(g["depth_loss_total"]), rtol=1e-6)
    np.testing.assert_allclose(depth.grad.numpy(), g["depth_grad_depth"], rtol=1e-5, atol=1e-9)


def test_canonical_axis_terms_match_reference_golden():
    """ncn_b200.losses.canonical_axis_terms (the optional terms of losses.py:480-502, sync-free masked form) downstream of the
    reference's own cluster labels: values, and - together with the three cluster terms of the oracle - the gradient that
    reaches the rendered depth (tests/golden/cluster_can_a.npz, generator oracle/gen_golden_can.py)"""
    import os
    import numpy as np
    import torch
    import ncn_b200  # noqa: F401
    from ncn_b200.losses import canonical_axis_terms
    from oracle import cluster_loss as cl
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cluster_can_a.npz"))
    rays_d = torch.from_numpy(g["rays_d"])
    depth = torch.from_numpy(g["depth"]).requires_grad_(True)
    x123 = {k: torch.from_numpy(g["tri"][i]) for i, k in enumerate(("x1", "x2", "x3"))}
    normals = cl.normals_from_rays(rays_d, rays_d, depth, x123)
    labels = torch.zeros(normals.shape[0], dtype=torch.int64)
    labels[torch.from_numpy(g["valid"])] = torch.from_numpy(g["labels"])
    can_dot, can_l1, has = canonical_axis_terms(normals, labels, float(g["tres"]))
    assert bool(has)
    np.testing.assert_allclose(float(g["w_can_dot"]) * float(can_dot), float(g["loss_can_dot"]), rtol=2e-5)
    np.testing.assert_allclose(float(g["w_can_l1"]) * float(can_l1), float(g["loss_can_l1"]), rtol=2e-5)
    ort, dot, l1 = cl.cluster_terms(normals, labels)
    w = float(g["w_clu"])
    np.testing.assert_allclose([w * float(ort), w * float(dot), w * float(l1)], [float(g["loss_ort"]), float(g["loss_dot"]), float(g["loss_l1"])], rtol=2e-5)
    (float(g["w_can_dot"]) * can_dot + float(g["w_can_l1"]) * can_l1 + w * (ort + dot + l1)).backward()
    np.testing.assert_allclose(depth.grad.numpy(), g["grad_depth"], rtol=2e-3, atol=2e-7)
    # no cluster mean near an axis -> flag off (the reference then adds no term)
    q, _ = torch.linalg.qr(torch.tensor([[1.0, 2.0, 3.0], [-2.0, 1.0, 0.5], [0.3, -1.0, 2.0]]))
    n = q[:, [0, 1, 2]].T.repeat_interleave(5, 0)
    lab = torch.tensor([1] * 5 + [2] * 5 + [3] * 5)
    assert not bool(canonical_axis_terms(n, lab, 0.01)[2])


def test_gt_normal_terms_match_reference_golden():
    """ncn_b200.losses.gt_normal_terms (losses.py:387-409, masked sync-free form) on the oracle's depth-derived normals against
    the reference's own losses.py (tests/golden/gt_normals_a.npz): both values and the gradient reaching the rendered depth"""
    import os
    import numpy as np
    import torch
    import ncn_b200  # noqa: F401
    from ncn_b200.losses import gt_normal_terms, triangle_indices
    from oracle import cluster_loss as cl
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gt_normals_a.npz"))
    rays_d = torch.from_numpy(g["rays_d"])
    depth = torch.from_numpy(g["depth"]).requires_grad_(True)
    x123 = triangle_indices(depth.shape[0], "all_images_triang", None, "cpu")
    normals = cl.normals_from_rays(rays_d, rays_d, depth, x123)                 # rays_o := rays_d (rendering.py:227)
    l1, dot, has = gt_normal_terms(normals, torch.from_numpy(g["normals_gt"])[x123["x1"]])
    assert bool(has)
    np.testing.assert_allclose(float(g["w_l1"]) * float(l1), float(g["loss_l1"]), rtol=1e-5)
    np.testing.assert_allclose(float(g["w_dot"]) * float(dot), float(g["loss_dot"]), rtol=1e-5)
    (float(g["w_l1"]) * l1 + float(g["w_dot"]) * dot).backward()
    np.testing.assert_allclose(depth.grad.numpy(), g["grad_depth"], rtol=2e-3, atol=2e-7)
    assert not bool(gt_normal_terms(normals.detach(), torch.zeros_like(normals))[2])
