"""TEST INFRASTRUCTURE ONLY.  Torch restatement of the reference's Manhattan normal-clustering path:

  normals_from_rays      datasets/hypersim_src/utils.py:505-541 (_extract_normals_from_ray_batch)
  spherical_kmeans       faiss.Kmeans(3, K, niter, spherical=True) as called at losses.py:86-92
                         (faiss is absent and unpinned -> PARITY UNPINNED for the k-means engine; this is
                         the stand-in documented in SURVEY.md Appendix C, used only to produce (assign, centroids))
  select_clusters        losses.py:97-166 given (assign, centroids)
  cluster_terms          losses.py:441-478: L_ort_dot, L_centr_dot, L_centr_L1 (differentiable)

Pinned by tests/golden/cluster_loss_*.npz, which were produced by running the reference's OWN losses.py
(imported unchanged from /root/reference behind import stubs) - see oracle/gen_golden_loss.py.
"""
import numpy as np
import torch
import torch.nn.functional as F


def normals_from_rays(rays_o, rays_d, depth, x123):
    P = rays_o + rays_d * depth.unsqueeze(-1)
    P1, P2, P3 = P[x123["x1"]], P[x123["x2"]], P[x123["x3"]]
    return F.normalize(torch.cross(P2 - P1, P3 - P1, dim=-1), p=2.0, dim=-1)


def valid_rows(n):
    bad = (n.abs().sum(-1) == 0) | (torch.isnan(n).sum(-1) > 0) | (torch.isinf(n).sum(-1) > 0)   # losses.py:427-429
    return ~bad


def spherical_kmeans(x, k, niter, seed=1234, max_points_per_centroid=256):
    """numpy spherical k-means in the shape of faiss.Kmeans. x (n,3) float32 -> (centroids, assign)"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = len(x)
    rng = np.random.RandomState(seed)
    xt = x[rng.permutation(n)[:max_points_per_centroid * k]] if n > max_points_per_centroid * k else x
    c = xt[rng.permutation(len(xt))[:k]].copy()
    c /= np.maximum(np.linalg.norm(c, axis=1, keepdims=True), 1e-20)
    for _ in range(niter):
        a = (xt @ c.T).argmax(1)
        for j in range(k):
            m = a == j
            if m.any():
                c[j] = xt[m].mean(0)
        cnt = np.bincount(a, minlength=k)
        for j in range(k):
            if cnt[j] == 0:
                big = int(cnt.argmax())
                sg = np.array([1, -1, 1], np.float32) / 1024
                c[j] = c[big] * (1 + sg); c[big] = c[big] * (1 - sg)
                cnt[j] = cnt[big] // 2; cnt[big] -= cnt[j]
        c /= np.maximum(np.linalg.norm(c, axis=1, keepdims=True), 1e-20)
    return c.astype(np.float32), (x @ c.T).argmax(1).astype(np.int64)


def select_clusters(assign, centrs, t_similar):
    """losses.py:97-166 -> (labels in {-3..3}, (c1,c2,c3)). assign (n) int64, centrs (K,3) float32 torch."""
    K = centrs.shape[0]
    sim = centrs @ centrs.T
    sim_abs = sim.abs()
    sizes = torch.bincount(assign, minlength=K)
    c1 = int(torch.argmax(sizes))
    crit = sim_abs[:, c1].unsqueeze(1) + sim_abs[c1, :].unsqueeze(0) + sim_abs
    mins, min_idx = torch.min(crit, dim=0)
    c2 = int(torch.argmin(mins))
    c3 = int(min_idx[c2])
    labels = torch.zeros_like(assign)
    lab_of = torch.zeros(K, dtype=torch.int64)
    for q, c in enumerate((c1, c2, c3)):
        lab_of[sim[c] > t_similar] = q + 1
    for q, c in enumerate((c1, c2, c3)):
        o = int(torch.argmin(sim[c]))
        if -sim[c][o] > t_similar:
            lab_of[sim[o] > t_similar] = -(q + 1)
    labels = lab_of[assign]
    return labels, (c1, c2, c3)


def cluster_terms(normals, labels):
    """losses.py:441-478 -> (L_ort_dot, L_centr_dot, L_centr_L1); differentiable w.r.t. normals."""
    keep = labels != 0
    n = normals[keep]
    l = labels[keep]
    n = torch.where((l < 0).unsqueeze(-1), -n, n)
    l = l.abs()
    cl = [n[l == k] for k in (1, 2, 3)]
    c = [F.normalize(x.mean(dim=0, keepdim=True), p=2.0, dim=-1) for x in cl]
    ort = ((c[0] * c[1]).sum().abs() + (c[0] * c[2]).sum().abs() + (c[1] * c[2]).sum().abs()) / 3.0
    dot = sum(1.0 - (x * ck).sum(-1).mean() for x, ck in zip(cl, c)) / 3.0
    l1 = sum((x - ck).abs().sum(-1).mean() for x, ck in zip(cl, c)) / 3.0
    return ort, dot, l1


def normals_from_depth_image(depth, ray_dirs_cc, poses):
    """datasets/hypersim_src/utils.py:544-611 (_extract_normals_from_depth_batch), restated: P = dirs * depth in the camera
    frame; n(y,x) = normalize(cross(P(y-1,x) - P(y,x), P(y,x-1) - P(y,x))) (F.normalize eps 1e-12); rotate by poses[:, :3, :3];
    zero on the one-pixel border (ZeroPad2d) and where the pixel's own depth is 0 / NaN / Inf (:604-606)."""
    B, H, W = depth.shape
    P = ray_dirs_cc.view(1, H, W, 3) * depth.view(B, H, W, 1)
    P1, P2, P3 = P[:, 1:-1, 1:-1], P[:, :-2, 1:-1], P[:, 1:-1, :-2]
    n = torch.cross(P2 - P1, P3 - P1, dim=-1)
    n = n / n.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    n = torch.einsum("bij,bhwj->bhwi", poses[:, :3, :3], n)
    out = torch.zeros(B, H, W, 3, dtype=depth.dtype)
    out[:, 1:-1, 1:-1] = n
    bad = (depth == 0.0) | torch.isnan(depth) | torch.isinf(depth)
    out[bad] = 0.0
    return out


def rotation_from_centroids(centrs_new, R_offset):
    """train_nerf.py:505-517 restated: columns +-c_k matched to the columns of R_offset by largest dot product, projected to SO(3)
    with scipy's Rotation.from_matrix (the reference's choice)."""
    from scipy.spatial.transform import Rotation
    c = centrs_new.double().T
    both = torch.cat([c, -1.0 * c], dim=1)
    sim = (R_offset.double().unsqueeze(1) * both.unsqueeze(-1)).sum(0)
    rot = both[:, torch.argmax(sim, dim=0)]
    return torch.from_numpy(Rotation.from_matrix(rot.numpy()).as_matrix())
