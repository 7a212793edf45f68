"""Host-side wrappers of the normal-clustering kernels (csrc/loss.cu).

  normals_from_depth   datasets/hypersim_src/utils.py:505-541 (_extract_normals_from_ray_batch), differentiable in depth
  kmeans_spherical     faiss.Kmeans(3, K, niter, spherical=True) as called at losses.py:86-92 - on the GPU, no host sync
  cluster_select       losses.py:97-166 (orthogonal triple, merge, opposite) - on the GPU, no .item()
  cluster_loss         losses.py:441-478 (L_ort, L_dot, L_L1) with its analytic gradient
  normals_clustering   drop-in for losses._normals_clustering(normals, ...) on device tensors
  normals_from_depth_image   datasets/hypersim_src/utils.py:544-611 (_extract_normals_from_depth_batch), evaluation loop
  rotation_from_normals      train_nerf.py:491-517: scene rotation recovered from the clustered normals of the rendered views
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, stream


class _NormalsFromDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, origin, direction, depth, x1, x2, x3):
        m = x1.shape[0]
        out = torch.empty(m, 3, dtype=torch.float32, device=depth.device)
        check(_lib.lib().ncn_normals_from_depth_fw(ptr(origin), ptr(direction), ptr(depth), ptr(x1), ptr(x2), ptr(x3), m,
                                                   ptr(out), stream()), "normals_from_depth_fw")
        ctx.save_for_backward(origin, direction, depth, x1, x2, x3)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        origin, direction, depth, x1, x2, x3 = ctx.saved_tensors
        gd = torch.zeros_like(depth)
        check(_lib.lib().ncn_normals_from_depth_bw(ptr(origin), ptr(direction), ptr(depth), ptr(x1), ptr(x2), ptr(x3),
                                                   ptr(g.contiguous().float()), x1.shape[0], ptr(gd), stream()),
              "normals_from_depth_bw")
        return None, None, gd, None, None, None


def normals_from_depth(rays_o, rays_d, depth, x123_idx):
    """normals (M,3) = normalize(cross(P2-P1, P3-P1)), P = rays_o + rays_d*depth; grad flows to depth."""
    f = lambda t: t.detach().float().contiguous()
    i = lambda t: t.to(torch.int64).contiguous()
    return _NormalsFromDepth.apply(f(rays_o), f(rays_d), depth.float().contiguous(), i(x123_idx["x1"]),
                                   i(x123_idx["x2"]), i(x123_idx["x3"]))


def kmeans_spherical(x, k=20, niter=20, seed=1234, max_points_per_centroid=256, spherical=True):
    """x (M,3) f32 CUDA (invalid rows - zero / NaN / Inf - are skipped) -> centroids (k,3), assign (M) i32 (-1 = skipped), n_valid (1) i32."""
    x = x.detach().float().contiguous()
    m = x.shape[0]
    dev = x.device
    p = _lib.KmeansParams(int(k), int(niter), int(seed), int(max_points_per_centroid), 1 if spherical else 0)
    cent = torch.empty(k, 3, dtype=torch.float32, device=dev)
    assign = torch.empty(m, dtype=torch.int32, device=dev)
    nv = torch.empty(1, dtype=torch.int32, device=dev)
    L = _lib.lib()
    ws = torch.empty(L.ncn_kmeans_workspace_bytes(m, k), dtype=torch.uint8, device=dev)
    check(L.ncn_kmeans_spherical(ptr(x), m, C.byref(p), ptr(cent), ptr(assign), ptr(nv), ptr(ws), ws.numel(), stream()),
          "kmeans_spherical")
    return cent, assign, nv


def cluster_select(centroids, assign, t_similar):
    """-> labels (M) i32 in {-3..3} (0 = unused), sel (3) i32 = (c1, c2, c3)"""
    m = assign.shape[0]
    labels = torch.empty(m, dtype=torch.int32, device=assign.device)
    sel = torch.empty(3, dtype=torch.int32, device=assign.device)
    check(_lib.lib().ncn_cluster_select(ptr(centroids), ptr(assign), m, centroids.shape[0], float(t_similar),
                                        ptr(labels), ptr(sel), stream()), "cluster_select")
    return labels, sel


class _ClusterLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, normals, labels):
        dev = normals.device
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        stats = torch.empty(32, dtype=torch.float32, device=dev)
        check(_lib.lib().ncn_cluster_loss_fw(ptr(normals), ptr(labels), normals.shape[0], ptr(losses), ptr(stats), stream()),
              "cluster_loss_fw")
        ctx.save_for_backward(normals, labels, stats)
        return losses

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        normals, labels, stats = ctx.saved_tensors
        w = torch.nan_to_num(g.float()).contiguous()
        dn = torch.empty_like(normals)
        check(_lib.lib().ncn_cluster_loss_bw(ptr(normals), ptr(labels), normals.shape[0], ptr(stats), ptr(w), ptr(dn), stream()),
              "cluster_loss_bw")
        return dn, None


def cluster_loss(normals, labels):
    """-> tensor (3) = [L_ort_dot, L_centr_dot, L_centr_L1] (unweighted; NaN when a cluster is empty)."""
    return _ClusterLoss.apply(normals.float().contiguous(), labels.contiguous())


def normals_clustering(normals, K=10, niter=10, t_similar=0.99, seed=1234):
    """Device-resident equivalent of losses._normals_clustering: returns (labels (M) in {-3..3} with 0 for
    unused AND invalid rows, assign (M), centroids of the selected triple (3,3))."""
    cent, assign, _ = kmeans_spherical(normals, K, niter, seed=seed)
    labels, sel = cluster_select(cent, assign, t_similar)
    return labels, assign, cent[sel.long()]


@torch.no_grad()
def normals_from_depth_image(depth, ray_dirs_cc, poses):
    """_extract_normals_from_depth_batch: depth (B,H,W), ray_dirs_cc (H*W,3), poses (B,3|4,4) -> world-frame normals (B,H,W,3);
    zero on the border and where the pixel's depth is 0 / NaN / Inf."""
    depth = depth.float().contiguous(); ray_dirs_cc = ray_dirs_cc.float().contiguous(); poses = poses.float().contiguous()
    B, H, W = depth.shape
    if ray_dirs_cc.shape != (H * W, 3) or poses.shape[0] != B or poses.shape[1] not in (3, 4) or poses.shape[2] != 4:
        raise RuntimeError("normals_from_depth_image: expected depth (B,H,W), ray_dirs_cc (H*W,3), poses (B,3|4,4)")
    out = torch.empty(B, H, W, 3, dtype=torch.float32, device=depth.device)
    check(_lib.lib().ncn_normals_from_depth_image(ptr(depth), ptr(ray_dirs_cc), ptr(poses), poses.shape[1], B, H, W, ptr(out), stream()),
          "normals_from_depth_image")
    return out


def rotation_from_centroids(centrs_new, R_offset):
    """train_nerf.py:505-517 downstream of the clustering: the three selected centroids (rows) and their negatives are matched
    to the columns of R_offset by largest dot product, and the result is projected onto SO(3) (scipy Rotation.from_matrix,
    as the reference).  3x3 host arithmetic; returns a float64 (3,3) tensor."""
    from scipy.spatial.transform import Rotation
    c = centrs_new.detach().cpu().double().T                           # columns = cluster directions
    both = torch.cat([c, -c], 1)                                        # (3,6)
    sim = (R_offset.detach().cpu().double().unsqueeze(1) * both.unsqueeze(-1)).sum(0)      # (6,3)
    rot = both[:, torch.argmax(sim, 0)]
    return torch.from_numpy(Rotation.from_matrix(rot.numpy()).as_matrix())


@torch.no_grad()
def rotation_from_normals(normals, R_offset, K=30, niter=30, t_similar=0.99, seed=1234):
    """validation_epoch_end (train_nerf.py:491-517): cluster all rendered-view normals (invalid rows skipped on the device),
    take the most orthogonal triple and read the scene rotation off it.  normals: (...,3) CUDA tensor."""
    n = normals.reshape(-1, 3).float()
    # the reference drops rows whose components SUM to zero (train_nerf.py:496) - a superset of the all-zero invalid code;
    # zeroing them lets the device-side validity filter of the k-means skip exactly those rows
    n = torch.where((n.sum(-1) != 0.0).unsqueeze(-1), n, torch.zeros_like(n))
    _, _, centrs = normals_clustering(n, K=K, niter=niter, t_similar=t_similar, seed=seed)
    return rotation_from_centroids(centrs, R_offset)
