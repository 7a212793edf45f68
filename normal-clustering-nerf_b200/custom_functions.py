"""Autograd surface of the hot path - same class names, ``.apply`` signatures and return tuples
as the reference's models/custom_functions.py (RayAABBIntersector :8, RaySphereIntersector :32,
RayMarcher :55, VolumeRenderer :115, TruncExp :162), so models/rendering.py and models/ngp_mt.py
can import it unchanged.  Every op lands in the sm_100a kernels of libncn.so via ``ncn_b200.vren``.
"""
import torch
from torch.amp import custom_bwd, custom_fwd

from . import vren
from . import _lib
from ._lib import check, ptr, stream

_fwd32 = custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = custom_bwd(device_type="cuda")


class RayAABBIntersector(torch.autograd.Function):
    """rays (N,3) x boxes (V,3) -> hits_cnt (N), hits_t (N,max_hits,2) near->far (-1 = none), hits_voxel_idx."""

    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, center, half_size, max_hits):
        return tuple(vren.ray_aabb_intersect(rays_o, rays_d, center, half_size, max_hits))


class RaySphereIntersector(torch.autograd.Function):
    """rays (N,3) x spheres -> hits_cnt, hits_t, hits_sphere_idx."""

    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, center, radii, max_hits):
        return tuple(vren.ray_sphere_intersect(rays_o, rays_d, center, radii, max_hits))


def segment_sum(src, indptr):
    """per-segment row sum (torch_scatter.segment_csr with reduce='sum')."""
    src = src.contiguous().float()
    indptr = indptr.contiguous().to(torch.int64)
    n_seg = indptr.shape[0] - 1
    d = src.shape[1] if src.dim() > 1 else 1
    out = torch.empty((n_seg, d) if src.dim() > 1 else (n_seg,), dtype=torch.float32, device=src.device)
    check(_lib.lib().ncn_segment_csr_sum(ptr(src), ptr(indptr), n_seg, d, ptr(out), stream()), "segment_csr")
    return out


class RayMarcher(torch.autograd.Function):
    """Occupancy-grid march.  Returns rays_a (N_rays,3) [ray_idx,start_idx,N_samples], xyzs (N,3),
    dirs (N,3), deltas (N), ts (N), total_samples (0-dim)."""

    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, grid_size,
                max_samples):
        noise = torch.rand_like(rays_o[:, 0])        # perturb the first sample of each ray
        rays_a, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(
            rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise, grid_size, max_samples)
        ctx.save_for_backward(rays_a, ts)
        return rays_a, xyzs, dirs, deltas, ts, counter[0]

    @staticmethod
    @_bwd
    def backward(ctx, g_rays_a, g_xyzs, g_dirs, g_deltas, g_ts, g_total):
        rays_a, ts = ctx.saved_tensors
        # rays_a rows are in ray order with start_idx an exclusive prefix sum -> a valid CSR pointer
        indptr = torch.cat([rays_a[:, 1], rays_a[-1:, 1] + rays_a[-1:, 2]])
        g_o = segment_sum(g_xyzs, indptr)
        g_d = segment_sum(g_xyzs * ts[:, None] + g_dirs, indptr)
        return g_o, g_d, None, None, None, None, None, None, None


class VolumeRenderer(torch.autograd.Function):
    """Training-time compositing of ragged per-ray samples: (sigmas (N), raws (N,C), deltas, ts, rays_a,
    T_threshold) -> total_samples (0-dim), opacity (R), depth (R), rend (R,C), ws (N)."""

    @staticmethod
    @_fwd32
    def forward(ctx, sigmas, raws, deltas, ts, rays_a, T_threshold):
        total, opacity, depth, rend, ws = vren.composite_train_multi_fw(sigmas, raws, deltas, ts, rays_a, T_threshold)
        ctx.save_for_backward(sigmas, raws, deltas, ts, rays_a, opacity, depth, rend, ws)
        ctx.T_threshold = T_threshold
        return total.sum(), opacity, depth, rend, ws

    @staticmethod
    @_bwd
    def backward(ctx, g_total, g_opacity, g_depth, g_rend, g_ws):
        sigmas, raws, deltas, ts, rays_a, opacity, depth, rend, ws = ctx.saved_tensors
        g_sigmas, g_raws = vren.composite_train_multi_bw(
            g_opacity.contiguous(), g_depth.contiguous(), g_rend.contiguous(), g_ws.contiguous(), sigmas, raws, ws,
            deltas, ts, rays_a, opacity, depth, rend, ctx.T_threshold)
        return g_sigmas, g_raws, None, None, None, None


class TruncExp(torch.autograd.Function):
    """exp with the backward clamped to exp(clamp(x,-15,15))."""

    @staticmethod
    @_fwd32
    def forward(ctx, x):
        x = x.float()                     # fp32 also outside an autocast region (custom_fwd only casts under autocast)
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @_bwd
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))
