"""Drop-in for the reference's ``vren`` extension module (models/csrc/binding.cpp:330-350).

Same 15 function names, argument order, return tuples and error behaviour (inputs must be
CUDA + contiguous, otherwise ``RuntimeError("<name> must be a CUDA tensor")`` /
``"... must be contiguous"`` exactly like CHECK_INPUT, models/csrc/include/utils.h:4-6),
implemented by hand-written sm_100a kernels behind the C-ABI of include/ncn.h.
Kernels are enqueued on torch's *current* stream (the reference uses the legacy default
stream).  Outputs are freshly allocated torch tensors; nothing is zero-filled on the host.

Documented deviations (all are layouts the reference itself can produce or supersets):
  * raymarching_train returns exact-length sample arrays (``counter[0]`` rows) instead of
    ``N_rays*max_samples`` zero-filled rows that the caller slices
    (custom_functions.py:91-96) - one host read of the sample count happens here.
  * rays_a rows are in ray order with start_idx an exclusive prefix sum (deterministic);
    the reference's order depends on an atomicAdd race (raymarching.cu:237-241).
"""
import torch

from . import _lib
from ._lib import check, ptr, stream

_workspaces = {}


def _chk(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _chk_all(**kw):
    for k, v in kw.items():
        _chk(v, k)


def _workspace(device, nbytes):
    """Grow-only per-device scratch (torch-owned)."""
    key = (device.type, device.index)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ----------------------------------------------------------------------------- intersection
def ray_aabb_intersect(rays_o, rays_d, centers, half_sizes, max_hits):
    """binding.cpp:12-24 -> [hit_cnt (R) i32, hits_t (R,max_hits,2) f32, hits_voxel_idx (R,max_hits) i64]"""
    _chk_all(rays_o=rays_o, rays_d=rays_d, centers=centers, half_sizes=half_sizes)
    R, V = rays_o.shape[0], centers.shape[0]
    dev = rays_o.device
    hit_cnt = torch.empty(R, dtype=torch.int32, device=dev)
    hits_t = torch.empty(R, max_hits, 2, dtype=torch.float32, device=dev)
    hits_idx = torch.empty(R, max_hits, dtype=torch.int64, device=dev)
    check(_lib.lib().ncn_ray_aabb_intersect(ptr(rays_o), ptr(rays_d), ptr(centers), ptr(half_sizes), R, V,
                                            int(max_hits), ptr(hit_cnt), ptr(hits_t), ptr(hits_idx), stream()),
          "ray_aabb_intersect")
    return [hit_cnt, hits_t, hits_idx]


def ray_sphere_intersect(rays_o, rays_d, centers, radii, max_hits):
    """binding.cpp:27-41"""
    _chk_all(rays_o=rays_o, rays_d=rays_d, centers=centers, radii=radii)
    R, V = rays_o.shape[0], centers.shape[0]
    dev = rays_o.device
    hit_cnt = torch.empty(R, dtype=torch.int32, device=dev)
    hits_t = torch.empty(R, max_hits, 2, dtype=torch.float32, device=dev)
    hits_idx = torch.empty(R, max_hits, dtype=torch.int64, device=dev)
    check(_lib.lib().ncn_ray_sphere_intersect(ptr(rays_o), ptr(rays_d), ptr(centers), ptr(radii), R, V,
                                              int(max_hits), ptr(hit_cnt), ptr(hits_t), ptr(hits_idx), stream()),
          "ray_sphere_intersect")
    return [hit_cnt, hits_t, hits_idx]


# ----------------------------------------------------------------------------- occupancy grid
def packbits(density_grid, density_threshold, density_bitfield):
    """binding.cpp:44-56: bit i of byte n = density_grid[8n+i] > threshold; writes density_bitfield in place."""
    _chk_all(density_grid=density_grid, density_bitfield=density_bitfield)
    if density_grid.dtype != torch.float32:
        raise RuntimeError("packbits: density_grid must be float32")
    check(_lib.lib().ncn_packbits(ptr(density_grid), density_bitfield.shape[0], float(density_threshold),
                                  ptr(density_bitfield), stream()), "packbits")


def morton3D(coords):
    """binding.cpp:59-65: (n,3) i32 -> (n) i32"""
    _chk(coords, "coords")
    out = torch.empty(coords.shape[0], dtype=coords.dtype, device=coords.device)
    check(_lib.lib().ncn_morton3d(ptr(coords), coords.shape[0], ptr(out), stream()), "morton3D")
    return out


def morton3D_invert(indices):
    """binding.cpp:68-74: (n) i32 -> (n,3) i32"""
    _chk(indices, "indices")
    out = torch.empty(indices.shape[0], 3, dtype=indices.dtype, device=indices.device)
    check(_lib.lib().ncn_morton3d_invert(ptr(indices), indices.shape[0], ptr(out), stream()), "morton3D_invert")
    return out


# ----------------------------------------------------------------------------- marching
def raymarching_train(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise,
                      grid_size, max_samples):
    """binding.cpp:78-99 -> [rays_a (R,3) i64, xyzs (N,3), dirs (N,3), deltas (N), ts (N), counter (2) i32]"""
    _chk_all(rays_o=rays_o, rays_d=rays_d, hits_t=hits_t, density_bitfield=density_bitfield, noise=noise)
    R = rays_o.shape[0]
    dev = rays_o.device
    L = _lib.lib()
    nbytes = L.ncn_march_train_workspace_bytes(R, int(max_samples))
    ws = _workspace(dev, nbytes)
    rays_a = torch.empty(R, 3, dtype=torch.int64, device=dev)
    counter = torch.empty(2, dtype=torch.int32, device=dev)
    st = stream()
    check(L.ncn_march_train_count(ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(density_bitfield), int(cascades),
                                  float(scale), float(exp_step_factor), ptr(noise), int(grid_size),
                                  int(max_samples), R, ptr(rays_a), ptr(counter), ptr(ws), ws.numel(), st),
          "raymarching_train")
    n = int(counter[0].item())   # the one host read the reference surface forces (custom_functions.py:91)
    xyzs = torch.empty(n, 3, dtype=torch.float32, device=dev)
    dirs = torch.empty(n, 3, dtype=torch.float32, device=dev)
    deltas = torch.empty(n, dtype=torch.float32, device=dev)
    ts = torch.empty(n, dtype=torch.float32, device=dev)
    check(L.ncn_march_train_expand(ptr(rays_o), ptr(rays_d), ptr(rays_a), float(exp_step_factor), float(scale),
                                   int(grid_size), int(max_samples), R, n, ptr(xyzs), ptr(dirs), ptr(deltas),
                                   ptr(ts), ptr(ws), ws.numel(), st), "raymarching_train")
    return [rays_a, xyzs, dirs, deltas, ts, counter]


def raymarching_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale, exp_step_factor,
                     grid_size, max_samples, N_samples):
    """binding.cpp:102-124 -> [xyzs (A,S,3), dirs (A,S,3), deltas (A,S), ts (A,S), N_eff (A) i32]; hits_t[:,0] advanced in place."""
    _chk_all(rays_o=rays_o, rays_d=rays_d, hits_t=hits_t, alive_indices=alive_indices,
             density_bitfield=density_bitfield)
    A, S = alive_indices.shape[0], int(N_samples)
    dev = rays_o.device
    xyzs = torch.empty(A, S, 3, dtype=torch.float32, device=dev)
    dirs = torch.empty(A, S, 3, dtype=torch.float32, device=dev)
    deltas = torch.empty(A, S, dtype=torch.float32, device=dev)
    ts = torch.empty(A, S, dtype=torch.float32, device=dev)
    n_eff = torch.empty(A, dtype=torch.int32, device=dev)
    check(_lib.lib().ncn_march_test(ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(alive_indices), ptr(density_bitfield),
                                    int(cascades), float(scale), float(exp_step_factor), int(grid_size),
                                    int(max_samples), S, A, ptr(xyzs), ptr(dirs), ptr(deltas), ptr(ts), ptr(n_eff),
                                    stream()), "raymarching_test")
    return [xyzs, dirs, deltas, ts, n_eff]


# ----------------------------------------------------------------------------- compositing
def _composite_fw(sigmas, raws, deltas, ts, rays_a, thr, what):
    R, N = rays_a.shape[0], sigmas.shape[0]
    C = raws.shape[1]
    dev = sigmas.device
    total_samples = torch.empty(R, dtype=torch.int64, device=dev)
    opacity = torch.empty(R, dtype=torch.float32, device=dev)
    depth = torch.empty(R, dtype=torch.float32, device=dev)
    rend = torch.empty(R, C, dtype=torch.float32, device=dev)
    ws = torch.empty(N, dtype=torch.float32, device=dev)
    check(_lib.lib().ncn_composite_train_fw(ptr(sigmas), ptr(raws), ptr(deltas), ptr(ts), ptr(rays_a), float(thr),
                                            R, N, C, ptr(total_samples), ptr(opacity), ptr(depth), ptr(rend),
                                            ptr(ws), stream()), what)
    return [total_samples, opacity, depth, rend, ws]


def composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, opacity_threshold):
    """binding.cpp:134-150 -> [total_samples (R) i64, opacity (R), depth (R), rgb (R,3), ws (N)]"""
    _chk_all(sigmas=sigmas, rgbs=rgbs, deltas=deltas, ts=ts, rays_a=rays_a)
    return _composite_fw(sigmas, rgbs, deltas, ts, rays_a, opacity_threshold, "composite_train_fw")


def composite_train_multi_fw(sigmas, raws, deltas, ts, rays_a, opacity_threshold):
    """binding.cpp:153-169 (N-channel) -> [total_samples, opacity, depth, rend (R,C), ws]"""
    _chk_all(sigmas=sigmas, raws=raws, deltas=deltas, ts=ts, rays_a=rays_a)
    return _composite_fw(sigmas, raws, deltas, ts, rays_a, opacity_threshold, "composite_train_multi_fw")


def _composite_bw(dL_dopacity, dL_ddepth, dL_drend, dL_dws, sigmas, raws, ws, deltas, ts, rays_a, opacity, depth,
                  rend, thr, what):
    R, N = rays_a.shape[0], sigmas.shape[0]
    C = raws.shape[1]
    dev = sigmas.device
    dL_dsigmas = torch.empty(N, dtype=torch.float32, device=dev)
    dL_draws = torch.empty(N, C, dtype=torch.float32, device=dev)
    check(_lib.lib().ncn_composite_train_bw(ptr(dL_dopacity), ptr(dL_ddepth), ptr(dL_drend), ptr(dL_dws),
                                            ptr(sigmas), ptr(raws), ptr(ws), ptr(deltas), ptr(ts), ptr(rays_a),
                                            ptr(opacity), ptr(depth), ptr(rend), float(thr), R, N, C,
                                            ptr(dL_dsigmas), ptr(dL_draws), stream()), what)
    return [dL_dsigmas, dL_draws]


def composite_train_bw(dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, ws, deltas, ts, rays_a, opacity,
                       depth, rgb, opacity_threshold):
    """binding.cpp:172-205 -> [dL_dsigmas (N), dL_drgbs (N,3)]"""
    _chk_all(dL_dopacity=dL_dopacity, dL_ddepth=dL_ddepth, dL_drgb=dL_drgb, dL_dws=dL_dws, sigmas=sigmas,
             rgbs=rgbs, ws=ws, deltas=deltas, ts=ts, rays_a=rays_a, opacity=opacity, depth=depth, rgb=rgb)
    return _composite_bw(dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, ws, deltas, ts, rays_a, opacity,
                         depth, rgb, opacity_threshold, "composite_train_bw")


def composite_train_multi_bw(dL_dopacity, dL_ddepth, dL_drend, dL_dws, sigmas, raws, ws, deltas, ts, rays_a,
                             opacity, depth, rend, opacity_threshold):
    """binding.cpp:208-241 -> [dL_dsigmas (N), dL_draws (N,C)]"""
    _chk_all(dL_dopacity=dL_dopacity, dL_ddepth=dL_ddepth, dL_drend=dL_drend, dL_dws=dL_dws, sigmas=sigmas,
             raws=raws, ws=ws, deltas=deltas, ts=ts, rays_a=rays_a, opacity=opacity, depth=depth, rend=rend)
    return _composite_bw(dL_dopacity, dL_ddepth, dL_drend, dL_dws, sigmas, raws, ws, deltas, ts, rays_a, opacity,
                         depth, rend, opacity_threshold, "composite_train_multi_bw")


def _composite_test(sigmas, raws, deltas, ts, hits_t, alive_indices, thr, N_eff_samples, opacity, depth, rend,
                    what):
    A, S = sigmas.shape[0], sigmas.shape[1]
    C = raws.shape[2]
    check(_lib.lib().ncn_composite_test_fw(ptr(sigmas), ptr(raws), ptr(deltas), ptr(ts), ptr(alive_indices),
                                           float(thr), ptr(N_eff_samples), A, S, C, ptr(opacity), ptr(depth),
                                           ptr(rend), stream()), what)


def composite_test_fw(sigmas, rgbs, deltas, ts, hits_t, alive_indices, T_threshold, N_eff_samples, opacity, depth,
                      rgb):
    """binding.cpp:244-269: in-place update of opacity/depth/rgb, alive_indices[n] = -1 for finished rays."""
    _chk_all(sigmas=sigmas, rgbs=rgbs, deltas=deltas, ts=ts, hits_t=hits_t, alive_indices=alive_indices,
             N_eff_samples=N_eff_samples, opacity=opacity, depth=depth, rgb=rgb)
    _composite_test(sigmas, rgbs, deltas, ts, hits_t, alive_indices, T_threshold, N_eff_samples, opacity, depth,
                    rgb, "composite_test_fw")


def composite_test_multi_fw(sigmas, raws, deltas, ts, hits_t, alive_indices, T_threshold, N_eff_samples, opacity,
                            depth, rend):
    """binding.cpp:272-297 (N-channel)"""
    _chk_all(sigmas=sigmas, raws=raws, deltas=deltas, ts=ts, hits_t=hits_t, alive_indices=alive_indices,
             N_eff_samples=N_eff_samples, opacity=opacity, depth=depth, rend=rend)
    _composite_test(sigmas, raws, deltas, ts, hits_t, alive_indices, T_threshold, N_eff_samples, opacity, depth,
                    rend, "composite_test_multi_fw")


# ----------------------------------------------------------------------------- distortion loss
def distortion_loss_fw(ws, deltas, ts, rays_a):
    """binding.cpp:301-311 -> [loss (R), ws_inclusive_scan (N), wts_inclusive_scan (N)]"""
    _chk_all(ws=ws, deltas=deltas, ts=ts, rays_a=rays_a)
    R, N = rays_a.shape[0], ws.shape[0]
    dev = ws.device
    loss = torch.zeros(R, dtype=torch.float32, device=dev)
    ws_inc = torch.zeros(N, dtype=torch.float32, device=dev)
    wts_inc = torch.zeros(N, dtype=torch.float32, device=dev)
    check(_lib.lib().ncn_distortion_fw(ptr(ws), ptr(deltas), ptr(ts), ptr(rays_a), R, N, ptr(loss), ptr(ws_inc),
                                       ptr(wts_inc), stream()), "distortion_loss_fw")
    return [loss, ws_inc, wts_inc]


def distortion_loss_bw(dL_dloss, ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts, rays_a):
    """binding.cpp:314-327 -> dL_dws (N)"""
    _chk_all(dL_dloss=dL_dloss, ws_inclusive_scan=ws_inclusive_scan, wts_inclusive_scan=wts_inclusive_scan, ws=ws,
             deltas=deltas, ts=ts, rays_a=rays_a)
    R, N = rays_a.shape[0], ws.shape[0]
    dL_dws = torch.zeros(N, dtype=torch.float32, device=ws.device)
    check(_lib.lib().ncn_distortion_bw(ptr(dL_dloss), ptr(ws_inclusive_scan), ptr(wts_inclusive_scan), ptr(ws),
                                       ptr(deltas), ptr(ts), ptr(rays_a), R, N, ptr(dL_dws), stream()),
          "distortion_loss_bw")
    return dL_dws
