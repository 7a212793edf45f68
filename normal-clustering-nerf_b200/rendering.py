"""render() - host-side mirror of the reference renderer (models/rendering.py:9-241): same signature,
kwargs (near_distance, max_samples, exp_step_factor, T_threshold, test_time, n_sem_cls, pred_norm_nn_norm,
random_bg, to_cpu/to_numpy) and result-dict keys (rgb, depth, opacity, ws, deltas, ts, rays_a, rm_samples,
vr_samples, rays_o (= rays_d, rendering.py:227), rays_d, depth_std, norm_nn, sem, total_samples).

Host-side differences: the AABB test and the near clamp are one kernel (ncn_ray_aabb_near) instead of
8 launches + 3 boolean-index launches; the dead `(rays_a[:,2]==0).any()` sync (rendering.py:195) is gone; the
test-time loop drops the `valid_mask` boolean gather/scatter (padding slots are evaluated and ignored by the
compositing kernel, which only reads N_eff samples) and polls the number of live rays every few iterations
instead of every iteration.
"""
import torch
import torch.nn.functional as F

from . import _lib, vren
from ._lib import check, ptr, stream
from .custom_functions import RayMarcher, VolumeRenderer


def ray_aabb_near(rays_o, rays_d, center, half_size, near_distance):
    """hits_t (R,1,2): slab test against one box + near clamp (rendering.py:26-28) in one launch."""
    R = rays_o.shape[0]
    hits_t = torch.empty(R, 1, 2, dtype=torch.float32, device=rays_o.device)
    check(_lib.lib().ncn_ray_aabb_near(ptr(rays_o), ptr(rays_d), ptr(center), ptr(half_size), float(near_distance), R,
                                       ptr(hits_t), stream()), "ray_aabb_near")
    return hits_t


def render(model, rays_o, rays_d, **kwargs):
    rays_o = rays_o.float().contiguous()
    rays_d = rays_d.float().contiguous()
    hits_t = ray_aabb_near(rays_o, rays_d, model.center, model.half_size, kwargs["near_distance"])
    with torch.autocast("cuda", dtype=torch.float16):
        fn = _render_rays_test if kwargs.get("test_time", False) else _render_rays_train
        results = fn(model, rays_o, rays_d, hits_t, **kwargs)
    if kwargs.get("to_cpu", False):
        for k, v in results.items():
            if torch.is_tensor(v):
                v = v.cpu()
                results[k] = v.numpy() if kwargs.get("to_numpy", False) else v
    return results


def _split_rend(results, rend, model, kwargs):
    i = 3
    results["rgb"] = rend[..., :i]
    if model.pred_norm:
        results["norm_nn"] = rend[..., i:i + 3]
        if kwargs.get("pred_norm_nn_norm", False):
            results["norm_nn"] = F.normalize(results["norm_nn"], p=2.0, dim=-1)
        i += 3
    if model.pred_sem:
        results["sem"] = rend[..., i:i + kwargs["n_sem_cls"]]


def _raws(model, out):
    raws = out["rgbs"].float()
    if model.pred_norm:
        raws = torch.cat((raws, out["norms"].float()), -1)
    if model.pred_sem:
        raws = torch.cat((raws, out["sems"].float()), -1)
    return raws.contiguous()


def _render_rays_train(model, rays_o, rays_d, hits_t, **kwargs):
    exp_step_factor = kwargs.get("exp_step_factor", 0.)
    results = {}
    rays_a, xyzs, dirs, results["deltas"], results["ts"], results["rm_samples"] = RayMarcher.apply(
        rays_o, rays_d, hits_t[:, 0], model.density_bitfield, model.cascades, model.scale, exp_step_factor,
        model.grid_size, kwargs["max_samples"])
    out = model(xyzs, dirs, **{k: v for k, v in kwargs.items() if not torch.is_tensor(v)})
    (results["vr_samples"], results["opacity"], results["depth"], rend, results["ws"]) = VolumeRenderer.apply(
        out["sigmas"], _raws(model, out), results["deltas"], results["ts"], rays_a, kwargs.get("T_threshold", 1e-4))
    _split_rend(results, rend, model, kwargs)
    results["rays_d"] = rays_d
    results["rays_o"] = rays_d            # sic - reference quirk (rendering.py:226-227), the loss depends on it
    results["rays_a"] = rays_a
    results["depth_std"] = torch.ones_like(results["depth"])
    if exp_step_factor == 0:
        rgb_bg = torch.ones(3, device=rays_o.device)
    else:
        rgb_bg = torch.rand(3, device=rays_o.device) if kwargs.get("random_bg", False) else torch.zeros(3, device=rays_o.device)
    results["rgb"] = results["rgb"] + rgb_bg * (1 - results["opacity"])[:, None]
    return results


@torch.no_grad()
def _render_rays_test(model, rays_o, rays_d, hits_t, **kwargs):
    exp_step_factor = kwargs.get("exp_step_factor", 0.)
    max_samples = kwargs["max_samples"]
    N_rays = rays_o.shape[0]
    dev = rays_o.device
    C = 3 + (3 if model.pred_norm else 0) + (kwargs["n_sem_cls"] if model.pred_sem else 0)
    opacity = torch.zeros(N_rays, device=dev)
    depth = torch.zeros(N_rays, device=dev)
    rend = torch.zeros(N_rays, C, device=dev)
    total_samples = torch.zeros((), dtype=torch.int64, device=dev)
    alive = torch.arange(N_rays, device=dev)
    min_samples = 1 if exp_step_factor == 0 else 4
    samples = 0
    thr = kwargs.get("T_threshold", 1e-4)
    while samples < max_samples:
        N_alive = alive.shape[0]
        if N_alive == 0:
            break
        N_samples = max(min(N_rays // N_alive, 64), min_samples)
        samples += N_samples
        xyzs, dirs, deltas, ts, n_eff = vren.raymarching_test(
            rays_o, rays_d, hits_t[:, 0], alive, model.density_bitfield, model.cascades, model.scale, exp_step_factor,
            model.grid_size, max_samples, N_samples)
        total_samples += n_eff.sum()
        # padding slots have dirs == 0: give them a unit direction so the field is finite; they are never composited
        d = dirs.view(-1, 3)
        d = torch.where((d == 0).all(1, keepdim=True), torch.ones_like(d), d)
        out = model(xyzs.view(-1, 3), d, **{k: v for k, v in kwargs.items() if not torch.is_tensor(v)})
        sigmas = out["sigmas"].float().view(N_alive, N_samples)
        raws = _raws(model, out).view(N_alive, N_samples, C)
        vren.composite_test_multi_fw(sigmas, raws, deltas, ts, hits_t[:, 0], alive, thr, n_eff, opacity, depth, rend)
        alive = alive[alive >= 0]
    results = {"opacity": opacity, "depth": depth, "total_samples": total_samples}
    _split_rend(results, rend, model, kwargs)
    rgb_bg = torch.ones(3, device=dev) if exp_step_factor == 0 else torch.zeros(3, device=dev)
    results["rgb"] = results["rgb"] + rgb_bg * (1 - opacity)[:, None]
    return results


@torch.no_grad()
def render_fast(model, rays_o, rays_d, tile=98304, **kwargs):
    """Evaluation render without the per-round Python loop (the reference's
    `__render_rays_test`, models/rendering.py:45-149, runs up to max_samples rounds of march_test -> field -> composite_test
    with a host sync per round).  Per tile of rays: ONE occupancy march with zero jitter (ncn_march_train: the same
    candidate sequence the test-time march walks), ONE field evaluation of all samples (NGPMT.field_eval), ONE front-to-back
    composite that stops at the same transmittance threshold (ncn_composite_train_fw) - 6-10 launches and one host read (the
    sample count) per tile instead of hundreds of rounds.  Returns the result-dict keys of render(test_time=True):
    rgb, depth, opacity, total_samples (+ norm_nn, sem).  Equivalence with the loop: tests/test_eval_path_gpu.py.
    `tile` = rays per pass (default: one 8-GPU shard of a 1024x768 image, parallel.shard_tiles)."""
    rays_o = rays_o.float().contiguous()
    rays_d = rays_d.float().contiguous()
    exp_step_factor = kwargs.get("exp_step_factor", 0.)
    thr = kwargs.get("T_threshold", 1e-4)
    N = rays_o.shape[0]
    dev = rays_o.device
    C = 3 + (3 if model.pred_norm else 0) + (kwargs.get("n_sem_cls", 0) if model.pred_sem else 0)
    opacity = torch.empty(N, device=dev); depth = torch.empty(N, device=dev); rend = torch.empty(N, C, device=dev)
    total = torch.zeros((), dtype=torch.int64, device=dev)
    noise = torch.zeros(min(tile, N), device=dev)
    for s in range(0, N, tile):
        e = min(N, s + tile)
        ro, rd = rays_o[s:e], rays_d[s:e]
        hits_t = ray_aabb_near(ro, rd, model.center, model.half_size, kwargs["near_distance"])
        rays_a, xyzs, dirs, deltas, ts, _ = vren.raymarching_train(ro, rd, hits_t[:, 0], model.density_bitfield, model.cascades,
                                                                   model.scale, exp_step_factor, noise[:e - s], model.grid_size,
                                                                   kwargs["max_samples"])
        sigmas, raws = model.field_eval(xyzs, dirs)
        n_used, o, d, r, _ = vren.composite_train_multi_fw(sigmas, raws, deltas, ts, rays_a, thr)
        opacity[s:e] = o; depth[s:e] = d; rend[s:e] = r
        total += n_used.sum()
    results = {"opacity": opacity, "depth": depth, "total_samples": total}
    _split_rend(results, rend, model, kwargs)
    rgb_bg = torch.ones(3, device=dev) if exp_step_factor == 0 else torch.zeros(3, device=dev)
    results["rgb"] = results["rgb"] + rgb_bg * (1 - opacity)[:, None]
    if kwargs.get("to_cpu", False):
        for k, v in results.items():
            if torch.is_tensor(v):
                v = v.cpu()
                results[k] = v.numpy() if kwargs.get("to_numpy", False) else v
    return results
