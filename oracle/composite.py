"""TEST INFRASTRUCTURE ONLY.  Torch restatement of composite_train_multi_fw (volumerendering.cu:97-137):
per ray, front to back: a = 1 - exp(-sigma*delta); w = a*T; rend += w*raw; depth += w*t; opacity += w; T *= 1-a;
stop after the first sample that drives T <= threshold (that sample is composited but not counted).
The reference has NO CPU compositing path (CHECK_CUDA, utils.h:4); the backward here is torch autograd through
this restatement (the reference's closed form is volumerendering.cu:349-359) - checked against the reference's
own CUDA kernels in tests/test_vren_parity_gpu.py via the product kernels."""
import torch


def composite_train(sigmas, raws, deltas, ts, rays_a, T_threshold=1e-4):
    """-> total_samples (R) i64, opacity (R), depth (R), rend (R,C), ws (N); differentiable in sigmas, raws."""
    R = rays_a.shape[0]
    n = rays_a[:, 2]
    start = rays_a[:, 1]
    ray_idx = rays_a[:, 0]
    nmax = int(n.max()) if R > 0 else 0
    N = sigmas.shape[0]
    C = raws.shape[1]
    if nmax == 0:
        z = sigmas.new_zeros
        return torch.zeros(R, dtype=torch.int64), z(R), z(R), z(R, C), z(N)
    k = torch.arange(nmax, device=sigmas.device)[None, :]
    mask = k < n[:, None]
    idx = (start[:, None] + k).clamp(max=max(N - 1, 0))
    a = torch.where(mask, 1.0 - torch.exp(-sigmas[idx] * deltas[idx]), torch.zeros((), dtype=sigmas.dtype))
    om = 1.0 - a
    T_after = torch.cumprod(om, dim=1)
    T_before = torch.cat([torch.ones(R, 1, dtype=sigmas.dtype), T_after[:, :-1]], 1)
    # a sample is composited iff no EARLIER sample already drove T <= thr
    stopped_before = torch.cat([torch.zeros(R, 1, dtype=torch.bool), (T_after <= T_threshold)[:, :-1]], 1)
    stopped_before = torch.cummax(stopped_before.to(torch.int8), dim=1)[0].bool()
    live = mask & ~stopped_before
    w = torch.where(live, a * T_before, torch.zeros((), dtype=sigmas.dtype))
    counted = live & ~(T_after <= T_threshold)
    total = torch.zeros(R, dtype=torch.int64).index_put_((ray_idx,), counted.sum(1))
    opacity = sigmas.new_zeros(R).index_put((ray_idx,), w.sum(1))
    depth = sigmas.new_zeros(R).index_put((ray_idx,), (w * ts[idx]).sum(1))
    rend = sigmas.new_zeros(R, C).index_put((ray_idx,), (w[..., None] * raws[idx]).sum(1))
    ws = sigmas.new_zeros(N).index_put((idx[mask],), w[mask])
    return total, opacity, depth, rend, ws
