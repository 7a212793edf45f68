"""Developer tool: A/B of the training compositing kernels by lanes-per-ray (ncn_set_composite_width): CUDA-event time per launch
at the bench's regime (8192 patch rays through the synthetic room, ~33 samples per ray) and at a dense regime (early training),
plus a consistency check between the widths (ws / total_samples identical, sums close).  Not part of the product."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np
import torch
import ncn_b200
from ncn_b200 import _lib, synth, vren

L = _lib.lib()
dev = torch.device("cuda")
R = 8192
b = synth.patch_batch(R, seed=1000)
rays_o = torch.from_numpy(b["rays_o"]).to(dev); rays_d = torch.from_numpy(b["rays_d"]).to(dev)
center = torch.zeros(1, 3, device=dev); half = torch.full((1, 3), 0.5, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
res = {}
for regime in ("room", "dense"):
    if regime == "room":
        grid = torch.from_numpy(synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))).to(dev)
    else:
        grid = (torch.rand(1, 128 ** 3, device=dev, generator=g) < 0.5).float() * 10
    bits = torch.zeros(128 ** 3 // 8, dtype=torch.uint8, device=dev)
    vren.packbits(grid, 5.9, bits)
    _, hits_t, _ = vren.ray_aabb_intersect(rays_o, rays_d, center, half, 1)
    hits_t[(hits_t[:, 0, 0] >= 0) & (hits_t[:, 0, 0] < 0.01), 0, 0] = 0.01
    noise = torch.rand(R, device=dev, generator=g)
    rays_a, xyzs, dirs, deltas, ts, cnt = vren.raymarching_train(rays_o, rays_d, hits_t[:, 0].contiguous(), bits, 1, 0.5, 0.0, noise, 128, 1024)
    N = xyzs.shape[0]
    sig = torch.rand(N, device=dev, generator=g) * (30 if regime == "room" else 3)
    out = {"samples": N, "samples_per_ray_mean": N / R, "samples_per_ray_max": int(rays_a[:, 2].max())}
    for C in (3, 9):
        raws = torch.rand(N, C, device=dev, generator=g)
        ref = None
        for w in (32, 16, 8, 4):
            L.ncn_set_composite_width(w)
            def fw():
                return vren.composite_train_multi_fw(sig, raws, deltas, ts, rays_a, 1e-4)
            tot, opa, dep, rend, ws = fw()
            dO = torch.rand(R, device=dev, generator=torch.Generator(device=dev).manual_seed(1)); dD = torch.rand(R, device=dev, generator=torch.Generator(device=dev).manual_seed(3)); dR = torch.rand(R, C, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
            dW = torch.zeros_like(ws)
            def bw():
                return vren.composite_train_multi_bw(dO, dD, dR, dW, sig, raws, ws, deltas, ts, rays_a, opa, dep, rend, 1e-4)
            ds, dr = bw()
            if ref is None:
                sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
                import build_ref
                vr = build_ref.load()
                r_fw = vr.composite_train_multi_fw(sig, raws, deltas, ts, rays_a, 1e-4)
                r_bw = vr.composite_train_multi_bw(dO, dD, dR, dW, sig, raws, r_fw[4], deltas, ts, rays_a, r_fw[1], r_fw[2], r_fw[3], 1e-4)
                ref = (r_fw[0], r_fw[1], r_fw[2], r_fw[3], r_fw[4], r_bw[0], r_bw[1])
            assert torch.equal(tot, ref[0]) and torch.equal(ws, ref[4]), (regime, C, w)
            names_ = ("opacity", "depth", "rend", "dL_dsigmas", "dL_draws")
            for nm, a_, b_ in zip(names_, (opa, dep, rend, ds, dr), (ref[1], ref[2], ref[3], ref[5], ref[6])):
                err = float((a_ - b_).abs().max()); mag = float(b_.abs().max())
                out[f"C{C}_W{w}_maxerr_{nm}"] = [err, mag]
            # kernel-only timing through the C ABI with preallocated outputs
            from ncn_b200._lib import ptr, stream
            ts_ = {}
            for name, fn in (("fw", lambda: L.ncn_composite_train_fw(ptr(sig), ptr(raws), ptr(deltas), ptr(ts), ptr(rays_a), 1e-4, R, N, C, ptr(tot), ptr(opa), ptr(dep), ptr(rend), ptr(ws), stream())),
                             ("bw", lambda: L.ncn_composite_train_bw(ptr(dO), ptr(dD), ptr(dR), None, ptr(sig), ptr(raws), ptr(ws), ptr(deltas), ptr(ts), ptr(rays_a), ptr(opa), ptr(dep), ptr(rend), 1e-4, R, N, C, ptr(ds), ptr(dr), stream())),
                             ("bw_raws_only", lambda: L.ncn_composite_train_bw(None, None, ptr(dR), None, ptr(sig), ptr(raws), ptr(ws), ptr(deltas), ptr(ts), ptr(rays_a), ptr(opa), ptr(dep), ptr(rend), 1e-4, R, N, C, None, ptr(dr), stream()))):
                for _ in range(5):
                    fn()
                torch.cuda.synchronize()
                v = []
                for _ in range(30):
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                    v.append(e0.elapsed_time(e1) * 1e3)
                v.sort()
                ts_[name] = v[len(v) // 2]
            out[f"C{C}_W{w}_us"] = ts_
    res[regime] = out
L.ncn_set_composite_width(16)
print(json.dumps(res, indent=1))
