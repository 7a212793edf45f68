"""Developer tool: where the table backward spends its time.  Samples = the march of one 8192-ray training batch of the synthetic
room (as bench.py); per level range [a, b) the launch time of ncn_grid_bwd_levels by CUDA events (mean of 20, warm L2, gradient
zeroed outside the timed region), plus the whole kernel in fp32 and fp16 gradient mode.

    python tools/time_grid_bwd.py
"""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ncn_b200  # noqa: F401
from ncn_b200 import _lib, synth, vren
from ncn_b200 import tinycudann as tcnn
from ncn_b200._lib import check, ptr, stream
from ncn_b200.rendering import ray_aabb_near

dev = torch.device("cuda", 0)
L = _lib.lib()
scale, G, R = 0.5, 128, 8192
grid = torch.from_numpy(synth.density_grid_from_occupancy(synth.room_occupancy(G, scale, seed=0))).to(dev)
bits = torch.zeros(G ** 3 // 8, dtype=torch.uint8, device=dev)
vren.packbits(grid, 5.9, bits)
b = synth.patch_batch(R, seed=1000)
ro = torch.from_numpy(b["rays_o"]).to(dev); rd = torch.from_numpy(b["rays_d"]).to(dev)
center = torch.zeros(1, 3, device=dev); half = torch.full((1, 3), scale, device=dev)
hits_t = ray_aabb_near(ro, rd, center, half, 0.01)
noise = torch.rand(R, device=dev)
rays_a, xyzs, dirs, deltas, ts, _ = vren.raymarching_train(ro, rd, hits_t[:, 0], bits, 1, scale, 0.0, noise, G, 1024)
n = xyzs.shape[0]
x01 = ((xyzs + scale) / (2 * scale)).contiguous()
bsc = float(np.exp(np.log(2048 * scale / 16) / 15))
enc = tcnn.Encoding(3, dict(otype="Grid", type="Hash", n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                            base_resolution=16, per_level_scale=bsc, interpolation="Linear")).to(dev)
dfeat = (torch.randn(n, 32, device=dev) * 1e-2).to(torch.float16)
grad = torch.zeros(enc.params.numel(), dtype=torch.float32, device=dev)
grad16 = torch.zeros(enc.params.numel(), dtype=torch.float16, device=dev)
st = stream()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(reps):
        grad.zero_(); grad16.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return 1e3 * tot / reps


res = {"samples": n}
L.ncn_set_grid_bwd_occupancy(6)
res["whole_fp32_occ6_us"] = timed(lambda: check(L.ncn_grid_bwd(C.byref(enc.desc), ptr(x01), ptr(dfeat), n, ptr(grad), 1.0, None, None, st)))
res["levels_0_8_occ6_us"] = timed(lambda: check(L.ncn_grid_bwd_levels(C.byref(enc.desc), ptr(x01), ptr(dfeat), n, ptr(grad), 1.0, None, None, 0, 8, 8, st)))
res["levels_8_16_occ6_us"] = timed(lambda: check(L.ncn_grid_bwd_levels(C.byref(enc.desc), ptr(x01), ptr(dfeat), n, ptr(grad), 1.0, None, None, 8, 16, 8, st)))
L.ncn_set_grid_bwd_occupancy(5)
res["whole_fp32_us"] = timed(lambda: check(L.ncn_grid_bwd(C.byref(enc.desc), ptr(x01), ptr(dfeat), n, ptr(grad), 1.0, None, None, st)))
res["whole_fp16_us"] = timed(lambda: check(L.ncn_grid_bwd_f16(C.byref(enc.desc), ptr(x01), ptr(dfeat), n, ptr(grad16), 1.0, None, None, st)))
for a, e in [(0, 4), (4, 8), (8, 12), (12, 16), (0, 8), (8, 16)]:
    res[f"levels_{a}_{e}_us"] = timed(lambda: check(L.ncn_grid_bwd_levels(C.byref(enc.desc), ptr(x01), ptr(dfeat), n, ptr(grad), 1.0, None, None, a, e, 8, st)))
print(json.dumps(res, indent=1))
