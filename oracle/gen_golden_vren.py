#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/vren_ref_*.npz from the reference's OWN csrc kernels
(oracle/_ref/vren_ref.so, built by oracle/build_ref.py from /root/reference/models/csrc, unmodified) on a GPU.

Run on the GPU box:   gpurun -- 'python oracle/gen_golden_vren.py gpurun_out/golden'
then copy gpurun_out/golden/*.npz to tests/golden/ (done once; the fixtures are committed).
The fixtures pin the CPU oracles (oracle/march_ref.c, oracle/composite.py) in the `-m "not gpu"` suite."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
import build_ref  # noqa: E402
import ncn_b200  # noqa: E402,F401
from ncn_b200 import synth  # noqa: E402


def canon(rays_a, arrays):
    order = torch.argsort(rays_a[:, 0])
    ra = rays_a[order]
    n, starts = ra[:, 2], ra[:, 1]
    within = torch.arange(int(n.sum()), device=n.device) - torch.repeat_interleave(torch.cumsum(n, 0) - n, n)
    src = torch.repeat_interleave(starts, n) + within
    ra = ra.clone(); ra[:, 1] = torch.cumsum(n, 0) - n
    return ra, [a[src] for a in arrays]


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = build_ref.load()
    assert ref is not None, "vren_ref.so missing"
    dev = "cuda"
    occ = synth.room_occupancy(128, 0.5, seed=0)
    grid = synth.density_grid_from_occupancy(occ)
    bits = torch.zeros(128 ** 3 // 8, dtype=torch.uint8, device=dev)
    ref.packbits(torch.from_numpy(grid).to(dev), 5.9, bits)
    center = torch.zeros(1, 3, device=dev); half = torch.full((1, 3), 0.5, device=dev)
    for case, (n, esf, cam, seed) in {"a": (512, 0.0, "hypersim", 0), "b": (384, 1.0 / 256, "scannet", 5)}.items():
        b = synth.random_batch(n, cam=cam, seed=seed)
        rays_o = torch.from_numpy(b["rays_o"]).to(dev); rays_d = torch.from_numpy(b["rays_d"]).to(dev).clone()
        rays_d[:4, 0] = 0.0
        rays_o[4:8] *= 4.0
        _, hits_t, _ = ref.ray_aabb_intersect(rays_o, rays_d, center, half, 1)
        hits_raw = hits_t.clone()
        m = (hits_t[:, 0, 0] >= 0) & (hits_t[:, 0, 0] < 0.01)
        hits_t[m, 0, 0] = 0.01
        noise = torch.rand(n, device=dev, generator=torch.Generator(device=dev).manual_seed(seed))
        rays_a, xyzs, dirs, deltas, ts, counter = ref.raymarching_train(rays_o, rays_d, hits_t[:, 0], bits, 1, 0.5, esf, noise, 128, 1024)
        tot = int(counter[0])
        ra, (xyzs, dirs, deltas, ts) = canon(rays_a, [xyzs[:tot], dirs[:tot], deltas[:tot], ts[:tot]])
        g = torch.Generator(device=dev).manual_seed(seed + 1)
        sigmas = torch.rand(tot, device=dev, generator=g) ** 4 * 600
        raws = torch.rand(tot, 6, device=dev, generator=g)
        total_s, opacity, depth, rend, ws = ref.composite_train_multi_fw(sigmas, raws, deltas, ts, ra, 1e-4)
        dO = torch.randn(n, device=dev, generator=g); dD = torch.randn(n, device=dev, generator=g)
        dR = torch.randn(n, 6, device=dev, generator=g); dW = torch.zeros(tot, device=dev)
        d_sig, d_raws = ref.composite_train_multi_bw(dO, dD, dR, dW, sigmas, raws, ws, deltas, ts, ra, opacity, depth, rend, 1e-4)
        coords = torch.randint(0, 128, (64, 3), dtype=torch.int32, device=dev, generator=g)
        c = lambda t: t.detach().cpu().numpy()
        np.savez_compressed(os.path.join(out_dir, f"vren_ref_{case}.npz"),
                            rays_o=c(rays_o), rays_d=c(rays_d), hits_raw=c(hits_raw[:, 0]), hits_t=c(hits_t[:, 0]), noise=c(noise),
                            esf=esf, rays_a=c(ra), xyzs=c(xyzs), dirs=c(dirs), deltas=c(deltas), ts=c(ts),
                            sigmas=c(sigmas), raws=c(raws), total_samples=c(total_s), opacity=c(opacity), depth=c(depth),
                            rend=c(rend), ws=c(ws), dO=c(dO), dD=c(dD), dR=c(dR), d_sigmas=c(d_sig), d_raws=c(d_raws),
                            coords=c(coords), morton=c(ref.morton3D(coords)))
        print(case, "rays", n, "samples", tot)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
