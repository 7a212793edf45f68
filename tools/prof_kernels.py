"""Small driver for ncu captures of individual hot-path kernels (k-means, march, ...). Not part of the product."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ncn_b200
from ncn_b200 import synth, vren, clustering

dev = "cuda"
occ = synth.room_occupancy(128, 0.5, seed=0)
grid = torch.from_numpy(synth.density_grid_from_occupancy(occ)).to(dev)
bits = torch.zeros(128 ** 3 // 8, dtype=torch.uint8, device=dev)
vren.packbits(grid, 5.9, bits)
b = synth.patch_batch(8192, seed=0)
ro = torch.from_numpy(b["rays_o"]).to(dev); rd = torch.from_numpy(b["rays_d"]).to(dev)
center = torch.zeros(1, 3, device=dev); half = torch.full((1, 3), 0.5, device=dev)
x, q = synth.manhattan_normals(6272, seed=0)
xt = torch.from_numpy(x).to(dev)
for it in range(3):
    _, hits, _ = vren.ray_aabb_intersect(ro, rd, center, half, 1)
    noise = torch.rand(8192, device=dev)
    out = vren.raymarching_train(ro, rd, hits[:, 0], bits, 1, 0.5, 0.0, noise, 128, 1024)
    cent, assign, nv = clustering.kmeans_spherical(xt, 20, 20)
    labels, sel = clustering.cluster_select(cent, assign, 0.99)
torch.cuda.synchronize()
print("ok", int(out[5][0]), int(nv))
