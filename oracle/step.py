"""TEST INFRASTRUCTURE ONLY.  CPU restatement of one full ngp_mt training step - the reference's hot path
(train_nerf.py:314-367 -> rendering.py:152-241 -> ngp_mt.py:196-229 -> losses.py:244-587) - assembled from
the oracles in this package, in plain torch on the host cores:

    AABB + near clamp + occupancy march   oracle/march_ref.c            (intersection.cu, raymarching.cu)
    hash-grid encode                      oracle/hashgrid.py            (tcnn GridEncoding)
    sigma / rgb MLPs, TruncExp            oracle/mlp.py                 (tcnn FullyFusedMLP, custom_functions.py:162-173)
    compositing                           oracle/composite.py           (volumerendering.cu:97-137), autograd backward
    photometric + opacity + clustering    oracle/cluster_loss.py        (losses.py:347-361, 419-509)
    Adam                                  torch.optim.Adam-equivalent update (apex FusedAdam semantics)

Used by bench.py as the CPU baseline (`cpu_baseline.port`, `--impl reference`) and by tests/test_step_oracle_gpu.py as the
end-to-end checker of the GPU step (losses and per-group gradient norms on 512 rays).
"""
import numpy as np
import torch

from . import cluster_loss, composite, hashgrid, march, mlp


class CpuField:
    def __init__(self, scale=0.5, log2_T=19, seed=1337):
        L, F, N_min = 16, 2, 16
        self.scale = scale
        self.b = float(np.exp(np.log(2048 * scale / N_min) / (L - 1)))
        self.levels, total = hashgrid.grid_levels(L, F, log2_T, N_min, self.b)
        g = torch.Generator().manual_seed(seed)
        self.table = ((torch.rand(total * F, generator=g) * 2 - 1) * 1e-4).requires_grad_(True)
        def xavier(shapes):
            return torch.cat([((torch.rand(o, i, generator=g) * 2 - 1) * (6.0 / (i + o)) ** 0.5).reshape(-1) for o, i in shapes])
        self.sigma_w = xavier([(64, 32), (16, 64)]).requires_grad_(True)
        self.rgb_w = xavier([(64, 32), (64, 64), (16, 64)]).requires_grad_(True)

    def params(self):
        return [self.table, self.sigma_w, self.rgb_w]

    def forward(self, xyzs, dirs):
        x = (xyzs + self.scale) / (2 * self.scale)
        feat = hashgrid.forward(x, self.table.view(-1, 2), self.levels, out_dtype=None)
        h = mlp.forward(feat, self.sigma_w, 32, 16, 1, "None", emulate_half=False)
        sigmas = torch.exp(h[:, 0])
        d = dirs / torch.norm(dirs, dim=1, keepdim=True)
        rgbs = mlp.forward(torch.cat([d, h], 1), self.rgb_w, 19, 3, 2, "Sigmoid", emulate_half=False)
        return sigmas, rgbs


def train_step(field, bitfield, rays_o, rays_d, target_rgb, tri, step=3000, hp=None, opt_state=None, lr=1e-2, noise=None, centroids=None):
    """one CPU training step on numpy rays; returns (loss dict, n_samples).  `centroids` (K,3): skip the k-means stand-in and
    assign every valid normal to its most similar given centroid (what faiss' index.search does, losses.py:90) - lets a test
    feed the centroids another engine found, so that everything downstream is compared on equal footing."""
    hp = hp or {}
    R = len(rays_o)
    hits = march.aabb(rays_o, rays_d, [0, 0, 0], [field.scale] * 3, hp.get("near", 0.01))
    if noise is None:
        noise = np.random.RandomState(step).rand(R).astype(np.float32)
    ra, xyzs, dirs, deltas, ts = march.march_train(rays_o, rays_d, hits, bitfield, 1, field.scale, 0.0, noise, 128, 1024)
    t = torch.from_numpy
    sigmas, rgbs = field.forward(t(xyzs), t(dirs))
    total, opacity, depth, rend, ws = composite.composite_train(sigmas, rgbs, t(deltas), t(ts), t(ra), 1e-4)
    rgb = rend + (1 - opacity)[:, None]
    loss = {"rgb": ((rgb - target_rgb) ** 2).mean()}
    o = opacity + 1e-10
    loss["opacity"] = hp.get("opacity_w", 1e-3) * (-o * torch.log(o)).mean()
    rd = t(rays_d)
    x123 = {k: t(tri[i]) for i, k in enumerate(("x1", "x2", "x3"))}
    normals = cluster_loss.normals_from_rays(rd, rd, depth, x123)
    valid = cluster_loss.valid_rows(normals.detach())
    if centroids is None:
        cent, assign = cluster_loss.spherical_kmeans(normals.detach()[valid].numpy(), 20, 20)
    else:
        cent = np.ascontiguousarray(centroids, dtype=np.float32)
        assign = (normals.detach()[valid].numpy() @ cent.T).argmax(1).astype(np.int64)
    labels, _ = cluster_loss.select_clusters(t(assign), t(cent), 1.0 - 0.01)
    ort, dot, l1 = cluster_loss.cluster_terms(normals[valid], labels)
    w = max(0.0, min(2e-3, (step - 500) * (2e-3 / 2500)))
    for k, v in (("norm_D_C_ort_dot", ort), ("norm_D_C_centr_dot", dot), ("norm_D_C_centr_L1", l1)):
        loss[k] = w * torch.nan_to_num(v)
    loss["total"] = sum(loss.values())
    for p in field.params():
        p.grad = None
    loss["total"].backward()
    # Adam (adam_w_mode, eps 1e-15; weight decay 0 on the table, 1e-6 on the nets) after clip_grad_norm_(0.05)
    if opt_state is not None:
        opt_state["t"] = opt_state.get("t", 0) + 1
        gn = torch.sqrt(sum((p.grad ** 2).sum() for p in field.params()))
        coef = min(1.0, 0.05 / (float(gn) + 1e-6))
        with torch.no_grad():
            for i, p in enumerate(field.params()):
                g = p.grad * coef
                m = opt_state.setdefault(("m", i), torch.zeros_like(p)); v = opt_state.setdefault(("v", i), torch.zeros_like(p))
                m.mul_(0.9).add_(g, alpha=0.1); v.mul_(0.999).addcmul_(g, g, value=0.001)
                bc1, bc2 = 1 - 0.9 ** opt_state["t"], 1 - 0.999 ** opt_state["t"]
                upd = (m / bc1) / ((v / bc2).sqrt() + 1e-15) + (0.0 if i == 0 else 1e-6) * p
                p.add_(upd, alpha=-lr)
    return loss, len(ts)
