"""NGPMT - host-side mirror of the reference model (models/ngp_mt.py:10-368): same constructor, buffers,
parameter names (``xyz_encoder.params``, ``sigma_net.params``, ``rgb_net.params``, ``sem_net.params``,
``norm_net.params``), ``density`` / ``forward`` / ``update_density_grid`` / ``mark_invisible_cells`` surface,
built on the sm_100a hash-grid / fused-MLP / occupancy kernels.

What differs from the reference (all host-side, results equivalent):
  * ``density_grid`` and ``grid_coords`` are registered by the model itself (the reference's trainer does it,
    train_nerf.py:153-157);
  * ``update_density_grid`` never synchronises: the occupied-cell sampler uses a cumulative count +
    searchsorted instead of ``nonzero`` (ngp_mt.py:262), and the decay/max update, the mean of the positive
    cells and packbits run as two fused kernels reading the threshold on the device (no ``.item()``, :365).
"""
import math

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib, vren
from . import tinycudann as tcnn
from ._lib import check, ptr, stream
from .custom_functions import TruncExp


class NGPMT(nn.Module):
    def __init__(self, scale, grid_size, rgb_act="Sigmoid", pred_sem=False, pred_norm=False, log2_T=19, **kwargs):
        super().__init__()
        self.pred_sem, self.pred_norm, self.rgb_act = pred_sem, pred_norm, rgb_act
        self.scale = scale
        self.register_buffer("center", torch.zeros(1, 3))
        self.register_buffer("xyz_min", -torch.ones(1, 3) * scale)
        self.register_buffer("xyz_max", torch.ones(1, 3) * scale)
        self.register_buffer("half_size", (self.xyz_max - self.xyz_min) / 2)
        self.cascades = max(1 + int(np.ceil(np.log2(2 * scale))), 1)       # ngp_mt.py:34
        self.grid_size = grid_size
        G = grid_size
        self.register_buffer("density_bitfield", torch.zeros(self.cascades * G ** 3 // 8, dtype=torch.uint8))
        self.register_buffer("density_grid", torch.zeros(self.cascades, G ** 3))
        ax = torch.arange(G, dtype=torch.int32)
        # kornia create_meshgrid3d(G,G,G,False) ordering (train_nerf.py:156): (d,h,w) grid of (x,y,z)=(w,h,d)
        zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")
        self.register_buffer("grid_coords", torch.stack([xx, yy, zz], -1).reshape(-1, 3).contiguous())

        L, F, N_min = 16, 2, 16
        b = float(np.exp(np.log(2048 * scale / N_min) / (L - 1)))          # ngp_mt.py:40-41
        self.xyz_encoder = tcnn.Encoding(3, dict(otype="Grid", type="Hash", n_levels=L, n_features_per_level=F,
                                                 log2_hashmap_size=log2_T, base_resolution=N_min, per_level_scale=b,
                                                 interpolation="Linear"))
        mlp = lambda n_hidden, act: dict(otype="FullyFusedMLP", activation="ReLU", output_activation=act,
                                         n_neurons=64, n_hidden_layers=n_hidden)
        self.sigma_net = tcnn.Network(L * F, 16, mlp(1, "None"))
        self.dir_encoder = tcnn.Encoding(3, dict(otype="SphericalHarmonics", degree=4))   # built, never called (:94-101,208)
        self.rgb_net = tcnn.Network(19, 3, mlp(2, rgb_act))
        if pred_sem:
            self.sem_net = tcnn.Network(16, kwargs["n_sem_cls"], mlp(2, "None"))
        if pred_norm:
            self.norm_net = tcnn.Network(16, 3, mlp(2, "None"))
        if rgb_act == "None":
            for i in range(3):
                setattr(self, f"tonemapper_net_{i}", tcnn.Network(1, 1, mlp(1, "Sigmoid")))

    # ------------------------------------------------------------------ field
    def density(self, x, return_feat=False):
        x = (x - self.xyz_min) / (self.xyz_max - self.xyz_min)
        h = self.sigma_net(self.xyz_encoder(x))
        sigmas = TruncExp.apply(h[:, 0])
        return (sigmas, h) if return_feat else sigmas

    def log_radiance_to_rgb(self, log_radiances, **kwargs):
        log_exposure = torch.log(kwargs["exposure"]) if "exposure" in kwargs else 0
        return torch.cat([getattr(self, f"tonemapper_net_{i}")(log_radiances[:, i:i + 1] + log_exposure) for i in range(3)], 1)

    def forward(self, x, d, **kwargs):
        sigmas, h = self.density(x, return_feat=True)
        d = d / torch.norm(d, dim=1, keepdim=True)
        rgbs = self.rgb_net(torch.cat([d, h], 1))
        if self.rgb_act == "None":
            rgbs = TruncExp.apply(rgbs) if kwargs.get("output_radiance", False) else self.log_radiance_to_rgb(rgbs, **kwargs)
        out = {"sigmas": sigmas, "rgbs": rgbs}
        if self.pred_sem:
            out["sems"] = self.sem_net(h)
        if self.pred_norm:
            out["norms"] = self.norm_net(h)
        return out

    @torch.no_grad()
    def field_eval(self, x, d):
        """Inference form of forward() for the evaluation renderer: hash-grid encode (input
        normalisation fused) + ONE launch for density trunk, TruncExp, [h | d/|d| | 1] and colour head
        (ncn_field_mlp_fwd), then the optional semantic / normal heads; nothing is saved for a backward.
        x, d (N,3) f32 -> sigmas (N) f32, raws (N, 3 [+3] [+n_cls]) f32 in the channel order of rendering.py:203-208."""
        if self.rgb_act != "Sigmoid":
            raise NotImplementedError("field_eval covers the sigmoid colour head (ngp_mt with rgb_act='Sigmoid')")
        from .tinycudann import _half_copy
        L = _lib.lib()
        st = stream()
        n = x.shape[0]
        dev = x.device
        x = x.float().contiguous(); d = d.float().contiguous()
        n_cls = self.sem_net.n_output_dims if self.pred_sem else 0
        Ct = 3 + (3 if self.pred_norm else 0) + n_cls
        f16 = dict(dtype=torch.float16, device=dev)
        feat, h = torch.empty(n, 32, **f16), torch.empty(n, 16, **f16)
        sigmas, raws = torch.empty(n, dtype=torch.float32, device=dev), torch.empty(n, Ct, dtype=torch.float32, device=dev)
        if n == 0:
            return sigmas, raws
        xform = (C.c_float * 6)(*([float(self.xyz_min[0, i]) for i in range(3)] + [float((self.xyz_max - self.xyz_min)[0, i]) for i in range(3)])) \
            if getattr(self, "_xform", None) is None else self._xform
        self._xform = xform
        check(L.ncn_grid_fwd(C.byref(self.xyz_encoder.desc), ptr(x), ptr(_half_copy(self.xyz_encoder)), n, ptr(feat), xform, None, st), "grid_fwd")
        check(L.ncn_field_mlp_fwd(ptr(feat), ptr(d), ptr(_half_copy(self.sigma_net)), ptr(_half_copy(self.rgb_net)), n, None, ptr(sigmas),
                                  ptr(raws), Ct, ptr(h), None, None, None, None, st), "field_mlp_fwd")
        if self.pred_norm or self.pred_sem:      # both extra heads in one launch, straight into their raws columns
            sem_off = 3 + (3 if self.pred_norm else 0)
            check(L.ncn_field_heads_fwd(ptr(h), n, None, ptr(raws), Ct,
                                        ptr(_half_copy(self.norm_net)) if self.pred_norm else None, 3, 3, None, None,
                                        ptr(_half_copy(self.sem_net)) if self.pred_sem else None, sem_off, max(n_cls, 1), None, None, st),
                  "field_heads_fwd")
        return sigmas, raws

    # ------------------------------------------------------------------ occupancy grid
    @torch.no_grad()
    def get_all_cells(self):
        indices = vren.morton3D(self.grid_coords).long()
        return [(indices, self.grid_coords)] * self.cascades

    @torch.no_grad()
    def sample_uniform_and_occupied_cells(self, M, density_threshold):
        cells = []
        G3 = self.grid_size ** 3
        dev = self.density_grid.device
        for c in range(self.cascades):
            coords1 = torch.randint(self.grid_size, (M, 3), dtype=torch.int32, device=dev)
            indices1 = vren.morton3D(coords1).long()
            occ = (self.density_grid[c] > density_threshold)
            csum = torch.cumsum(occ, 0, dtype=torch.int32)
            total = csum[-1]
            r = (torch.rand(M, device=dev) * total).to(torch.int32)
            indices2 = torch.searchsorted(csum, r, right=True).clamp_(max=G3 - 1)
            coords2 = vren.morton3D_invert(indices2.int())
            cells += [(torch.cat([indices1, indices2]), torch.cat([coords1, coords2]))]
        return cells

    @torch.no_grad()
    def mark_invisible_cells(self, K, dev, poses, img_wh, near_distance, chunk=64 ** 3):
        """density -1 for the cells no camera covers, once before training (ngp_mt.py:273-337; same argument list, including the
        `dev` the reference's caller passes, train_nerf.py:307-312).  K is either a (3,3) pinhole matrix or the Hypersim tuple
        (M_ndc_from_cam (4,4), M_uv_from_ndc, shift, scale) of the tilt-shift camera model (ngp_mt.py:290-296, 314-321)."""
        N_cams = poses.shape[0]
        self.count_grid = torch.zeros_like(self.density_grid)
        w2c_R = poses[:, :3, :3].transpose(1, 2)
        w2c_T = -w2c_R @ poses[:, :3, 3:]
        if isinstance(K, torch.Tensor):
            K = K.to(dev)
            project = None
        elif isinstance(K, tuple):
            M_ndc_from_cam, M_uv_from_ndc, scene_scale = K[0].to(dev), K[1].to(dev), K[3]
            project = (M_ndc_from_cam, M_uv_from_ndc, scene_scale)
        else:
            raise AssertionError("mark_invisible_cells: K must be a (3,3) tensor or the Hypersim projection tuple")
        cells = self.get_all_cells()
        for c in range(self.cascades):
            indices, coords = cells[c]
            s = min(2 ** (c - 1), self.scale)
            half_grid_size = s / self.grid_size
            for i in range(0, len(indices), chunk):
                sl = slice(i, i + chunk)
                xyzs_w = ((coords[sl] / (self.grid_size - 1) * 2 - 1) * (s - half_grid_size)).T            # (3, chunk)
                xyzs_c = w2c_R @ xyzs_w + w2c_T                                                           # (N_cams, 3, chunk)
                if project is None:
                    uvd = K @ xyzs_c
                    uv = uvd[:, :2] / uvd[:, 2:]
                else:       # back to metric scale, homogeneous clip -> ndc -> uv (depth = the uv matrix's third row)
                    xyzs_c *= 2 * project[2]
                    xyzs_h = torch.cat((xyzs_c, torch.ones_like(xyzs_c)[:, :1, :]), 1)
                    xyz_clip = project[0] @ xyzs_h
                    uvd = project[1] @ (xyz_clip / xyz_clip[:, 3:])
                    uv = uvd[:, :2]
                in_image = (uvd[:, 2] >= 0) & (uv[:, 0] >= 0) & (uv[:, 0] < img_wh[0]) & (uv[:, 1] >= 0) & (uv[:, 1] < img_wh[1])
                covered = (uvd[:, 2] >= near_distance) & in_image
                self.count_grid[c, indices[sl]] = count = covered.sum(0) / N_cams
                too_near = ((uvd[:, 2] < near_distance) & in_image).any(0)
                self.density_grid[c, indices[sl]] = torch.where((count > 0) & (~too_near), 0., -1.)

    @torch.no_grad()
    def update_density_grid(self, density_threshold, warmup=False, decay=0.95, erode=False):
        tmp = torch.zeros_like(self.density_grid)
        cells = self.get_all_cells() if warmup else \
            self.sample_uniform_and_occupied_cells(self.grid_size ** 3 // 4, density_threshold)
        for c in range(self.cascades):
            indices, coords = cells[c]
            s = min(2 ** (c - 1), self.scale)
            half_grid_size = s / self.grid_size
            xyzs_w = (coords / (self.grid_size - 1) * 2 - 1) * (s - half_grid_size)
            xyzs_w += (torch.rand_like(xyzs_w) * 2 - 1) * half_grid_size
            tmp[c, indices] = self.density(xyzs_w)
        if erode:
            raise NotImplementedError("erode (colmap datasets) is outside the hot-path scope")
        L = _lib.lib()
        stats = torch.zeros(2, dtype=torch.float32, device=tmp.device)
        check(L.ncn_density_grid_update(ptr(self.density_grid), ptr(tmp), self.density_grid.numel(), float(decay),
                                        ptr(stats), stream()), "density_grid_update")
        check(L.ncn_packbits_auto(ptr(self.density_grid), self.density_bitfield.numel(), ptr(stats),
                                  float(density_threshold), ptr(self.density_bitfield), stream()), "packbits_auto")
