"""Training step either side of the hot path: the part of
NeRFSystem.training_step (train_nerf.py:314-367) + configure_optimizers (:262-291) + Lightning's AMP / clip /
DDP plumbing (:949-955) that the path needs in order to be measured as a whole:

  * parameters live in ONE flat fp32 buffer (hash table first, then the MLPs) with ONE flat fp32 gradient
    buffer; module ``params`` are views, so state-dict keys stay those of the reference;
  * optimizer = apex FusedAdam semantics (adam_w_mode, eps 1e-15, weight decay 0 for ``xyz_encoder`` / 1e-6 for the
    nets), static loss scale with non-finite skip (GradScaler), clip_grad_norm_(0.05), cosine LR per epoch - as
    three streaming kernels (sum of squares, clip coefficient, fused Adam + fp16 refresh + gradient zero);
  * data parallel: one process per GPU, one ncclAllReduce(sum) of the flat gradient on the compute stream
    (libncn's NCCL binding) - torch.distributed is only used to ship the 128-byte NCCL id and for barriers;
  * occupancy-grid upkeep every 16 steps (warm-up 256 steps), sync-free.
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import check, ptr, stream
from .losses import NeRFMTLoss
from .ngp import NGPMT
from .rendering import render


def default_hparams(**over):
    """paper defaults that fix hot-path shapes (experiments/hypersim/hyperparameters.py:20-65, opt.py)"""
    hp = dict(scale=0.5, grid_size=128, rend_max_samples=1024, rend_near_dist=0.01, batch_size=8192, lr=1e-2,
              num_epochs=30, steps_per_epoch=1000, density_tresh_decay=1.0, update_interval=16, warmup_steps=256,
              grad_clip=0.05, loss_scale=1024.0,
              ray_sampling_strategy="all_images_triang_patch", pred_norm_depth=True, pred_norm_nn=False, pred_sem=False,
              loss_opacity_w=1e-3, loss_distortion_w=0, loss_depth_w=0, loss_sem_w=0, loss_norm_can_tres=0.01,
              loss_norm_D_C_ort_dot_w=2e-3, loss_norm_D_C_centr_dot_w=2e-3, loss_norm_D_C_centr_L1_w=2e-3,
              loss_norm_can_start=500, loss_norm_can_grow=2500, loss_norm_can_end=-1, exp_step_factor=0.0)
    hp.update(over)
    return hp


class Communicator:
    """libncn NCCL communicator (one per process); None-op at world_size 1."""

    def __init__(self, rank=0, world_size=1):
        self.rank, self.world_size, self.handle = rank, world_size, None
        if world_size > 1:
            import torch.distributed as dist
            L = _lib.lib()
            buf = (C.c_ubyte * 128)()
            if rank == 0:
                check(L.ncn_comm_unique_id(buf), "comm_unique_id")
            t = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, 0)
            raw = bytes(t.cpu().tolist())
            h = C.c_void_p()
            rc = L.ncn_comm_init(C.byref(h), raw, world_size, rank)
            if rc != 0:
                raise RuntimeError("ncn_comm_init: " + L.ncn_comm_last_error().decode())
            self.handle = h

    def allreduce_sum_(self, flat):
        if self.handle is not None:
            L = _lib.lib()
            rc = L.ncn_comm_allreduce_sum_f32(self.handle, ptr(flat), flat.numel(), stream())
            if rc != 0:
                raise RuntimeError("ncn_comm_allreduce: " + L.ncn_comm_last_error().decode())

    def close(self):
        if self.handle is not None:
            _lib.lib().ncn_comm_destroy(self.handle)
            self.handle = None


class FlatAdam:
    """Flat-buffer Adam over [(name, parameter, weight_decay)] groups (apex FusedAdam adam_w_mode semantics)."""

    def __init__(self, named_params, lr, eps=1e-15, betas=(0.9, 0.999), loss_scale=1.0, grad_clip=0.05, world_size=1):
        named_params = [(n, p) for n, p in named_params if p.numel() > 0]
        dev = named_params[0][1].device
        enc = [(n, p) for n, p in named_params if "xyz_encoder" in n]          # train_nerf.py:264-274
        net = [(n, p) for n, p in named_params if "xyz_encoder" not in n]
        self.groups = []
        total = sum((p.numel() + 3) // 4 * 4 for _, p in enc + net)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for grp, wd in ((enc, 0.0), (net, 1e-6)):
            start = off
            for n, p in grp:
                k = p.numel()
                self.flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + k].view_as(p.data)
                p.grad = self.grad[off:off + k].view_as(p.data)
                off += (k + 3) // 4 * 4
            if off > start:
                self.groups.append((start, off - start, wd))
        # fp16 working copy of every parameter, refreshed by the Adam kernel itself (the tcnn binding re-casts the
        # whole 11.4 M-entry table on every forward call)
        self.flat16 = self.flat.to(torch.float16)
        self.lr, self.eps, self.betas = lr, eps, betas
        self.loss_scale, self.grad_clip, self.world_size = loss_scale, grad_clip, world_size
        self.step_count = 0
        self.grad_div = torch.tensor([loss_scale * world_size], dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.coef = torch.ones(1, dtype=torch.float32, device=dev)

    def adopt_half_copies(self, model):
        """point every tcnn-style module's fp16 working copy at this optimizer's flat16 buffer"""
        base = self.flat.data_ptr()
        for mod in model.modules():
            p = getattr(mod, "params", None)
            if p is None or p.numel() == 0 or not hasattr(mod, "_half_key"):
                continue
            o = (p.data_ptr() - base) // 4
            mod._half = self.flat16[o:o + p.numel()]
            mod._half_key = "flat"
            mod._flat_version = p._version        # in sync now; a later user-side in-place write bumps the version

    def step(self, lr=None):
        L = _lib.lib()
        st = stream()
        self.step_count += 1
        self.sumsq.zero_(); self.flag.zero_()
        check(L.ncn_grad_sumsq(ptr(self.grad), self.grad.numel(), ptr(self.grad_div), ptr(self.sumsq), ptr(self.flag), st), "grad_sumsq")
        coef = None
        if self.grad_clip and self.grad_clip > 0:
            check(L.ncn_clip_coef(ptr(self.sumsq), float(self.grad_clip), ptr(self.coef), st), "clip_coef")
            coef = self.coef
        for (start, n, wd) in self.groups:
            sl = slice(start, start + n)
            check(L.ncn_adam_step(ptr(self.flat[sl]), ptr(self.grad[sl]), ptr(self.m[sl]), ptr(self.v[sl]), ptr(self.flat16[sl]), n,
                                  float(lr if lr is not None else self.lr), self.betas[0], self.betas[1], self.eps, wd,
                                  self.step_count, ptr(self.grad_div), ptr(self.flag), ptr(coef), None, st), "adam_step")


class NeRFTrainer:
    """One process = one GPU = one ray shard.  ``train_step(batch)`` is the unit bench.py times."""

    def __init__(self, hparams=None, device="cuda", rank=0, world_size=1, seed=0, n_sem_cls=0, log2_T=19):
        self.hp = hp = default_hparams(**(hparams or {}))
        self.device = torch.device(device)
        self.rank, self.world_size = rank, world_size
        kw = {"n_sem_cls": n_sem_cls} if hp["pred_sem"] else {}
        self.model = NGPMT(scale=hp["scale"], grid_size=hp["grid_size"], rgb_act="Sigmoid", pred_sem=hp["pred_sem"],
                           pred_norm=hp["pred_norm_nn"], log2_T=log2_T, **kw).to(self.device)
        self.loss = NeRFMTLoss(hp)
        self.comm = Communicator(rank, world_size)
        self.opt = FlatAdam(list(self.model.named_parameters()), lr=hp["lr"], loss_scale=hp["loss_scale"],
                            grad_clip=hp["grad_clip"], world_size=world_size)
        self.opt.adopt_half_copies(self.model)
        self.global_step = 0
        self.fused = None
        self.render_kwargs = dict(near_distance=hp["rend_near_dist"], max_samples=hp["rend_max_samples"],
                                  exp_step_factor=hp["exp_step_factor"], n_sem_cls=n_sem_cls,
                                  pred_norm_nn_norm=False)
        self.poses = None
        self.directions = None

    # dataset tensors that NeRFSystem keeps on the device (train_nerf.py:239-240)
    def set_cameras(self, poses, directions):
        self.poses = torch.as_tensor(poses, dtype=torch.float32, device=self.device)
        self.directions = torch.as_tensor(directions, dtype=torch.float32, device=self.device)

    def rays_from_batch(self, img_idxs, pix_idxs):
        """get_rays (datasets/ray_utils.py:46-71) on the device: rays_d = directions @ R^T, rays_o = t"""
        c2w = self.poses[img_idxs]
        d = self.directions[pix_idxs]
        rays_d = (d[:, None, :] * c2w[:, :, :3]).sum(-1)
        rays_o = c2w[:, :, 3]
        return rays_o.contiguous(), rays_d.contiguous()

    def lr_now(self):
        epoch = self.global_step // self.hp["steps_per_epoch"]
        return 0.5 * self.hp["lr"] * (1 + math.cos(math.pi * epoch / self.hp["num_epochs"]))

    def maybe_update_grid(self):
        hp = self.hp
        if self.global_step % hp["update_interval"] == 0:
            thr = 0.01 * hp["rend_max_samples"] / 3 ** 0.5 * hp["density_tresh_decay"]
            self.model.update_density_grid(thr, warmup=self.global_step < hp["warmup_steps"])

    def forward_loss(self, rays_o, rays_d, target):
        results = render(self.model, rays_o, rays_d, global_step=self.global_step, **self.render_kwargs)
        loss_d = self.loss(results, target, global_step=self.global_step)
        return results, loss_d

    def fused_step(self, capacity_per_ray=64, use_graph=True, fuse_fwd="mlp"):
        """the sync-free CUDA-graph step (ncn_b200.fused.FusedStep)"""
        if self.fused is None:
            from .fused import FusedStep
            self.fused = FusedStep(self, capacity_per_ray=capacity_per_ray, use_graph=use_graph, fuse_fwd=fuse_fwd)
            self.opt.grad_div.fill_(float(self.world_size))     # fused gradients are already un-scaled
        return self.fused

    def train_step_fused(self, rays_o=None, rays_d=None, target_rgb=None, tri=None, update_grid=True, noise=None, packed=None,
                         grid_restore=None, sem_target=None):
        """same step as train_step, through the fused path; returns nothing (stats via self.fused.stats_host())"""
        fs = self.fused_step()
        if tri is not None and (fs.tri is None or fs.tri.data_ptr() != tri.data_ptr()):
            fs.set_triangles(tri)
        if update_grid and self.global_step % self.hp["update_interval"] == 0:
            fs.flush()                            # the grid update evaluates the field: it must see the latest parameters
            if self.global_step < self.hp["warmup_steps"]:
                self.maybe_update_grid()
            else:
                fs.update_grid(restore=grid_restore)
        fs.step(rays_o, rays_d, target_rgb, noise=noise, packed=packed, sem_target=sem_target)

    def train_step_from_pixels(self, img_idx, pix_idx, target_rgb, update_grid=True, grid_restore=None):
        """public end-to-end entry: (image, pixel) indices + target colours on the device -> one fused training step"""
        fs = self.fused_step()
        fs.rays_from_pixels(img_idx, pix_idx)
        fs.target.copy_(target_rgb)
        self.train_step_fused(update_grid=update_grid, grid_restore=grid_restore)

    def train_step(self, rays_o, rays_d, target, update_grid=True):
        if update_grid:
            self.maybe_update_grid()
        results, loss_d = self.forward_loss(rays_o, rays_d, target)
        (loss_d["total"] * self.hp["loss_scale"]).backward()
        self.comm.allreduce_sum_(self.opt.grad)
        self.opt.step(self.lr_now())
        self.global_step += 1
        return results, loss_d
