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


class _DeviceMemory:
    """__cuda_array_interface__ view of memory owned by libncn (torch.as_tensor wraps it without a copy)"""

    def __init__(self, address, n, typestr, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(address), False), "version": 2,
                                         "strides": None}


class PeerLink:
    """Sharded data-parallel optimizer over NVLink peer memory (include/ncn.h ncn_peer_*): this rank's gradient and fp16
    parameter buffers are allocated by libncn, exported with CUDA IPC and mapped by every other rank; ``step`` replaces
    all-reduce + ||g||^2 + Adam with two kernels (no NCCL in the step, so it stays one CUDA graph).
    torch.distributed only ships the 192-byte handle blocks at start-up."""

    def __init__(self, rank, world_size, n_params, device):
        L = _lib.lib()
        self.rank, self.world_size, self.n = rank, world_size, int(n_params)
        self.handle = None
        self.external_zero = False
        import os
        if os.environ.get("NCN_PEER_LOADS"):       # developer A/B only
            L.ncn_peer_set_loads(int(os.environ["NCN_PEER_LOADS"]))
        if os.environ.get("NCN_PEER_EARLY_LOADS"):
            L.ncn_peer_set_early_loads(int(os.environ["NCN_PEER_EARLY_LOADS"]))
        if os.environ.get("NCN_PEER_BATCH") or os.environ.get("NCN_PEER_CTAS"):
            L.ncn_peer_set_shape(int(os.environ.get("NCN_PEER_BATCH", "4")), int(os.environ.get("NCN_PEER_CTAS", "2")))
        h = C.c_void_p()
        if world_size == 1:
            check(L.ncn_peer_create(C.byref(h), rank, world_size, self.n), "peer_create")
            self.handle = h
        else:
            # every phase ends with an agreement (MIN all-reduce) so that a failure on one rank raises on ALL ranks instead of
            # leaving the others inside a collective - the caller then falls back to the NCCL exchange everywhere
            import torch.distributed as dist
            on_dev = dist.get_backend() == "nccl"

            def agree(ok, what):
                t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device if on_dev else "cpu")
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                if int(t.item()) == 0:
                    self.close()
                    raise RuntimeError(f"PeerLink: {what} failed on at least one rank")

            buf = (C.c_ubyte * 192)()
            ok = L.ncn_peer_create(C.byref(h), rank, world_size, self.n) == 0
            if ok:
                self.handle = h
                ok = L.ncn_peer_handles(h, buf) == 0
            agree(ok, "buffer allocation / cudaIpcGetMemHandle")
            mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
            if on_dev:
                mine = mine.to(device)
            every = [torch.empty_like(mine) for _ in range(world_size)]
            dist.all_gather(every, mine)
            raw = b"".join(bytes(t.cpu().tolist()) for t in every)
            agree(L.ncn_peer_connect(h, raw) == 0, "cudaIpcOpenMemHandle")
        self.grad = torch.as_tensor(_DeviceMemory(L.ncn_peer_grad(h), self.n, "<f4", self), device=device)
        self.p16 = torch.as_tensor(_DeviceMemory(L.ncn_peer_p16(h), self.n, "<f2", self), device=device)
        self.cut = self.n
        self.shard = self.segments(rank)[0]

    def segments(self, q):
        """rank q's shard of the flat vector: [(lo, hi) of the late range, (lo, hi) of the early range (empty without a cut)]"""
        seg = (C.c_int64 * 4)()
        check(_lib.lib().ncn_peer_segments(self.handle, int(q), seg), "peer_segments")
        return [(seg[0], seg[1]), (seg[2], seg[3])]

    def set_cut(self, cut):
        """[cut, n) becomes the early range (include/ncn.h ncn_peer_set_cut); same value on every rank"""
        check(_lib.lib().ncn_peer_set_cut(self.handle, int(cut)), "peer_set_cut")
        self.cut = int(cut)
        self.shard = self.segments(self.rank)[0]

    def early(self, grad_div, st):
        """reduce this rank's slice of the early range now (its gradient is complete on stream `st`)"""
        check(_lib.lib().ncn_peer_early(self.handle, ptr(grad_div), st), "peer_early")

    def step(self, flat, m, v, groups, betas, eps, grad_div, flag, lr_bc, sumsq_out, st):
        check(_lib.lib().ncn_peer_step(self.handle, ptr(flat), ptr(m), ptr(v), C.byref(groups), betas[0], betas[1], eps, ptr(grad_div),
                                       ptr(flag), ptr(lr_bc), ptr(sumsq_out), st), "peer_step")

    def set_external_zero(self, on=True):
        """the caller zeroes the gradient buffer after every step() itself (off the exchange's critical path)"""
        check(_lib.lib().ncn_peer_set_external_zero(self.handle, 1 if on else 0), "peer_set_external_zero")
        self.external_zero = bool(on)

    def poll(self):
        """error word from mapped host memory, no device synchronisation (0 = ok, 1 + phase of the wait that timed out)"""
        return int(_lib.lib().ncn_peer_poll(self.handle)) if self.handle is not None else 0

    def error(self):
        e = C.c_uint32()
        check(_lib.lib().ncn_peer_error(self.handle, C.byref(e)), "peer_error")
        return e.value

    def close(self):
        if self.handle is not None:
            self.grad = self.p16 = None
            _lib.lib().ncn_peer_destroy(self.handle)
            self.handle = None


class FlatAdam:
    """Flat-buffer Adam over [(name, parameter, weight_decay)] groups (apex FusedAdam adam_w_mode semantics)."""

    def __init__(self, named_params, lr, eps=1e-15, betas=(0.9, 0.999), loss_scale=1.0, grad_clip=0.05, world_size=1, rank=0,
                 shard=False):
        named_params = [(n, p) for n, p in named_params if p.numel() > 0]
        dev = named_params[0][1].device
        enc = [(n, p) for n, p in named_params if "xyz_encoder" in n]          # train_nerf.py:264-274
        net = [(n, p) for n, p in named_params if "xyz_encoder" not in n]
        self.groups = []
        total = sum((p.numel() + 3) // 4 * 4 for _, p in enc + net)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        # shard=True: gradient + fp16 parameter buffers live in IPC-shared memory and the update runs sharded over NVLink peer
        # memory (PeerLink); the fp32 master / m / v are then current only inside this rank's slice (self.peer.shard)
        self.peer = PeerLink(rank, world_size, total, dev) if shard else None
        self.grad = self.peer.grad if shard else torch.zeros(total, dtype=torch.float32, device=dev)
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
        if shard:
            self.flat16 = self.peer.p16
            self.flat16.copy_(self.flat)
        else:
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
        if self.peer is not None:
            t = self.step_count
            lr_bc = torch.tensor([float(lr if lr is not None else self.lr), 1 - self.betas[0] ** t, 1 - self.betas[1] ** t],
                                 dtype=torch.float32, device=self.flat.device)
            groups = _lib.AdamGroups()
            groups.n_groups = len(self.groups)
            for q, (start, _, wd) in enumerate(self.groups):
                groups.start[q] = start
                groups.weight_decay[q] = wd
            groups.max_norm = float(self.grad_clip or 0.0)
            self.peer.step(self.flat, self.m, self.v, groups, self.betas, self.eps, self.grad_div, self.flag, lr_bc, self.sumsq, st)
            return
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

    def __init__(self, hparams=None, device="cuda", rank=0, world_size=1, seed=0, n_sem_cls=0, log2_T=19, shard_optimizer=None):
        """shard_optimizer: None = sharded peer-memory optimizer whenever world_size > 1 (falls back to the NCCL all-reduce +
        replicated Adam when CUDA IPC / peer access is not available on every rank); True / False force it."""
        self.hp = hp = default_hparams(**(hparams or {}))
        self.device = torch.device(device)
        self.rank, self.world_size = rank, world_size
        kw = {"n_sem_cls": n_sem_cls} if hp["pred_sem"] else {}
        self.model = NGPMT(scale=hp["scale"], grid_size=hp["grid_size"], rgb_act="Sigmoid", pred_sem=hp["pred_sem"],
                           pred_norm=hp["pred_norm_nn"], log2_T=log2_T, **kw).to(self.device)
        self.loss = NeRFMTLoss(hp)
        self.comm = Communicator(rank, world_size)
        shard = world_size > 1 if shard_optimizer is None else bool(shard_optimizer)
        if shard and world_size > 1:
            shard = self._peer_access_everywhere()
        mk = lambda sh: FlatAdam(list(self.model.named_parameters()), lr=hp["lr"], loss_scale=hp["loss_scale"],
                                 grad_clip=hp["grad_clip"], world_size=world_size, rank=rank, shard=sh)
        try:
            self.opt = mk(shard)
        except RuntimeError as e:
            if not (shard and shard_optimizer is None and "PeerLink" in str(e)):
                raise
            import sys
            print(f"[ncn_b200] rank {rank}: {e}; using the NCCL all-reduce exchange", file=sys.stderr)
            self.opt = mk(False)
        self.peer = self.opt.peer
        self.opt.adopt_half_copies(self.model)
        # sharded optimizer on W > 1 ranks: the fp32 master / m / v outside this rank's slice go stale with the first step.
        # state_dict() (checkpoints) rebuilds them from their owners first - a collective, like the checkpoint itself.
        self.master_stale = False
        if self.peer is not None and world_size > 1:
            self.model.register_state_dict_pre_hook(lambda *a, **k: self.gather_master_params())
        self.global_step = 0
        self.fused = None
        self.render_kwargs = dict(near_distance=hp["rend_near_dist"], max_samples=hp["rend_max_samples"],
                                  exp_step_factor=hp["exp_step_factor"], n_sem_cls=n_sem_cls,
                                  pred_norm_nn_norm=False)
        self.poses = None
        self.directions = None

    def _peer_access_everywhere(self):
        """every rank can map every other rank's device (one GPU per rank on one NVSwitch box); agreed on by all ranks"""
        import torch.distributed as dist
        ok = 1
        try:
            me = self.device.index if self.device.index is not None else torch.cuda.current_device()
            n_dev = torch.cuda.device_count()
            ok = int(n_dev >= self.world_size and all(torch.cuda.can_device_access_peer(me, q) for q in range(self.world_size) if q != me))
        except Exception:  # noqa: BLE001
            ok = 0
        t = torch.tensor([ok], dtype=torch.int32, device=self.device if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))

    def gather_master_params(self):
        """sharded optimizer: make the fp32 master copy complete on every rank (checkpointing / evaluation of fp32 state);
        each rank owns opt.peer.shard, the rest is rebuilt from the owners"""
        if self.peer is None or self.world_size == 1:
            return
        import torch.distributed as dist
        if self.fused is not None:
            self.fused.flush()
        self.master_stale = False
        L = _lib.lib()
        for q in range(self.world_size):
            for lo, hi in self.peer.segments(q):
                if hi > lo:
                    for buf in (self.opt.flat, self.opt.m, self.opt.v):
                        dist.broadcast(buf[lo:hi], q)

    # dataset tensors that NeRFSystem keeps on the device (train_nerf.py:239-240)
    def set_cameras(self, poses, directions, random_poses=None):
        """poses (P,3,4), directions (H*W,3); random_poses (Q,3,4) = the generated poses of --random_tr_poses
        (ncn_b200.batches.generate_random_poses; NeRFSystem keeps them in a buffer of their own, train_nerf.py:241-242) are
        appended behind the training poses, so one pose table serves both halves of a batch: image index P + q = generated pose q"""
        poses = torch.as_tensor(poses, dtype=torch.float32, device=self.device)[:, :3, :]
        self.n_train_poses, self.n_random_poses = int(poses.shape[0]), 0
        if random_poses is not None:
            rp = torch.as_tensor(random_poses, dtype=torch.float32, device=self.device)[:, :3, :]
            self.n_random_poses = int(rp.shape[0])
            poses = torch.cat([poses, rp], 0)
        self.poses = poses.contiguous()
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

    def fused_step(self, capacity_per_ray=None, use_graph=True, fuse_fwd="mlp"):
        """the sync-free CUDA-graph step (ncn_b200.fused.FusedStep).  It carries its own gradient divisor (its gradients are
        already un-scaled); FlatAdam's divisor (loss_scale * world_size) stays what the module path's train_step needs, so the
        two paths can be mixed on one trainer."""
        if self.fused is None:
            from .fused import FusedStep
            self.fused = FusedStep(self, capacity_per_ray=capacity_per_ray, use_graph=use_graph, fuse_fwd=fuse_fwd)
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
        fs.target[:target_rgb.shape[0]].copy_(target_rgb)      # random_tr_poses: colours exist for the first half of the rays only
        self.train_step_fused(update_grid=update_grid, grid_restore=grid_restore)

    def train_step(self, rays_o, rays_d, target, update_grid=True):
        if self.fused is not None:
            self.fused.flush()                    # a deferred fused update must land before this path reads the parameters
        if update_grid:
            self.maybe_update_grid()
        results, loss_d = self.forward_loss(rays_o, rays_d, target)
        (loss_d["total"] * self.hp["loss_scale"]).backward()
        if self.peer is None:
            self.comm.allreduce_sum_(self.opt.grad)
        self.opt.step(self.lr_now())
        if self.peer is not None and self.world_size > 1:
            self.master_stale = True
        self.global_step += 1
        return results, loss_d
