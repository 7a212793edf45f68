"""FusedStep - the sync-free, CUDA-graph-captured training step (the B200-native form of
NeRFSystem.training_step, train_nerf.py:314-367, for the ngp_mt configurations: RGB+depth with the opacity and
Manhattan normal-clustering losses, optionally with the semantic head (+ cross-entropy, losses.py:569-573) and the
normal head (--pred_sem / --pred_norm_nn: extra compositing channels, rendering.py:203-224)).

Same arithmetic as the module path (ncn_b200.rendering.render + ncn_b200.losses.NeRFMTLoss + autograd +
FlatAdam), but expressed as one linear sequence of libncn kernels on pre-allocated arenas:

  aabb+near -> march (count, scan, expand into a capacity-sized arena; the sample count stays ON THE DEVICE) ->
  grid encode (input normalisation fused) -> sigma MLP -> [d/|d|, h, 1] + TruncExp -> rgb MLP -> raws ->
  composite fw -> photometric+opacity loss (+gradient) -> normals-from-depth -> k-means -> triple selection ->
  cluster loss (+gradient) -> normals bw -> composite bw -> rgb MLP bw -> dh -> sigma MLP bw -> grid bw ->
  [NCCL all-reduce] -> sum-of-squares -> clip coefficient -> fused Adam (x2 groups)
  (multi-rank default: the bracket and everything after it = ncn_peer_step, the sharded exchange over NVLink peer memory)

Optional nodes: the batch itself (use_device_sampling: indices + target gather) and ray generation (use_pixel_batches) in
front; the semantic / normal heads (ncn_field_heads_fwd, ncn_semantic_ce_loss, semantic-head backward) when the model has them.

No torch op, no autograd graph, no allocation and no host synchronisation happen inside the step, so it is
captured once into a CUDA graph and replayed (the reference's step has >= 6 host syncs).
Host-visible scalars (losses, sample counts) are read back lazily from a small stats buffer.
"""
import ctypes as C
import math
import os

import torch

from . import _lib
from ._lib import check, ptr

GSCALE = 131072.0      # 2^17: loss scale carried by the fp16 dL/dy tensors (the module path uses 1024 * 128)


class FusedStep:
    def __init__(self, trainer, capacity_per_ray=None, use_graph=True, fuse_fwd="mlp"):
        self.tr = trainer
        self.model = m = trainer.model
        self.opt = trainer.opt
        self.hp = trainer.hp
        # rendered channels (rendering.py:203-224): rgb | norm_nn (3) | sem (n_cls)
        self.n_cls = int(trainer.render_kwargs.get("n_sem_cls", 0)) if m.pred_sem else 0
        if m.pred_sem and not 1 <= self.n_cls <= 16:
            raise NotImplementedError("FusedStep: the semantic head is padded to 16 outputs (n_sem_cls <= 16)")
        self.norm_off = 3
        self.sem_off = 3 + (3 if m.pred_norm else 0)
        self.Ct = Ct = self.sem_off + self.n_cls
        self.sem_w = float(self.hp.get("loss_sem_w", 0.0)) if m.pred_sem else 0.0
        for k in ("loss_norm_D_C_can_dot_w", "loss_norm_D_C_can_L1_w", "loss_depth_w", "loss_distortion_w", "loss_reg_depth_w",
                  "loss_norm_depth_L1_w", "loss_norm_depth_dot_w"):
            if float(self.hp.get(k, 0) or 0) > 0:
                raise NotImplementedError(f"FusedStep: {k} > 0 is carried by the module path only (0 in every shipped experiment)")
        self.dev = dev = trainer.device
        self.R = R = self.hp["batch_size"]
        # --random_tr_poses (datasets/base.py:106-126, train_nerf.py:169-172, losses.py:265-297): the batch is [n_gt rays of training
        # views WITH a target colour | the same pixels seen from generated poses]; photometric / semantic terms on the first half,
        # the opacity term on every ray, the normal-clustering terms on the second half only
        self.rtp = bool(self.hp.get("random_tr_poses", False))
        if self.rtp and R % 2:
            raise RuntimeError("FusedStep: random_tr_poses needs an even batch_size (two halves of equal length)")
        self.n_gt = R // 2 if self.rtp else R
        self.u0 = self.n_gt if self.rtp else 0       # first ray of the set the geometric regularisers see (losses.py:286 unsup_start)
        self.use_graph = use_graph
        self.graph = None
        self.L = _lib.lib()
        f32 = dict(dtype=torch.float32, device=dev)
        E = lambda *s, **k: torch.empty(*s, **k)
        # static inputs: ONE (3,R,3) buffer [rays_o | rays_d | target rgb] so a step needs a single input copy
        self.inp = E(3, R, 3, **f32)
        self.rays_o, self.rays_d, self.target = self.inp[0], self.inp[1], self.inp[2]
        self.dev_sampling = None             # use_device_sampling(): the batch itself is drawn inside the step
        self.pix_inputs = False              # use_pixel_batches(): rays are generated inside the step from (image, pixel) indices
        self.noise = E(R, **f32)
        self.gen_noise = True                # march jitter drawn inside the step (torch.rand_like of custom_functions.py:83)
        self.tri = None                      # (3, M) int64
        # marching
        self.hits_t = E(R, 1, 2, **f32)
        self.rays_a = E(R, 3, dtype=torch.int64, device=dev)
        self.counter = torch.zeros(2, dtype=torch.int32, device=dev)
        ws_bytes = self.L.ncn_march_train_workspace_bytes(R, self.hp["rend_max_samples"])
        self.march_ws = E(ws_bytes, dtype=torch.uint8, device=dev)
        # sample arena.  The reference sizes its sample arrays exactly after a host sync (raymarching.cu:302-305, worst case
        # N_rays * max_samples right after the warm-up grid); this step never syncs, so the arena has a fixed capacity:
        #   capacity_per_ray=None  -> the worst case (max_samples per ray: overflow is impossible) whenever that costs < 30 % of the
        #                             device memory (12 GB for 8192 rays on a 180 GB B200), else 64 per ray
        #   an explicit number     -> that many rows per ray.
        # Whenever capacity < worst case the step is GUARDED: ncn_step_guard turns a step whose march overflowed into a skipped
        # step on the device (never a silently truncated one) and the host, polling a pinned copy of the guard state without a
        # sync, regrows the arena and re-captures the graph (single rank; a multi-rank job raises - size it for the worst case).
        worst = int(self.hp["rend_max_samples"])
        if capacity_per_ray is None:
            total_mem = torch.cuda.get_device_properties(dev).total_memory
            capacity_per_ray = worst if R * worst * self._row_bytes() <= 0.30 * total_mem else 64
        self.cap_max = R * worst
        self.guard = torch.zeros(3, dtype=torch.int32, device=dev)           # [overflowed, n overflowed steps, max samples seen]
        self.guard_host = torch.zeros(3, dtype=torch.int32).pin_memory()
        self.ovf_seen = 0
        self.total_samples = E(R, dtype=torch.int64, device=dev)
        self.opacity, self.depth, self.rend = E(R, **f32), E(R, **f32), E(R, Ct, **f32)
        self.rgb = E(R, 3, **f32)
        self.zeros = torch.zeros(8, **f32)          # [0:2] photometric sums, [2] grad sumsq, [3] non-finite flag (as int bits), [4:6] CE sum, valid rays
        self.d_rend, self.d_opacity, self.d_depth = E(R, Ct, **f32), E(R, **f32), torch.zeros(R, **f32)
        self.losses, self.stats, self.weights = E(3, **f32), E(32, **f32), torch.zeros(3, **f32)
        if m.pred_sem:
            self.sem_target = torch.zeros(R, dtype=torch.int64, device=dev)      # labels in [0, n_cls], 0 = void
        # False: encoder, density trunk, glue, colour head, glue (5 launches); "mlp": encoder + ONE launch for both MLPs; True: ONE
        # launch for everything.  The fused kernels build x_rgb / dx_rgb in the [h | d | 1] column order (ncn_mlp_bwd_src.perm)
        self.fuse_fwd = fuse_fwd
        self.fuse_photo = not os.environ.get("NCN_NO_FUSE_PHOTO")      # env: developer A/B only
        self.fuse_chain = not os.environ.get("NCN_NO_FUSE_CHAIN")      # env: developer A/B only
        self._alloc_arena(min(int(R * capacity_per_ray), self.cap_max))
        self.side_stream = torch.cuda.Stream(device=dev)
        self.ev_fork, self.ev_join = torch.cuda.Event(), torch.cuda.Event()
        # fp16 parameter copies (owned by the optimizer, refreshed by its Adam kernel each step)
        self.flat16 = self.opt.flat16
        # per-step schedule scalars (lr, bc1, bc2, w_ort, w_dot, w_l1): a ring of pinned host slots feeds one device
        # slot with an async H2D per step; the ring + an event throttle keep the host from overwriting a slot
        # whose copy has not executed yet
        self.RING = 64
        self.host_sched = torch.zeros(self.RING, 12, dtype=torch.float32).pin_memory()   # [this step 6 | previous step 6]
        self.ring_events = [None] * self.RING
        self.ring_pos = 0
        self.dev_sched = torch.zeros(12, **f32)
        self.prev_sched = [0.0] * 6
        self.flag_init = torch.zeros(1, dtype=torch.int32, device=dev)
        self.pending = False                 # deferred mode: a gradient is waiting for its optimizer pass
        # deferred optimizer: graph mode, single rank (the gradient all-reduce stays an eager NCCL call between two graphs;
        # capturing it inside the step graph dead-locked on 2 GPUs)
        # ... unless the optimizer is the sharded peer-memory one (trainer.peer): its two kernels carry the exchange themselves
        self.peer = trainer.peer
        # sharded exchange: the 4 B/param gradient memset leaves the optimizer kernel and becomes a node of its own that overlaps the
        # field forward (it only has to land before the first gradient-writing backward kernel)
        self.ext_zero = self.peer is not None and not os.environ.get("NCN_PEER_ZERO_IN_KERNEL")      # env: developer A/B only
        if self.peer is not None:
            self.peer.set_external_zero(self.ext_zero)
        self.ev_zero = torch.cuda.Event()
        self._zero_forked = False            # True while capturing / running the deferred form: _run_field must join ev_zero
        # sharded exchange on > 1 rank: the table backward runs as two launches, fine levels first; the fine levels' gradient
        # (and the MLPs', complete before it) is the exchange's EARLY range, pulled over NVLink by ncn_peer_early on the side
        # stream while the coarse levels are still being computed (set up in _setup_peer_early once the offsets are known)
        self.peer_early = None
        self.ev_fork3, self.ev_join3 = torch.cuda.Event(), torch.cuda.Event()
        self.nccl = trainer.world_size > 1 and self.peer is None
        self.defer = use_graph and not self.nccl and not os.environ.get("NCN_NO_DEFER")      # env: developer A/B only
        # multi-rank: the same overlap with three graphs on two streams and an EAGER all-reduce in between
        #   opt stream : [all-reduce(prev grads) -> graph(adam)]      main stream: graph(march) -> join -> graph(field)
        self.defer_multi = use_graph and self.nccl
        # developer A/B knob (NCN_SUMSQ_TAIL=1): take ||g||^2 at the TAIL of the step that produced the gradient (L2-hot) instead of
        # in front of the next step's Adam.  Measured: 0.556 vs 0.553 ms/step, e2e 0.581 vs 0.569 - no gain (the optimizer branch
        # is not what bounds the head of the graph once it co-runs with the march), so the default stays the head placement.
        self.sumsq_tail = self.defer and self.peer is None and os.environ.get("NCN_SUMSQ_TAIL", "0") == "1"
        self.opt_stream = torch.cuda.Stream(device=dev)
        self.ev_fork2, self.ev_join2 = torch.cuda.Event(), torch.cuda.Event()
        self.coef = torch.ones(1, **f32)
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._flat_mods = [mod for mod in self.model.modules() if getattr(mod, "_half_key", None) == "flat"]
        self.adam_groups = _lib.AdamGroups()
        self.adam_groups.n_groups = len(self.opt.groups)
        for q, (start, _, wd) in enumerate(self.opt.groups):
            self.adam_groups.start[q] = start
            self.adam_groups.weight_decay[q] = wd
        self.adam_groups.max_norm = float(self.hp["grad_clip"])
        self.grad_div = torch.tensor([float(trainer.world_size)], **f32)
        self.xform = (C.c_float * 6)(*([float(m.xyz_min[0, i]) for i in range(3)] + [float((m.xyz_max - m.xyz_min)[0, i]) for i in range(3)]))
        self.bg = (C.c_float * 3)(1.0, 1.0, 1.0) if self.hp["exp_step_factor"] == 0 else (C.c_float * 3)(0.0, 0.0, 0.0)
        self.km_params = _lib.KmeansParams(20, 20, 1234, 256, 1)
        self._offsets()
        self._alloc_tri(None)

    def _row_bytes(self):
        """bytes of arena per sample row (all per-sample buffers of _alloc_arena)"""
        m = self.model
        Ct = self.Ct
        b = 12 + 12 + 4 + 4                      # xyzs, dirs, deltas, ts
        b += 64 + 32 + 128 + 64 + 32 + 256       # feat, h, sig_acts, x_rgb, rgb_out, rgb_acts
        b += 4 + 4 * Ct + 4                      # sigmas, raws, ws
        b += 4 + 4 * Ct + 32 + 64 + 32 + 64      # d_sigmas, d_raws, dout_rgb, dx_rgb, dh, dfeat
        b += 2 * (16 + 2 * 64) * 2               # two MLP-backward workspaces
        if m.pred_norm:
            b += 32
        if m.pred_sem:
            b += 32 + 256 + 32 + 32
        return b

    def _alloc_arena(self, cap):
        """every per-sample buffer of the step, for `cap` rows (called once, and again by _regrow)"""
        m, dev, Ct, R = self.model, self.dev, self.Ct, self.R
        self.cap = cap = int(cap)
        f32 = dict(dtype=torch.float32, device=dev); f16 = dict(dtype=torch.float16, device=dev)
        E = lambda *s, **k: torch.empty(*s, **k)
        self.xyzs, self.dirs = torch.zeros(cap, 3, **f32), torch.zeros(cap, 3, **f32)
        self.deltas, self.ts = torch.zeros(cap, **f32), torch.zeros(cap, **f32)
        # field
        self.feat, self.h = E(cap, 32, **f16), E(cap, 16, **f16)
        cap_t = (cap + 127) // 128 * 128                         # saved activations: whole 128-row tiles (ncn_mlp_acts_bytes)
        self.sig_acts = E(1, cap_t, 64, **f16)
        self.x_rgb, self.rgb_out, self.rgb_acts = E(cap, 32, **f16), E(cap, 16, **f16), E(2, cap_t, 64, **f16)
        self.sigmas, self.raws = E(cap, **f32), E(cap, Ct, **f32)
        self.ws = E(cap, **f32)
        self.d_sigmas, self.d_raws = E(cap, **f32), E(cap, Ct, **f32)
        self.dout_rgb, self.dx_rgb, self.dh, self.dfeat = E(cap, 16, **f16), E(cap, 32, **f16), E(cap, 16, **f16), E(cap, 32, **f16)
        nb = max(self.L.ncn_mlp_bwd_workspace_bytes(C.byref(m.rgb_net.desc), cap),
                 self.L.ncn_mlp_bwd_workspace_bytes(C.byref(m.sigma_net.desc), cap))
        # extra heads on h (ngp_mt.py:217-224).  The reference puts no loss on the rendered norm_nn channels (losses.py only
        # slices them, :279/:294), so norm_net gets an exactly-zero gradient: its backward is not launched.
        if m.pred_norm:
            self.norm_out = E(cap, 16, **f16)
        if m.pred_sem:
            self.sem_out, self.sem_acts, self.dh_sem = E(cap, 16, **f16), E(2, cap_t, 64, **f16), E(cap, 16, **f16)
            self.sem_dout = E(cap, 16, **f16) if self.n_cls > 3 else None
            nb = max(nb, self.L.ncn_mlp_bwd_workspace_bytes(C.byref(m.sem_net.desc), cap))
        self.mlp_ws = E(nb, dtype=torch.uint8, device=dev)
        self.mlp_ws2 = E(nb, dtype=torch.uint8, device=dev)      # the colour-head backward runs concurrently on a side stream
        fuse_fwd = self.fuse_fwd
        self.src_rgb = _lib.MlpBwdSrc(1, ptr(self.d_raws), Ct, 0, 3, None, None, None, 1.0, 1 if fuse_fwd else 0)
        self.src_sig = _lib.MlpBwdSrc(2, None, 0, 0, 0, ptr(self.dx_rgb), ptr(self.d_sigmas), ptr(self.h), 1.0, 2 if fuse_fwd else 0,
                                      ptr(self.dh_sem) if m.pred_sem else None)
        self.src_sem = _lib.MlpBwdSrc(1, ptr(self.d_raws), Ct, self.sem_off, self.n_cls, None, None, None, 1.0, 0) if m.pred_sem else None
        self.graph = None
        self.grid_graph = None

    def _regrow(self, need):
        """the march produced `need` > capacity samples (those steps were skipped on the device): grow and re-capture"""
        import warnings
        if self.tr.world_size > 1:
            raise RuntimeError(f"FusedStep: the march produced {need} samples for an arena of {self.cap} rows on rank {self.tr.rank}; the "
                               "overflowed steps were skipped on every rank.  A multi-rank job cannot re-capture on one rank alone: "
                               "construct it with capacity_per_ray=rend_max_samples (the default when it fits in memory)")
        new_cap = min(self.cap_max, max(2 * self.cap, (int(need * 1.25) + 1023) // 1024 * 1024))
        warnings.warn(f"FusedStep: sample arena overflow ({need} samples > capacity {self.cap}); {int(self.guard_host[1])} step(s) were skipped. "
                      f"Growing the arena to {new_cap} rows and re-capturing the step graph.")
        self.flush()
        torch.cuda.synchronize()
        self._alloc_arena(new_cap)

    def _poll(self):
        """host-side, sync-free health check before a step: peer-exchange time-outs are fatal, arena overflows regrow"""
        if self.peer is not None and self.peer.poll():
            raise RuntimeError(f"peer-memory gradient exchange: a cross-GPU wait timed out (phase {self.peer.poll() - 1}) on rank "
                               f"{self.tr.rank}; the replicated parameters are no longer trustworthy - restart from a checkpoint")
        n_ovf = int(self.guard_host[1])
        if n_ovf > self.ovf_seen:
            self.ovf_seen = n_ovf
            if int(self.guard_host[2]) > self.cap:
                self._regrow(int(self.guard_host[2]))

    @property
    def overflow_steps(self):
        """number of steps skipped because the march overflowed the arena (synchronises)"""
        torch.cuda.synchronize()
        return int(self.guard[1])

    def _offsets(self):
        """slices of the flat parameter / gradient buffers per module (hash table first: FlatAdam order)"""
        opt, m = self.opt, self.model
        base = opt.flat.data_ptr()
        self.off = {}
        for name in ("xyz_encoder", "sigma_net", "rgb_net", "sem_net", "norm_net"):
            if not hasattr(m, name):
                continue
            p = getattr(m, name).params
            o = (p.data_ptr() - base) // 4
            self.off[name] = (o, p.numel())

    def _alloc_tri(self, tri):
        dev = self.dev
        M = 0 if tri is None else tri.shape[1]
        self.tri = tri
        self.M = M
        self.normals = torch.empty(max(M, 1), 3, dtype=torch.float32, device=dev)
        self.dn = torch.empty(max(M, 1), 3, dtype=torch.float32, device=dev)
        self.assign = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
        self.labels = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
        self.centroids = torch.empty(20, 3, dtype=torch.float32, device=dev)
        self.n_valid = torch.empty(1, dtype=torch.int32, device=dev)
        self.sel = torch.empty(3, dtype=torch.int32, device=dev)
        self.km_ws = torch.empty(self.L.ncn_kmeans_workspace_bytes(max(M, 1), 20), dtype=torch.uint8, device=dev)

    # ------------------------------------------------------------------ the kernel sequence
    def _w16(self, name):
        o, n = self.off[name]
        return self.flat16[o:o + n]

    def _g32(self, name):
        o, n = self.off[name]
        return self.opt.grad[o:o + n]

    def _run(self):
        self._run_march()
        self._run_field()

    def _run_march(self):
        """the part of the step that does not depend on the parameters: jitter, AABB, occupancy march"""
        L, m, hp = self.L, self.model, self.hp
        st = torch.cuda.current_stream().cuda_stream
        R, cap = self.R, self.cap
        ck = check
        self.zeros[0:2].zero_()
        if self.n_cls:
            self.zeros[4:6].zero_()
        self.d_depth.zero_()
        if self.dev_sampling is not None:    # BaseDataset.__getitem__ (datasets/base.py:94-183) on the device: indices + target gather
            sm = self.dev_sampling
            n_gt = self.n_gt                 # random_tr_poses: the drawn half; the other half repeats its pixels from generated poses
            ck(L.ncn_sample_ray_batch_ex(sm["strategy"], ptr(sm["seed"]), n_gt, sm["n_poses"], sm["H"], sm["W"], sm["patch"], sm["max_expand"],
                                         ptr(self.b_img), ptr(self.b_pix), st), "sample_ray_batch")
            if self.rtp:
                ck(L.ncn_sample_random_pose_half(sm["strategy"], ptr(sm["seed"]), n_gt, sm["n_random"], sm["n_poses"], sm["patch"],
                                                 ptr(self.b_img), ptr(self.b_pix), st), "sample_random_pose_half")
            ck(L.ncn_gather_pixels(ptr(sm["images"]), ptr(self.b_img), ptr(self.b_pix), n_gt, sm["H"] * sm["W"], 3, ptr(self.target), st), "gather_rgb")
            if sm["labels"] is not None:
                ck(L.ncn_gather_pixels(ptr(sm["labels"]), ptr(self.b_img), ptr(self.b_pix), n_gt, sm["H"] * sm["W"], 2, ptr(self.sem_target), st),
                   "gather_labels")
        if self.pix_inputs:                  # NeRFSystem.forward gather + get_rays (train_nerf.py:167-182) as the step's first node
            tr = self.tr
            ck(L.ncn_rays_from_pixels(ptr(tr.poses), ptr(tr.directions), ptr(self.b_img), ptr(self.b_pix), R, ptr(self.rays_o),
                                      ptr(self.rays_d), st), "rays_from_pixels")
        if self.gen_noise:
            self.noise.uniform_()
        ck(L.ncn_ray_aabb_near(ptr(self.rays_o), ptr(self.rays_d), ptr(m.center), ptr(m.half_size), float(hp["rend_near_dist"]), R,
                               ptr(self.hits_t), st), "aabb")
        ck(L.ncn_march_train(ptr(self.rays_o), ptr(self.rays_d), ptr(self.hits_t), ptr(m.density_bitfield), m.cascades,
                             float(m.scale), float(hp["exp_step_factor"]), ptr(self.noise), m.grid_size, hp["rend_max_samples"], R,
                             cap, ptr(self.rays_a), ptr(self.xyzs), ptr(self.dirs), ptr(self.deltas), ptr(self.ts), ptr(self.counter),
                             ptr(self.march_ws), self.march_ws.numel(), st), "march")

    def _run_field(self):
        """field forward, compositing, losses and the whole backward pass (needs the current parameters)"""
        L, m, hp = self.L, self.model, self.hp
        st = torch.cuda.current_stream().cuda_stream
        R, cap, Ct = self.R, self.cap, self.Ct
        n_dev = ptr(self.counter)
        ck = check
        enc, sg, rgbn = m.xyz_encoder, m.sigma_net, m.rgb_net
        if self.fuse_fwd == "mlp":      # encoder, then both MLPs in one launch
            ck(L.ncn_grid_fwd(C.byref(enc.desc), ptr(self.xyzs), ptr(self._w16("xyz_encoder")), cap, ptr(self.feat), self.xform, n_dev, st), "grid_fwd")
            ck(L.ncn_field_mlp_fwd(ptr(self.feat), ptr(self.dirs), ptr(self._w16("sigma_net")), ptr(self._w16("rgb_net")), cap, n_dev,
                                   ptr(self.sigmas), ptr(self.raws), Ct, ptr(self.h), ptr(self.sig_acts), ptr(self.x_rgb), ptr(self.rgb_acts),
                                   ptr(self.rgb_out), st), "field_mlp_fwd")
        elif self.fuse_fwd:
            ck(L.ncn_field_fwd(C.byref(enc.desc), ptr(self.xyzs), ptr(self.dirs), ptr(self._w16("xyz_encoder")), ptr(self._w16("sigma_net")),
                               ptr(self._w16("rgb_net")), cap, n_dev, self.xform, ptr(self.sigmas), ptr(self.raws), Ct, ptr(self.feat),
                               ptr(self.h), ptr(self.sig_acts), ptr(self.x_rgb), ptr(self.rgb_acts), ptr(self.rgb_out), st), "field_fwd")
        else:
            ck(L.ncn_grid_fwd(C.byref(enc.desc), ptr(self.xyzs), ptr(self._w16("xyz_encoder")), cap, ptr(self.feat), self.xform, n_dev, st), "grid_fwd")
            ck(L.ncn_mlp_fwd(C.byref(sg.desc), ptr(self.feat), ptr(self._w16("sigma_net")), cap, ptr(self.h), ptr(self.sig_acts), n_dev, st), "sigma_fwd")
            ck(L.ncn_field_prepare_rgb(ptr(self.dirs), ptr(self.h), cap, n_dev, ptr(self.x_rgb), ptr(self.sigmas), st), "prepare_rgb")
            ck(L.ncn_mlp_fwd(C.byref(rgbn.desc), ptr(self.x_rgb), ptr(self._w16("rgb_net")), cap, ptr(self.rgb_out), ptr(self.rgb_acts), n_dev, st), "rgb_fwd")
            ck(L.ncn_field_head_out(ptr(self.rgb_out), 16, cap, n_dev, ptr(self.raws), Ct, 0, 3, st), "head_out")
        if (m.pred_norm or m.pred_sem) and self.fuse_fwd:      # both extra heads in ONE launch, outputs written into their raws columns
            ck(L.ncn_field_heads_fwd(ptr(self.h), cap, n_dev, ptr(self.raws), Ct,
                                     ptr(self._w16("norm_net")) if m.pred_norm else None, self.norm_off, 3, None, None,
                                     ptr(self._w16("sem_net")) if m.pred_sem else None, self.sem_off, max(self.n_cls, 1),
                                     ptr(self.sem_acts) if m.pred_sem else None, ptr(self.sem_out) if m.pred_sem else None, st), "heads_fwd")
        elif m.pred_norm:               # raws[:, 3:6] = norm_net(h)   (ngp_mt.py:221-224, rendering.py:203-206)
            ck(L.ncn_mlp_fwd(C.byref(m.norm_net.desc), ptr(self.h), ptr(self._w16("norm_net")), cap, ptr(self.norm_out), None, n_dev, st), "norm_fwd")
            ck(L.ncn_field_head_out(ptr(self.norm_out), 16, cap, n_dev, ptr(self.raws), Ct, self.norm_off, 3, st), "norm_head_out")
        if m.pred_sem and not self.fuse_fwd:   # raws[:, sem_off:] = sem_net(h)  (ngp_mt.py:217-220, rendering.py:207-208)
            ck(L.ncn_mlp_fwd(C.byref(m.sem_net.desc), ptr(self.h), ptr(self._w16("sem_net")), cap, ptr(self.sem_out), ptr(self.sem_acts), n_dev, st), "sem_fwd")
            ck(L.ncn_field_head_out(ptr(self.sem_out), 16, cap, n_dev, ptr(self.raws), Ct, self.sem_off, self.n_cls, st), "sem_head_out")
        # ---- compositing + losses (+ their gradients w.r.t. the rendered quantities; channels without a loss get a zero gradient)
        # random_tr_poses: the *_gt entry points (target colours on the first n_gt rays only); otherwise the plain ones (n_gt = R)
        if self.fuse_photo and Ct in (3, 6, 9):      # ONE launch: the lane that finishes a ray also evaluates its photometric / opacity terms
            head = (ptr(self.sigmas), ptr(self.raws), ptr(self.deltas), ptr(self.ts), ptr(self.rays_a), 1e-4, R, cap, Ct, ptr(self.total_samples),
                    ptr(self.opacity), ptr(self.depth), ptr(self.rend), ptr(self.ws), ptr(self.target))
            tail = (self.bg, float(hp["loss_opacity_w"]), GSCALE, ptr(self.rgb), ptr(self.zeros), ptr(self.d_rend), ptr(self.d_opacity), st)
            if self.rtp:
                ck(L.ncn_composite_train_fw_photometric_gt(*head, self.n_gt, *tail), "composite_fw_photometric")
            else:
                ck(L.ncn_composite_train_fw_photometric(*head, *tail), "composite_fw_photometric")
        else:
            ck(L.ncn_composite_train_fw(ptr(self.sigmas), ptr(self.raws), ptr(self.deltas), ptr(self.ts), ptr(self.rays_a), 1e-4, R, cap, Ct,
                                        ptr(self.total_samples), ptr(self.opacity), ptr(self.depth), ptr(self.rend), ptr(self.ws), st), "composite_fw")
            tail = (Ct, self.bg, float(hp["loss_opacity_w"]), GSCALE, ptr(self.rgb), ptr(self.zeros), ptr(self.d_rend), ptr(self.d_opacity), st)
            if self.rtp:
                ck(L.ncn_photometric_loss_gt(ptr(self.rend), ptr(self.opacity), ptr(self.target), R, self.n_gt, *tail), "photometric")
            else:
                ck(L.ncn_photometric_loss(ptr(self.rend), ptr(self.opacity), ptr(self.target), R, *tail), "photometric")
        if m.pred_sem and self.sem_w > 0:
            # pred['sem'][:gt_l] (losses.py:277): rays of generated poses keep the zero gradient the photometric pass wrote
            ck(L.ncn_semantic_ce_loss(ptr(self.rend), Ct, self.sem_off, self.n_cls, ptr(self.sem_target), self.n_gt, self.sem_w * GSCALE,
                                      ptr(self.zeros[4:6]), ptr(self.d_rend), st), "semantic_ce")
        inv = 1.0 / GSCALE
        # Two independent branches from here (forked onto a side stream; inside a CUDA graph they become parallel
        # branches): (A) the colour head's backward needs only dL/draws = dL/drend * w, (B) the normal-clustering
        # chain (normals -> k-means on an 8-CTA cluster -> selection -> cluster loss -> dL/ddepth) and then the
        # density part of the compositing backward.  They join at dL/dh.
        main = torch.cuda.current_stream()
        side = main if getattr(self, "serial", False) else self.side_stream      # serial: instrumented passes time every call alone
        self.ev_fork.record(main)
        side.wait_event(self.ev_fork)
        with torch.cuda.stream(side):
            sst = side.cuda_stream
            ck(L.ncn_composite_train_bw(None, None, ptr(self.d_rend), None, ptr(self.sigmas), ptr(self.raws), ptr(self.ws),
                                        ptr(self.deltas), ptr(self.ts), ptr(self.rays_a), ptr(self.opacity), ptr(self.depth), ptr(self.rend), 1e-4,
                                        R, cap, Ct, None, ptr(self.d_raws), sst), "composite_bw_raws")
            if self._zero_forked:
                side.wait_event(self.ev_zero)     # first gradient writer of this branch: the forked memset must have landed
            ck(L.ncn_mlp_bwd_src_fused(C.byref(rgbn.desc), C.byref(self.src_rgb), ptr(self.x_rgb), ptr(self._w16("rgb_net")), ptr(self.rgb_out),
                                       ptr(self.rgb_acts), cap, ptr(self._g32("rgb_net")), ptr(self.dx_rgb), inv, ptr(self.mlp_ws2),
                                       self.mlp_ws2.numel(), n_dev, sst), "rgb_bwd")
            if m.pred_sem:              # semantic head backward: dL/dh of this head joins the density trunk's dL/dout (src_sig.dx_extra)
                semn = m.sem_net
                if self.sem_dout is None:
                    ck(L.ncn_mlp_bwd_src_fused(C.byref(semn.desc), C.byref(self.src_sem), ptr(self.h), ptr(self._w16("sem_net")), ptr(self.sem_out),
                                               ptr(self.sem_acts), cap, ptr(self._g32("sem_net")), ptr(self.dh_sem), inv, ptr(self.mlp_ws2),
                                               self.mlp_ws2.numel(), n_dev, sst), "sem_bwd")
                else:
                    ck(L.ncn_field_head_dout(ptr(self.d_raws), Ct, self.sem_off, self.n_cls, 1.0, cap, n_dev, ptr(self.sem_dout), 16, sst), "sem_head_dout")
                    ck(L.ncn_mlp_bwd(C.byref(semn.desc), ptr(self.h), ptr(self._w16("sem_net")), ptr(self.sem_out), ptr(self.sem_acts),
                                     ptr(self.sem_dout), cap, ptr(self._g32("sem_net")), ptr(self.dh_sem), inv, ptr(self.mlp_ws2),
                                     self.mlp_ws2.numel(), n_dev, sst), "sem_bwd")
            self.ev_join.record(side)
        if self.M > 0:
            x1, x2, x3 = ptr(self.tri[0]), ptr(self.tri[1]), ptr(self.tri[2])
            t_sim = 1.0 - float(hp["loss_norm_can_tres"])
            # rays_o := rays_d (rendering.py:227 quirk).  The triangles index the rays from u0 on (all rays, or the generated-pose half)
            u0 = self.u0
            rd_u, depth_u, ddepth_u = ptr(self.rays_d[u0:]), ptr(self.depth[u0:]), ptr(self.d_depth[u0:])
            if self.fuse_chain and self.M <= 262144 and self.km_params.k <= 32:
                # normals -> k-means -> selection -> cluster statistics + losses in ONE cluster launch, then dL/dnormals + dL/ddepth
                ck(L.ncn_cluster_chain(rd_u, rd_u, depth_u, x1, x2, x3, self.M, C.byref(self.km_params), t_sim,
                                       ptr(self.normals), ptr(self.centroids), ptr(self.assign), ptr(self.n_valid), ptr(self.labels), ptr(self.sel),
                                       ptr(self.losses), ptr(self.stats), ptr(self.km_ws), self.km_ws.numel(), st), "cluster_chain")
                ck(L.ncn_cluster_bw_depth(ptr(self.normals), ptr(self.labels), self.M, ptr(self.stats), ptr(self.dev_sched[3:6]), ptr(self.dn),
                                          rd_u, rd_u, depth_u, x1, x2, x3, ddepth_u, st), "cluster_bw_depth")
            else:
                ck(L.ncn_normals_from_depth_fw(rd_u, rd_u, depth_u, x1, x2, x3, self.M, ptr(self.normals), st), "normals_fw")
                ck(L.ncn_kmeans_spherical(ptr(self.normals), self.M, C.byref(self.km_params), ptr(self.centroids), ptr(self.assign),
                                          ptr(self.n_valid), ptr(self.km_ws), self.km_ws.numel(), st), "kmeans")
                # selection + cluster losses (one single-CTA launch), dL/dnormals + dL/ddepth (one multi-CTA launch)
                ck(L.ncn_cluster_tail(ptr(self.centroids), ptr(self.assign), self.M, 20, t_sim,
                                      ptr(self.labels), ptr(self.sel), ptr(self.normals), ptr(self.losses), ptr(self.stats),
                                      ptr(self.dev_sched[3:6]), ptr(self.dn), rd_u, rd_u, depth_u,
                                      x1, x2, x3, ddepth_u, st), "cluster_tail")
        ck(L.ncn_composite_train_bw(ptr(self.d_opacity), ptr(self.d_depth), ptr(self.d_rend), None, ptr(self.sigmas), ptr(self.raws), ptr(self.ws),
                                    ptr(self.deltas), ptr(self.ts), ptr(self.rays_a), ptr(self.opacity), ptr(self.depth), ptr(self.rend), 1e-4,
                                    R, cap, Ct, ptr(self.d_sigmas), None, st), "composite_bw_sigma")
        main.wait_event(self.ev_join)
        if self._zero_forked:
            main.wait_event(self.ev_zero)
        ck(L.ncn_mlp_bwd_src_fused(C.byref(sg.desc), C.byref(self.src_sig), ptr(self.feat), ptr(self._w16("sigma_net")), ptr(self.h),
                                   ptr(self.sig_acts), cap, ptr(self._g32("sigma_net")), ptr(self.dfeat), inv, ptr(self.mlp_ws),
                                   self.mlp_ws.numel(), n_dev, st), "sigma_bwd")
        if self.peer_early is None:
            self._setup_peer_early()
        lc = self.peer_early
        if lc and not getattr(self, "serial", False):
            ck(L.ncn_grid_bwd_levels(C.byref(enc.desc), ptr(self.xyzs), ptr(self.dfeat), cap, ptr(self._g32("xyz_encoder")), inv, self.xform, n_dev,
                                     lc, int(enc.desc.n_levels), 8, st), "grid_bwd_fine")
            self.ev_fork3.record(main)
            self.side_stream.wait_event(self.ev_fork3)
            with torch.cuda.stream(self.side_stream):
                self.peer.early(self.grad_div, self.side_stream.cuda_stream)
                self.ev_join3.record(self.side_stream)
            # one CTA slot per SM less than the kernel could fill: the exchange's CTAs must become resident beside it
            ck(L.ncn_grid_bwd_levels(C.byref(enc.desc), ptr(self.xyzs), ptr(self.dfeat), cap, ptr(self._g32("xyz_encoder")), inv, self.xform, n_dev,
                                     0, lc, int(os.environ.get("NCN_PEER_EARLY_COARSE_CTAS", "4")), st), "grid_bwd_coarse")
            main.wait_event(self.ev_join3)
        else:
            ck(L.ncn_grid_bwd(C.byref(enc.desc), ptr(self.xyzs), ptr(self.dfeat), cap, ptr(self._g32("xyz_encoder")), inv, self.xform, n_dev, st), "grid_bwd")
            if lc:                               # instrumented (serial) pass: same exchange, no overlap
                self.peer.early(self.grad_div, st)
        if cap < self.cap_max:          # overflow is possible: a truncated step must become a skipped step, never a silent one
            ck(L.ncn_step_guard(n_dev, cap, ptr(self.guard), ptr(self.opt.grad), st), "step_guard")
            self.guard_host.copy_(self.guard, non_blocking=True)       # pinned: the host polls it without synchronising

    def _setup_peer_early(self):
        """decide once whether the exchange gets an early range: > 1 rank on the peer-memory exchange, the hash table is the FIRST
        tensor of the flat vector (so that [cut, n) = fine levels + every MLP is one contiguous range), NCN_PEER_EARLY=1 (opt-in:
        see DESIGN.md section 6 for what it measured)"""
        self.peer_early = 0
        tr = self.tr
        force = os.environ.get("NCN_PEER_EARLY_FORCE") == "1"      # developer: the one-rank form (local reads), to look at the co-scheduling
        if self.peer is None or (tr.world_size < 2 and not force) or os.environ.get("NCN_PEER_EARLY", "0") != "1":
            return
        enc = self.model.xyz_encoder
        o, n = self.off["xyz_encoder"]
        lc = int(os.environ.get("NCN_PEER_EARLY_LEVEL", "8"))
        if o != 0 or int(enc.desc.n_features) != 2 or not (0 < lc < int(enc.desc.n_levels)):
            return
        self.peer.set_cut(int(enc.desc.level_offset[lc]) * 2)
        self.peer_early = lc

    def _optimizer(self, sched_off=0):
        """sum of squares -> clip coefficient -> fused Adam (both parameter groups); sched_off selects the schedule slot
        (0 = this step's lr / bias corrections, 6 = the previous step's, for the deferred update)"""
        L, opt = self.L, self.opt
        st = torch.cuda.current_stream().cuda_stream
        sumsq = self.zeros[2:3]
        sumsq.zero_()
        self.flag.copy_(self.flag_init)          # 0, or 1 to skip the (empty) update of the very first deferred step
        if self.peer is not None:                # gradient exchange + norm + clip + Adam + fp16 refresh over NVLink peer memory
            self.peer.step(opt.flat, opt.m, opt.v, self.adam_groups, opt.betas, opt.eps, self.grad_div, self.flag,
                           self.dev_sched[sched_off:sched_off + 3], sumsq, st)
            if self.ext_zero and not self._zero_forked:      # sequential forms (eager step, flush): zero right behind the exchange
                opt.grad.zero_()
            return
        check(L.ncn_grad_sumsq(ptr(opt.grad), opt.grad.numel(), ptr(self.grad_div), ptr(sumsq), ptr(self.flag), st), "sumsq")
        # both parameter groups (hash table wd 0 / MLPs wd 1e-6) and the clip coefficient in ONE launch
        check(L.ncn_adam_step_groups(ptr(opt.flat), ptr(opt.grad), ptr(opt.m), ptr(opt.v), ptr(self.flat16), opt.flat.numel(),
                                     C.byref(self.adam_groups), opt.betas[0], opt.betas[1], opt.eps, ptr(self.grad_div), ptr(self.flag),
                                     ptr(sumsq), ptr(self.dev_sched[sched_off:sched_off + 3]), st), "adam")

    def _sumsq(self):
        """||g||^2 (+ non-finite flag) of the gradient that is complete on the current stream"""
        opt = self.opt
        sumsq = self.zeros[2:3]
        sumsq.zero_()
        self.flag.zero_()
        check(self.L.ncn_grad_sumsq(ptr(opt.grad), opt.grad.numel(), ptr(self.grad_div), ptr(sumsq), ptr(self.flag),
                                    torch.cuda.current_stream().cuda_stream), "sumsq")

    def _adam_only(self, sched_off=0):
        """clip + Adam of both parameter groups from an already computed ||g||^2 / skip flag (sumsq_tail mode)"""
        opt = self.opt
        check(self.L.ncn_adam_step_groups(ptr(opt.flat), ptr(opt.grad), ptr(opt.m), ptr(opt.v), ptr(self.flat16), opt.flat.numel(),
                                          C.byref(self.adam_groups), opt.betas[0], opt.betas[1], opt.eps, ptr(self.grad_div), ptr(self.flag),
                                          ptr(self.zeros[2:3]), ptr(self.dev_sched[sched_off:sched_off + 3]),
                                          torch.cuda.current_stream().cuda_stream), "adam")

    def _run_deferred(self, multi):
        """One replay = [apply the PREVIOUS step's update] || [jitter + AABB + march of THIS step] -> field/backward.
        The optimizer pass (dense, HBM bound) and the march (serial, latency bound) do not depend on each other, so they
        run as parallel graph branches; the sequence of updates is unchanged (Adam_k still precedes forward_k+1),
        `flush()` applies the last pending update."""
        main = torch.cuda.current_stream()
        opt_stream = self.opt_stream
        self.ev_fork2.record(main)
        opt_stream.wait_event(self.ev_fork2)
        forked = self.ext_zero
        with torch.cuda.stream(opt_stream):
            if multi:
                self.tr.comm.allreduce_sum_(self.opt.grad)
            self._zero_forked = forked
            if self.sumsq_tail:
                self._adam_only(sched_off=6)      # norm + skip flag were left by the previous replay's tail (or by step() / flush())
            else:
                self._optimizer(sched_off=6)
            self.ev_join2.record(opt_stream)
            if forked:                            # the parameters are published: the forward may start; the memset runs beside it
                self.opt.grad.zero_()
                self.ev_zero.record(opt_stream)
        self._run_march()
        main.wait_event(self.ev_join2)
        self._run_field()
        self._zero_forked = False
        if self.sumsq_tail:
            self._sumsq()

    # ------------------------------------------------------------------ public
    def set_triangles(self, tri):
        tri = tri.to(self.dev, torch.int64).contiguous()
        if self.tri is None or tri.shape != self.tri.shape:
            self._alloc_tri(tri)
            self.graph = None
        else:
            self.tri.copy_(tri)

    def _schedule(self):
        tr, hp, opt = self.tr, self.hp, self.opt
        opt.step_count += 1
        t = opt.step_count
        ls = tr.loss
        step = tr.global_step
        on = (step <= ls.can_sched_end) or (ls.can_sched_end == -1)
        slot = self.ring_pos % self.RING
        self.ring_pos += 1
        if self.ring_events[slot] is not None:
            self.ring_events[slot].synchronize()        # the copy issued RING steps ago has executed
        h = self.host_sched[slot]
        h[0] = tr.lr_now(); h[1] = 1 - opt.betas[0] ** t; h[2] = 1 - opt.betas[1] ** t
        h[3] = ls.w_sched(ls.w_ort, step) * GSCALE if on else 0.0
        h[4] = ls.w_sched(ls.w_dot, step) * GSCALE if on else 0.0
        h[5] = ls.w_sched(ls.w_l1, step) * GSCALE if on else 0.0
        for i in range(6):
            h[6 + i] = self.prev_sched[i]
        self.prev_sched = [float(h[i]) for i in range(6)]
        self.dev_sched.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.ring_events[slot] = ev

    def step(self, rays_o=None, rays_d=None, target_rgb=None, noise=None, packed=None, sem_target=None):
        """one training step.  Inputs are device tensors copied into the graph's static buffers: either
        rays_o/rays_d/target_rgb (R,3) each, or `packed` (3,R,3) = [rays_o, rays_d, rgb] (one copy), or nothing when the
        caller filled self.inp in place (e.g. rays_from_pixels).  With random_tr_poses the rays are [n_gt training-view rays |
        n_gt rays of generated poses] and target_rgb / sem_target may be given for the first n_gt = R/2 rays only."""
        tr = self.tr
        self._poll()
        if packed is not None:
            self.inp.copy_(packed)
        elif rays_o is not None:
            self.rays_o.copy_(rays_o); self.rays_d.copy_(rays_d)
            if target_rgb.shape[0] not in (self.n_gt, self.R):
                raise RuntimeError(f"FusedStep.step: target_rgb has {target_rgb.shape[0]} rows, expected {self.n_gt}")
            self.target[:target_rgb.shape[0]].copy_(target_rgb)
        if sem_target is not None:
            self.sem_target[:sem_target.shape[0]].copy_(sem_target)
        if (noise is None) != self.gen_noise:
            self.gen_noise = noise is None
            self.graph = None                 # noise source is part of the captured sequence
        if noise is not None:
            self.noise.copy_(noise)
        stale = [mod for mod in self._flat_mods if mod.params._version != mod._flat_version]
        if stale:                             # parameters written from outside (load_state_dict, tests): refresh the fp16 working copy
            if tr.master_stale:
                # sharded optimizer: outside this rank's slice the fp32 master is stale - rebuild it from the owners before it
                # overwrites the (correct) fp16 copy.  Collective: replicated parameters must be written on every rank alike.
                keep = {id(mod): mod.params.detach().clone() for mod in stale}
                tr.gather_master_params()
                for mod in stale:             # the caller's write wins inside the written tensors (same write on every rank)
                    mod.params.data.copy_(keep[id(mod)])
            for mod in stale:
                mod._half.copy_(mod.params.detach())
                mod._flat_version = mod.params._version
        self._schedule()
        multi = self.nccl
        if not self.use_graph:
            self._run()
            if multi:
                tr.comm.allreduce_sum_(self.opt.grad)
            self._optimizer()
        elif self.defer_multi:
            if self.graph is None:
                self._capture_multi()
            g_march, g_adam, g_field = self.graph
            main = torch.cuda.current_stream()
            self.opt_stream.wait_stream(main)             # the previous step's backward (gradients) is complete
            if self.pending:
                with torch.cuda.stream(self.opt_stream):
                    tr.comm.allreduce_sum_(self.opt.grad)
                    g_adam.replay()
            g_march.replay()
            main.wait_stream(self.opt_stream)
            g_field.replay()
            self.pending = True
        elif self.defer:
            if self.graph is None:
                self._capture(multi)
            if not self.pending:
                # nothing to apply yet: the optimizer branch of this replay is a no-op
                (self.flag if self.sumsq_tail else self.flag_init).fill_(1)
            self.graph[0].replay()
            if not self.pending:
                self.flag_init.zero_()
                self.pending = True
        else:
            if self.graph is None:
                self._capture(multi)
            self.graph[0].replay()
            if multi:
                tr.comm.allreduce_sum_(self.opt.grad)
                self.graph[1].replay()
        if self.peer is not None and tr.world_size > 1:
            tr.master_stale = True
        tr.global_step += 1

    def _new_graph(self):
        if getattr(self, "census", None) is not None:      # bench.py: keep the cudaGraph_t to count its nodes
            try:
                return torch.cuda.CUDAGraph(keep_graph=True)
            except TypeError:
                pass
        return torch.cuda.CUDAGraph()

    def _census(self, name, g):
        """node counts [kernel, memcpy, memset, other] of a captured graph into self.census[name] (only when self.census is a dict)"""
        if getattr(self, "census", None) is None or not hasattr(g, "raw_cuda_graph"):
            return
        try:
            raw = g.raw_cuda_graph()
            c = (C.c_int * 4)()
            check(self.L.ncn_graph_node_counts(C.c_void_p(int(raw)), c), "graph_node_counts")
            self.census[name] = list(c)
            g.instantiate()
        except Exception as e:  # noqa: BLE001  (beta API: a failed census must not break the step)
            self.census[name] = f"unavailable: {type(e).__name__}: {e}"

    def _capture_multi(self):
        self.opt.grad.zero_()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        keep = (self.opt.flat.clone(), self.opt.m.clone(), self.opt.v.clone(), self.flat16.clone(), self.guard.clone())
        with torch.cuda.stream(s):
            self._run()
            self._optimizer()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.guard.copy_(keep[4]); self.guard_host.copy_(self.guard)      # the warm-up pass is not a step: undo its guard record
        self.opt.flat.copy_(keep[0]); self.opt.m.copy_(keep[1]); self.opt.v.copy_(keep[2]); self.flat16.copy_(keep[3])
        self.opt.grad.zero_()
        graphs = []
        for name, fn in (("march", self._run_march), ("optimizer", lambda: self._optimizer(sched_off=6)), ("field", self._run_field)):
            g = self._new_graph()
            with torch.cuda.graph(g):
                fn()
            self._census(name, g)
            graphs.append(g)
        self.graph = tuple(graphs)

    def flush(self):
        """deferred mode: apply the pending optimizer pass of the last step (call before reading parameters / evaluating)"""
        if (self.defer or self.defer_multi) and self.pending:
            if self.nccl:
                self.tr.comm.allreduce_sum_(self.opt.grad)
            if self.sumsq_tail and self.defer:
                self._adam_only(sched_off=0)      # the replay's tail already took the norm of this gradient
            else:
                self._optimizer(sched_off=0)      # slot 0 still holds the last step's schedule
            self.pending = False

    def use_pixel_batches(self, on=True):
        """End-to-end input form (what a DataLoader hands over, base.py:94-173): ONE record of 28 bytes per ray
        [img_idx i64 (R) | pix_idx i64 (R) | target rgb f32 (R,3)] = a single host->device copy per step into ``self.batch``;
        ray generation runs inside the step graph.  ``step_pixels(record)`` copies and steps.  Needs trainer.set_cameras()."""
        R = self.R
        if on:
            if self.tr.poses is None or self.tr.directions is None:
                raise RuntimeError("use_pixel_batches: call trainer.set_cameras(poses, directions) first")
            if getattr(self, "batch", None) is None:
                self.batch = torch.zeros(R * 28, dtype=torch.uint8, device=self.dev)
                self.b_img = self.batch[:8 * R].view(torch.int64)
                self.b_pix = self.batch[8 * R:16 * R].view(torch.int64)
            self.target = self.batch[16 * R:].view(torch.float32).view(R, 3)
        else:
            self.target = self.inp[2]
        if on != self.pix_inputs:
            self.pix_inputs = on
            self.graph = None                 # the captured sequence and the target pointer change

    STRATEGIES = {"all_images_triang_patch": 0, "same_image_triang_patch": 1, "all_images_triang": 2, "same_image_triang": 3}

    def use_device_sampling(self, images, height, width, strategy="all_images_triang_patch", patch_size=8, sem_labels=None, seed=0,
                            max_expand=0):
        """Draw every batch on the device (BaseDataset.__getitem__, datasets/base.py:94-183):
        images (P, H*W, 3) f32 resident in HBM (the reference keeps `rays` there too, train_nerf.py:239-240), optional
        sem_labels (P, H*W) i64.  Indices, target gather and ray generation become the first nodes of the step graph;
        ``step()`` then needs no input.  Also sets the triangle topology the strategy implies (losses.py:294-313)."""
        images = images.to(self.dev, torch.float32).contiguous()
        P = images.shape[0]
        if images.shape != (P, height * width, 3):
            raise RuntimeError("use_device_sampling: images must be (P, H*W, 3)")
        if sem_labels is not None:
            if not self.n_cls:
                raise RuntimeError("use_device_sampling: sem_labels given but the model has no semantic head")
            sem_labels = sem_labels.to(self.dev, torch.int64).contiguous()
        self.use_pixel_batches(True)
        n_random = 0
        if self.rtp:
            # the generated poses follow the P training poses in the trainer's pose table (trainer.set_cameras(random_poses=...)),
            # as in the reference's torch.cat((poses, random_poses[rnd_img_idxs])) (train_nerf.py:170)
            n_random = int(getattr(self.tr, "n_random_poses", 0))
            if n_random < 1 or int(self.tr.n_train_poses) != P:
                raise RuntimeError("use_device_sampling: random_tr_poses needs trainer.set_cameras(poses, directions, random_poses=...) with as "
                                   "many training poses as images")
            group = patch_size * patch_size if strategy.endswith("patch") else 3
            if self.n_gt % group:
                raise RuntimeError(f"use_device_sampling: random_tr_poses needs batch_size / 2 to be a whole number of {group}-ray groups")
        self.dev_sampling = dict(images=images, labels=sem_labels, H=int(height), W=int(width), n_poses=P, patch=int(patch_size),
                                 max_expand=int(max_expand),      # triang_max_expand of the triangle strategies (base.py:130-141)
                                 n_random=n_random,
                                 strategy=self.STRATEGIES[strategy], seed=torch.full((1,), int(seed), dtype=torch.int64, device=self.dev))
        self.set_triangles(self.batch_triangles(self.R - self.u0, strategy, patch_size))
        self.graph = None

    @staticmethod
    def batch_triangles(n_rays, strategy, patch_size=8):
        """(3, M) ray indices (x1, x2, x3) of every triangle of a batch: losses.py:294-299 for the triangle strategies,
        :301-313 with the local offsets of datasets/base.py:51-58 (x1 = (i,j), x2 = (i-1,j), x3 = (i,j-1), i,j >= 1) for patches."""
        if strategy in ("all_images_triang", "same_image_triang"):
            t = torch.arange(n_rays // 3 * 3).view(-1, 3)
            return t.t().contiguous()
        p = patch_size
        loc = torch.arange(p * p).view(p, p)
        offs = torch.stack([loc[1:, 1:].reshape(-1), loc[:-1, 1:].reshape(-1), loc[1:, :-1].reshape(-1)])       # (3, (p-1)^2)
        base = (torch.arange(n_rays // (p * p)) * p * p).view(1, -1, 1)
        return (base + offs.view(3, 1, -1)).reshape(3, -1).contiguous()

    @staticmethod
    def pack_pixel_batch(img_idx, pix_idx, rgb, pin=True, rnd_img_idx=None, n_train_poses=None):
        """host-side record for step_pixels: uint8 (R*28) = [img_idx i64 | pix_idx i64 | rgb f32].
        random_tr_poses: pass the batch of the training views (n_gt rows) plus rnd_img_idx (n_gt) and n_train_poses; the record
        then holds R = 2 n_gt rows - the same pixels again with image index n_train_poses + rnd_img_idx (the row of the generated
        pose in trainer.poses, train_nerf.py:169-172) and zero colours that no loss term reads."""
        if rnd_img_idx is not None:
            img_idx = torch.cat([img_idx.to(torch.int64), rnd_img_idx.to(torch.int64) + int(n_train_poses)])
            pix_idx = torch.cat([pix_idx.to(torch.int64), pix_idx.to(torch.int64)])
            rgb = torch.cat([rgb.to(torch.float32), torch.zeros_like(rgb, dtype=torch.float32)])
        rec = torch.cat([img_idx.to(torch.int64).contiguous().view(torch.uint8).reshape(-1),
                         pix_idx.to(torch.int64).contiguous().view(torch.uint8).reshape(-1),
                         rgb.to(torch.float32).contiguous().view(torch.uint8).reshape(-1)])
        return rec.pin_memory() if pin and not rec.is_cuda else rec

    def step_pixels(self, record, **kw):
        """one training step from a packed pixel batch (host pinned or device): one copy + one graph replay"""
        if not self.pix_inputs:
            self.use_pixel_batches(True)
        self.batch.copy_(record, non_blocking=True)
        self.step(**kw)

    def rays_from_pixels(self, img_idx, pix_idx):
        """fill the static rays_o / rays_d buffers from (image, pixel) indices (one kernel)"""
        tr = self.tr
        check(self.L.ncn_rays_from_pixels(ptr(tr.poses), ptr(tr.directions), ptr(img_idx), ptr(pix_idx), self.R, ptr(self.rays_o),
                                          ptr(self.rays_d), torch.cuda.current_stream().cuda_stream), "rays_from_pixels")

    def _update_grid_impl(self, thr, decay=0.95):
        """models/ngp_mt.py:339-368 (warmup=False) on the step's arenas: one sampling kernel, the encoder + density trunk
        through the C-ABI (no module temporaries), one scatter, then decay/max + packbits.  Multi-cascade grids take the
        module path (ncn_b200.ngp.NGPMT.update_density_grid)."""
        m = self.model
        if m.cascades != 1:
            m.update_density_grid(thr, warmup=False, decay=decay)
            return
        L = self.L
        st = torch.cuda.current_stream().cuda_stream
        G = m.grid_size
        G3 = G ** 3
        M = G3 // 4
        if getattr(self, "grid_tmp", None) is None:
            dev = self.dev
            self.grid_tmp = torch.zeros(G3, dtype=torch.float32, device=dev)
            self.grid_idx = torch.empty(2 * M, dtype=torch.int32, device=dev)
            self.grid_xyz = torch.empty(2 * M, 3, dtype=torch.float32, device=dev)
            self.grid_seed = torch.full((1,), 20240531, dtype=torch.int64, device=dev)
            self.grid_stats = torch.zeros(2, dtype=torch.float32, device=dev)
        csum = torch.cumsum(m.density_grid[0] > thr, 0, dtype=torch.int32)
        self.grid_tmp.zero_()
        s = min(2.0 ** -1, float(m.scale))
        check(L.ncn_grid_sample_cells(ptr(csum), G, M, s, ptr(self.grid_seed), ptr(self.grid_idx), ptr(self.grid_xyz), st), "grid_sample_cells")
        enc, sg = m.xyz_encoder, m.sigma_net
        cap = self.cap
        for c0 in range(0, 2 * M, cap):
            n = min(cap, 2 * M - c0)
            last = c0 + n >= 2 * M
            check(L.ncn_grid_fwd(C.byref(enc.desc), ptr(self.grid_xyz[c0:]), ptr(self._w16("xyz_encoder")), n, ptr(self.feat), self.xform,
                                 None, st), "grid_fwd(update)")
            check(L.ncn_mlp_fwd(C.byref(sg.desc), ptr(self.feat), ptr(self._w16("sigma_net")), n, ptr(self.h), None, None, st), "sigma_fwd(update)")
            check(L.ncn_grid_scatter_density(ptr(self.h), 16, ptr(self.grid_idx[c0:]), n, ptr(self.grid_tmp),
                                             ptr(self.grid_seed) if last else None, st), "grid_scatter")
        self.grid_stats.zero_()
        check(L.ncn_density_grid_update(ptr(m.density_grid), ptr(self.grid_tmp), G3, float(decay), ptr(self.grid_stats), st), "density_grid_update")
        check(L.ncn_packbits_auto(ptr(m.density_grid), m.density_bitfield.numel(), ptr(self.grid_stats), float(thr),
                                  ptr(m.density_bitfield), st), "packbits_auto")

    def update_grid(self, restore=None):
        """occupancy-grid upkeep (models/ngp_mt.py:339-368), replayed as its own CUDA graph (steady state: warmup=False)"""
        hp = self.hp
        thr = 0.01 * hp["rend_max_samples"] / 3 ** 0.5 * hp["density_tresh_decay"]
        if not self.use_graph:
            self._update_grid_impl(thr)
            if restore is not None:
                self.model.density_grid.copy_(restore[0]); self.model.density_bitfield.copy_(restore[1])
            return
        if getattr(self, "grid_graph", None) is None:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._update_grid_impl(thr)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = self._new_graph()
            with torch.cuda.graph(g):
                self._update_grid_impl(thr)
                if restore is not None:
                    self.model.density_grid.copy_(restore[0]); self.model.density_bitfield.copy_(restore[1])
            self._census("grid_update", g)
            self.grid_graph = g
        self.grid_graph.replay()

    def _capture(self, multi):
        # warm-up on a side stream (sets function attributes, touches every buffer), then capture
        self.flush()                              # a gradient still waiting for its update must not be dropped by the re-capture
        if self.peer is not None and self.tr.world_size > 1:
            import torch.distributed as dist
            dist.barrier()                        # the peer step waits for every rank on the device: enter it together
        self.opt.grad.zero_()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        keep = (self.opt.flat.clone(), self.opt.m.clone(), self.opt.v.clone(), self.flat16.clone(), self.guard.clone())
        with torch.cuda.stream(s):
            self._run()
            self._optimizer()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.guard.copy_(keep[4]); self.guard_host.copy_(self.guard)      # the warm-up pass is not a step: undo its guard record
        # undo the warm-up update
        self.opt.flat.copy_(keep[0]); self.opt.m.copy_(keep[1]); self.opt.v.copy_(keep[2]); self.flat16.copy_(keep[3])
        self.opt.grad.zero_()
        if self.defer:
            g0 = self._new_graph()
            with torch.cuda.graph(g0):
                self._run_deferred(multi)
            self._census("step", g0)
            self.graph = (g0, None)
            return
        g0 = self._new_graph()
        with torch.cuda.graph(g0):
            self._run()
            if not multi:
                self._optimizer()
        self._census("step", g0)
        g1 = None
        if multi:
            g1 = self._new_graph()
            with torch.cuda.graph(g1):
                self._optimizer()
            self._census("optimizer", g1)
        self.graph = (g0, g1)

    def stats_host(self):
        """(loss dict, n_samples) - synchronises; call outside the timed region"""
        self.flush()
        torch.cuda.synchronize()
        R = self.R
        z = self.zeros.cpu()
        d = {"rgb": float(z[0]) / (3 * self.n_gt), "opacity": float(self.hp["loss_opacity_w"]) * float(z[1]) / R}
        if self.M > 0:
            l = torch.nan_to_num(self.losses.cpu())
            w = self.dev_sched[3:6].cpu() / GSCALE
            d.update(norm_D_C_ort_dot=float(w[0] * l[0]), norm_D_C_centr_dot=float(w[1] * l[1]), norm_D_C_centr_L1=float(w[2] * l[2]))
        if self.n_cls and self.sem_w > 0:
            d["sem"] = self.sem_w * float(z[4]) / float(z[5]) if float(z[5]) > 0 else 0.0
        d["total"] = sum(d.values())
        return d, int(self.counter[0])
