"""GPU: the fused step from global step 0 through the occupancy warm-up (train_nerf.py:315-320: all G^3 cells every 16 steps for
256 steps), and the sample-arena overflow guard.

Right after the step-0 grid update about half of the cells are occupied, so the march yields hundreds of samples per ray -
the regime the reference handles by sizing its sample arrays after a host sync (raymarching.cu:302-305, worst case
N_rays * max_samples).  The fused step must (a) survive it with the default (worst-case) arena, loss going down on a textured
synthetic room, zero overflow; (b) with a deliberately small arena never produce a silently truncated step: the overflowed
step is skipped on the device (parameters untouched, gradient zeroed), the host regrows and re-captures, training continues.
"""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H, W, P = 768, 1024, 8


def _textured_room(tr):
    """target images of a synthetic scene: the walls of the room [-0.4,0.4]^3 with a smooth colour texture, seen by P cameras"""
    from ncn_b200 import synth
    poses = torch.from_numpy(synth.camera_poses(P, 0)).cuda()
    dirs = torch.from_numpy(synth.pixel_directions("hypersim")).cuda()
    tr.set_cameras(poses, dirs)
    images = torch.empty(P, H * W, 3, device="cuda")
    pix = torch.arange(H * W, device="cuda")
    for i in range(P):
        ro, rd = tr.rays_from_batch(torch.full((H * W,), i, device="cuda"), pix)
        t = torch.where(rd > 0, (0.4 - ro) / rd, (-0.4 - ro) / rd).min(-1)[0]
        p = ro + rd * t[:, None]
        images[i] = 0.5 + 0.5 * torch.sin(9.0 * p + torch.tensor([0.0, 2.0, 4.0], device="cuda"))
    return images


def _trainer(R, seed=0):
    from ncn_b200.trainer import NeRFTrainer
    torch.manual_seed(seed)
    tr = NeRFTrainer(dict(batch_size=R), device="cuda")
    assert tr.global_step == 0 and float(tr.model.density_grid.abs().max()) == 0.0
    return tr


def test_fused_training_from_step_zero_through_the_warmup():
    R = 4096
    tr = _trainer(R)
    images = _textured_room(tr)
    fs = tr.fused_step()                                   # default arena: worst case, overflow impossible
    assert fs.cap == R * tr.hp["rend_max_samples"]
    fs.use_device_sampling(images, H, W, strategy="all_images_triang_patch", seed=1)
    hist, spr = [], []
    n_steps = 320                                          # 256 warm-up steps (16 all-cell grid updates) + 4 steady-state updates
    for i in range(n_steps):
        tr.train_step_fused()
        if i % 16 == 15:
            d, n = fs.stats_host()
            hist.append(d["rgb"]); spr.append(n / R)
            assert np.isfinite(d["total"])
    assert tr.global_step == n_steps
    assert fs.overflow_steps == 0
    assert max(spr) > 64, spr                              # the old fixed 64-rows-per-ray arena would have truncated these steps
    assert hist[-1] < 0.35 * hist[0], hist                 # the photometric loss really goes down
    assert torch.isfinite(tr.opt.flat).all()
    # the occupancy grid has been carved: fewer cells are occupied than right after the first warm-up update
    bits = tr.model.density_bitfield
    occ = float(sum(((bits >> b) & 1).sum() for b in range(8))) / (128 ** 3)
    assert 0.0 < occ < 0.6, occ


def test_overflowed_step_is_skipped_not_truncated_and_the_arena_regrows():
    R = 1024
    tr = _trainer(R)
    images = _textured_room(tr)
    fs = tr.fused_step(capacity_per_ray=8)                 # far too small for the dense warm-up grid
    fs.use_device_sampling(images, H, W, strategy="all_images_triang_patch", seed=2)
    p0 = tr.opt.flat.clone(); p16 = tr.opt.flat16.clone()
    with warnings.catch_warnings(record=True) as wrn:
        warnings.simplefilter("always")
        tr.train_step_fused()                              # step 0: grid update (warm-up) + a march that overflows the arena
        fs.flush(); torch.cuda.synchronize()
        need = int(fs.counter[0])
        assert need > fs.cap == 8 * R
        assert int(fs.guard[0]) == 1 and int(fs.guard[1]) == 1 and int(fs.guard[2]) == need
        # skipped: parameters, fp16 working copy and Adam state untouched; the poisoned gradient was zeroed
        assert torch.equal(tr.opt.flat, p0) and torch.equal(tr.opt.flat16, p16)
        assert float(tr.opt.m.abs().max()) == 0.0 and float(tr.opt.grad.abs().max()) == 0.0
        for _ in range(12):                                # the host notices on the next step, regrows and re-captures
            tr.train_step_fused()
        fs.flush(); torch.cuda.synchronize()
    assert any("arena overflow" in str(w.message) for w in wrn)
    assert fs.cap >= need and fs.cap > 8 * R
    assert int(fs.guard[0]) == 0                            # the last step fitted
    assert fs.overflow_steps >= 1
    assert not torch.equal(tr.opt.flat, p0) and torch.isfinite(tr.opt.flat).all()
    d, n = fs.stats_host()
    assert np.isfinite(d["total"]) and n <= fs.cap


def test_guard_kernel_states():
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    L = _lib.lib()
    counter = torch.tensor([100, 7], dtype=torch.int32, device="cuda")
    state = torch.zeros(3, dtype=torch.int32, device="cuda")
    grad = torch.ones(4, device="cuda")
    check(L.ncn_step_guard(ptr(counter), 100, ptr(state), ptr(grad), stream()))          # == capacity: fits
    assert state.tolist() == [0, 0, 100] and torch.equal(grad, torch.ones(4, device="cuda"))
    counter[0] = 101
    check(L.ncn_step_guard(ptr(counter), 100, ptr(state), ptr(grad), stream()))
    assert state.tolist() == [1, 1, 101] and torch.isnan(grad[0]) and float(grad[1]) == 1.0
    counter[0] = 50
    check(L.ncn_step_guard(ptr(counter), 100, ptr(state), None, stream()))
    assert state.tolist() == [0, 1, 101]
