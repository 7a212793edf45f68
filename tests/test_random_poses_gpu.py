"""GPU: the --random_tr_poses form of the batch and of the photometric terms (datasets/base.py:106-126,
148-159, train_nerf.py:169-172, losses.py:265-297): the *_gt kernels against torch, the device sampler's second half, and a
CUDA-graph step whose only input is the device sampler.  (The whole step against the reference's own render + NeRFMTLoss lives in
tests/test_reference_on_shims_gpu.py::test_random_tr_poses_reference_vs_mirror_vs_fused.)"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("Ct,n_gt", [(3, 700), (6, 1024), (9, 0), (3, 1024)])
def test_photometric_gt_kernel_matches_torch(Ct, n_gt):
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    L = _lib.lib()
    R, w_op, gs = 1024, 1e-3, 64.0
    g = torch.Generator(device="cuda").manual_seed(Ct * 7 + n_gt)
    rend = torch.rand(R, Ct, device="cuda", generator=g)
    op = torch.rand(R, device="cuda", generator=g).clamp(0.01, 0.99)
    tgt = torch.rand(max(n_gt, 1), 3, device="cuda", generator=g)
    bg = (C.c_float * 3)(1.0, 0.5, 0.25)
    sums = torch.zeros(2, device="cuda"); rgb = torch.empty(R, 3, device="cuda")
    d_rend = torch.full((R, Ct), 7.0, device="cuda"); d_op = torch.full((R,), 7.0, device="cuda")
    check(L.ncn_photometric_loss_gt(ptr(rend), ptr(op), ptr(tgt) if n_gt else None, R, n_gt, Ct, bg, w_op, gs, ptr(rgb), ptr(sums), ptr(d_rend),
                                    ptr(d_op), stream()))
    # torch: losses.py:347-361 on rendering.py:231-241's rgb
    rend_t = rend.clone().double().requires_grad_(True); op_t = op.clone().double().requires_grad_(True)
    bgt = torch.tensor([1.0, 0.5, 0.25], device="cuda", dtype=torch.float64)
    rgb_t = rend_t[:, :3] + bgt * (1 - op_t[:, None])
    l_rgb = ((rgb_t[:n_gt] - tgt[:n_gt].double()) ** 2).mean() if n_gt else rgb_t.sum() * 0
    o = op_t + 1e-10
    l_op = w_op * (-o * torch.log(o)).mean()
    ((l_rgb + l_op) * gs).backward()
    torch.testing.assert_close(rgb, rgb_t.detach().float(), rtol=1e-6, atol=1e-6)
    if n_gt:
        assert abs(float(sums[0]) / (3 * n_gt) - float(l_rgb)) <= 1e-5 * float(l_rgb)
    else:
        assert float(sums[0]) == 0.0
    assert abs(w_op * float(sums[1]) / R - float(l_op)) <= 1e-5 * float(l_op)
    torch.testing.assert_close(d_rend, rend_t.grad.float(), rtol=1e-4, atol=1e-9)
    # dL/dopacity = -sum_c bg_c dL/drgb_c + the entropy term: the two cancel to any degree, so the bound is absolute (fp32 rounding of
    # terms of magnitude gs * 2 |e| / (3 n_gt))
    torch.testing.assert_close(d_op, op_t.grad.float(), rtol=1e-4, atol=1e-7 * gs)
    assert float(d_rend[n_gt:].abs().max() if n_gt < R else 0.0) == 0.0


@pytest.mark.parametrize("Ct", [3, 6, 9])
def test_fused_composite_photometric_gt_equals_the_two_calls(Ct, scene):
    """ncn_composite_train_fw_photometric_gt == ncn_composite_train_fw + ncn_photometric_loss_gt (same arithmetic, one launch)"""
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib, synth, vren
    from ncn_b200._lib import check, ptr, stream
    L = _lib.lib()
    R, n_gt = 1024, 448
    b = synth.patch_batch(R, seed=2)
    ro, rd = torch.from_numpy(b["rays_o"]).cuda(), torch.from_numpy(b["rays_d"]).cuda()
    _, hits_t, _ = vren.ray_aabb_intersect(ro, rd, scene["center"], scene["half_size"], 1)
    noise = torch.rand(R, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    rays_a, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(ro, rd, hits_t[:, 0], scene["bitfield"], 1, 0.5, 0.0, noise, 128, 1024)
    N = int(counter[0])
    g = torch.Generator(device="cuda").manual_seed(3)
    sig = torch.rand(N, device="cuda", generator=g) * 30; raws = torch.rand(N, Ct, device="cuda", generator=g)
    tgt = torch.rand(n_gt, 3, device="cuda", generator=g)
    bg = (C.c_float * 3)(1.0, 1.0, 1.0)
    mk = lambda: dict(ts_=torch.empty(R, dtype=torch.int64, device="cuda"), op=torch.empty(R, device="cuda"), dp=torch.empty(R, device="cuda"),
                      rend=torch.empty(R, Ct, device="cuda"), ws=torch.zeros(N, device="cuda"), rgb=torch.empty(R, 3, device="cuda"),
                      sums=torch.zeros(2, device="cuda"), d_rend=torch.empty(R, Ct, device="cuda"), d_op=torch.empty(R, device="cuda"))
    a, c = mk(), mk()
    check(L.ncn_composite_train_fw(ptr(sig), ptr(raws), ptr(deltas), ptr(ts), ptr(rays_a), 1e-4, R, N, Ct, ptr(a["ts_"]), ptr(a["op"]), ptr(a["dp"]),
                                   ptr(a["rend"]), ptr(a["ws"]), stream()))
    check(L.ncn_photometric_loss_gt(ptr(a["rend"]), ptr(a["op"]), ptr(tgt), R, n_gt, Ct, bg, 1e-3, 8.0, ptr(a["rgb"]), ptr(a["sums"]), ptr(a["d_rend"]),
                                    ptr(a["d_op"]), stream()))
    check(L.ncn_composite_train_fw_photometric_gt(ptr(sig), ptr(raws), ptr(deltas), ptr(ts), ptr(rays_a), 1e-4, R, N, Ct, ptr(c["ts_"]), ptr(c["op"]),
                                                  ptr(c["dp"]), ptr(c["rend"]), ptr(c["ws"]), ptr(tgt), n_gt, bg, 1e-3, 8.0, ptr(c["rgb"]),
                                                  ptr(c["sums"]), ptr(c["d_rend"]), ptr(c["d_op"]), stream()))
    assert torch.equal(a["ts_"], c["ts_"]) and torch.equal(a["ws"], c["ws"])
    for k in ("op", "dp", "rend", "rgb", "d_rend", "d_op"):
        torch.testing.assert_close(a[k], c[k], rtol=1e-5, atol=1e-7, msg=lambda m, k=k: f"{k}: {m}")
    torch.testing.assert_close(a["sums"], c["sums"], rtol=1e-4, atol=1e-6)
    assert float(c["d_rend"][n_gt:].abs().max()) == 0.0 and float(c["d_rend"][:n_gt, :3].abs().max()) > 0


@pytest.mark.parametrize("strategy", ["all_images_triang_patch", "same_image_triang_patch", "all_images_triang", "same_image_triang"])
def test_random_pose_half_of_the_device_sampler(strategy):
    """ncn_sample_random_pose_half: rows [n_gt, 2 n_gt) = the same pixels, pose rows P + U[0, Q) per patch / triangle (one for the
    whole batch with the same_image strategies) - datasets/base.py:106-126, 148-159, train_nerf.py:169-172"""
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    from ncn_b200.fused import FusedStep
    L = _lib.lib()
    H, W, P, Q, p = 48, 64, 7, 23, 8
    sid = FusedStep.STRATEGIES[strategy]
    group = p * p if sid <= 1 else 3
    n_gt = group * 400
    seed = torch.full((1,), 5, dtype=torch.int64, device="cuda")
    img = torch.full((2 * n_gt,), -1, dtype=torch.int64, device="cuda"); pix = torch.full((2 * n_gt,), -1, dtype=torch.int64, device="cuda")
    check(L.ncn_sample_ray_batch(sid, ptr(seed), n_gt, P, H, W, p, ptr(img), ptr(pix), stream()))
    assert int(img[n_gt:].max()) == -1                                 # the first call draws the first half only
    check(L.ncn_sample_random_pose_half(sid, ptr(seed), n_gt, Q, P, p, ptr(img), ptr(pix), stream()))
    assert int(seed) == 6                                              # read, not advanced
    assert torch.equal(pix[n_gt:], pix[:n_gt])
    rnd = img[n_gt:].view(-1, group)
    assert int(rnd.min()) >= P and int(rnd.max()) < P + Q and bool((rnd == rnd[:, :1]).all())
    assert int(img[:n_gt].max()) < P
    if sid & 1:
        assert bool((rnd == rnd[0, 0]).all())
    else:
        cnt = torch.bincount(rnd[:, 0] - P, minlength=Q).float()
        assert float(cnt.min()) > 0.2 * float(cnt.mean()) and float(cnt.max()) < 2.2 * float(cnt.mean())
        # independent of the training-image draw of the same patch
        joint = torch.bincount(img[:n_gt].view(-1, group)[:, 0] * Q + (rnd[:, 0] - P), minlength=P * Q)
        assert int((joint > 0).sum()) > 0.7 * P * Q
    firsts = {int(rnd[0, 0])}
    img2 = img.clone()
    for _ in range(8):                                                 # the following batches draw other poses
        check(L.ncn_sample_ray_batch(sid, ptr(seed), n_gt, P, H, W, p, ptr(img2), ptr(pix), stream()))
        check(L.ncn_sample_random_pose_half(sid, ptr(seed), n_gt, Q, P, p, ptr(img2), ptr(pix), stream()))
        firsts.add(int(img2[n_gt]))
        assert torch.equal(pix[n_gt:], pix[:n_gt])
    assert len(firsts) >= 4 and int(seed) == 14


def test_graph_step_fed_by_the_device_sampler_with_random_poses():
    """a CUDA-graph training step with random_tr_poses whose only input is the device sampler: first half = training views with
    gathered targets, second half = the same pixels seen from poses generated by ncn_b200.batches.generate_random_poses"""
    import ncn_b200  # noqa: F401
    from ncn_b200 import batches, synth, vren
    from ncn_b200.trainer import NeRFTrainer
    R, P, Q = 1536, 50, 200
    torch.manual_seed(4)
    tr = NeRFTrainer(dict(batch_size=R, random_tr_poses=True), device="cuda")
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    tr.model.density_grid.copy_(torch.from_numpy(grid).cuda())
    vren.packbits(tr.model.density_grid, 5.9, tr.model.density_bitfield)
    n = tr.model.xyz_encoder.params.numel()
    tr.opt.flat[:n].copy_(torch.randn(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) * 0.3)
    tr.opt.flat16.copy_(tr.opt.flat)
    tr.global_step = 3000
    poses = synth.camera_poses(P, 0)
    pos = poses[:, :3, 3]
    rnd_poses, _ = batches.generate_random_poses(poses, pos.min(0), pos.max(0), Q, rng=np.random.RandomState(0))
    dirs = torch.from_numpy(synth.pixel_directions("hypersim")).cuda()
    tr.set_cameras(torch.from_numpy(poses), dirs, random_poses=rnd_poses)
    assert tr.poses.shape == (P + Q, 3, 4) and tr.n_train_poses == P and tr.n_random_poses == Q
    fs = tr.fused_step(use_graph=True)
    imgs = torch.rand(P, 768 * 1024, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    fs.use_device_sampling(imgs, 768, 1024, strategy="all_images_triang_patch", patch_size=8, seed=3)
    n_gt = R // 2
    assert fs.M == n_gt // 64 * 49 and fs.n_gt == n_gt
    p0 = tr.opt.flat.clone()
    fs.step(); torch.cuda.synchronize()
    b_img, b_pix = fs.b_img.clone(), fs.b_pix.clone()
    assert int(b_img[:n_gt].max()) < P and int(b_img[n_gt:].min()) >= P and int(b_img[n_gt:].max()) < P + Q
    assert torch.equal(b_pix[n_gt:], b_pix[:n_gt])
    assert torch.equal(fs.target[:n_gt], imgs[b_img[:n_gt], b_pix[:n_gt]])
    ro, rd = tr.rays_from_batch(b_img, b_pix)                          # get_rays (ray_utils.py:46-71) on the concatenated pose table
    torch.testing.assert_close(fs.rays_o, ro, rtol=0, atol=0)
    torch.testing.assert_close(fs.rays_d, rd, rtol=1e-6, atol=1e-7)
    assert torch.equal(fs.rays_o[n_gt:], tr.poses[b_img[n_gt:], :, 3])  # the generated cameras' positions
    fs.step(); torch.cuda.synchronize()
    assert not torch.equal(b_img[n_gt:], fs.b_img[n_gt:])
    d, n_s = fs.stats_host()
    assert all(np.isfinite(v) for v in d.values()) and n_s > R and d["rgb"] > 0 and "norm_D_C_centr_dot" in d
    assert float((tr.opt.flat - p0).abs().max()) > 0                   # the optimizer moved the parameters


def test_host_batcher_feeds_step_pixels():
    """the host data path of random_tr_poses end to end: HostBatcher (reference index draw + target gather + pinned record) ->
    FusedStep.step_pixels (one H2D copy + one graph replay); the step sees exactly the batch the host drew"""
    import ncn_b200  # noqa: F401
    from ncn_b200 import batches, synth, vren
    from ncn_b200.trainer import NeRFTrainer
    R, P, Q, H, W = 1024, 6, 40, 768, 1024
    torch.manual_seed(5)
    tr = NeRFTrainer(dict(batch_size=R, random_tr_poses=True), device="cuda")
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    tr.model.density_grid.copy_(torch.from_numpy(grid).cuda())
    vren.packbits(tr.model.density_grid, 5.9, tr.model.density_bitfield)
    n = tr.model.xyz_encoder.params.numel()
    tr.opt.flat[:n].copy_(torch.randn(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) * 0.3)
    tr.opt.flat16.copy_(tr.opt.flat)
    tr.global_step = 3000
    poses = synth.camera_poses(P, 0)
    rnd_poses, _ = batches.generate_random_poses(poses, poses[:, :3, 3].min(0), poses[:, :3, 3].max(0), Q, rng=np.random.RandomState(0))
    tr.set_cameras(torch.from_numpy(poses), torch.from_numpy(synth.pixel_directions("hypersim")).cuda(), random_poses=rnd_poses)
    images = torch.rand(P, H * W, 3, generator=torch.Generator().manual_seed(0))
    strategy = "all_images_triang_patch"
    hb = batches.HostBatcher(images, H, W, strategy, R, patch_size=8, random_tr_poses=True, n_random_poses=Q, ring=3,
                             rng=np.random.RandomState(3))
    assert hb.n_rays == R and hb.n_gt == R // 2 and hb.ring[0].is_pinned()
    fs = tr.fused_step(use_graph=True)
    fs.use_pixel_batches(True)
    fs.set_triangles(fs.batch_triangles(R - fs.u0, strategy, 8))
    for it in range(3):
        rec, _ = hb.next()
        fs.step_pixels(rec)
        torch.cuda.synchronize()
        img, pix, rgb = hb.views(rec)
        assert torch.equal(fs.b_img.cpu(), img) and torch.equal(fs.b_pix.cpu(), pix)
        assert torch.equal(fs.target[:R // 2].cpu(), images[img[:R // 2], pix[:R // 2]])
        assert int(img[R // 2:].min()) >= P and torch.equal(pix[R // 2:], pix[:R // 2])
        ro, rd = tr.rays_from_batch(fs.b_img, fs.b_pix)
        torch.testing.assert_close(fs.rays_o, ro, rtol=0, atol=0)
        torch.testing.assert_close(fs.rays_d, rd, rtol=1e-6, atol=1e-7)
    d, n_s = fs.stats_host()
    assert all(np.isfinite(v) for v in d.values()) and n_s > R and d["rgb"] > 0
