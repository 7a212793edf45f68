"""GPU: the fused CUDA-graph training step (ncn_b200.fused.FusedStep) against the module path
(render + NeRFMTLoss + autograd) on identical parameters, rays and march noise: same losses, same gradients
(the module path is itself pinned kernel-by-kernel against the oracles / reference kernels)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(R=2048, seed=0, hp=None, n_sem_cls=0):
    import ncn_b200
    from ncn_b200 import synth, vren
    from ncn_b200.trainer import NeRFTrainer
    torch.manual_seed(seed)
    tr = NeRFTrainer(dict(batch_size=R, **(hp or {})), device="cuda", n_sem_cls=n_sem_cls)
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    tr.model.density_grid.copy_(torch.from_numpy(grid).cuda())
    vren.packbits(tr.model.density_grid, 5.9, tr.model.density_bitfield)
    # non-degenerate field: larger table values so densities / colours vary
    g = torch.Generator(device="cuda").manual_seed(1)
    o, n = 0, tr.model.xyz_encoder.params.numel()
    tr.opt.flat[:n].copy_(torch.randn(n, device="cuda", generator=g) * 0.3)
    tr.opt.flat16.copy_(tr.opt.flat)
    tr.global_step = 3000
    b = synth.patch_batch(R, seed=seed)
    rays_o = torch.from_numpy(b["rays_o"]).cuda(); rays_d = torch.from_numpy(b["rays_d"]).cuda()
    tri = torch.from_numpy(b["tri"]).cuda()
    rgb = torch.rand(R, 3, device="cuda", generator=g)
    target = {"rgb": rgb, "patch_area": 64, "x1_offsets_local": tri[0][:49] % 64, "x2_offsets_local": tri[1][:49] % 64,
              "x3_offsets_local": tri[2][:49] % 64}
    return tr, rays_o, rays_d, tri, rgb, target


@pytest.mark.parametrize("fuse_fwd", [True, False, "mlp"])
def test_fused_matches_module_path(fuse_fwd):
    from ncn_b200.fused import GSCALE
    tr, rays_o, rays_d, tri, rgb, target = _setup()
    R = rays_o.shape[0]
    # module path
    torch.manual_seed(123)
    results, loss_d = tr.forward_loss(rays_o, rays_d, target)
    (loss_d["total"] * tr.hp["loss_scale"]).backward()
    g_mod = (tr.opt.grad / tr.hp["loss_scale"]).clone()
    tr.opt.grad.zero_()
    # fused path, eager (no graph), same noise
    torch.manual_seed(123)
    noise = torch.rand(R, device="cuda")
    fs = tr.fused_step(use_graph=False, fuse_fwd=fuse_fwd)
    fs.set_triangles(tri)
    fs.rays_o.copy_(rays_o); fs.rays_d.copy_(rays_d); fs.target.copy_(rgb); fs.noise.copy_(noise)
    fs.gen_noise = False
    fs._schedule()
    fs._run()
    torch.cuda.synchronize()
    assert int(fs.counter[0]) == int(results["rm_samples"])
    assert torch.equal(fs.rays_a, results["rays_a"])
    torch.testing.assert_close(fs.depth, results["depth"].detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(fs.opacity, results["opacity"].detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(fs.rgb, results["rgb"].detach().float(), rtol=1e-3, atol=1e-3)
    d, n = fs.stats_host()
    for k in ("rgb", "opacity", "norm_D_C_ort_dot", "norm_D_C_centr_dot", "norm_D_C_centr_L1"):
        assert abs(d[k] - float(loss_d[k])) <= 2e-3 * abs(float(loss_d[k])) + 1e-7, (k, d[k], float(loss_d[k]))
    g_fus = tr.opt.grad.clone()
    for name in ("rgb_net", "sigma_net", "xyz_encoder"):
        o, k = fs.off[name]
        a, b = g_fus[o:o + k], g_mod[o:o + k]
        assert torch.isfinite(a).all()
        rel = float((a - b).norm() / b.norm().clamp_min(1e-20))
        assert rel <= 3e-2, (name, rel, float(b.norm()))


@pytest.mark.parametrize("n_cls,fuse_fwd", [(3, "mlp"), (13, "mlp"), (3, False)])
def test_fused_extra_heads_match_module_path(n_cls, fuse_fwd):
    """config 3 (--pred_sem --pred_norm_nn, loss_sem_w > 0): semantic + normal heads as extra compositing channels, semantic
    cross-entropy on the rendered logits; n_cls = 3 takes the on-the-fly dL/dout rows (tcgen05 mode 1), 13 the head_dout path.
    The reference puts no loss on norm_nn, so norm_net's gradient must be exactly zero on both paths."""
    hp = dict(pred_sem=True, pred_norm_nn=True, loss_sem_w=4e-2)
    tr, rays_o, rays_d, tri, rgb, target = _setup(hp=hp, n_sem_cls=n_cls)
    R = rays_o.shape[0]
    g = torch.Generator(device="cuda").manual_seed(5)
    for name in ("sem_net", "norm_net"):                      # larger head weights so the logits are not all ~0
        p = getattr(tr.model, name).params
        p.data.copy_(torch.randn(p.numel(), device="cuda", generator=g) * 0.2)
    tr.opt.flat16.copy_(tr.opt.flat)
    sem = torch.randint(0, n_cls + 1, (R,), device="cuda", generator=g)          # 0 = void (ignored)
    target["semantics"] = sem
    torch.manual_seed(123)
    results, loss_d = tr.forward_loss(rays_o, rays_d, target)
    (loss_d["total"] * tr.hp["loss_scale"]).backward()
    g_mod = (tr.opt.grad / tr.hp["loss_scale"]).clone()
    tr.opt.grad.zero_()
    torch.manual_seed(123)
    noise = torch.rand(R, device="cuda")
    fs = tr.fused_step(use_graph=False, fuse_fwd=fuse_fwd)
    fs.set_triangles(tri)
    fs.rays_o.copy_(rays_o); fs.rays_d.copy_(rays_d); fs.target.copy_(rgb); fs.noise.copy_(noise); fs.sem_target.copy_(sem)
    fs.gen_noise = False
    fs._schedule()
    fs._run()
    torch.cuda.synchronize()
    assert int(fs.counter[0]) == int(results["rm_samples"])
    assert fs.Ct == 6 + n_cls
    torch.testing.assert_close(fs.depth, results["depth"].detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(fs.rend[:, 3:6], results["norm_nn"].detach().float(), rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(fs.rend[:, 6:], results["sem"].detach().float(), rtol=2e-3, atol=2e-3)
    d, n = fs.stats_host()
    for k in ("rgb", "opacity", "sem", "norm_D_C_ort_dot", "norm_D_C_centr_dot", "norm_D_C_centr_L1"):
        assert abs(d[k] - float(loss_d[k])) <= 2e-3 * abs(float(loss_d[k])) + 1e-7, (k, d[k], float(loss_d[k]))
    g_fus = tr.opt.grad.clone()
    for name in ("sem_net", "rgb_net", "sigma_net", "xyz_encoder"):
        o, k = fs.off[name]
        a, b = g_fus[o:o + k], g_mod[o:o + k]
        assert torch.isfinite(a).all()
        rel = float((a - b).norm() / b.norm().clamp_min(1e-20))
        assert rel <= 3e-2, (name, rel, float(b.norm()))
    o, k = fs.off["norm_net"]
    assert float(g_fus[o:o + k].abs().max()) == 0.0 and float(g_mod[o:o + k].abs().max()) == 0.0


def test_semantic_ce_kernel_matches_torch():
    """ncn_semantic_ce_loss vs torch.nn.CrossEntropyLoss(ignore_index=-1)(logits, label - 1) (losses.py:240-242): value,
    gradient, untouched neighbouring columns, and the all-void batch (NaN loss in the reference -> term dropped, zero gradient)"""
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(0)
    for R, n_cls, c_off, Ct in ((8192, 3, 6, 9), (1000, 40, 3, 45), (1, 1, 0, 1)):
        rend = torch.randn(R, Ct, device="cuda", generator=g) * 3
        lab = torch.randint(0, n_cls + 1, (R,), device="cuda", generator=g)
        sums = torch.zeros(2, device="cuda"); d = torch.full((R, Ct), 7.0, device="cuda")
        check(L.ncn_semantic_ce_loss(ptr(rend), Ct, c_off, n_cls, ptr(lab), R, 2.5, ptr(sums), ptr(d), stream()))
        x = rend[:, c_off:c_off + n_cls].clone().requires_grad_(True)
        n_valid = int((lab > 0).sum())
        assert int(sums[1]) == n_valid
        if n_valid == 0:
            assert float(d[:, c_off:c_off + n_cls].abs().max()) == 0.0
            continue
        ce = torch.nn.functional.cross_entropy(x, lab - 1, ignore_index=-1)
        (2.5 * ce).backward()
        torch.testing.assert_close(sums[0] / sums[1], ce.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(d[:, c_off:c_off + n_cls], x.grad, rtol=1e-4, atol=1e-8)
        keep = torch.ones(Ct, dtype=torch.bool, device="cuda"); keep[c_off:c_off + n_cls] = False
        assert bool((d[:, keep] == 7.0).all())
    lab = torch.zeros(64, dtype=torch.int64, device="cuda")
    rend = torch.randn(64, 9, device="cuda", generator=g); sums = torch.zeros(2, device="cuda"); d = torch.ones(64, 9, device="cuda")
    check(L.ncn_semantic_ce_loss(ptr(rend), 9, 6, 3, ptr(lab), 64, 1.0, ptr(sums), ptr(d), stream()))
    assert float(sums[1]) == 0.0 and float(d[:, 6:].abs().max()) == 0.0


def test_fused_graph_replay_trains():
    """graph replay == eager fused step, and the loss goes down over a few steps on a fixed batch"""
    tr, rays_o, rays_d, tri, rgb, target = _setup(R=1024, seed=1)
    tr2, *_ = _setup(R=1024, seed=1)
    noise = torch.rand(1024, device="cuda")
    fs = tr.fused_step(use_graph=True); fs.set_triangles(tri)
    fs2 = tr2.fused_step(use_graph=False); fs2.set_triangles(tri)
    for i in range(3):
        fs.step(rays_o, rays_d, rgb, noise=noise)
        fs2.step(rays_o, rays_d, rgb, noise=noise)
    fs.flush()
    torch.cuda.synchronize()
    rel = float((tr.opt.flat - tr2.opt.flat).norm() / tr2.opt.flat.norm())
    assert rel < 3e-3, rel      # run-to-run noise of the atomically accumulated backward through Adam's sign-like step (tools/check_peer.py)
    l0 = None
    for i in range(30):
        fs.step(rays_o, rays_d, rgb, noise=noise)
        if i == 0:
            l0 = fs.stats_host()[0]["rgb"]
    l1 = fs.stats_host()[0]["rgb"]
    assert l1 < l0, (l0, l1)


def test_rays_from_pixels_matches_get_rays():
    """ncn_rays_from_pixels == the reference's gather + get_rays (datasets/ray_utils.py:46-71) restated in torch"""
    from ncn_b200 import synth
    tr, *_ = _setup(R=1024, seed=2)
    poses = torch.from_numpy(synth.camera_poses(50, 0)).cuda(); dirs = torch.from_numpy(synth.pixel_directions("hypersim")).cuda()
    tr.set_cameras(poses, dirs)
    g = torch.Generator(device="cuda").manual_seed(0)
    img = torch.randint(0, 50, (1024,), device="cuda", generator=g); pix = torch.randint(0, 1024 * 768, (1024,), device="cuda", generator=g)
    fs = tr.fused_step(use_graph=False)
    fs.rays_from_pixels(img, pix)
    c2w = poses[img]
    rays_d = (dirs[pix][:, None, :] @ c2w[..., :3].transpose(1, 2))[:, 0]
    torch.testing.assert_close(fs.rays_d, rays_d, rtol=1e-6, atol=1e-7)
    assert torch.equal(fs.rays_o, c2w[..., 3])


def test_pixel_batch_step_equals_ray_step():
    """step_pixels (ONE packed [img | pix | rgb] record copied per step, ray generation as the first node of the step graph)
    against rays_from_pixels + step on a second trainer: identical rays, targets and sample counts, same losses; also from a
    pinned HOST record, and switching the mode off restores the (3,R,3) input buffer."""
    from ncn_b200 import synth
    R = 1024
    tr, _, _, tri, rgb, _ = _setup(R=R, seed=2)
    tr2, *_ = _setup(R=R, seed=2)
    poses = torch.from_numpy(synth.camera_poses(50, 0)).cuda(); dirs = torch.from_numpy(synth.pixel_directions("hypersim")).cuda()
    b = synth.patch_batch(R, seed=5)
    img = torch.from_numpy(b["img_idx"]).cuda(); pix = torch.from_numpy(b["pix_idx"]).cuda()
    noise = torch.rand(R, device="cuda")
    for t in (tr, tr2):
        t.set_cameras(poses, dirs)
    fs = tr.fused_step(use_graph=True); fs.set_triangles(tri)
    fs2 = tr2.fused_step(use_graph=True); fs2.set_triangles(tri)
    rec = fs.pack_pixel_batch(img.cpu(), pix.cpu(), rgb.cpu())
    assert rec.is_pinned() and rec.numel() == R * 28
    fs.step_pixels(rec, noise=noise)
    fs2.rays_from_pixels(img, pix); fs2.target.copy_(rgb)
    fs2.step(noise=noise)
    torch.cuda.synchronize()
    assert torch.equal(fs.rays_o, fs2.rays_o) and torch.equal(fs.rays_d, fs2.rays_d) and torch.equal(fs.target, fs2.target)
    assert torch.equal(fs.rays_a, fs2.rays_a) and int(fs.counter[0]) == int(fs2.counter[0]) > R
    d1, _ = fs.stats_host(); d2, _ = fs2.stats_host()
    for k in d2:
        assert abs(d1[k] - d2[k]) <= 1e-4 * abs(d2[k]) + 1e-9, (k, d1[k], d2[k])
    fs.use_pixel_batches(False)
    assert fs.target.data_ptr() == fs.inp[2].data_ptr() and fs.graph is None


def test_rays_from_pixels_matches_reference_golden():
    """ncn_rays_from_pixels against the reference's own get_ray_directions + get_rays (datasets/ray_utils.py:8-71) used as
    NeRFSystem.forward does (tests/golden/get_rays_a.npz, generator oracle/gen_golden_rays.py); the synthetic camera table of
    ncn_b200.synth follows the same pixel-centre convention."""
    import os
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib, synth
    from ncn_b200._lib import check, ptr, stream
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "get_rays_a.npz"))
    poses = torch.from_numpy(g["poses"]).cuda().contiguous(); dirs = torch.from_numpy(g["directions"]).cuda().contiguous()
    img = torch.from_numpy(g["img_idx"]).cuda(); pix = torch.from_numpy(g["pix_idx"]).cuda()
    n = img.shape[0]
    ro = torch.empty(n, 3, device="cuda"); rd = torch.empty(n, 3, device="cuda")
    check(_lib.lib().ncn_rays_from_pixels(ptr(poses), ptr(dirs), ptr(img), ptr(pix), n, ptr(ro), ptr(rd), stream()))
    assert torch.equal(ro.cpu(), torch.from_numpy(g["rays_o"]))
    torch.testing.assert_close(rd.cpu(), torch.from_numpy(g["rays_d"]), rtol=1e-6, atol=2e-7)
    # synth.pixel_directions: same (u - cx + 0.5) / fx convention, normalised
    k = synth.CAMERAS["hypersim"]
    d = synth.pixel_directions("hypersim").reshape(k["H"], k["W"], 3)
    v, u = 100, 200
    want = np.array([(u - k["cx"] + 0.5) / k["fx"], (v - k["cy"] + 0.5) / k["fy"], 1.0])
    np.testing.assert_allclose(d[v, u], want / np.linalg.norm(want), rtol=1e-6)


def test_grid_update_sampling_kernels():
    """ncn_grid_sample_cells / ncn_grid_scatter_density (models/ngp_mt.py:254-271, 345-357): index ranges, the second half
    hits occupied cells only and uniformly, every point lies inside its cell, the seed advances, the scatter writes exp(h0)."""
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib, vren
    from ncn_b200._lib import check, ptr, stream
    L = _lib.lib()
    dev = "cuda"
    G, s = 128, 0.5
    G3, M = G ** 3, G ** 3 // 4
    g = torch.Generator(device=dev).manual_seed(0)
    grid = torch.rand(G3, device=dev, generator=g)
    thr = 0.9
    occ = grid > thr
    csum = torch.cumsum(occ, 0, dtype=torch.int32)
    idx = torch.empty(2 * M, dtype=torch.int32, device=dev); xyz = torch.empty(2 * M, 3, device=dev)
    seed = torch.full((1,), 7, dtype=torch.int64, device=dev)
    check(L.ncn_grid_sample_cells(ptr(csum), G, M, s, ptr(seed), ptr(idx), ptr(xyz), stream()))
    assert int(idx.min()) >= 0 and int(idx.max()) < G3
    assert bool(occ[idx[M:].long()].all())                                  # occupied half: occupied cells only
    hist = torch.bincount(idx[M:].long(), minlength=G3)[occ].float()        # ... drawn uniformly among them
    assert abs(float(hist.mean()) - M / int(occ.sum())) < 1e-3 and float(hist.std()) < 3.0 * float(hist.mean().sqrt())
    coords = vren.morton3D_invert(idx).float()
    assert abs(float(coords[:M].mean()) - (G - 1) / 2) < 0.5                # uniform half
    half = s / G
    centre = (coords / (G - 1) * 2 - 1) * (s - half)
    assert float((xyz - centre).abs().max()) <= half * (1 + 1e-5)           # jitter stays inside the cell
    assert float((xyz - centre).abs().mean()) > 0.4 * half                  # ... and is not degenerate
    # a second draw with the advanced seed differs; the scatter writes exp(h0) and advances the seed
    h = torch.randn(2 * M, 16, device=dev, generator=g).half()
    tmp = torch.zeros(G3, device=dev)
    check(L.ncn_grid_scatter_density(ptr(h), 16, ptr(idx), 2 * M, ptr(tmp), ptr(seed), stream()))
    assert int(seed) == 8
    k = int(idx[12345])
    cand = torch.exp(h[(idx == k).nonzero().flatten(), 0].float())
    assert bool((cand == tmp[k]).any())
    assert float(tmp[(torch.bincount(idx.long(), minlength=G3) == 0)].abs().max()) == 0.0
    idx2 = torch.empty_like(idx); xyz2 = torch.empty_like(xyz)
    check(L.ncn_grid_sample_cells(ptr(csum), G, M, s, ptr(seed), ptr(idx2), ptr(xyz2), stream()))
    assert not torch.equal(idx, idx2)


def test_fused_grid_update_matches_module_path_statistically():
    """FusedStep.update_grid (sampling kernel + C-ABI field + scatter) against NGPMT.update_density_grid (the reference's
    torch sequence): different random streams, so compare the resulting occupancy statistics."""
    tr, rays_o, rays_d, tri, rgb, target = _setup(R=2048)
    fs = tr.fused_step(use_graph=False)
    hp = tr.hp
    thr = 0.01 * hp["rend_max_samples"] / 3 ** 0.5 * hp["density_tresh_decay"]
    g0, b0 = tr.model.density_grid.clone(), tr.model.density_bitfield.clone()
    torch.manual_seed(0)
    tr.model.update_density_grid(thr, warmup=False)
    ga, ba = tr.model.density_grid.clone(), tr.model.density_bitfield.clone()
    tr.model.density_grid.copy_(g0); tr.model.density_bitfield.copy_(b0)
    fs.update_grid()
    gb, bb = tr.model.density_grid.clone(), tr.model.density_bitfield.clone()
    pop = lambda b: float(sum(((b >> i) & 1).sum() for i in range(8)))
    assert abs(pop(ba) - pop(bb)) <= 0.02 * max(pop(ba), 1.0)
    assert abs(float(ga.clamp(min=0).mean()) - float(gb.clamp(min=0).mean())) <= 0.05 * float(ga.clamp(min=0).mean())
    assert float((gb != g0).float().mean()) > 0.2                            # the update touched a large part of the grid


@pytest.mark.parametrize("strategy", ["all_images_triang_patch", "same_image_triang_patch", "all_images_triang", "same_image_triang"])
def test_device_batch_sampling(strategy):
    """ncn_sample_ray_batch + ncn_gather_pixels against the index arithmetic of BaseDataset.__getitem__ (datasets/base.py:94-183):
    ranges, patch / triangle structure (incl. the corner-INDEX quirk), one image for the same_image strategies, uniform draws,
    the seed advancing, gathered targets; ragged batch tail; then a fused step fed by nothing but the device sampler."""
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib, synth
    from ncn_b200._lib import check, ptr, stream
    from ncn_b200.fused import FusedStep
    L = _lib.lib()
    H, W, P, p = 48, 64, 7, 8
    R = 64 * 300 + 5                                   # ragged: 5 rays beyond the last whole patch / 2 beyond the last triangle
    sid = FusedStep.STRATEGIES[strategy]
    seed = torch.full((1,), 11, dtype=torch.int64, device="cuda")
    img = torch.empty(R, dtype=torch.int64, device="cuda"); pix = torch.empty(R, dtype=torch.int64, device="cuda")
    check(L.ncn_sample_ray_batch(sid, ptr(seed), R, P, H, W, p, ptr(img), ptr(pix), stream()))
    assert int(seed) == 12
    group = p * p if sid <= 1 else 3
    n_used = R // group * group
    assert int(img[n_used:].abs().sum()) == 0 and int(pix[n_used:].abs().sum()) == 0
    im, px = img[:n_used].view(-1, group), pix[:n_used].view(-1, group)
    assert int(im.min()) >= 0 and int(im.max()) < P and bool((im == im[:, :1]).all())
    assert int(px.min()) >= 0 and int(px.max()) < H * W
    if sid & 1:
        assert bool((im == im[0, 0]).all())
    else:
        cnt = torch.bincount(im[:, 0], minlength=P).float()
        assert float(cnt.min()) > 0.5 * float(cnt.mean())
    if sid <= 1:
        dy, dx = torch.meshgrid(torch.arange(p, device="cuda"), torch.arange(p, device="cuda"), indexing="ij")
        assert torch.equal(px - px[:, :1], (dy * W + dx).reshape(1, -1).expand_as(px))
        assert int(px[:, 0].max()) < (H - p + 1) * (W - p + 1)            # corner INDEX, not pixel id (base.py:164-166)
        assert int(px[:, 0].max()) > 0.9 * (H - p + 1) * (W - p + 1)
    else:
        x1, x2, x3 = px[:, 0], px[:, 1], px[:, 2]
        assert torch.equal(x2, x1 - W) and torch.equal(x3, x1 - 1)
        y, x = x1 // W, x1 % W
        assert int(y.min()) >= 1 and int(y.max()) <= H - 2 and int(x.min()) >= 1 and int(x.max()) <= W - 2
        assert int(y.max()) == H - 2 and int(x.max()) == W - 2 and int(y.min()) == 1 and int(x.min()) == 1
    img2 = torch.empty_like(img); pix2 = torch.empty_like(pix)
    check(L.ncn_sample_ray_batch(sid, ptr(seed), R, P, H, W, p, ptr(img2), ptr(pix2), stream()))
    assert not torch.equal(pix, pix2)
    g = torch.Generator(device="cuda").manual_seed(0)
    images = torch.rand(P, H * W, 3, device="cuda", generator=g)
    out = torch.empty(R, 3, device="cuda")
    check(L.ncn_gather_pixels(ptr(images), ptr(img), ptr(pix), R, H * W, 3, ptr(out), stream()))
    assert torch.equal(out, images[img, pix])
    labels = torch.randint(0, 4, (P, H * W), device="cuda", generator=g)
    lab = torch.empty(R, dtype=torch.int64, device="cuda")
    check(L.ncn_gather_pixels(ptr(labels), ptr(img), ptr(pix), R, H * W, 2, ptr(lab), stream()))
    assert torch.equal(lab, labels[img, pix])
    # a fused step whose only input is the device sampler
    tr, *_ = _setup(R=1536, seed=4)
    poses = torch.from_numpy(synth.camera_poses(50, 0)).cuda(); dirs = torch.from_numpy(synth.pixel_directions("hypersim")).cuda()
    tr.set_cameras(poses, dirs)
    fs = tr.fused_step(use_graph=True)
    imgs = torch.rand(50, 768 * 1024, 3, device="cuda", generator=g)
    fs.use_device_sampling(imgs, 768, 1024, strategy=strategy, patch_size=8, seed=3)
    assert fs.M == (1536 // 64 * 49 if sid <= 1 else 512)
    fs.step(); torch.cuda.synchronize()
    first = fs.b_pix.clone()
    assert torch.equal(fs.target, imgs[fs.b_img, fs.b_pix])
    fs.step(); torch.cuda.synchronize()
    assert not torch.equal(first, fs.b_pix) and int(fs.dev_sampling["seed"]) >= 5
    d, n = fs.stats_host()
    assert np.isfinite(d["total"]) and n > 1536


@pytest.mark.parametrize("strategy,expand", [("all_images_triang", 3), ("same_image_triang", 40), ("all_images_triang", 0)])
def test_device_batch_sampling_triangle_expansion(strategy, expand):
    """`triang_max_expand` of the triangle strategies (datasets/base.py:130-141): with the same seed the expanded batch must be the
    reference's arithmetic applied to the unit triangles - x1 + e W if it stays below H W, x2 - e W if it stays >= 0, x3 - e if
    it stays in its row (numpy floor division)."""
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    from ncn_b200.fused import FusedStep
    L = _lib.lib()
    H, W, P = 48, 64, 5
    R = 3 * 4000
    sid = FusedStep.STRATEGIES[strategy]
    out = {}
    for e in (0, expand):
        seed = torch.full((1,), 77, dtype=torch.int64, device="cuda")
        img = torch.empty(R, dtype=torch.int64, device="cuda"); pix = torch.empty(R, dtype=torch.int64, device="cuda")
        check(L.ncn_sample_ray_batch_ex(sid, ptr(seed), R, P, H, W, 8, e, ptr(img), ptr(pix), stream()))
        out[e] = (img.cpu().numpy(), pix.cpu().numpy().reshape(-1, 3))
    assert np.array_equal(out[0][0], out[expand][0])                       # the images do not depend on the expansion
    x1, x2, x3 = out[0][1][:, 0], out[0][1][:, 1], out[0][1][:, 2]
    assert np.array_equal(x2, x1 - W) and np.array_equal(x3, x1 - 1)
    N = H * W
    x1n = np.where(x1 + expand * W < N, x1 + expand * W, x1)              # base.py:132-133
    x2n = np.where(x2 - expand * W >= 0, x2 - expand * W, x2)             # :135-136
    x3n = np.where((x3 - expand) // W == x3 // W, x3 - expand, x3)        # :138-140
    got = out[expand][1]
    assert np.array_equal(got[:, 0], x1n) and np.array_equal(got[:, 1], x2n) and np.array_equal(got[:, 2], x3n)
    if expand:
        assert (x1n != x1).any() and (x1n == x1).any() and (x3n != x3).any() and (x3n == x3).any()      # both branches occur


def test_programmatic_dependent_launch_changes_no_result():
    """ncn_set_pdl(0/1): the forward of one eager fused step (everything up to the losses is deterministic: no atomics) is
    bit-identical with and without programmatic dependent launch, and so is a CUDA-graph replay of it.  dL/dsigma sits behind
    dL/ddepth, which the normals' backward accumulates with fp32 atomics (up to three triangles meet in a ray, in any order), so it
    is reproducible to rounding only - a missed dependency would show as whole missing contributions, far above that bound."""
    from ncn_b200 import _lib
    L = _lib.lib()
    outs = {}
    old = L.ncn_set_pdl(1)
    try:
        for mode in ((1, False), (0, False), (1, True), (0, True)):
            L.ncn_set_pdl(mode[0])
            tr, rays_o, rays_d, tri, rgb, target = _setup(R=1024, seed=3)
            fs = tr.fused_step(use_graph=mode[1]); fs.set_triangles(tri)
            noise = torch.rand(1024, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
            fs.step(rays_o, rays_d, rgb, noise=noise)
            fs.flush()
            torch.cuda.synchronize()
            n_live = int(fs.counter[0])
            outs[mode] = (fs.rend.clone(), fs.depth.clone(), fs.opacity.clone(), fs.losses.clone(), fs.d_sigmas[:n_live].clone(), n_live)
    finally:
        L.ncn_set_pdl(old)
    ref = outs[(0, False)]
    assert ref[5] > 1024 and torch.isfinite(ref[3]).all()
    for mode, got in outs.items():
        for a, b in zip(ref[:4], got[:4]):
            assert torch.equal(a, b), mode
        assert got[5] == ref[5]
        torch.testing.assert_close(got[4], ref[4], rtol=1e-4, atol=1e-5 * float(ref[4].abs().max()), msg=lambda m, mode=mode: f"{mode}: {m}")
