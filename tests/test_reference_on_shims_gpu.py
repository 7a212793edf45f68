"""GPU: the reference's UNCHANGED Python hot path executed on a B200 through the drop-in shims (north_star
"models/rendering.py run[s] unchanged"), and used as the oracle for the host mirrors and the fused step:

    reference models/rendering.py:render  (its own @autocast, RayAABBIntersector + near clamp, RayMarcher, VolumeRenderer)
    reference models/ngp_mt.py:NGPMT      (built on shims/tinycudann -> libncn hash grid + MLPs; update_density_grid, mark_invisible_cells)
    reference losses.py:NeRFMTLoss        (faiss.Kmeans -> shims/faiss -> libncn k-means; its own selection / merge / loss code)

against ncn_b200.rendering.render / ncn_b200.ngp.NGPMT / ncn_b200.losses.NeRFMTLoss (same parameters, rays, torch RNG state) and
against FusedStep (same parameters, rays, march jitter).  The files are the byte-for-byte copies oracle/build_ref.py stages into
git-ignored oracle/_ref/py/ (tests skip when that tree is absent).
"""
import contextlib
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HP = dict(loss_opacity_w=1e-3, loss_norm_can_tres=0.01, loss_norm_D_C_ort_dot_w=2e-3, loss_norm_D_C_centr_dot_w=2e-3,
          loss_norm_D_C_centr_L1_w=2e-3, loss_norm_can_start=500, loss_norm_can_grow=2500, loss_norm_can_end=-1,
          ray_sampling_strategy="all_images_triang_patch", random_tr_poses=False, pred_norm_nn=False, pred_norm_depth=True)


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_py
    ns = ref_py.load(faiss="shim", want_models=True)
    if ns is None:
        pytest.skip("oracle/_ref/py not staged (run oracle/build_ref.py where /root/reference exists)")
    return ns


def _trainer(R, hp=None, n_sem_cls=0, table_std=0.3, seed=0):
    from ncn_b200 import synth, vren
    from ncn_b200.trainer import NeRFTrainer
    torch.manual_seed(seed)
    tr = NeRFTrainer(dict(batch_size=R, **(hp or {})), device="cuda", n_sem_cls=n_sem_cls)
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    tr.model.density_grid.copy_(torch.from_numpy(grid).cuda())
    vren.packbits(tr.model.density_grid, 5.9, tr.model.density_bitfield)
    g = torch.Generator(device="cuda").manual_seed(1)
    n = tr.model.xyz_encoder.params.numel()
    tr.opt.flat[:n].copy_(torch.randn(n, device="cuda", generator=g) * table_std)
    for name in ("sem_net", "norm_net"):
        if hasattr(tr.model, name):
            p = getattr(tr.model, name).params
            p.data.copy_(torch.randn(p.numel(), device="cuda", generator=g) * 0.2)
    tr.opt.flat16.copy_(tr.opt.flat)
    tr.global_step = 3000
    return tr


def _ref_model(ref, tr, n_sem_cls=0):
    """the reference's NGPMT on the shims with OUR parameters / occupancy state (buffers as train_nerf.py:153-157 registers them)"""
    m = tr.model
    kw = dict(n_sem_cls=n_sem_cls) if m.pred_sem else {}
    with contextlib.redirect_stdout(io.StringIO()):
        rm = ref.ngp_mt.NGPMT(scale=0.5, grid_size=128, rgb_act="Sigmoid", pred_sem=m.pred_sem, pred_norm=m.pred_norm, **kw).cuda()
    rm.register_buffer("density_grid", m.density_grid.clone())
    rm.register_buffer("grid_coords", m.grid_coords.clone())
    rm.density_bitfield.copy_(m.density_bitfield)
    ours = dict(m.named_parameters())
    with torch.no_grad():
        for k, p in rm.named_parameters():
            assert p.shape == ours[k].shape, k
            p.copy_(ours[k])
    assert type(rm.xyz_encoder).__module__.endswith("tinycudann") and rm.cascades == m.cascades
    return rm


def _batch(R, seed=0):
    from ncn_b200 import synth
    b = synth.patch_batch(R, seed=seed)
    rays_o = torch.from_numpy(b["rays_o"]).cuda(); rays_d = torch.from_numpy(b["rays_d"]).cuda()
    tri = torch.from_numpy(b["tri"]).cuda()
    g = torch.Generator(device="cuda").manual_seed(7)
    rgb = torch.rand(R, 3, device="cuda", generator=g)
    target = {"rgb": rgb, "patch_area": 64, "x1_offsets_local": tri[0][:49] % 64, "x2_offsets_local": tri[1][:49] % 64,
              "x3_offsets_local": tri[2][:49] % 64}
    return rays_o, rays_d, tri, rgb, target


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


LOSS_SCALE = 1024.0     # what the AMP GradScaler of the reference's Trainer(precision=16) does; without it fp16 dL/dy underflows


def _ref_forward_backward(ref, rm, hp, rays_o, rays_d, target, kwargs, step, seed=123):
    loss_fn = ref.losses.NeRFMTLoss(hp)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        results = ref.rendering.render(rm, rays_o.clone(), rays_d.clone(), global_step=step, **kwargs)
        loss_d = loss_fn(results, target, global_step=step)
    for p in rm.parameters():
        p.grad = None
    (loss_d["total"] * LOSS_SCALE).backward()
    grads = {k: (p.grad.detach().clone() / LOSS_SCALE if p.grad is not None else torch.zeros_like(p)) for k, p in rm.named_parameters()}
    return results, loss_d, grads


@pytest.mark.parametrize("R,heads", [(2048, False), (1024, True)])
def test_reference_render_and_loss_on_shims_equal_the_mirrors(ref, R, heads):
    hp = dict(HP)
    n_cls = 0
    if heads:
        hp.update(pred_sem=True, pred_norm_nn=True, loss_sem_w=4e-2, load_sem_gt=True, load_sem_WF_gt=False)
        n_cls = 3
    tr = _trainer(R, hp={k: v for k, v in hp.items() if k in ("pred_sem", "pred_norm_nn", "loss_sem_w")}, n_sem_cls=n_cls)
    rm = _ref_model(ref, tr, n_cls)
    rays_o, rays_d, tri, rgb, target = _batch(R)
    if heads:
        target["semantics"] = torch.randint(0, n_cls + 1, (R,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    kwargs = dict(tr.render_kwargs)
    step = tr.global_step
    # ---- the reference's own code on the shims
    res_r, loss_r, g_ref = _ref_forward_backward(ref, rm, hp, rays_o, rays_d, target, kwargs, step)
    # ---- our mirrors, same RNG state (the march jitter is torch.rand_like in both RayMarcher classes)
    torch.manual_seed(123)
    res_o, loss_o = tr.forward_loss(rays_o, rays_d, target)
    (loss_o["total"] * tr.hp["loss_scale"]).backward()
    g_our = {k: (p.grad / tr.hp["loss_scale"]).clone() for k, p in tr.model.named_parameters() if p.numel()}
    tr.opt.grad.zero_()
    # marching: integer / index outputs exactly, sample positions bit for bit
    assert int(res_r["rm_samples"]) == int(res_o["rm_samples"]) > R
    assert torch.equal(res_r["rays_a"], res_o["rays_a"])
    for k in ("ts", "deltas"):
        assert torch.equal(res_r[k], res_o[k]), k
    assert int(res_r["vr_samples"]) == int(res_o["vr_samples"])
    # rendered quantities: the same kernels on the same inputs
    for k in ("opacity", "depth", "ws", "rgb") + (("norm_nn", "sem") if heads else ()):
        torch.testing.assert_close(res_r[k].detach().float(), res_o[k].detach().float(), rtol=1e-5, atol=1e-6, msg=lambda m, k=k: f"{k}: {m}")
    assert torch.equal(res_r["rays_o"], res_r["rays_d"]) and torch.equal(res_o["rays_o"], res_o["rays_d"])      # rendering.py:227
    # losses: the reference's selection / merge / opposite / loss code vs the libncn kernels
    # (the reference also emits norm_D_C_can_dot / norm_D_C_can_L1 whenever a cluster centre happens to lie near a canonical axis,
    #  losses.py:491-502 - a data-dependent key whose value is w_sched(0) * loss = 0 in every shipped configuration; the sync-free
    #  mirror emits those keys only when their weights are non-zero)
    extra = set(loss_r) - set(loss_o)
    assert extra <= {"norm_D_C_can_dot", "norm_D_C_can_L1"} and all(float(loss_r[k]) == 0.0 for k in extra), extra
    assert set(loss_o) <= set(loss_r), (sorted(loss_r), sorted(loss_o))
    for k in loss_o:
        a, b = float(loss_r[k]), float(loss_o[k])
        assert abs(a - b) <= 1e-3 * abs(a) + 1e-8, (k, a, b)
    for k in ("norm_D_C_ort_dot", "norm_D_C_centr_dot", "norm_D_C_centr_L1"):
        assert float(loss_r[k]) > 0
    # gradients on every parameter group (fp32 atomics in a different order; borderline k-means assignments may flip a label)
    for k, gr in g_ref.items():
        if gr.numel() == 0:
            continue
        if k == "norm_net.params":            # no loss touches norm_nn (losses.py:279/294 only slice it): exactly zero on both sides
            assert float(gr.abs().max()) == 0.0 and float(g_our[k].abs().max()) == 0.0
            continue
        assert float(gr.norm()) > 0, k
        assert _rel(g_our[k], gr) <= 2e-2, (k, _rel(g_our[k], gr))


@pytest.mark.parametrize("fuse_fwd", ["mlp", True])
def test_fused_step_matches_the_reference_on_shims(ref, fuse_fwd):
    """FusedStep (one CUDA-graph-able kernel sequence) vs the reference's render + NeRFMTLoss + autograd on the shims"""
    R = 2048
    tr = _trainer(R)
    rm = _ref_model(ref, tr)
    rays_o, rays_d, tri, rgb, target = _batch(R)
    res_r, loss_r, g_ref = _ref_forward_backward(ref, rm, dict(HP), rays_o, rays_d, target, dict(tr.render_kwargs), tr.global_step)
    torch.manual_seed(123)
    noise = torch.rand(R, device="cuda")            # what RayMarcher.forward drew (custom_functions.py:83)
    fs = tr.fused_step(use_graph=False, fuse_fwd=fuse_fwd)
    fs.set_triangles(tri)
    fs.rays_o.copy_(rays_o); fs.rays_d.copy_(rays_d); fs.target.copy_(rgb); fs.noise.copy_(noise)
    fs.gen_noise = False
    fs._schedule()
    fs._run()
    torch.cuda.synchronize()
    N = int(res_r["rm_samples"])
    assert int(fs.counter[0]) == N
    assert torch.equal(fs.rays_a, res_r["rays_a"])
    assert torch.equal(fs.ts[:N], res_r["ts"]) and torch.equal(fs.deltas[:N], res_r["deltas"])
    torch.testing.assert_close(fs.depth, res_r["depth"].detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(fs.opacity, res_r["opacity"].detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(fs.rgb, res_r["rgb"].detach().float(), rtol=1e-3, atol=1e-3)
    d, n = fs.stats_host()
    for k in ("rgb", "opacity", "norm_D_C_ort_dot", "norm_D_C_centr_dot", "norm_D_C_centr_L1"):
        assert abs(d[k] - float(loss_r[k])) <= 2e-3 * abs(float(loss_r[k])) + 1e-7, (k, d[k], float(loss_r[k]))
    for name in ("rgb_net", "sigma_net", "xyz_encoder"):
        o, k = fs.off[name]
        rel = _rel(tr.opt.grad[o:o + k], g_ref[name + ".params"])
        assert rel <= 3e-2, (name, rel)


# ------------------------------------------------------------------ a17: occupancy-grid upkeep
def _thr(tr):
    hp = tr.hp
    return 0.01 * hp["rend_max_samples"] / 3 ** 0.5 * hp["density_tresh_decay"]


@pytest.mark.parametrize("table_std", [0.3, 4.0])
def test_update_density_grid_warmup_reference_vs_mirror(ref, table_std):
    """warm-up branch (ngp_mt.py:341-342, all G^3 cells): same cells, same torch RNG -> same jittered positions -> the mirror's
    fused decay/max + packbits must reproduce the reference's torch.where / mean().item() / packbits.  table_std 0.3 leaves the
    mean density below the threshold (packbits threshold = the mean), 4.0 above it (threshold = density_threshold)."""
    tr = _trainer(1024, table_std=table_std)
    rm = _ref_model(ref, tr)
    g = torch.Generator(device="cuda").manual_seed(9)
    start = torch.rand(tr.model.density_grid.shape, device="cuda", generator=g) * 12.0
    start[:, ::5] = -1.0                                   # cells mark_invisible_cells disabled stay at -1 (ngp_mt.py:360-363)
    thr = _thr(tr)
    for rounds in range(2):
        tr.model.density_grid.copy_(start); rm.density_grid.copy_(start)
        torch.manual_seed(77 + rounds)
        with torch.autocast("cuda", dtype=torch.float16):      # Lightning runs training_step under AMP (train_nerf.py precision=16):
            rm.update_density_grid(thr, warmup=True)          # the reference's TruncExp only returns fp32 under autocast (custom_fwd)
        torch.manual_seed(77 + rounds)
        tr.model.update_density_grid(thr, warmup=True)
        torch.cuda.synchronize()
        assert torch.equal(rm.density_grid < 0, tr.model.density_grid < 0)
        assert (rm.density_grid[:, ::5] == -1).all()
        torch.testing.assert_close(tr.model.density_grid, rm.density_grid, rtol=0, atol=0)
        mean = float(rm.density_grid[rm.density_grid > 0].mean())
        diff = (tr.model.density_bitfield ^ rm.density_bitfield)
        n_diff = int(sum(((diff >> b) & 1).sum() for b in range(8)))
        if mean > thr:                                      # threshold is the constant: bit-exact
            assert n_diff == 0
        else:                                               # threshold is the mean (two reduction orders): only cells within 1e-5 of it may flip
            near = ((rm.density_grid - mean).abs() <= 1e-5 * mean).sum()
            assert n_diff <= int(near), (n_diff, int(near))
        assert (mean > thr) == (table_std > 1.0), (mean, thr)
        start = rm.density_grid.clone()                     # second round starts from an updated grid


def test_cell_sampling_reference_vs_mirror_vs_fused(ref):
    """steady-state branch (ngp_mt.py:248-270): M uniform cells + M cells drawn from {density > threshold}.  The three samplers
    (reference: randint + nonzero; mirror: randint + cumsum/searchsorted; fused: ncn_grid_sample_cells) use different random streams,
    so the comparison is by the exact properties (index<->coordinate consistency, occupied half inside the occupied set) and by
    the first two moments of the per-cell hit histogram."""
    import ctypes as C
    from ncn_b200 import _lib, vren
    from ncn_b200._lib import check, ptr, stream
    tr = _trainer(1024)
    rm = _ref_model(ref, tr)
    G, M = 128, 128 ** 3 // 4
    thr = _thr(tr)
    occ = tr.model.density_grid[0] > thr
    n_occ = int(occ.sum())
    assert 0 < n_occ < G ** 3
    torch.manual_seed(5)
    (idx_r, coords_r), = rm.sample_uniform_and_occupied_cells(M, thr)
    (idx_m, coords_m), = tr.model.sample_uniform_and_occupied_cells(M, thr)
    csum = torch.cumsum(occ, 0, dtype=torch.int32)
    idx_f = torch.empty(2 * M, dtype=torch.int32, device="cuda"); xyz_f = torch.empty(2 * M, 3, device="cuda")
    seed = torch.full((1,), 4242, dtype=torch.int64, device="cuda")
    check(_lib.lib().ncn_grid_sample_cells(ptr(csum), G, M, 0.5, ptr(seed), ptr(idx_f), ptr(xyz_f), stream()), "grid_sample_cells")
    hists = {}
    for name, idx, coords in (("reference", idx_r, coords_r), ("mirror", idx_m, coords_m), ("fused", idx_f.long(), None)):
        assert idx.shape[0] == 2 * M
        if coords is not None:
            assert torch.equal(vren.morton3D(coords.int().contiguous()).long(), idx.long()), name
        assert occ[idx[M:].long()].all(), f"{name}: occupied half left the occupied set"
        assert int(idx.min()) >= 0 and int(idx.max()) < G ** 3
        h_u = torch.bincount(idx[:M].long(), minlength=G ** 3).double()
        h_o = torch.bincount(idx[M:].long(), minlength=G ** 3).double()[occ]
        hists[name] = (float(h_u.mean()), float(h_u.var()), float(h_o.mean()), float(h_o.var()), float((h_o == 0).double().mean()))
    r = hists["reference"]
    for name in ("mirror", "fused"):
        h = hists[name]
        assert abs(h[0] - r[0]) < 1e-9 and abs(h[2] - r[2]) < 1e-9                     # same number of draws per half
        assert abs(h[1] - r[1]) <= 0.03 * r[1], (name, "uniform variance", h[1], r[1])  # Poisson(M/G^3): variance ~ mean
        assert abs(h[3] - r[3]) <= 0.05 * r[3], (name, "occupied variance", h[3], r[3])
        assert abs(h[4] - r[4]) <= 0.02 + 0.1 * r[4], (name, "unvisited occupied cells", h[4], r[4])
    # the fused sampler's positions: cell centre + jitter inside the cell (ngp_mt.py:352-355)
    half = 0.5 / G
    coords_f = vren.morton3D_invert(idx_f.contiguous()).float()
    centre = (coords_f / (G - 1) * 2 - 1) * (0.5 - half)
    assert float((xyz_f - centre).abs().max()) <= half * (1 + 1e-5)
    assert float((xyz_f - centre).abs().mean()) > 0.3 * half


def test_update_density_grid_steady_reference_vs_mirror_vs_fused(ref):
    """one steady-state update from the same grid: every cell either decayed exactly (x0.95) or was raised to a sampled density;
    the three implementations agree on the exact part and statistically on the sampled part"""
    tr = _trainer(8192, table_std=1.0)
    rm = _ref_model(ref, tr)
    thr = _thr(tr)
    g = torch.Generator(device="cuda").manual_seed(9)
    start = torch.rand(tr.model.density_grid.shape, device="cuda", generator=g) * 12.0
    start[:, ::7] = -1.0
    out = {}
    rm.density_grid.copy_(start)
    torch.manual_seed(3)
    with torch.autocast("cuda", dtype=torch.float16):
        rm.update_density_grid(thr, warmup=False)
    out["reference"] = (rm.density_grid.clone(), rm.density_bitfield.clone())
    tr.model.density_grid.copy_(start)
    torch.manual_seed(3)
    tr.model.update_density_grid(thr, warmup=False)
    out["mirror"] = (tr.model.density_grid.clone(), tr.model.density_bitfield.clone())
    tr.model.density_grid.copy_(start)
    fs = tr.fused_step(use_graph=False)
    fs.update_grid()
    torch.cuda.synchronize()
    out["fused"] = (tr.model.density_grid.clone(), tr.model.density_bitfield.clone())
    decayed = start * 0.95
    stats = {}
    for name, (grid, bits) in out.items():
        assert torch.equal(grid < 0, start < 0), name
        assert torch.equal(grid[start < 0], start[start < 0]), name
        pos = start >= 0
        assert (grid[pos] >= decayed[pos]).all(), name                                   # max(grid * decay, sampled)
        raised = (grid > decayed) & pos
        stats[name] = (float(raised.double().mean()), float(grid[pos].double().mean()), float(grid[raised].double().mean()),
                       int(sum(((bits >> b) & 1).sum() for b in range(8))))
    r = stats["reference"]
    for name in ("mirror", "fused"):
        s = stats[name]
        assert abs(s[0] - r[0]) <= 0.02 * r[0] + 1e-4, (name, "fraction of raised cells", s[0], r[0])
        assert abs(s[1] - r[1]) <= 0.01 * r[1], (name, "mean density", s[1], r[1])
        assert abs(s[2] - r[2]) <= 0.03 * r[2], (name, "mean raised density", s[2], r[2])
        assert abs(s[3] - r[3]) <= 0.01 * r[3], (name, "occupied bits", s[3], r[3])


@pytest.mark.parametrize("variant", ["pinhole", "hypersim_tuple"])
def test_mark_invisible_cells_reference_vs_mirror(ref, variant):
    """ngp_mt.py:273-337 (run once before training, train_nerf.py:306-312): cells outside every camera frustum or closer than
    `near` to a camera get density -1.  Same argument list on both sides, both intrinsics forms."""
    from ncn_b200 import synth
    tr = _trainer(1024)
    rm = _ref_model(ref, tr)
    poses = torch.from_numpy(synth.camera_poses(12, 0)).cuda()
    k = synth.CAMERAS["hypersim"]
    img_wh = (k["W"], k["H"])
    if variant == "pinhole":
        K = torch.tensor([[k["fx"], 0, k["cx"]], [0, k["fy"], k["cy"]], [0, 0, 1]], dtype=torch.float32)
        # OpenCV-style poses (z forward): depth = z
    else:
        # Hypersim form (datasets/hypersim.py:102-105, hypersim_src/cam_model.py:67-80): OpenGL projection (camera looks down -z,
        # y up) + NDC->uv matrix whose third row yields the depth test value; scene scale 1.7
        n_, f_ = 0.1, 100.0
        fx = 2 * k["fx"] / k["W"]; fy = 2 * k["fy"] / k["H"]
        M_proj = torch.tensor([[fx, 0, 0.05, 0], [0, fy, -0.02, 0], [0, 0, -(f_ + n_) / (f_ - n_), -2 * f_ * n_ / (f_ - n_)], [0, 0, -1, 0]],
                              dtype=torch.float32)
        W, H = img_wh
        M_uv = torch.tensor([[0.5 * (W - 1), 0, 0, 0.5 * (W - 1)], [0, -0.5 * (H - 1), 0, 0.5 * (H - 1)], [0, 0, 0.5, 0.5], [0, 0, 0, 1.0]],
                            dtype=torch.float32)
        K = (M_proj, M_uv, [0.0, 0.0, 0.0], 1.7)
        flip = torch.diag(torch.tensor([1.0, -1.0, -1.0], device="cuda"))               # OpenCV -> OpenGL camera axes
        poses = torch.cat([poses[:, :, :3] @ flip, poses[:, :, 3:]], -1).contiguous()
    near = 0.05 if variant == "pinhole" else 0.3
    for chunk in (64 ** 3, 100003):
        rm.density_grid.zero_(); tr.model.density_grid.zero_()
        rm.mark_invisible_cells(K, torch.device("cuda"), poses, img_wh, near, chunk=chunk)
        tr.model.mark_invisible_cells(K, torch.device("cuda"), poses, img_wh, near, chunk=chunk)
        assert torch.equal(rm.density_grid, tr.model.density_grid)
        assert torch.equal(rm.count_grid, tr.model.count_grid)
        frac = float((rm.density_grid < 0).float().mean())
        assert 0.02 < frac < 0.98, frac                                                   # both outcomes occur


# ------------------------------------------------------------------ f4: --random_tr_poses
def _batch_rtp(R):
    """[R/2 rays of training views with a target colour | R/2 rays of other (generated) poses], patch topology in both halves"""
    ro_a, rd_a, tri, rgb, target = _batch(R // 2, seed=0)
    ro_b, rd_b, _, _, _ = _batch(R // 2, seed=5)
    return torch.cat([ro_a, ro_b]).contiguous(), torch.cat([rd_a, rd_b]).contiguous(), tri, rgb, target


def test_random_tr_poses_reference_vs_mirror_vs_fused(ref):
    """--random_tr_poses (losses.py:265-297, train_nerf.py:169-172, 338-344): photometric term on the first half of the batch only,
    opacity on every ray, the normal-clustering terms on the rays of the generated poses only.  The reference's own render +
    NeRFMTLoss on the shims is the oracle for the module-path mirror and for FusedStep (n_gt / u0 split, ncn_*_gt kernels).
    R = 4096: the generated-pose half yields the 1568 normals of the R = 2048 tests above (same k-means regime)."""
    R = 4096
    hp = dict(HP, random_tr_poses=True)
    tr = _trainer(R, hp=dict(random_tr_poses=True))
    rm = _ref_model(ref, tr)
    rays_o, rays_d, tri, rgb, target = _batch_rtp(R)
    assert rgb.shape[0] == R // 2 and int(tri.max()) < R // 2
    res_r, loss_r, g_ref = _ref_forward_backward(ref, rm, hp, rays_o, rays_d, target, dict(tr.render_kwargs), tr.global_step)
    # module-path mirror
    torch.manual_seed(123)
    res_o, loss_o = tr.forward_loss(rays_o, rays_d, target)
    (loss_o["total"] * tr.hp["loss_scale"]).backward()
    g_our = {k: (p.grad / tr.hp["loss_scale"]).clone() for k, p in tr.model.named_parameters() if p.numel()}
    tr.opt.grad.zero_()
    for k in loss_o:
        a, b = float(loss_r[k]), float(loss_o[k])
        assert abs(a - b) <= 1e-3 * abs(a) + 1e-8, (k, a, b)
    for k, gr in g_ref.items():
        if gr.numel():
            assert float(gr.norm()) > 0 and _rel(g_our[k], gr) <= 2e-2, (k, _rel(g_our[k], gr))
    # the split matters: the loss over ALL rays against an R-row target would differ
    assert res_r["rgb"].shape[0] == R
    # fused step
    torch.manual_seed(123)
    noise = torch.rand(R, device="cuda")
    fs = tr.fused_step(use_graph=False)
    assert fs.rtp and fs.n_gt == R // 2 and fs.u0 == R // 2
    fs.set_triangles(tri)
    fs.gen_noise = False
    fs.rays_o.copy_(rays_o); fs.rays_d.copy_(rays_d); fs.target[:R // 2].copy_(rgb); fs.target[R // 2:].fill_(123.0)      # never read
    fs.noise.copy_(noise)
    fs._schedule()
    fs._run()
    torch.cuda.synchronize()
    assert int(fs.counter[0]) == int(res_r["rm_samples"]) and torch.equal(fs.rays_a, res_r["rays_a"])
    torch.testing.assert_close(fs.depth, res_r["depth"].detach(), rtol=1e-4, atol=1e-5)
    d, n = fs.stats_host()
    for k in ("rgb", "opacity", "norm_D_C_ort_dot", "norm_D_C_centr_dot", "norm_D_C_centr_L1"):
        assert abs(d[k] - float(loss_r[k])) <= 2e-3 * abs(float(loss_r[k])) + 1e-7, (k, d[k], float(loss_r[k]))
    assert float(fs.d_rend[R // 2:].abs().max()) == 0.0 and float(fs.d_rend[:R // 2].abs().max()) > 0      # no colour gradient on the generated half
    assert float(fs.d_depth[:R // 2].abs().max()) == 0.0 and float(fs.d_depth[R // 2:].abs().max()) > 0    # no cluster gradient on the training half
    for name in ("rgb_net", "sigma_net", "xyz_encoder"):
        o, k = fs.off[name]
        rel = _rel(tr.opt.grad[o:o + k], g_ref[name + ".params"])
        assert rel <= 3e-2, (name, rel)
