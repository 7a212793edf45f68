"""GPU: the test-time rendering loop (render(test_time=True): ncn_march_test + ncn_composite_test_fw in rounds over the
still-alive rays, reference models/rendering.py:73-168) against the training kernels (ncn_march_train with zero
jitter + ncn_composite_train_fw, reference :171-239) on the same field: both integrate the same samples front to back
and stop at the same transmittance threshold, so colour / depth / opacity must agree."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_rays", [1024, 777])
def test_test_time_loop_matches_train_kernels(n_rays):
    import ncn_b200  # noqa: F401
    from ncn_b200 import synth, vren
    from ncn_b200.rendering import render, ray_aabb_near, _raws
    from ncn_b200.trainer import NeRFTrainer
    torch.manual_seed(0)
    tr = NeRFTrainer(dict(batch_size=n_rays), device="cuda")
    model = tr.model
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    model.density_grid.copy_(torch.from_numpy(grid).cuda())
    vren.packbits(model.density_grid, 5.9, model.density_bitfield)
    g = torch.Generator(device="cuda").manual_seed(1)
    n = model.xyz_encoder.params.numel()
    tr.opt.flat[:n].copy_(torch.randn(n, device="cuda", generator=g) * 0.5)      # dense enough that most rays saturate
    tr.opt.flat16.copy_(tr.opt.flat)
    b = synth.patch_batch(1024, seed=3)
    rays_o = torch.from_numpy(b["rays_o"]).cuda()[:n_rays].contiguous()
    rays_d = torch.from_numpy(b["rays_d"]).cuda()[:n_rays].contiguous()
    kw = dict(near_distance=0.0, max_samples=1024, exp_step_factor=0.0, T_threshold=1e-4, n_sem_cls=0)
    with torch.no_grad():
        res = render(model, rays_o, rays_d, test_time=True, **kw)
        hits_t = ray_aabb_near(rays_o, rays_d, model.center, model.half_size, 0.0)
        noise = torch.zeros(n_rays, device="cuda")
        rays_a, xyzs, dirs, deltas, ts, _ = vren.raymarching_train(rays_o, rays_d, hits_t[:, 0], model.density_bitfield, model.cascades,
                                                                   model.scale, 0.0, noise, model.grid_size, 1024)
        out = model(xyzs, dirs, **kw)
        raws = _raws(model, out)
        n_used, opacity, depth, rend, ws = vren.composite_train_multi_fw(out["sigmas"].float().contiguous(), raws, deltas, ts, rays_a, 1e-4)
        rgb = rend[:, :3] + (1 - opacity)[:, None]
    assert int(res["total_samples"]) > 0
    hit = opacity > 0
    assert hit.float().mean() > 0.5                       # the synthetic room is in view
    torch.testing.assert_close(res["opacity"], opacity, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(res["depth"], depth, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(res["rgb"], rgb, rtol=1e-4, atol=5e-4)
    # the loop marches whole rounds of N_samples, so it may visit more samples than the ones composited in training, never fewer
    assert int(res["total_samples"]) >= int(n_used.sum())


@pytest.mark.parametrize("heads,n_cls", [((False, False), 0), ((True, True), 5)])
def test_render_fast_matches_test_time_loop(heads, n_cls):
    """render_fast (one march + one field pass + one composite per tile) against the reference-shaped
    round loop render(test_time=True): same rgb / depth / opacity (+ norm_nn, sem), with a tile size that does not divide
    the ray count (ragged last tile) and rays that miss the box."""
    import ncn_b200  # noqa: F401
    from ncn_b200 import synth, vren
    from ncn_b200.rendering import render, render_fast
    from ncn_b200.trainer import NeRFTrainer
    torch.manual_seed(0)
    pred_sem, pred_norm = heads
    tr = NeRFTrainer(dict(batch_size=1024, pred_sem=pred_sem, pred_norm_nn=pred_norm), device="cuda", n_sem_cls=n_cls)
    model = tr.model
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    model.density_grid.copy_(torch.from_numpy(grid).cuda())
    vren.packbits(model.density_grid, 5.9, model.density_bitfield)
    g = torch.Generator(device="cuda").manual_seed(1)
    n = model.xyz_encoder.params.numel()
    tr.opt.flat[:n].copy_(torch.randn(n, device="cuda", generator=g) * 0.5)
    for name in ("sem_net", "norm_net"):
        if hasattr(model, name):
            p = getattr(model, name).params
            p.data.copy_(torch.randn(p.numel(), device="cuda", generator=g) * 0.2)
    tr.opt.flat16.copy_(tr.opt.flat)
    b = synth.patch_batch(1024, seed=3)
    rays_o = torch.from_numpy(b["rays_o"]).cuda()[:1000].contiguous()
    rays_d = torch.from_numpy(b["rays_d"]).cuda()[:1000].contiguous()
    rays_o[::50] += 10.0                                  # some rays start far outside and miss the box
    kw = dict(near_distance=0.01, max_samples=1024, exp_step_factor=0.0, T_threshold=1e-4, n_sem_cls=n_cls)
    ref = render(model, rays_o, rays_d, test_time=True, **kw)
    out = render_fast(model, rays_o, rays_d, tile=384, **kw)
    assert (ref["opacity"] > 0).float().mean() > 0.5
    for k in ("opacity", "depth", "rgb") + (("norm_nn",) if pred_norm else ()) + (("sem",) if pred_sem else ()):
        torch.testing.assert_close(out[k], ref[k].float(), rtol=1e-3, atol=1e-3, msg=lambda m, k=k: f"{k}: {m}")
    assert int(ref["total_samples"]) >= int(out["total_samples"]) > 0
    empty = render_fast(model, rays_o[:0], rays_d[:0], **kw)
    assert empty["rgb"].shape == (0, 3)
