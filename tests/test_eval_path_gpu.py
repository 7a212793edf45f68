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
