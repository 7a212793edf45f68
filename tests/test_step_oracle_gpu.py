"""GPU: one whole training step (AABB -> march -> hash grid -> MLPs -> compositing -> photometric + opacity + normal-clustering loss ->
backward) of the module path AND of FusedStep against oracle/step.py - the CPU restatement of the same step assembled from the
pinned oracles (C march, torch hash grid / MLP / compositing / cluster loss; fp32 throughout) - on 512 rays, identical
parameters (fp16-representable values), rays, target colours and march jitter.

The k-means ENGINE is unpinned on both sides (faiss absent), so the centroids the GPU engine found are handed to
the oracle, which then assigns, selects, merges and evaluates the loss with its own code (oracle.step.train_step(centroids=...)).

Tolerances: sample counts exact; photometric / opacity / total loss 0.5 %, cluster terms 3 % (fp16 activations on the GPU vs fp32
on the CPU; a borderline normal may change cluster);
per-group gradient norm 5 %, direction (cosine) >= 0.99.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

R = 512


def _make():
    import ncn_b200  # noqa: F401
    from ncn_b200 import synth, vren, clustering
    from ncn_b200.trainer import NeRFTrainer
    from oracle import march, step as ostep
    torch.manual_seed(0)
    tr = NeRFTrainer(dict(batch_size=R), device="cuda")
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    tr.model.density_grid.copy_(torch.from_numpy(grid).cuda())
    vren.packbits(tr.model.density_grid, 5.9, tr.model.density_bitfield)
    bits = march.packbits(grid, 5.9)
    assert np.array_equal(bits, tr.model.density_bitfield.cpu().numpy())
    # a field with structure: table N(0, 0.3), fp16-representable so that both sides hold the SAME numbers
    g = torch.Generator(device="cuda").manual_seed(1)
    n = tr.model.xyz_encoder.params.numel()
    tr.opt.flat[:n].copy_(torch.randn(n, device="cuda", generator=g) * 0.3)
    tr.opt.flat.copy_(tr.opt.flat.half().float())
    tr.opt.flat16.copy_(tr.opt.flat)
    field = ostep.CpuField()
    m = tr.model
    assert field.table.numel() == m.xyz_encoder.params.numel()
    with torch.no_grad():
        field.table.copy_(m.xyz_encoder.params.detach().cpu())
        field.sigma_w.copy_(m.sigma_net.params.detach().cpu())
        field.rgb_w.copy_(m.rgb_net.params.detach().cpu())
    b = synth.patch_batch(R, seed=4)
    noise = None
    rgb = torch.rand(R, 3, generator=torch.Generator().manual_seed(2))
    tr.global_step = 3000
    return tr, field, bits, b, noise, rgb, clustering


def _cos(a, b):
    return float((a.double() * b.double()).sum() / (a.double().norm() * b.double().norm()).clamp_min(1e-30))


def _check_losses(got, want, keys, rtol):
    for k in keys:
        a, b = float(got[k]), float(want[k])
        assert abs(a - b) <= rtol * abs(b) + 1e-8, (k, a, b)


def test_module_step_and_fused_step_match_the_cpu_oracle_step(monkeypatch):
    from oracle import step as ostep
    tr, field, bits, b, noise, rgb, clustering = _make()
    dev = "cuda"
    rays_o = torch.from_numpy(b["rays_o"]).to(dev); rays_d = torch.from_numpy(b["rays_d"]).to(dev)
    tri = torch.from_numpy(b["tri"]).to(dev)
    target = {"rgb": rgb.to(dev), "patch_area": 64, "x1_offsets_local": tri[0][:49] % 64, "x2_offsets_local": tri[1][:49] % 64,
              "x3_offsets_local": tri[2][:49] % 64}
    # ---- GPU module path; the jitter RayMarcher.forward draws (torch.rand_like after this seed) is handed to the oracle
    torch.manual_seed(123)
    noise_dev = torch.rand(R, device=dev)
    noise = noise_dev.cpu().numpy()
    captured = {}
    orig_km = clustering.kmeans_spherical

    def km(*a, **k):
        out = orig_km(*a, **k)
        captured["cent"] = out[0].detach().cpu().numpy().copy()
        return out

    monkeypatch.setattr(clustering, "kmeans_spherical", km)
    torch.manual_seed(123)
    results, loss_g = tr.forward_loss(rays_o, rays_d, target)
    (loss_g["total"] * tr.hp["loss_scale"]).backward()
    monkeypatch.undo()
    g_mod = {k: (p.grad / tr.hp["loss_scale"]).detach().cpu().clone() for k, p in tr.model.named_parameters() if p.numel()}
    tr.opt.grad.zero_()
    # ---- CPU oracle step on the same inputs, the GPU engine's centroids
    loss_c, n_samples = ostep.train_step(field, bits, b["rays_o"], b["rays_d"], rgb, b["tri"], step=3000, noise=noise,
                                         centroids=captured["cent"])
    g_cpu = {"xyz_encoder.params": field.table.grad, "sigma_net.params": field.sigma_w.grad, "rgb_net.params": field.rgb_w.grad}
    assert int(results["rm_samples"]) == n_samples > R
    keys = ("rgb", "opacity", "total")
    ckeys = ("norm_D_C_ort_dot", "norm_D_C_centr_dot", "norm_D_C_centr_L1")
    _check_losses(loss_g, loss_c, keys, 5e-3)
    _check_losses(loss_g, loss_c, ckeys, 3e-2)
    assert float(loss_c["norm_D_C_centr_dot"]) > 0 and float(loss_c["norm_D_C_ort_dot"]) > 0
    for k, gc in g_cpu.items():
        gm = g_mod[k]
        assert abs(float(gm.norm()) - float(gc.norm())) <= 5e-2 * float(gc.norm()), (k, float(gm.norm()), float(gc.norm()))
        assert _cos(gm, gc) >= 0.99, (k, _cos(gm, gc))
    # ---- FusedStep on the same inputs
    fs = tr.fused_step(use_graph=False)
    fs.set_triangles(tri)
    fs.rays_o.copy_(rays_o); fs.rays_d.copy_(rays_d); fs.target.copy_(target["rgb"]); fs.noise.copy_(noise_dev)
    fs.gen_noise = False
    fs._schedule()
    fs._run()
    torch.cuda.synchronize()
    assert int(fs.counter[0]) == n_samples
    d, _ = fs.stats_host()
    _check_losses(d, loss_c, keys, 5e-3)
    _check_losses(d, loss_c, ckeys, 3e-2)
    for name in ("xyz_encoder", "sigma_net", "rgb_net"):
        o, k = fs.off[name]
        gf = tr.opt.grad[o:o + k].detach().cpu()
        gc = g_cpu[name + ".params"]
        assert abs(float(gf.norm()) - float(gc.norm())) <= 5e-2 * float(gc.norm()), (name, float(gf.norm()), float(gc.norm()))
        assert _cos(gf, gc) >= 0.99, (name, _cos(gf, gc))
