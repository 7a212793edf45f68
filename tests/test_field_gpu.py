"""GPU parity of the tcnn drop-ins (hash-grid Encoding, fused MLP Network) against the torch
restatements in oracle/ (tcnn itself is absent: parity unpinned, see oracle/hashgrid.py).

Tolerances: features / MLP outputs are fp16 -> |err| <= 2e-3*scale + rtol 1e-2;
gradients are compared with autograd through the fp32/fp64 restatement on identical fp16-rounded
parameters, rtol 2e-2 of the gradient scale (fp16 dL/dy with loss scale 128)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(otype="Grid", type="Hash", n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
           base_resolution=16, per_level_scale=float(np.exp(np.log(2048 * 0.5 / 16) / 15)), interpolation="Linear")


def _enc(log2_T=19, std=0.05, seed=0):
    from ncn_b200 import tinycudann as tcnn
    cfg = dict(CFG, log2_hashmap_size=log2_T)
    enc = tcnn.Encoding(3, cfg).cuda()
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():
        enc.params.copy_((torch.randn(enc.params.numel(), device="cuda", generator=g) * std).half().float())
    return enc, cfg


def _levels(cfg, enc=None):
    """oracle level table; must equal the library's (ncn_grid_desc_init) bit for bit"""
    from oracle import hashgrid
    levels, total = hashgrid.grid_levels(cfg["n_levels"], cfg["n_features_per_level"], cfg["log2_hashmap_size"],
                                         cfg["base_resolution"], cfg["per_level_scale"])
    if enc is not None:
        for a, b in zip(levels, enc.level_table()):
            assert a["res"] == b["res"] and a["size"] == b["size"] and a["offset"] == b["offset"]
            assert np.float32(a["scale"]) == np.float32(b["scale"]), (a, b)
    return levels, total


@pytest.mark.parametrize("log2_T,n", [(19, 100000), (14, 4099), (22, 65536), (19, 1)])
def test_grid_forward(ncn, log2_T, n):
    from oracle import hashgrid
    enc, cfg = _enc(log2_T)
    levels, total = _levels(cfg, enc)
    assert total * 2 == enc.params.numel()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, 3, device="cuda", generator=g)
    if n > 16:
        x[:8] = torch.tensor([[0, 0, 0], [1, 1, 1], [0, 1, 0], [1, 0, 0], [0.5, 0.5, 0.5], [1, 1, 0], [0, 0, 1], [0.25, 1, 0.75]], device="cuda")
    out = enc(x)
    assert out.dtype == torch.float16 and out.shape == (n, 32)
    ref = hashgrid.forward(x, enc.params.detach().view(-1, 2), levels, out_dtype=None)
    err = (out.float() - ref).abs()
    assert err.max() <= 2e-3 * ref.abs().max() + 1e-2 * 0, f"max err {err.max()} vs scale {ref.abs().max()}"


def test_grid_backward_params_and_input(ncn):
    from oracle import hashgrid
    enc, cfg = _enc(15)
    levels, _ = _levels(cfg, enc)
    g = torch.Generator(device="cuda").manual_seed(2)
    n = 20000
    x = torch.rand(n, 3, device="cuda", generator=g).requires_grad_(True)
    dy = torch.randn(n, 32, device="cuda", generator=g)
    out = enc(x)
    (out.float() * dy).sum().backward()
    gp, gx = enc.params.grad.clone(), x.grad.clone()
    xr = x.detach().clone().requires_grad_(True)
    tab = enc.params.detach().clone().view(-1, 2).requires_grad_(True)
    ref = hashgrid.forward(xr, tab, levels, out_dtype=None)
    (ref * dy.half().float()).sum().backward()
    sp = tab.grad.abs().max()
    assert (gp.view(-1, 2) - tab.grad).abs().max() <= 2e-2 * sp
    sx = xr.grad.abs().max()
    assert (gx - xr.grad).abs().max() <= 2e-2 * sx


def test_grid_backward_ray_coherent_samples(ncn):
    """ray-ordered samples (1.7e-3 apart): exercises the warp-level run merging of the scatter; also a device-side count"""
    from oracle import hashgrid
    enc, cfg = _enc(19)
    levels, _ = _levels(cfg, enc)
    g = torch.Generator(device="cuda").manual_seed(5)
    n_rays, per = 700, 41
    o = torch.rand(n_rays, 1, 3, device="cuda", generator=g) * 0.5 + 0.1
    d = torch.nn.functional.normalize(torch.randn(n_rays, 1, 3, device="cuda", generator=g), dim=-1)
    t = torch.arange(per, device="cuda").view(1, per, 1) * 1.6915e-3
    x = (o + t * d).clamp(0, 1).reshape(-1, 3).contiguous()
    n = x.shape[0]
    dy = torch.randn(n, 32, device="cuda", generator=g)
    dy[n // 2: n // 2 + 300] = 0           # a block of dead samples
    out = enc(x)
    (out.float() * dy).sum().backward()
    gp = enc.params.grad.clone()
    tab = enc.params.detach().clone().view(-1, 2).requires_grad_(True)
    ref = hashgrid.forward(x, tab, levels, out_dtype=None)
    (ref * dy.half().float()).sum().backward()
    assert (gp.view(-1, 2) - tab.grad).abs().max() <= 1e-2 * tab.grad.abs().max()
    # merged and direct scatter agree
    from ncn_b200 import _lib
    old = _lib.lib().ncn_set_grid_bwd_merge(0)
    enc.params.grad = None
    (enc(x).float() * dy).sum().backward()
    _lib.lib().ncn_set_grid_bwd_merge(old)
    assert (enc.params.grad - gp).abs().max() <= 1e-3 * gp.abs().max()


def test_grid_double_backward(ncn):
    """d/d(dy) and d/d(params) of <dL/dx, v> (the path a density-gradient normal would use)."""
    from oracle import hashgrid
    enc, cfg = _enc(14, std=0.5)
    levels, _ = _levels(cfg, enc)
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 3000
    x = torch.rand(n, 3, device="cuda", generator=g).requires_grad_(True)
    proj = torch.randn(32, device="cuda", generator=g)
    v = torch.randn(n, 3, device="cuda", generator=g)
    out = enc(x)
    s = (out.float() * proj).sum()
    (dx,) = torch.autograd.grad(s, x, create_graph=True)
    enc.params.grad = None
    (dx * v).sum().backward()
    gp = enc.params.grad.clone()
    xr = x.detach().clone().requires_grad_(True)
    tab = enc.params.detach().clone().view(-1, 2).requires_grad_(True)
    ref = hashgrid.forward(xr, tab, levels, out_dtype=None)
    (dxr,) = torch.autograd.grad((ref * proj).sum(), xr, create_graph=True)
    torch.testing.assert_close(dx.detach(), dxr.detach(), rtol=2e-2, atol=2e-2 * float(dxr.abs().max()))
    (dxr * v).sum().backward()
    assert (gp.view(-1, 2) - tab.grad).abs().max() <= 2e-2 * tab.grad.abs().max()


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("n_in,n_out,n_hidden,act", [(32, 16, 1, "None"), (32, 3, 2, "Sigmoid")])
def test_mlp_backward_implementations_agree(ncn, impl, n_in, n_out, n_hidden, act):
    """ncn_set_mlp_bwd_impl: the warp-MMA kernels and the tcgen05 kernel against the same restatement, ragged tail and
    enough rows that every persistent CTA walks several tiles (prefetch + accumulate paths)."""
    from ncn_b200 import _lib, tinycudann as tcnn
    from oracle import mlp
    n = 148 * 3 * 128 * 3 + 77
    net = tcnn.Network(n_in, n_out, dict(otype="FullyFusedMLP", activation="ReLU", output_activation=act,
                                         n_neurons=64, n_hidden_layers=n_hidden)).cuda()
    with torch.no_grad():
        net.params.copy_(net.params.half().float())
    g = torch.Generator(device="cuda").manual_seed(impl + 10)
    x = (torch.randn(n, n_in, device="cuda", generator=g)).half().float().requires_grad_(True)
    dy = torch.randn(n, n_out, device="cuda", generator=g)
    old = _lib.lib().ncn_set_mlp_bwd_impl(impl)
    try:
        (net(x).float() * dy).sum().backward()
    finally:
        _lib.lib().ncn_set_mlp_bwd_impl(old)
    xr = x.detach().clone().requires_grad_(True)
    pr = net.params.detach().clone().requires_grad_(True)
    refo = mlp.forward(xr, pr, n_in, n_out, n_hidden, act, emulate_half=True)
    (refo.float() * dy.half().float()).sum().backward()
    for got, want in ((x.grad, xr.grad), (net.params.grad, pr.grad)):
        assert (got - want).norm() <= 1e-2 * want.norm() + 1e-5
        assert (got - want).abs().max() <= 8e-2 * want.abs().max() + 1e-4


def test_grid_backward_level_ranges_sum_to_the_whole(ncn):
    """ncn_grid_bwd_levels over [0, 11) and [11, 16) accumulates exactly what one ncn_grid_bwd launch does."""
    import ctypes as C
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    enc, cfg = _enc(19, std=0.1)
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 20000
    x = torch.rand(n, 3, device="cuda", generator=g)
    dy = torch.randn(n, 32, device="cuda", generator=g).half()
    ga = torch.zeros_like(enc.params); gb = torch.zeros_like(enc.params)
    check(L.ncn_grid_bwd(C.byref(enc.desc), ptr(x), ptr(dy), n, ptr(ga), 1.0, None, None, stream()))
    check(L.ncn_grid_bwd_levels(C.byref(enc.desc), ptr(x), ptr(dy), n, ptr(gb), 1.0, None, None, 0, 11, 8, stream()))
    lo = int(enc.desc.level_offset[11]) * 2
    assert float(gb[lo:].abs().max()) == 0.0 and float(gb[:lo].abs().max()) > 0.0
    check(L.ncn_grid_bwd_levels(C.byref(enc.desc), ptr(x), ptr(dy), n, ptr(gb), 1.0, None, None, 11, 16, 6, stream()))
    assert (ga - gb).abs().max() <= 1e-5 * ga.abs().max()


@pytest.mark.parametrize("coherent", [True, False])
def test_grid_backward_fp16_gradient_mode(ncn, coherent):
    """ncn_grid_bwd_f16 (packed fp16 reductions into a __half2 table, tiny-cuda-nn's own accumulation type for F = 2) against
    the fp32 table of ncn_grid_bwd on the same inputs: ray-ordered samples (merge paths) and random ones (direct path).
    Tolerance: fp16 rounding of every accumulated term - entries agree to 2^-10 relative of the level's largest entry
    plus the accumulated rounding of the most-hit coarse entries (1 %)."""
    import ctypes as C
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    enc, cfg = _enc(19, std=0.1)
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(7)
    n = 40000
    if coherent:                                  # 1250 rays x 32 samples, 1.7e-3 apart: long runs on the coarse levels
        o = torch.rand(n // 32, 1, 3, device="cuda", generator=g) * 0.8 + 0.1
        d = torch.randn(n // 32, 1, 3, device="cuda", generator=g); d = d / d.norm(dim=-1, keepdim=True)
        t = torch.arange(32, device="cuda").view(1, 32, 1) * 1.7e-3
        x = (o + t * d).clamp(0, 1).reshape(n, 3).contiguous()
    else:
        x = torch.rand(n, 3, device="cuda", generator=g)
    dy = torch.randn(n, 32, device="cuda", generator=g).half()
    dy[::7] = 0                                   # terminated samples carry exact zeros
    g32 = torch.zeros_like(enc.params)
    g16 = torch.zeros(enc.params.numel(), dtype=torch.float16, device="cuda")
    check(L.ncn_grid_bwd(C.byref(enc.desc), ptr(x), ptr(dy), n, ptr(g32), 1.0, None, None, stream()))
    check(L.ncn_grid_bwd_f16(C.byref(enc.desc), ptr(x), ptr(dy), n, ptr(g16), 1.0, None, None, stream()))
    assert torch.isfinite(g16).all()
    assert (g32 != 0).sum() > 1000
    assert ((g16 != 0) & (g32 == 0)).sum() == 0                     # nothing lands outside the entries the fp32 pass touched
    for l in range(16):
        a, b = int(enc.desc.level_offset[l]) * 2, int(enc.desc.level_offset[l + 1]) * 2
        ref, got = g32[a:b], g16[a:b].float()
        assert (got - ref).abs().max() <= 1e-2 * ref.abs().max() + 1e-6, l
        assert (got - ref).norm() <= 4e-3 * ref.norm() + 1e-6, l
    # F != 2 is refused, not silently converted
    from ncn_b200 import tinycudann as tcnn
    enc4 = tcnn.Encoding(3, dict(cfg, n_features_per_level=4)).cuda()
    assert L.ncn_grid_bwd_f16(C.byref(enc4.desc), ptr(x), ptr(dy), 10, ptr(g16), 1.0, None, None, stream()) != 0


NETS = [(32, 16, 1, "None"), (19, 3, 2, "Sigmoid"), (16, 3, 2, "None"), (16, 40, 2, "None"), (1, 1, 1, "Sigmoid")]


@pytest.mark.parametrize("n_in,n_out,n_hidden,act", NETS)
@pytest.mark.parametrize("n", [1, 100, 16384 + 5])
def test_mlp_forward_backward(ncn, n_in, n_out, n_hidden, act, n):
    from ncn_b200 import tinycudann as tcnn
    from oracle import mlp
    net = tcnn.Network(n_in, n_out, dict(otype="FullyFusedMLP", activation="ReLU", output_activation=act,
                                         n_neurons=64, n_hidden_layers=n_hidden)).cuda()
    with torch.no_grad():
        net.params.copy_(net.params.half().float())
    g = torch.Generator(device="cuda").manual_seed(n)
    x = (torch.randn(n, n_in, device="cuda", generator=g)).half().float().requires_grad_(True)
    out = net(x)
    assert out.dtype == torch.float16 and out.shape == (n, n_out)
    ref = mlp.forward(x.detach(), net.params.detach(), n_in, n_out, n_hidden, act)
    torch.testing.assert_close(out.float(), ref.float(), rtol=1e-2, atol=2e-3)
    dy = torch.randn(n, n_out, device="cuda", generator=g)
    (out.float() * dy).sum().backward()
    gp, gx = net.params.grad.clone(), x.grad.clone()
    xr = x.detach().clone().requires_grad_(True)
    pr = net.params.detach().clone().requires_grad_(True)
    # autograd through the restatement WITH the fp16 rounding of the hidden states (casts are straight-through),
    # so the ReLU gates are the ones the kernel saw
    refo = mlp.forward(xr, pr, n_in, n_out, n_hidden, act, emulate_half=True)
    (refo.float() * dy.half().float()).sum().backward()
    # dL/dz is carried in fp16 between layers (loss scale 128): Frobenius error tight, max error looser
    for got, want in ((gx, xr.grad), (gp, pr.grad)):
        assert (got - want).norm() <= 1e-2 * want.norm() + 1e-5
        assert (got - want).abs().max() <= 8e-2 * want.abs().max() + 1e-4


@pytest.mark.parametrize("n,n_cls", [(1, 3), (1000, 3), (40000 + 7, 13)])
def test_field_heads_fwd_equals_separate_launches(ncn, n, n_cls):
    """ncn_field_heads_fwd (sem_net + norm_net on h in one launch, outputs straight into their raws columns) against
    ncn_mlp_fwd + ncn_field_head_out per head: same MMA chain, so outputs, saved activations and raws are identical;
    untouched raws columns stay untouched; a NULL head is skipped."""
    import ctypes as C
    from ncn_b200 import _lib
    from ncn_b200 import tinycudann as tcnn
    from ncn_b200._lib import check, ptr, stream
    L = _lib.lib()
    cfg = dict(otype="FullyFusedMLP", activation="ReLU", output_activation="None", n_neurons=64, n_hidden_layers=2)
    sem, nrm = tcnn.Network(16, n_cls, cfg).cuda(), tcnn.Network(16, 3, cfg, seed=7).cuda()
    g = torch.Generator(device="cuda").manual_seed(n)
    ws = (torch.randn(sem.params.numel(), device="cuda", generator=g) * 0.2).half()
    wn = (torch.randn(nrm.params.numel(), device="cuda", generator=g) * 0.2).half()
    h = torch.randn(n, 16, device="cuda", generator=g).half()
    Ct = 6 + n_cls
    n_t = (n + 127) // 128 * 128
    st = stream()
    # reference: separate launches
    raws_ref = torch.full((n, Ct), -7.0, device="cuda")
    out_s, out_n = torch.empty(n, 16, dtype=torch.float16, device="cuda"), torch.empty(n, 16, dtype=torch.float16, device="cuda")
    acts_s = torch.zeros(2, n_t, 64, dtype=torch.float16, device="cuda")
    check(L.ncn_mlp_fwd(C.byref(sem.desc), ptr(h), ptr(ws), n, ptr(out_s), ptr(acts_s), None, st))
    check(L.ncn_mlp_fwd(C.byref(nrm.desc), ptr(h), ptr(wn), n, ptr(out_n), None, None, st))
    check(L.ncn_field_head_out(ptr(out_n), 16, n, None, ptr(raws_ref), Ct, 3, 3, st))
    check(L.ncn_field_head_out(ptr(out_s), 16, n, None, ptr(raws_ref), Ct, 6, n_cls, st))
    # one launch
    raws = torch.full((n, Ct), -7.0, device="cuda")
    out_s2 = torch.empty_like(out_s); acts_s2 = torch.zeros_like(acts_s)
    check(L.ncn_field_heads_fwd(ptr(h), n, None, ptr(raws), Ct, ptr(wn), 3, 3, None, None, ptr(ws), 6, n_cls, ptr(acts_s2), ptr(out_s2), st))
    assert torch.equal(raws, raws_ref)
    assert torch.equal(out_s2, out_s) and torch.equal(acts_s2, acts_s)
    assert bool((raws[:, :3] == -7.0).all())
    # device-side row count + absent head
    n_dev = torch.tensor([n // 2], dtype=torch.int32, device="cuda")
    raws3 = torch.full((n, Ct), -7.0, device="cuda")
    check(L.ncn_field_heads_fwd(ptr(h), n, ptr(n_dev), ptr(raws3), Ct, None, 0, 0, None, None, ptr(ws), 6, n_cls, None, None, st))
    assert torch.equal(raws3[:n // 2, 6:], raws_ref[:n // 2, 6:]) and bool((raws3[n // 2:] == -7.0).all()) and bool((raws3[:, :6] == -7.0).all())


@pytest.mark.parametrize("n,n_live", [(1, None), (127, None), (128, None), (129, None), (40000 + 77, None), (20000, 12345), (4096, 0)])
def test_field_mlp_fwd_tcgen05_vs_warp_mma_vs_oracle(ncn, n, n_live):
    """ncn_field_mlp_fwd (density trunk -> TruncExp -> [h | d/|d| | 1] -> colour head, ngp_mt.py:157-229): the tcgen05 / TMEM
    implementation against the warp-MMA one (same fp16 rounding points: every output must agree to fp16 rounding of the hidden
    states) and against the torch restatement oracle/mlp.py; ragged tiles, a device-side live count below the capacity, saved
    activations in the tiled panel layout the backward reads, and the inference form (no saved tensors)."""
    import ctypes as C
    from ncn_b200 import _lib, tinycudann as tcnn
    from ncn_b200._lib import check, ptr, stream
    from oracle import mlp
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(n)
    sig = tcnn.Network(32, 16, dict(otype="FullyFusedMLP", activation="ReLU", output_activation="None", n_neurons=64, n_hidden_layers=1)).cuda()
    rgb = tcnn.Network(19, 3, dict(otype="FullyFusedMLP", activation="ReLU", output_activation="Sigmoid", n_neurons=64, n_hidden_layers=2)).cuda()
    w_sig = (sig.params.detach() * 1.5).half().contiguous(); w_rgb = (rgb.params.detach() * 1.5).half().contiguous()
    feat = (torch.randn(n, 32, device="cuda", generator=g) * 0.5).half().contiguous()
    dirs = torch.randn(n, 3, device="cuda", generator=g).contiguous() * 3.0          # not normalised: the kernel normalises
    n_dev = None if n_live is None else torch.tensor([n_live, 0], dtype=torch.int32, device="cuda")
    live = n if n_live is None else n_live
    Ct = 3
    cap_t = (n + 127) // 128 * 128

    def run(impl, save=True):
        old = L.ncn_set_field_fwd_impl(impl)
        try:
            o = dict(sigmas=torch.full((n,), -7.0, device="cuda"), raws=torch.full((n, Ct), -7.0, device="cuda"),
                     h=torch.zeros(n, 16, dtype=torch.float16, device="cuda"))
            if save:
                o.update(sig_acts=torch.zeros(1, cap_t, 64, dtype=torch.float16, device="cuda"), x_rgb=torch.zeros(n, 32, dtype=torch.float16, device="cuda"),
                         rgb_acts=torch.zeros(2, cap_t, 64, dtype=torch.float16, device="cuda"), rgb_out=torch.zeros(n, 16, dtype=torch.float16, device="cuda"))
            check(L.ncn_field_mlp_fwd(ptr(feat), ptr(dirs), ptr(w_sig), ptr(w_rgb), n, ptr(n_dev) if n_dev is not None else None, ptr(o["sigmas"]),
                                      ptr(o["raws"]), Ct, ptr(o["h"]), ptr(o["sig_acts"]) if save else None, ptr(o["x_rgb"]) if save else None,
                                      ptr(o["rgb_acts"]) if save else None, ptr(o["rgb_out"]) if save else None, stream()), "field_mlp_fwd")
            torch.cuda.synchronize()
            return o
        finally:
            L.ncn_set_field_fwd_impl(old)

    tc, wm = run(1), run(0)
    # rows past the live count are never written
    for o in (tc, wm):
        assert (o["sigmas"][live:] == -7.0).all() and (o["raws"][live:] == -7.0).all()
    if live == 0:
        return
    # the hidden states are rounded to fp16 at the same points: tile-by-tile agreement to a few fp16 ulps of the pre-activations
    torch.testing.assert_close(tc["h"][:live].float(), wm["h"][:live].float(), rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(tc["sigmas"][:live], wm["sigmas"][:live], rtol=4e-3, atol=1e-6)
    torch.testing.assert_close(tc["raws"][:live], wm["raws"][:live], rtol=0, atol=2e-3)
    torch.testing.assert_close(tc["x_rgb"][:live].float(), wm["x_rgb"][:live].float(), rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(tc["rgb_out"][:live].float(), wm["rgb_out"][:live].float(), rtol=0, atol=2e-3)
    # saved activations: same tiled panel layout (act_offset), same values to fp16 rounding; compare row-major views of the live rows
    def rows(a, layers):
        t = a.view(layers, cap_t // 128, 8, 128, 8).permute(0, 1, 3, 2, 4).reshape(layers, cap_t, 64)      # [tile][f/8][row][8] -> (rows, 64)
        return t[:, :live].float()
    torch.testing.assert_close(rows(tc["sig_acts"], 1), rows(wm["sig_acts"], 1), rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(rows(tc["rgb_acts"], 2), rows(wm["rgb_acts"], 2), rtol=2e-3, atol=3e-3)
    # against the oracle (tcnn semantics, App. B): h = sigma_net(feat); rgb = rgb_net([d/|d|, h])
    f32 = feat[:live].float()
    h_ref = mlp.forward(f32, w_sig.float(), 32, 16, 1, "None")
    torch.testing.assert_close(tc["h"][:live].float(), h_ref.float(), rtol=1e-2, atol=3e-3)
    d = dirs[:live] / dirs[:live].norm(dim=1, keepdim=True)
    rgb_ref = mlp.forward(torch.cat([d, tc["h"][:live].float()], 1), w_rgb.float(), 19, 3, 2, "Sigmoid")
    torch.testing.assert_close(tc["raws"][:live], rgb_ref.float(), rtol=0, atol=4e-3)
    torch.testing.assert_close(tc["sigmas"][:live], torch.exp(tc["h"][:live, 0].float()), rtol=1e-5, atol=0)
    # x_rgb = [h | d | 1...] in the fused column order
    torch.testing.assert_close(tc["x_rgb"][:live, 16:19].float(), d.half().float(), rtol=0, atol=1e-3)
    assert (tc["x_rgb"][:live, 19:] == 1).all() and torch.equal(tc["x_rgb"][:live, :16], tc["h"][:live])
    # inference form: nothing saved, same sigma / rgb
    inf = run(1, save=False)
    assert torch.equal(inf["sigmas"][:live], tc["sigmas"][:live]) and torch.equal(inf["raws"][:live], tc["raws"][:live])
