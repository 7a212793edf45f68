"""GPU parity of the vren drop-in (hand-written sm_100a kernels through the C-ABI) against the
reference's OWN csrc kernels compiled for sm_100 (oracle/_ref/vren_ref.so).

Bar: bit-exact for integer / index outputs and for every fp32 marching / intersection value;
compositing sums within rtol 1e-5 / atol 1e-6 (the warp-tree summation order differs), its
integer sample counts and ws exactly.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(scene, kind="patch", n=8192, seed=0, cam="hypersim"):
    from ncn_b200 import synth
    b = synth.patch_batch(n, cam=cam, seed=seed) if kind == "patch" else synth.random_batch(n, cam=cam, seed=seed)
    dev = scene["dev"]
    return torch.from_numpy(b["rays_o"]).to(dev), torch.from_numpy(b["rays_d"]).to(dev)


def _bits_equal(a, b):
    return torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))


# ------------------------------------------------------------------ integer kernels
@pytest.mark.parametrize("n", [0, 1, 3, 4, 7, 1000, 128 ** 3])
def test_morton_roundtrip_and_ref(ncn, vren_ref, n):
    from ncn_b200 import vren
    g = torch.Generator(device="cuda").manual_seed(n)
    coords = torch.randint(0, 1024, (n, 3), dtype=torch.int32, device="cuda", generator=g)
    idx = vren.morton3D(coords)
    back = vren.morton3D_invert(idx)
    assert torch.equal(back, coords)
    if n > 0:
        assert torch.equal(idx, vren_ref.morton3D(coords))
        assert torch.equal(back, vren_ref.morton3D_invert(idx))
        # arbitrary (also negative) codes: the reference shifts the signed value
        raw = torch.randint(-2 ** 31, 2 ** 31 - 1, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
        assert torch.equal(vren.morton3D_invert(raw), vren_ref.morton3D_invert(raw))


@pytest.mark.parametrize("n_bytes,thr", [(1, 0.5), (3, 0.0), (5, 0.5), (4096, 0.01), (128 ** 3 // 8, 5.9)])
def test_packbits(ncn, vren_ref, n_bytes, thr):
    from ncn_b200 import vren
    g = torch.Generator(device="cuda").manual_seed(n_bytes)
    grid = torch.randn(n_bytes * 8, device="cuda", generator=g) * 6
    grid[::7] = -1.0
    grid[::11] = thr            # equality must NOT set the bit
    ours = torch.zeros(n_bytes, dtype=torch.uint8, device="cuda")
    ref = torch.zeros(n_bytes, dtype=torch.uint8, device="cuda")
    vren.packbits(grid, thr, ours)
    vren_ref.packbits(grid, thr, ref)
    assert torch.equal(ours, ref)
    npb = np.packbits(grid.cpu().numpy() > np.float32(thr), bitorder="little")
    assert np.array_equal(ours.cpu().numpy(), npb)


# ------------------------------------------------------------------ intersection
def test_aabb_single_box(ncn, vren_ref, scene):
    from ncn_b200 import vren
    rays_o, rays_d = _batch(scene, "random", 65536, seed=3)
    rays_d = rays_d.clone()
    rays_d[:100, 0] = 0.0               # axis-parallel rays -> inf reciprocals
    rays_o[100:200] = rays_o[100:200] * 10   # origins outside the box (some miss)
    cnt, t, idx = vren.ray_aabb_intersect(rays_o, rays_d, scene["center"], scene["half_size"], 1)
    rcnt, rt, ridx = vren_ref.ray_aabb_intersect(rays_o, rays_d, scene["center"], scene["half_size"], 1)
    assert torch.equal(cnt, rcnt) and torch.equal(idx, ridx) and _bits_equal(t, rt)


def test_aabb_many_boxes_and_spheres(ncn, vren_ref, scene):
    from ncn_b200 import vren
    rays_o, rays_d = _batch(scene, "random", 4096, seed=4)
    g = torch.Generator(device="cuda").manual_seed(0)
    centers = (torch.rand(5, 3, device="cuda", generator=g) - 0.5) * 0.8
    halfs = torch.rand(5, 3, device="cuda", generator=g) * 0.2 + 0.02
    radii = torch.rand(5, device="cuda", generator=g) * 0.2 + 0.05
    for max_hits in (1, 3, 8):
        for ours_f, ref_f, ext in ((vren.ray_aabb_intersect, vren_ref.ray_aabb_intersect, halfs),
                                   (vren.ray_sphere_intersect, vren_ref.ray_sphere_intersect, radii)):
            cnt, t, idx = ours_f(rays_o, rays_d, centers, ext, max_hits)
            rcnt, rt, ridx = ref_f(rays_o, rays_d, centers, ext, max_hits)
            assert torch.equal(cnt, rcnt)
            full = cnt <= max_hits      # when more volumes hit than slots the reference keeps a racy subset
            # sorted t1 sequences must agree bit for bit (ties may permute idx)
            assert _bits_equal(t[full][..., 0], rt[full][..., 0])
            assert _bits_equal(torch.sort(t[full][..., 1], 1)[0], torch.sort(rt[full][..., 1], 1)[0])
            assert torch.equal(torch.sort(idx[full], 1)[0], torch.sort(ridx[full], 1)[0])


def test_check_input_errors(ncn):
    from ncn_b200 import vren
    o = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="rays_o must be a CUDA tensor"):
        vren.ray_aabb_intersect(o, o, o, o, 1)
    oc = torch.zeros(3, 4, device="cuda").t()
    with pytest.raises(RuntimeError, match="rays_o must be contiguous"):
        vren.ray_aabb_intersect(oc, oc, oc, oc, 1)


# ------------------------------------------------------------------ marching
def _hits(vren_mod, scene, rays_o, rays_d, near=0.01):
    _, hits_t, _ = vren_mod.ray_aabb_intersect(rays_o, rays_d, scene["center"], scene["half_size"], 1)
    m = (hits_t[:, 0, 0] >= 0) & (hits_t[:, 0, 0] < near)
    hits_t[m, 0, 0] = near
    return hits_t


def _canon(rays_a, arrays):
    """reorder a (possibly racy) reference layout into ray order"""
    order = torch.argsort(rays_a[:, 0])
    ra = rays_a[order]
    n = ra[:, 2]
    starts = ra[:, 1]
    within = torch.arange(int(n.sum()), device=n.device) - torch.repeat_interleave(torch.cumsum(n, 0) - n, n)
    src = torch.repeat_interleave(starts, n) + within
    return ra, [a[src] for a in arrays]


@pytest.mark.parametrize("kind,n,exp_step_factor,cam", [("patch", 8192, 0.0, "hypersim"),
                                                        ("random", 65536, 0.0, "hypersim"),
                                                        ("random", 8192, 1.0 / 256, "scannet"),
                                                        ("random", 5, 0.0, "hypersim")])
def test_march_train_bit_exact(ncn, vren_ref, scene, kind, n, exp_step_factor, cam):
    from ncn_b200 import vren
    rays_o, rays_d = _batch(scene, kind, n, seed=7, cam=cam)
    if n > 1000:
        rays_d = rays_d.clone(); rays_d[:64, 1] = 0.0      # axis-parallel
        rays_o = rays_o.clone(); rays_o[64:128] *= 5.0     # start outside / miss
    hits_t = _hits(vren_ref, scene, rays_o, rays_d)
    noise = torch.rand(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    args = (rays_o, rays_d, hits_t[:, 0], scene["bitfield"], scene["cascades"], scene["scale"], exp_step_factor,
            noise, scene["grid_size"], scene["max_samples"])
    ra, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(*args)
    rra, rxyzs, rdirs, rdeltas, rts, rcounter = vren_ref.raymarching_train(*args)
    assert counter.tolist() == rcounter.tolist()
    total = int(counter[0])
    assert xyzs.shape[0] == total
    cra, (cx, cd, cdt, ct) = _canon(rra, [rxyzs, rdirs, rdeltas, rts])
    assert torch.equal(ra[:, 0], cra[:, 0]) and torch.equal(ra[:, 2], cra[:, 2])
    assert torch.equal(ra[:, 1], torch.cumsum(ra[:, 2], 0) - ra[:, 2])
    assert _bits_equal(ts, ct) and _bits_equal(deltas, cdt) and _bits_equal(xyzs, cx) and _bits_equal(dirs, cd)
    if n > 1000:
        assert total > 5 * n      # the scene is not degenerate


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("kind,n", [("patch", 8192), ("random", 30000)])
def test_march_train_segment_modes(ncn, vren_ref, scene, kind, n, mode):
    """constant-step path: one lane per ray (0), four speculative lanes per ray (1, default) and four lanes with every
    segment re-marched from its predecessor's landing point (2, the repair path) - all bit-identical to the reference."""
    from ncn_b200 import _lib, vren
    rays_o, rays_d = _batch(scene, kind, n, seed=11, cam="hypersim")
    rays_d = rays_d.clone(); rays_d[:64, 1] = 0.0
    rays_o = rays_o.clone(); rays_o[64:128] *= 5.0
    hits_t = _hits(vren_ref, scene, rays_o, rays_d)
    noise = torch.rand(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    args = (rays_o, rays_d, hits_t[:, 0], scene["bitfield"], scene["cascades"], scene["scale"], 0.0,
            noise, scene["grid_size"], scene["max_samples"])
    old = _lib.lib().ncn_set_march_segments(mode)
    try:
        ra, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(*args)
    finally:
        _lib.lib().ncn_set_march_segments(old)
    rra, rxyzs, rdirs, rdeltas, rts, rcounter = vren_ref.raymarching_train(*args)
    assert counter.tolist() == rcounter.tolist()
    cra, (cx, cd, cdt, ct) = _canon(rra, [rxyzs, rdirs, rdeltas, rts])
    assert torch.equal(ra[:, 2], cra[:, 2])
    assert _bits_equal(ts, ct) and _bits_equal(deltas, cdt) and _bits_equal(xyzs, cx) and _bits_equal(dirs, cd)
    assert int(counter[0]) > 5 * n            # non-degenerate scene: rays cross several 256-candidate segments


def test_march_train_multi_cascade(ncn, vren_ref):
    """cascades > 1 (scale 2): mip selection from position and step size."""
    from ncn_b200 import vren, synth
    G, C, scale = 64, 3, 2.0
    g = torch.Generator(device="cuda").manual_seed(5)
    bitfield = (torch.rand(C * G ** 3 // 8, device="cuda", generator=g) < 0.15).to(torch.uint8) * \
        torch.randint(1, 256, (C * G ** 3 // 8,), device="cuda", generator=g).to(torch.uint8)
    b = synth.random_batch(16384, seed=9)
    rays_o = torch.from_numpy(b["rays_o"]).cuda() * 3
    rays_d = torch.from_numpy(b["rays_d"]).cuda()
    center = torch.zeros(1, 3, device="cuda"); half = torch.full((1, 3), scale, device="cuda")
    _, hits_t, _ = vren_ref.ray_aabb_intersect(rays_o, rays_d, center, half, 1)
    noise = torch.rand(16384, device="cuda", generator=g)
    for esf in (0.0, 1.0 / 256):
        args = (rays_o, rays_d, hits_t[:, 0], bitfield, C, scale, esf, noise, G, 256)
        ra, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(*args)
        rra, rxyzs, rdirs, rdeltas, rts, rcounter = vren_ref.raymarching_train(*args)
        assert counter.tolist() == rcounter.tolist()
        cra, (cx, cd, cdt, ct) = _canon(rra, [rxyzs, rdirs, rdeltas, rts])
        assert torch.equal(ra[:, 2], cra[:, 2])
        assert _bits_equal(ts, ct) and _bits_equal(deltas, cdt) and _bits_equal(xyzs, cx)


@pytest.mark.parametrize("n_samples,esf", [(1, 0.0), (4, 0.0), (64, 0.0), (8, 1.0 / 256)])
def test_march_test_bit_exact(ncn, vren_ref, scene, n_samples, esf):
    from ncn_b200 import vren
    rays_o, rays_d = _batch(scene, "random", 20000, seed=11)
    hits_a = _hits(vren_ref, scene, rays_o, rays_d)
    hits_b = hits_a.clone()
    alive = torch.arange(20000, device="cuda")[::2].contiguous()
    for it in range(3):   # resume: hits_t is advanced in place
        a = vren.raymarching_test(rays_o, rays_d, hits_a[:, 0], alive, scene["bitfield"], 1, 0.5, esf, 128, 1024, n_samples)
        b = vren_ref.raymarching_test(rays_o, rays_d, hits_b[:, 0], alive, scene["bitfield"], 1, 0.5, esf, 128, 1024, n_samples)
        for x, y in zip(a[:4], b[:4]):
            assert _bits_equal(x, y)
        assert torch.equal(a[4], b[4])
        assert _bits_equal(hits_a, hits_b)


# ------------------------------------------------------------------ compositing
def _samples(scene, vren_mod, n_rays=8192, C=3, seed=0, sigma_scale=30.0):
    rays_o, rays_d = _batch(scene, "patch", n_rays, seed=seed)
    hits_t = _hits(vren_mod, scene, rays_o, rays_d)
    noise = torch.rand(n_rays, device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed))
    from ncn_b200 import vren
    ra, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(
        rays_o, rays_d, hits_t[:, 0], scene["bitfield"], 1, 0.5, 0.0, noise, 128, 1024)
    N = xyzs.shape[0]
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    sigmas = torch.rand(N, device="cuda", generator=g) ** 4 * sigma_scale * 20
    raws = torch.rand(N, C, device="cuda", generator=g)
    return ra, sigmas, raws, deltas, ts


@pytest.mark.parametrize("C,sigma_scale", [(3, 30.0), (3, 0.5), (6, 30.0), (9, 5.0), (43, 30.0)])
def test_composite_train_fw_bw(ncn, vren_ref, scene, C, sigma_scale):
    from ncn_b200 import vren
    ra, sigmas, raws, deltas, ts = _samples(scene, vren_ref, C=C, sigma_scale=sigma_scale)
    ours = vren.composite_train_multi_fw(sigmas, raws, deltas, ts, ra, 1e-4)
    ref = vren_ref.composite_train_multi_fw(sigmas, raws, deltas, ts, ra, 1e-4)
    assert torch.equal(ours[0], ref[0])                 # total_samples per ray, exact
    assert _bits_equal(ours[4], ref[4])                 # ws: exact (sequential transmittance replay)
    for a, b in zip(ours[1:4], ref[1:4]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    assert (ours[0] < ra[:, 2]).any() or sigma_scale <= 5   # early termination is exercised
    if C == 3:
        o3 = vren.composite_train_fw(sigmas, raws, deltas, ts, ra, 1e-4)
        for a, b in zip(o3, ours):
            assert torch.equal(a, b)
    g = torch.Generator(device="cuda").manual_seed(3)
    R, N = ra.shape[0], sigmas.shape[0]
    dO = torch.randn(R, device="cuda", generator=g); dD = torch.randn(R, device="cuda", generator=g)
    dR = torch.randn(R, C, device="cuda", generator=g); dW = torch.randn(N, device="cuda", generator=g)
    _, opacity, depth, rend, ws = ref
    for dws in (dW, torch.zeros_like(dW)):
        go = vren.composite_train_multi_bw(dO, dD, dR, dws, sigmas, raws, ws, deltas, ts, ra, opacity, depth, rend, 1e-4)
        gr = vren_ref.composite_train_multi_bw(dO, dD, dR, dws, sigmas, raws, ws, deltas, ts, ra, opacity, depth, rend, 1e-4)
        torch.testing.assert_close(go[1], gr[1], rtol=1e-5, atol=1e-6)
        # dL_dsigmas carries the cancellation (D - d): compare relative to the per-ray gradient scale
        scale = gr[0].abs().max()
        assert (go[0] - gr[0]).abs().max() <= 2e-5 * scale + 1e-6


def test_composite_test_bit_exact(ncn, vren_ref, scene):
    from ncn_b200 import vren
    g = torch.Generator(device="cuda").manual_seed(0)
    R, A, S, C = 5000, 3000, 8, 6
    alive = torch.randperm(R, device="cuda", generator=g)[:A].contiguous()
    sig = torch.rand(A, S, device="cuda", generator=g) * 200
    raws = torch.rand(A, S, C, device="cuda", generator=g)
    deltas = torch.full((A, S), 1.7e-3, device="cuda"); ts = torch.rand(A, S, device="cuda", generator=g)
    n_eff = torch.randint(0, S + 1, (A,), device="cuda", generator=g, dtype=torch.int32)
    hits_t = torch.zeros(R, 2, device="cuda")
    st = [torch.rand(R, device="cuda", generator=g) * 0.5, torch.rand(R, device="cuda", generator=g),
          torch.rand(R, C, device="cuda", generator=g)]
    a = [x.clone() for x in st]; b = [x.clone() for x in st]
    al_a, al_b = alive.clone(), alive.clone()
    vren.composite_test_multi_fw(sig, raws, deltas, ts, hits_t, al_a, 1e-4, n_eff, *a)
    vren_ref.composite_test_multi_fw(sig, raws, deltas, ts, hits_t, al_b, 1e-4, n_eff, *b)
    assert torch.equal(al_a, al_b)
    for x, y in zip(a, b):
        assert _bits_equal(x, y)


# ------------------------------------------------------------------ distortion loss
def test_distortion(ncn, vren_ref, scene):
    from ncn_b200 import vren
    ra, sigmas, raws, deltas, ts = _samples(scene, vren_ref, n_rays=4096, sigma_scale=2.0)
    ws = vren_ref.composite_train_multi_fw(sigmas, raws, deltas, ts, ra, 1e-4)[4]
    ours = vren.distortion_loss_fw(ws, deltas, ts, ra)
    ref = vren_ref.distortion_loss_fw(ws, deltas, ts, ra)
    for a, b in zip(ours, ref):
        assert _bits_equal(a, b)
    dl = torch.randn(ra.shape[0], device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    go = vren.distortion_loss_bw(dl, ours[1], ours[2], ws, deltas, ts, ra)
    gr = vren_ref.distortion_loss_bw(dl, ref[1], ref[2], ws, deltas, ts, ra)
    torch.testing.assert_close(go, gr, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------ committed golden vectors (reference csrc on B200)
import glob
import os

_GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vren_ref_*.npz")))


@pytest.mark.parametrize("path", _GOLD)
def test_against_committed_golden(ncn, scene, path):
    from ncn_b200 import vren
    g = np.load(path)
    t = lambda k: torch.from_numpy(g[k]).cuda()
    _, hits, _ = vren.ray_aabb_intersect(t("rays_o"), t("rays_d"), scene["center"], scene["half_size"], 1)
    assert _bits_equal(hits[:, 0], t("hits_raw"))
    ra, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(t("rays_o"), t("rays_d"), t("hits_t"), scene["bitfield"], 1, 0.5,
                                                                 float(g["esf"]), t("noise"), 128, 1024)
    assert torch.equal(ra, t("rays_a"))
    assert _bits_equal(ts, t("ts")) and _bits_equal(deltas, t("deltas")) and _bits_equal(xyzs, t("xyzs")) and _bits_equal(dirs, t("dirs"))
    out = vren.composite_train_multi_fw(t("sigmas"), t("raws"), deltas, ts, ra, 1e-4)
    assert torch.equal(out[0], t("total_samples")) and _bits_equal(out[4], t("ws"))
    for a, k in zip(out[1:4], ("opacity", "depth", "rend")):
        torch.testing.assert_close(a, t(k), rtol=1e-5, atol=1e-6)
    assert torch.equal(vren.morton3D(t("coords")), t("morton"))
