"""GPU parity of the two small boundary pieces round 1 left untested:

  a2   the near clamp of models/rendering.py:28, which libncn fuses into the AABB kernel (ncn_ray_aabb_near), against the
       reference's OWN ray_aabb_intersect (oracle/_ref/vren_ref.so) followed by the reference's clamp expression, with cameras
       inside the box, outside it, grazing a face, and sitting closer to the entry face than `near`;
  a5b  RayMarcher.backward (models/custom_functions.py:102-112): torch_scatter.segment_csr over the per-ray sample segments,
       which libncn provides as ncn_segment_csr_sum, against torch.index_add_ - including rays without samples, a single
       ray, and the gradient routed through the real autograd class.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bits_equal(a, b):
    return torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))


def _ref_aabb_near(vren_ref, rays_o, rays_d, center, half_size, near):
    _, hits_t, _ = vren_ref.ray_aabb_intersect(rays_o, rays_d, center, half_size, 1)
    # models/rendering.py:28, verbatim expression
    hits_t[(hits_t[:, 0, 0] >= 0) & (hits_t[:, 0, 0] < near), 0, 0] = near
    return hits_t


def _camera_sets(scene, near, n=4096):
    """rays with origins inside the box, outside it, on a face, and within `near` of the entry face"""
    from ncn_b200 import synth
    dev = scene["dev"]
    b = synth.random_batch(n, seed=11)
    rays_o = torch.from_numpy(b["rays_o"]).to(dev)
    rays_d = torch.from_numpy(b["rays_d"]).to(dev)
    g = torch.Generator(device="cuda").manual_seed(5)
    sets = {"inside": (rays_o.clone(), rays_d.clone())}
    # outside: origins on a shell of radius 1.2..2 looking roughly at the box (t1 > near on hits, many misses)
    o = torch.randn(n, 3, device=dev, generator=g)
    o = o / o.norm(dim=1, keepdim=True) * (1.2 + 0.8 * torch.rand(n, 1, device=dev, generator=g))
    tgt = (torch.rand(n, 3, device=dev, generator=g) - 0.5) * 1.4          # some targets outside the box -> misses
    d = tgt - o
    d = d / d.norm(dim=1, keepdim=True)
    sets["outside"] = (o.contiguous(), d.contiguous())
    # grazing: origin exactly on the +x face / travelling inside the face plane / starting a hair outside it
    o = rays_o.clone(); d = rays_d.clone()
    o[:, 0] = 0.5
    d[: n // 4, 0] = 0.0                                                   # parallel to the face, inside its plane
    o[n // 4: n // 2, 0] = 0.5 + 1e-7
    o[n // 2: 3 * n // 4, 0] = 0.5 - 1e-7
    sets["grazing"] = (o.contiguous(), d.contiguous())
    # closer to the entry face than `near`: t1 in (0, near) must be lifted to `near`, t1 == 0 too, t1 > near must not move
    o = torch.zeros(n, 3, device=dev)
    d = torch.zeros(n, 3, device=dev); d[:, 0] = 1.0
    o[:, 0] = -0.5 - torch.linspace(0.0, 3.0 * max(near, 0.01), n, device=dev)   # entry distance 0 .. 3 near: both sides of the clamp
    o[:, 1:] = (torch.rand(n, 2, device=dev, generator=g) - 0.5) * 0.8
    sets["near_band"] = (o.contiguous(), d.contiguous())
    return sets


@pytest.mark.parametrize("near", [0.01, 0.05, 0.0])
def test_aabb_near_clamp_matches_reference_kernel_plus_reference_clamp(ncn, vren_ref, scene, near):
    from ncn_b200.rendering import ray_aabb_near
    for name, (rays_o, rays_d) in _camera_sets(scene, near).items():
        ours = ray_aabb_near(rays_o, rays_d, scene["center"], scene["half_size"], near)
        ref = _ref_aabb_near(vren_ref, rays_o, rays_d, scene["center"], scene["half_size"], near)
        assert ours.shape == ref.shape == (rays_o.shape[0], 1, 2), name
        assert _bits_equal(ours, ref), f"{name}: {(ours != ref).sum().item()} differing entries"
        hit = ref[:, 0, 0] >= 0
        if name == "outside":
            assert (~hit).any() and hit.any()                               # both hits and misses were exercised
            assert (ref[hit, 0, 0] > max(near, 0.05)).any()                 # entries beyond near stay untouched
        if name == "near_band" and near > 0:
            assert (ref[:, 0, 0] == near).sum() > 100 and (ref[:, 0, 0] > near).sum() > 100
        if name == "inside" and near > 0:
            assert (ref[hit, 0, 0] == near).all()                           # t1 = 0 inside the box -> near


def test_aabb_near_is_what_the_module_render_and_the_fused_step_march_from(ncn, vren_ref, scene):
    """the hits the marcher consumes (module path and FusedStep share ncn_ray_aabb_near) equal AABB + clamp of the reference
    for the bench's own ray batches"""
    from ncn_b200 import synth
    from ncn_b200.rendering import ray_aabb_near
    b = synth.patch_batch(8192, seed=1000)
    rays_o = torch.from_numpy(b["rays_o"]).cuda(); rays_d = torch.from_numpy(b["rays_d"]).cuda()
    ours = ray_aabb_near(rays_o, rays_d, scene["center"], scene["half_size"], 0.01)
    ref = _ref_aabb_near(vren_ref, rays_o, rays_d, scene["center"], scene["half_size"], 0.01)
    assert _bits_equal(ours, ref)


# ------------------------------------------------------------------ a5b
def _segments(n_rays, max_n, seed, zero_frac=0.3):
    g = torch.Generator(device="cuda").manual_seed(seed)
    n = torch.randint(0, max_n + 1, (n_rays,), device="cuda", generator=g)
    n[torch.rand(n_rays, device="cuda", generator=g) < zero_frac] = 0
    indptr = torch.zeros(n_rays + 1, dtype=torch.int64, device="cuda")
    indptr[1:] = torch.cumsum(n, 0)
    return n, indptr


@pytest.mark.parametrize("n_rays,max_n,dim", [(1, 5, 3), (1, 0, 3), (7, 3, 3), (4096, 70, 3), (8192, 33, 1), (513, 1100, 3), (64, 40, 7)])
def test_segment_csr_sum_vs_index_add(ncn, n_rays, max_n, dim):
    from ncn_b200.custom_functions import segment_sum
    n, indptr = _segments(n_rays, max_n, seed=n_rays + max_n)
    total = int(indptr[-1])
    g = torch.Generator(device="cuda").manual_seed(1)
    src = torch.randn(total, dim, device="cuda", generator=g) if dim > 1 else torch.randn(total, device="cuda", generator=g)
    out = segment_sum(src, indptr)
    seg = torch.repeat_interleave(torch.arange(n_rays, device="cuda"), n)
    want = torch.zeros((n_rays, dim) if dim > 1 else (n_rays,), dtype=torch.float64, device="cuda")
    want.index_add_(0, seg, src.double())
    assert out.shape == want.shape and out.dtype == torch.float32
    assert (out[n == 0] == 0).all()                                         # empty segments are exact zeros, not garbage
    torch.testing.assert_close(out.double(), want, rtol=1e-5, atol=1e-5)
    if max_n <= 32 and total > 0:                                           # short segments: a sequential fp32 sum is reproduced to 1 ulp scale
        want32 = torch.zeros_like(out).index_add_(0, seg, src)
        torch.testing.assert_close(out, want32, rtol=2e-6, atol=2e-6)


def test_segment_csr_shim_is_what_the_reference_imports(ncn):
    """`from torch_scatter import segment_csr` (custom_functions.py:4) resolves to the libncn kernel with the reference's call form"""
    import sys
    ncn.install_shims()
    for k in [k for k in sys.modules if k == "torch_scatter"]:
        del sys.modules[k]
    from torch_scatter import segment_csr
    n, indptr = _segments(300, 20, seed=3)
    src = torch.randn(int(indptr[-1]), 3, device="cuda")
    out = segment_csr(src, indptr)
    seg = torch.repeat_interleave(torch.arange(300, device="cuda"), n)
    want = torch.zeros(300, 3, device="cuda").index_add_(0, seg, src)
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)


def test_ray_marcher_backward_routes_sample_gradients_to_the_rays(ncn, scene):
    """custom_functions.py:102-112: dL/drays_o = segment_csr(dL/dxyzs), dL/drays_d = segment_csr(dL/dxyzs * ts + dL/ddirs);
    checked against index_add_ on the marcher's own rays_a, with rays that got no sample (misses) in the batch"""
    from ncn_b200 import synth
    from ncn_b200.custom_functions import RayMarcher
    from ncn_b200.rendering import ray_aabb_near
    b = synth.random_batch(4096, seed=21)
    rays_o = torch.from_numpy(b["rays_o"]).cuda().requires_grad_(True)
    rays_d = torch.from_numpy(b["rays_d"]).cuda()
    with torch.no_grad():                                                   # a third of the rays start outside and miss the box
        rays_o.data[::3] = rays_o.data[::3] * 0 + torch.tensor([3.0, 3.0, 3.0], device="cuda")
    rays_d = rays_d.clone().requires_grad_(True)
    hits_t = ray_aabb_near(rays_o.detach(), rays_d.detach(), scene["center"], scene["half_size"], 0.01)
    torch.manual_seed(0)
    rays_a, xyzs, dirs, deltas, ts, total = RayMarcher.apply(rays_o, rays_d, hits_t[:, 0], scene["bitfield"], 1, 0.5, 0.0, 128, 1024)
    N = int(total)
    assert N > 4096 and (rays_a[:, 2] == 0).any()
    g = torch.Generator(device="cuda").manual_seed(2)
    gx = torch.randn(N, 3, device="cuda", generator=g); gd = torch.randn(N, 3, device="cuda", generator=g)
    (xyzs[:N] * gx).sum().add((dirs[:N] * gd).sum()).backward()
    seg = torch.repeat_interleave(rays_a[:, 0], rays_a[:, 2])
    # rays_a is in ray order with contiguous segments (our marcher's canonical layout)
    assert torch.equal(rays_a[:, 0], torch.arange(4096, device="cuda"))
    want_o = torch.zeros(4096, 3, dtype=torch.float64, device="cuda").index_add_(0, seg, gx.double())
    want_d = torch.zeros(4096, 3, dtype=torch.float64, device="cuda").index_add_(0, seg, (gx * ts[:N, None] + gd).double())
    torch.testing.assert_close(rays_o.grad.double(), want_o, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rays_d.grad.double(), want_d, rtol=1e-5, atol=1e-5)
    assert (rays_o.grad[rays_a[:, 2] == 0] == 0).all()
    # and xyzs really is o + t d, so the analytic gradient above is the gradient of what was returned
    recon = rays_o.detach()[seg] + ts[:N, None] * rays_d.detach()[seg]
    torch.testing.assert_close(xyzs[:N], recon, rtol=0, atol=1e-6)
