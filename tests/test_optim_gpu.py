"""GPU: fused Adam / clip / sum-of-squares kernels against the apex FusedAdam (adam_w_mode) update restated in fp64
(train_nerf.py:262-285: eps 1e-15, weight decay 0 / 1e-6; clip_grad_norm_ 0.05, train_nerf.py:955)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_step(p, g, m, v, lr, b1, b2, eps, wd, t, gdiv, coef):
    g = g.double() / gdiv * coef
    m = b1 * m.double() + (1 - b1) * g
    v = b2 * v.double() + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    upd = (m / bc1) / ((v / bc2).sqrt() + eps) + wd * p.double()
    return p.double() - lr * upd, m, v


@pytest.mark.parametrize("n,wd", [(11445040, 0.0), (10243, 1e-6), (5, 1e-6)])
def test_adam_matches_reference_formula(ncn, n, wd):
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(n)
    n_pad = (n + 3) // 4 * 4
    p = torch.randn(n_pad, device="cuda", generator=g) * 0.1
    grad = torch.randn(n_pad, device="cuda", generator=g) * 3.0
    grad[::7] = 0
    m = torch.randn(n_pad, device="cuda", generator=g) * 0.01
    v = torch.rand(n_pad, device="cuda", generator=g) * 1e-4
    p16 = torch.zeros(n_pad, dtype=torch.float16, device="cuda")
    gdiv = torch.tensor([1024.0], device="cuda")
    sumsq = torch.zeros(1, device="cuda"); flag = torch.zeros(1, dtype=torch.int32, device="cuda"); coef = torch.ones(1, device="cuda")
    check(L.ncn_grad_sumsq(ptr(grad), n, ptr(gdiv), ptr(sumsq), ptr(flag), stream()))
    check(L.ncn_clip_coef(ptr(sumsq), 0.05, ptr(coef), stream()))
    want_norm = float((grad[:n].double() / 1024).norm())
    assert abs(math.sqrt(float(sumsq)) - want_norm) <= 1e-4 * want_norm and int(flag) == 0
    want_coef = min(1.0, 0.05 / (want_norm + 1e-6))
    assert abs(float(coef) - want_coef) <= 1e-4 * want_coef
    p_old = p.clone()
    rp, rm, rv = _ref_step(p[:n], grad[:n], m[:n], v[:n], 1e-2, 0.9, 0.999, 1e-15, wd, 7, 1024.0, float(coef))
    check(L.ncn_adam_step(ptr(p), ptr(grad), ptr(m), ptr(v), ptr(p16), n, 1e-2, 0.9, 0.999, 1e-15, wd, 7, ptr(gdiv), ptr(flag), ptr(coef),
                          None, stream()))
    # compare the applied update (fp32 storage of p bounds the absolute error by ulp(p))
    torch.testing.assert_close((p[:n] - p_old[:n]).double(), rp - p_old[:n].double(), rtol=1e-4, atol=2e-7)
    torch.testing.assert_close(m[:n].double(), rm, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(v[:n].double(), rv, rtol=1e-5, atol=1e-12)
    assert (grad[:n] == 0).all()                                   # the pass zeroes the gradient
    assert torch.equal(p16[:n], p[:n].half())                      # and refreshes the fp16 working copy
    # device-side schedule (graph replay) == host scalars
    p2 = p.clone(); m2 = m.clone(); v2 = v.clone(); g2 = torch.randn(n_pad, device="cuda", generator=g)
    p3 = p.clone(); m3 = m.clone(); v3 = v.clone(); g3 = g2.clone()
    check(L.ncn_adam_step(ptr(p2), ptr(g2), ptr(m2), ptr(v2), None, n, 3e-3, 0.9, 0.999, 1e-15, wd, 9, None, None, None, None, stream()))
    sched = torch.tensor([3e-3, 1 - 0.9 ** 9, 1 - 0.999 ** 9], device="cuda")
    check(L.ncn_adam_step(ptr(p3), ptr(g3), ptr(m3), ptr(v3), None, n, 0.0, 0.9, 0.999, 1e-15, wd, 1, None, None, None, ptr(sched), stream()))
    torch.testing.assert_close(p2 - p, p3 - p, rtol=1e-4, atol=1e-6)   # host powf vs python pow for the bias corrections


def test_adam_skips_on_nonfinite(ncn):
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    L = _lib.lib()
    n = 4096
    p = torch.ones(n, device="cuda"); grad = torch.ones(n, device="cuda"); grad[100] = float("inf")
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    sumsq = torch.zeros(1, device="cuda"); flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    check(L.ncn_grad_sumsq(ptr(grad), n, None, ptr(sumsq), ptr(flag), stream()))
    assert int(flag) == 1
    check(L.ncn_adam_step(ptr(p), ptr(grad), ptr(m), ptr(v), None, n, 1e-2, 0.9, 0.999, 1e-15, 0.0, 1, None, ptr(flag), None, None, stream()))
    assert (p == 1).all() and (m == 0).all() and (grad == 0).all()


def test_adam_groups_equals_per_group_launches(ncn):
    """ncn_adam_step_groups (both parameter groups + the clip coefficient in one launch) == ncn_clip_coef + one
    ncn_adam_step per group, bit for bit."""
    import ctypes as C
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    n0, n1 = 400004, 10244                       # hash-table group (wd 0) and MLP group (wd 1e-6), starts multiples of 4
    n = n0 + n1
    mk = lambda s: torch.randn(n, device="cuda", generator=g) * s
    p, grad, m, v = mk(0.1), mk(3.0), mk(0.01), mk(1e-2).abs() * 1e-2
    gdiv = torch.tensor([128.0], device="cuda")
    sched = torch.tensor([1e-2, 1 - 0.9 ** 5, 1 - 0.999 ** 5], device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    sumsq = torch.zeros(1, device="cuda"); coef = torch.ones(1, device="cuda")
    check(L.ncn_grad_sumsq(ptr(grad), n, ptr(gdiv), ptr(sumsq), ptr(flag), stream()))
    check(L.ncn_clip_coef(ptr(sumsq), 0.05, ptr(coef), stream()))
    assert float(coef) < 1.0                      # the clip is active
    a = [t.clone() for t in (p, grad, m, v)]; a16 = torch.zeros(n, dtype=torch.float16, device="cuda")
    b = [t.clone() for t in (p, grad, m, v)]; b16 = torch.zeros(n, dtype=torch.float16, device="cuda")
    for (start, cnt, wd) in ((0, n0, 0.0), (n0, n1, 1e-6)):
        sl = slice(start, start + cnt)
        check(L.ncn_adam_step(ptr(a[0][sl]), ptr(a[1][sl]), ptr(a[2][sl]), ptr(a[3][sl]), ptr(a16[sl]), cnt, 0.0, 0.9, 0.999, 1e-15, wd, 1,
                              ptr(gdiv), ptr(flag), ptr(coef), ptr(sched), stream()))
    grp = _lib.AdamGroups()
    grp.n_groups = 2
    grp.start[0], grp.start[1] = 0, n0
    grp.weight_decay[0], grp.weight_decay[1] = 0.0, 1e-6
    grp.max_norm = 0.05
    check(L.ncn_adam_step_groups(ptr(b[0]), ptr(b[1]), ptr(b[2]), ptr(b[3]), ptr(b16), n, C.byref(grp), 0.9, 0.999, 1e-15,
                                 ptr(gdiv), ptr(flag), ptr(sumsq), ptr(sched), stream()))
    torch.cuda.synchronize()
    for x, y in zip(a + [a16], b + [b16]):
        assert torch.equal(x, y)
    assert not torch.equal(b[0], p)


def test_sumsq_is_deterministic(ncn):
    """same gradient -> same bits, call after call (the data-parallel ranks derive the clip coefficient from it)"""
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(11)
    grad = torch.randn(11445040 + 10240, device="cuda", generator=g)
    outs = []
    for _ in range(6):
        out = torch.zeros(1, device="cuda"); flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        check(L.ncn_grad_sumsq(ptr(grad), grad.numel(), None, ptr(out), ptr(flag), stream()))
        outs.append(out.clone())
    torch.cuda.synchronize()
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    want = float((grad.double() ** 2).sum())
    assert abs(float(outs[0]) - want) <= 1e-5 * want


def test_peer_step_single_rank_equals_sumsq_plus_adam(ncn):
    """ncn_peer_step at world_size 1 (no peers: the shard is the whole vector) against ncn_grad_sumsq + ncn_adam_step_groups:
    same parameters / moments / fp16 copy (the norm is summed in a different block order -> the clip coefficient may differ
    in the last ulp), gradient buffer zeroed, non-finite gradient skips the update, epoch advances once per step."""
    import ctypes as C
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    from ncn_b200.trainer import PeerLink
    L = _lib.lib()
    n = 1 << 20
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device="cuda").manual_seed(0)
    link = PeerLink(0, 1, n, dev)
    assert link.shard == (0, n) and link.grad.numel() == n and link.p16.dtype == torch.float16
    p0 = torch.randn(n, device="cuda", generator=g) * 0.1
    groups = _lib.AdamGroups(); groups.n_groups = 2; groups.start[0] = 0; groups.start[1] = n - 4096
    groups.weight_decay[0] = 0.0; groups.weight_decay[1] = 1e-6; groups.max_norm = 0.05
    div = torch.tensor([2.0], device="cuda")
    state = {k: [p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")] for k in ("peer", "ref")}
    p16_ref = torch.zeros(n, dtype=torch.float16, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    for step in range(1, 4):
        grad = torch.randn(n, device="cuda", generator=g) * (10.0 ** -step)
        lr_bc = torch.tensor([1e-2, 1 - 0.9 ** step, 1 - 0.999 ** step], device="cuda")
        link.grad.copy_(grad)
        sumsq_p = torch.zeros(1, device="cuda")
        p, m, v = state["peer"]
        link.step(p, m, v, groups, (0.9, 0.999), 1e-15, div, flag, lr_bc, sumsq_p, stream())
        gr = grad.clone(); sumsq = torch.zeros(1, device="cuda")
        p, m, v = state["ref"]
        check(L.ncn_grad_sumsq(ptr(gr), n, ptr(div), ptr(sumsq), ptr(flag), stream()))
        check(L.ncn_adam_step_groups(ptr(p), ptr(gr), ptr(m), ptr(v), ptr(p16_ref), n, C.byref(groups), 0.9, 0.999, 1e-15, ptr(div), ptr(flag),
                                     ptr(sumsq), ptr(lr_bc), stream()))
        torch.testing.assert_close(sumsq_p, sumsq, rtol=1e-5, atol=0)
        for a, b in zip(state["peer"], state["ref"]):
            torch.testing.assert_close(a, b, rtol=2e-5, atol=1e-9)
        torch.testing.assert_close(link.p16.float(), p16_ref.float(), rtol=2e-3, atol=1e-7)
        assert float(link.grad.abs().max()) == 0.0
    assert link.error() == 0
    before = state["peer"][0].clone()
    link.grad.copy_(torch.full((n,), float("nan"), device="cuda"))
    link.step(*state["peer"], groups, (0.9, 0.999), 1e-15, div, flag, lr_bc, sumsq_p, stream())
    torch.cuda.synchronize()
    assert torch.equal(state["peer"][0], before) and float(link.grad.abs().max()) == 0.0 and not torch.isfinite(sumsq_p).all()
    link.close()


def test_fused_step_with_sharded_optimizer_single_rank():
    """the fused CUDA-graph step driven by the peer-memory optimizer (world_size 1) trains like the replicated one"""
    from test_fused_gpu import _setup
    from ncn_b200.trainer import NeRFTrainer
    tr, rays_o, rays_d, tri, rgb, target = _setup(R=1024, seed=1)
    torch.manual_seed(1)
    tr2 = NeRFTrainer(dict(batch_size=1024), device="cuda", shard_optimizer=True)
    assert tr2.peer is not None and tr2.opt.grad.data_ptr() == tr2.peer.grad.data_ptr()
    tr2.model.density_grid.copy_(tr.model.density_grid); tr2.model.density_bitfield.copy_(tr.model.density_bitfield)
    tr2.opt.flat.copy_(tr.opt.flat); tr2.opt.flat16.copy_(tr.opt.flat16); tr2.global_step = tr.global_step
    noise = torch.rand(1024, device="cuda")
    fs = tr.fused_step(use_graph=True); fs.set_triangles(tri)
    fs2 = tr2.fused_step(use_graph=True); fs2.set_triangles(tri)
    assert fs2.defer and not fs2.nccl
    p_start = tr2.opt.flat.clone()
    for i in range(4):
        fs.step(rays_o, rays_d, rgb, noise=noise)
        fs2.step(rays_o, rays_d, rgb, noise=noise)
    fs.flush(); fs2.flush()
    torch.cuda.synchronize()
    assert tr2.peer.error() == 0
    # Two runs of the SAME optimizer already differ at the 1e-3 level after a few steps (the backward accumulates with fp32
    # atomics; Adam with eps = 1e-15 turns the sign of a cancelling gradient into a +-lr step - tools/check_peer.py measures
    # that yardstick), so the check is: same trajectory within that noise, and the fp16 copy IS the rounded fp32 master.
    rel = float((tr.opt.flat - tr2.opt.flat).norm() / tr.opt.flat.norm())
    assert rel < 5e-3, rel
    moved = float((tr2.opt.flat - p_start).norm() / p_start.norm())
    assert moved > 5 * rel, (moved, rel)
    assert torch.equal(tr2.opt.flat16, tr2.opt.flat.to(torch.float16))


@pytest.mark.parametrize("cut_frac", [0.0, 0.37, 1.0])
def test_peer_step_with_early_range_single_rank(ncn, cut_frac):
    """ncn_peer_set_cut + ncn_peer_early + ncn_peer_step at world_size 1 against ncn_grad_sumsq + ncn_adam_step_groups: the
    early kernel reduces [cut, n) (a self-copy at one rank) and parks its share of the norm, the step's reduce kernel handles
    [0, cut) and folds the parked partials in - together the same norm, the same update, a zeroed gradient.  cut = n is the
    plain step and refuses ncn_peer_early; cut = 0 makes everything early."""
    import ctypes as C
    from ncn_b200 import _lib
    from ncn_b200._lib import check, ptr, stream
    from ncn_b200.trainer import PeerLink
    L = _lib.lib()
    n = (1 << 20) + 4 * 33
    cut = int(n * cut_frac) // 4 * 4
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device="cuda").manual_seed(3)
    link = PeerLink(0, 1, n, dev)
    link.set_cut(cut)
    assert link.segments(0) == [(0, cut), (cut, n)] and link.shard == (0, cut)
    p0 = torch.randn(n, device="cuda", generator=g) * 0.1
    groups = _lib.AdamGroups(); groups.n_groups = 2; groups.start[0] = 0; groups.start[1] = n - 4096
    groups.weight_decay[0] = 0.0; groups.weight_decay[1] = 1e-6; groups.max_norm = 0.05
    div = torch.tensor([2.0], device="cuda")
    state = {k: [p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")] for k in ("peer", "ref")}
    p16_ref = torch.zeros(n, dtype=torch.float16, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    for step in range(1, 4):
        grad = torch.randn(n, device="cuda", generator=g) * (10.0 ** -step)
        lr_bc = torch.tensor([1e-2, 1 - 0.9 ** step, 1 - 0.999 ** step], device="cuda")
        link.grad.copy_(grad)
        sumsq_p = torch.zeros(1, device="cuda")
        p, m, v = state["peer"]
        if cut < n:
            link.early(div, stream())
        else:
            assert L.ncn_peer_early(link.handle, ptr(div), stream()) != 0
        link.step(p, m, v, groups, (0.9, 0.999), 1e-15, div, flag, lr_bc, sumsq_p, stream())
        gr = grad.clone(); sumsq = torch.zeros(1, device="cuda")
        p, m, v = state["ref"]
        check(L.ncn_grad_sumsq(ptr(gr), n, ptr(div), ptr(sumsq), ptr(flag), stream()))
        check(L.ncn_adam_step_groups(ptr(p), ptr(gr), ptr(m), ptr(v), ptr(p16_ref), n, C.byref(groups), 0.9, 0.999, 1e-15, ptr(div), ptr(flag),
                                     ptr(sumsq), ptr(lr_bc), stream()))
        torch.testing.assert_close(sumsq_p, sumsq, rtol=1e-5, atol=0)
        # the norm is summed in a different order (parked early partials + final partials): the clip coefficient may differ in its
        # last ulps, i.e. the lr = 1e-2 update by a few 1e-9
        for a, b in zip(state["peer"], state["ref"]):
            torch.testing.assert_close(a, b, rtol=2e-5, atol=1e-8)
        torch.testing.assert_close(link.p16.float(), p16_ref.float(), rtol=2e-3, atol=1e-7)
        assert float(link.grad.abs().max()) == 0.0
    # a non-finite value in the EARLY range must poison the norm through the parked partials and skip the step
    if 0 < cut < n:
        before = state["peer"][0].clone()
        bad = torch.zeros(n, device="cuda"); bad[n - 5] = float("inf")
        link.grad.copy_(bad)
        link.early(div, stream())
        link.step(*state["peer"], groups, (0.9, 0.999), 1e-15, div, flag, lr_bc, sumsq_p, stream())
        torch.cuda.synchronize()
        assert torch.equal(state["peer"][0], before) and float(link.grad.abs().max()) == 0.0 and not torch.isfinite(sumsq_p).all()
    assert link.error() == 0
    link.close()


def test_fused_step_early_range_single_rank(monkeypatch):
    """FusedStep with the exchange's early range switched on (the one-rank form, NCN_PEER_EARLY_FORCE): the table backward runs as
    two launches with ncn_peer_early between them on the side stream; trains like the plain sharded step"""
    from test_fused_gpu import _setup
    from ncn_b200.trainer import NeRFTrainer
    tr, rays_o, rays_d, tri, rgb, target = _setup(R=1024, seed=1)
    monkeypatch.setenv("NCN_PEER_EARLY", "1"); monkeypatch.setenv("NCN_PEER_EARLY_FORCE", "1")
    torch.manual_seed(1)
    tr2 = NeRFTrainer(dict(batch_size=1024), device="cuda", shard_optimizer=True)
    tr2.model.density_grid.copy_(tr.model.density_grid); tr2.model.density_bitfield.copy_(tr.model.density_bitfield)
    tr2.opt.flat.copy_(tr.opt.flat); tr2.opt.flat16.copy_(tr.opt.flat16); tr2.global_step = tr.global_step
    noise = torch.rand(1024, device="cuda")
    fs = tr.fused_step(use_graph=True); fs.set_triangles(tri)
    fs2 = tr2.fused_step(use_graph=True); fs2.set_triangles(tri)
    p_start = tr2.opt.flat.clone()
    for i in range(4):
        fs.step(rays_o, rays_d, rgb, noise=noise)
        fs2.step(rays_o, rays_d, rgb, noise=noise)
    fs.flush(); fs2.flush()
    torch.cuda.synchronize()
    assert fs2.peer_early == 8 and tr2.peer.cut == int(tr2.model.xyz_encoder.desc.level_offset[8]) * 2
    assert tr2.peer.error() == 0
    rel = float((tr.opt.flat - tr2.opt.flat).norm() / tr.opt.flat.norm())
    assert rel < 5e-3, rel
    moved = float((tr2.opt.flat - p_start).norm() / p_start.norm())
    assert moved > 5 * rel, (moved, rel)
    assert torch.equal(tr2.opt.flat16, tr2.opt.flat.to(torch.float16))
    assert float(tr2.opt.grad.abs().max()) == 0.0
