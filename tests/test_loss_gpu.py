"""GPU parity of the loss kernels (csrc/loss.cu) against (i) golden vectors produced by the reference's own
losses.py and (ii) the torch restatement in oracle/cluster_loss.py on larger seeded inputs."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "cluster_loss_*.npz")))


@pytest.mark.parametrize("path", CASES)
def test_kernels_match_reference_golden(ncn, path):
    from ncn_b200 import clustering
    g = {k: v for k, v in np.load(path).items()}
    dev = "cuda"
    rays_d = torch.from_numpy(g["rays_d"]).to(dev)
    depth = torch.from_numpy(g["depth"]).to(dev).requires_grad_(True)
    x123 = {k: torch.from_numpy(g["tri"][i]).to(dev) for i, k in enumerate(("x1", "x2", "x3"))}
    normals = clustering.normals_from_depth(rays_d, rays_d, depth, x123)
    torch.testing.assert_close(normals.detach().cpu(), torch.from_numpy(g["normals"]), rtol=1e-5, atol=1e-6)
    # selection from the SAME k-means output the reference saw (the k-means engine itself is unpinned)
    assign_full = torch.full((normals.shape[0],), -1, dtype=torch.int32, device=dev)
    assign_full[torch.from_numpy(g["valid"]).to(dev)] = torch.from_numpy(g["kmeans_assign"]).to(dev).int()
    labels, sel = clustering.cluster_select(torch.from_numpy(g["kmeans_centroids"]).to(dev), assign_full, 1.0 - 0.01)
    assert np.array_equal(labels.cpu().numpy()[g["valid"]], g["labels"])          # integer labels: exact
    terms = clustering.cluster_loss(normals, labels)
    w = float(g["w_sched"])
    got = (w * terms).detach().cpu().numpy()
    want = np.array([g["loss_ort"], g["loss_dot"], g["loss_l1"]], dtype=np.float32)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-7)
    (w * terms.sum()).backward()
    torch.testing.assert_close(depth.grad.cpu(), torch.from_numpy(g["grad_depth"]), rtol=2e-3, atol=2e-7)


def test_photometric_matches_golden(ncn):
    import ctypes as C
    from ncn_b200 import _lib
    from ncn_b200._lib import ptr, stream, check
    g = {k: v for k, v in np.load(CASES[0]).items()}
    dev = "cuda"
    R = g["rgb"].shape[0]
    # golden rgb is the already-composited prediction: feed it as rend with opacity 1 (bg term vanishes) for the MSE,
    # and the opacity entropy from the golden opacity separately
    rend = torch.from_numpy(g["rgb"]).to(dev).contiguous()
    ones = torch.ones(R, device=dev)
    sums = torch.zeros(2, device=dev)
    d_rend = torch.empty_like(rend); d_op = torch.empty(R, device=dev)
    bg = (C.c_float * 3)(1.0, 1.0, 1.0)
    check(_lib.lib().ncn_photometric_loss(ptr(rend), ptr(ones), ptr(torch.from_numpy(g["target_rgb"]).to(dev)), R, 3, bg, 0.0, 1.0,
                                          None, ptr(sums), ptr(d_rend), ptr(d_op), stream()))
    assert abs(float(sums[0]) / (3 * R) - float(g["loss_rgb"])) <= 1e-5 * float(g["loss_rgb"])
    torch.testing.assert_close(d_rend.cpu(), torch.from_numpy(g["grad_rgb"]), rtol=1e-4, atol=1e-9)
    sums.zero_()
    op = torch.from_numpy(g["opacity"]).to(dev)
    check(_lib.lib().ncn_photometric_loss(ptr(rend), ptr(op), ptr(torch.from_numpy(g["target_rgb"]).to(dev)), R, 3, bg, 1e-3, 1.0,
                                          None, ptr(sums), None, None, stream()))
    assert abs(1e-3 * float(sums[1]) / R - float(g["loss_opacity"])) <= 1e-5 * float(g["loss_opacity"])


def test_gpu_kmeans_recovers_planted_frame(ncn):
    """config 1: 8192 normals, K=20, 20 iterations: the selected triple aligns with the planted axes within 1.5 deg,
    and the loss terms agree with the restatement evaluated on the kernel's own labels; run-to-run bit-reproducible."""
    from ncn_b200 import clustering, synth
    from oracle import cluster_loss as cl
    x, q = synth.manhattan_normals(8192, seed=0)
    xt = torch.from_numpy(x).cuda()
    labels, assign, axes = clustering.normals_clustering(xt, K=20, niter=20, t_similar=0.99)
    labels2, assign2, axes2 = clustering.normals_clustering(xt, K=20, niter=20, t_similar=0.99)
    assert torch.equal(labels, labels2) and torch.equal(axes, axes2)
    valid = cl.valid_rows(torch.from_numpy(x))
    assert torch.equal((assign >= 0).cpu(), valid)
    cos = np.abs(axes.cpu().numpy() @ q)
    assert (cos.max(1) > np.cos(np.deg2rad(1.5))).all(), cos
    assert sorted(cos.argmax(1).tolist()) == [0, 1, 2]
    xg = xt.clone().requires_grad_(True)
    terms = clustering.cluster_loss(xg, labels)
    xr = torch.from_numpy(x).requires_grad_(True)
    ort, dot, l1 = cl.cluster_terms(xr, labels.cpu().long())
    np.testing.assert_allclose(terms.detach().cpu().numpy(), np.array([float(ort), float(dot), float(l1)]), rtol=2e-4, atol=1e-6)
    wts = torch.tensor([0.7, 1.3, 0.4])
    (terms * wts.cuda()).sum().backward()
    (wts[0] * ort + wts[1] * dot + wts[2] * l1).backward()
    torch.testing.assert_close(xg.grad.cpu(), xr.grad, rtol=2e-3, atol=1e-7)
    # downstream of the SAME assignment the selection is exact
    cent, a2, _ = clustering.kmeans_spherical(xt, 20, 20)
    lab_ref, _ = cl.select_clusters(a2[a2 >= 0].cpu().long(), cent.cpu(), 0.99)
    lab_gpu, _ = clustering.cluster_select(cent, a2, 0.99)
    assert torch.equal(lab_gpu.cpu()[valid].long(), lab_ref)


def test_empty_cluster_is_nan_and_zero_grad(ncn):
    from ncn_b200 import clustering
    n = torch.nn.functional.normalize(torch.randn(100, 3, device="cuda"), dim=-1).requires_grad_(True)
    labels = torch.ones(100, dtype=torch.int32, device="cuda")      # clusters 2 and 3 empty
    terms = clustering.cluster_loss(n, labels)
    assert torch.isnan(terms).all()
    torch.nan_to_num(terms).sum().backward()
    assert (n.grad == 0).all()


def test_cluster_tail_equals_the_four_calls(ncn):
    """ncn_cluster_tail (two launches) against ncn_cluster_select -> _loss_fw -> _loss_bw -> ncn_normals_from_depth_bw."""
    import ctypes as C
    from ncn_b200 import _lib, clustering, synth
    from ncn_b200.vren import ptr, stream
    L = _lib.lib()
    dev = "cuda"
    b = synth.patch_batch(8192, seed=5)
    rays_d = torch.from_numpy(b["rays_d"]).to(dev)
    tri = torch.from_numpy(b["tri"]).to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    depth = (1.0 + 0.3 * torch.rand(8192, device=dev, generator=g)).contiguous()
    x123 = {k: tri[i] for i, k in enumerate(("x1", "x2", "x3"))}
    normals = clustering.normals_from_depth(rays_d, rays_d, depth, x123).detach().contiguous()
    M = normals.shape[0]
    cent, assign, nv = clustering.kmeans_spherical(normals, 20, 20)
    w = torch.tensor([0.3, 0.5, 0.7], device=dev)

    def run(tail):
        labels = torch.empty(M, dtype=torch.int32, device=dev); sel = torch.empty(3, dtype=torch.int32, device=dev)
        losses = torch.empty(3, device=dev); stats = torch.empty(32, device=dev)
        dn = torch.empty(M, 3, device=dev); dd = torch.zeros(8192, device=dev)
        st = stream()
        if tail:
            rc = L.ncn_cluster_tail(ptr(cent), ptr(assign), M, 20, 0.99, ptr(labels), ptr(sel), ptr(normals), ptr(losses), ptr(stats),
                                    ptr(w), ptr(dn), ptr(rays_d), ptr(rays_d), ptr(depth), ptr(tri[0]), ptr(tri[1]), ptr(tri[2]), ptr(dd), st)
            assert rc == 0
        else:
            assert L.ncn_cluster_select(ptr(cent), ptr(assign), M, 20, 0.99, ptr(labels), ptr(sel), st) == 0
            assert L.ncn_cluster_loss_fw(ptr(normals), ptr(labels), M, ptr(losses), ptr(stats), st) == 0
            assert L.ncn_cluster_loss_bw(ptr(normals), ptr(labels), M, ptr(stats), ptr(w), ptr(dn), st) == 0
            assert L.ncn_normals_from_depth_bw(ptr(rays_d), ptr(rays_d), ptr(depth), ptr(tri[0]), ptr(tri[1]), ptr(tri[2]), ptr(dn), M, ptr(dd), st) == 0
        torch.cuda.synchronize()
        return labels, sel, losses, stats[:28], dn, dd

    a, bb = run(False), run(True)
    assert torch.equal(a[0], bb[0]) and torch.equal(a[1], bb[1])
    assert (a[0] != 0).any()
    torch.testing.assert_close(a[2], bb[2], rtol=1e-6, atol=0, equal_nan=True)
    torch.testing.assert_close(a[3], bb[3], rtol=1e-6, atol=1e-7, equal_nan=True)
    torch.testing.assert_close(a[4], bb[4], rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(a[5], bb[5], rtol=1e-4, atol=1e-7)      # float atomics: order differs


@pytest.mark.parametrize("n_rays,zero_frac", [(8192, 0.02), (2048, 0.0), (512, 0.3), (64, 1.0)])
def test_cluster_chain_equals_the_four_calls(ncn, n_rays, zero_frac):
    """ncn_cluster_chain (normals -> k-means -> selection -> cluster statistics + losses in ONE thread-block-cluster launch) against
    ncn_normals_from_depth_fw -> ncn_kmeans_spherical -> ncn_cluster_select -> ncn_cluster_loss_fw: every integer output identical
    (assignments, labels, selected triple, valid count), normals / centroids bit-identical, loss terms to fp32 summation order.
    8192 rays -> 6272 normals crosses the 256*K sub-sampling branch; zero-depth rays make invalid (all-zero) normals; the last
    case has no valid normal at all."""
    import ctypes as C
    from ncn_b200 import _lib, clustering, synth
    from ncn_b200.vren import ptr, stream
    L = _lib.lib()
    dev = "cuda"
    b = synth.patch_batch(n_rays, seed=7)
    rays_d = torch.from_numpy(b["rays_d"]).to(dev)
    tri = torch.from_numpy(b["tri"]).to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    rays_o = torch.from_numpy(b["rays_o"]).to(dev)
    t_wall = torch.where(rays_d > 0, (0.4 - rays_o) / rays_d, (-0.4 - rays_o) / rays_d).min(-1)[0]      # a room: Manhattan normals
    depth = (t_wall + 0.004 * torch.randn(n_rays, device=dev, generator=g)).clamp_min(0.02)
    if zero_frac >= 1.0:
        depth.zero_()
    elif zero_frac > 0:
        n_patch = n_rays // 64
        patch = torch.arange(n_patch, device=dev) % max(2, int(round(1 / zero_frac))) == 1      # whole patches at depth 0 -> all-zero normals
        depth[patch.repeat_interleave(64)] = 0.0
    depth = depth.contiguous()
    M = tri.shape[1]
    p = _lib.KmeansParams(20, 20, 1234, 256, 1)
    ws = torch.empty(L.ncn_kmeans_workspace_bytes(M, 20), dtype=torch.uint8, device=dev)
    st = stream()
    # rays_o := rays_d (rendering.py:227): with depth 0 every vertex of a triangle is its own direction -> not degenerate, so the
    # invalid rows come from the TRUE origin here (all three vertices coincide at depth 0)
    org = torch.zeros_like(rays_d) if zero_frac > 0 else rays_d

    def alloc():
        return dict(normals=torch.empty(M, 3, device=dev), cent=torch.empty(20, 3, device=dev), assign=torch.empty(M, dtype=torch.int32, device=dev),
                    nv=torch.empty(1, dtype=torch.int32, device=dev), labels=torch.empty(M, dtype=torch.int32, device=dev),
                    sel=torch.empty(3, dtype=torch.int32, device=dev), losses=torch.empty(3, device=dev), stats=torch.zeros(32, device=dev))

    a = alloc()
    assert L.ncn_normals_from_depth_fw(ptr(org), ptr(rays_d), ptr(depth), ptr(tri[0]), ptr(tri[1]), ptr(tri[2]), M, ptr(a["normals"]), st) == 0
    assert L.ncn_kmeans_spherical(ptr(a["normals"]), M, C.byref(p), ptr(a["cent"]), ptr(a["assign"]), ptr(a["nv"]), ptr(ws), ws.numel(), st) == 0
    assert L.ncn_cluster_select(ptr(a["cent"]), ptr(a["assign"]), M, 20, 0.99, ptr(a["labels"]), ptr(a["sel"]), st) == 0
    assert L.ncn_cluster_loss_fw(ptr(a["normals"]), ptr(a["labels"]), M, ptr(a["losses"]), ptr(a["stats"]), st) == 0
    torch.cuda.synchronize()
    c = alloc()
    for rep in range(2):                       # twice: the second launch must reproduce the first bit for bit
        rc = L.ncn_cluster_chain(ptr(org), ptr(rays_d), ptr(depth), ptr(tri[0]), ptr(tri[1]), ptr(tri[2]), M, C.byref(p), 0.99, ptr(c["normals"]),
                                 ptr(c["cent"]), ptr(c["assign"]), ptr(c["nv"]), ptr(c["labels"]), ptr(c["sel"]), ptr(c["losses"]), ptr(c["stats"]),
                                 ptr(ws), ws.numel(), st)
        assert rc == 0
        torch.cuda.synchronize()
        if rep == 0:
            first = {k: v.clone() for k, v in c.items()}
        else:
            for k in c:
                assert torch.equal(torch.nan_to_num(c[k].float(), nan=-7.0), torch.nan_to_num(first[k].float(), nan=-7.0)), k
    nv = int(a["nv"])
    assert int(c["nv"]) == nv
    assert torch.equal(c["normals"], a["normals"])
    valid = a["assign"] >= 0
    assert int(valid.sum()) == nv
    if zero_frac >= 1.0:
        assert nv == 0 and torch.isnan(c["losses"]).all() and float(c["stats"][27]) == 0.0
        return
    if zero_frac > 0:
        assert 0 < nv < M
    assert torch.equal(c["cent"], a["cent"])
    assert torch.equal(c["assign"], a["assign"])
    assert torch.equal(c["labels"], a["labels"]) and torch.equal(c["sel"], a["sel"])
    assert (a["labels"] != 0).any() and (a["labels"][~valid] == 0).all()
    torch.testing.assert_close(c["losses"], a["losses"], rtol=2e-6, atol=1e-7, equal_nan=True)
    torch.testing.assert_close(c["stats"][:28], a["stats"][:28], rtol=2e-6, atol=1e-6, equal_nan=True)
    if n_rays >= 2048:
        assert float(a["stats"][27]) == 1.0 and torch.isfinite(a["losses"]).all()      # the room has three populated clusters


def test_normals_image_kernel_matches_reference_golden(ncn):
    """ncn_normals_from_depth_image vs the reference's _extract_normals_from_depth_batch (golden fixture), with (B,4,4) and
    (B,3,4) poses; then a full 768x1024 image against the oracle restatement"""
    import os
    from ncn_b200 import clustering
    from oracle import cluster_loss as cl
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "normals_image_a.npz"))
    depth, dirs, poses = (torch.from_numpy(g[k]).cuda() for k in ("depth", "dirs", "poses"))
    ref = g["normals"]
    fin = np.isfinite(ref).all(-1)
    for p in (poses, poses[:, :3].contiguous()):
        out = clustering.normals_from_depth_image(depth, dirs, p).cpu().numpy()
        assert not np.isfinite(out[~fin]).all(-1).any()
        # unit vectors from differences of nearly equal fp32 points (|P| / pixel spacing ~ 30): FMA contraction alone moves them by ~1e-6
        np.testing.assert_allclose(out[fin], ref[fin], rtol=0, atol=2e-5)
        zero = (ref[fin] == 0).all(-1)
        assert (out[fin][zero] == 0).all()
    gen = torch.Generator().manual_seed(0)
    H, W = 768, 1024
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    d = (1.5 + 0.2 * torch.sin(xx / 100) + 0.1 * torch.cos(yy / 80)).repeat(2, 1, 1); d[1] += 0.3
    d[0, 100:200, 300:400] = 0.0
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    dr = torch.stack([(xs - W / 2) / 886.81, (ys - H / 2) / 886.81, torch.ones_like(xs)], -1).reshape(-1, 3)
    q, _ = torch.linalg.qr(torch.randn(2, 3, 3, generator=gen))
    ps = torch.eye(4).repeat(2, 1, 1); ps[:, :3, :3] = q
    want = cl.normals_from_depth_image(d, dr, ps)
    got = clustering.normals_from_depth_image(d.cuda(), dr.cuda(), ps.cuda()).cpu()
    # full resolution: pixel spacing ~ 1e-3 of the depth, so fp32 rounding of the points is ~1e-4 of the edge vectors
    interior = torch.ones(2, H, W, dtype=torch.bool); interior[0, 99:202, 299:402] = False
    torch.testing.assert_close(got[interior], want[interior], rtol=0, atol=5e-4)
    assert float(got[0, 100:200, 300:400].abs().max()) == 0.0 and float(got[:, 0].abs().max()) == 0.0 and float(got[:, :, -1].abs().max()) == 0.0


def test_rotation_from_normals_recovers_planted_frame(ncn):
    """validation_epoch_end (train_nerf.py:491-517): K = 30, 30 iterations on planted Manhattan normals; the recovered rotation
    matches the planted frame within 1.5 degrees, and equals the oracle's read-out of the kernel's own centroids"""
    from ncn_b200 import clustering, synth
    from oracle import cluster_loss as cl
    x, q = synth.manhattan_normals(32768, seed=1)
    R = torch.from_numpy(q)
    xt = torch.from_numpy(x).cuda()
    rot = clustering.rotation_from_normals(xt, R, K=30, niter=30, t_similar=0.99)
    assert abs(float(torch.det(rot)) - 1.0) < 1e-9
    ang = np.rad2deg(np.arccos(np.clip((np.trace((rot.T @ R.double()).numpy()) - 1) / 2, -1, 1)))
    assert ang < 1.5, ang
    _, _, centrs = clustering.normals_clustering(xt, K=30, niter=30, t_similar=0.99)
    torch.testing.assert_close(rot, cl.rotation_from_centroids(centrs.cpu(), R), rtol=0, atol=1e-12)


def test_semantic_and_photometric_kernels_match_reference_sem_golden(ncn):
    """ncn_photometric_loss + ncn_semantic_ce_loss on a C = 3 + 3 + n_cls rendered row against the reference's own losses.py with
    the semantic head enabled (tests/golden/sem_loss_a.npz): the three loss values and the gradient on every channel - colour
    and logits as the reference computes them, the norm_nn channels exactly zero (the reference puts no loss on them);
    all-void batch: term dropped, zero gradient."""
    import ctypes as C
    import os
    from ncn_b200 import _lib
    from ncn_b200._lib import ptr, stream, check
    L = _lib.lib()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sem_loss_a.npz"))
    sem_w, op_w = float(g["sem_w"]), float(g["opacity_w"])
    bg = (C.c_float * 3)(0.0, 0.0, 0.0)                 # the golden rgb is the final colour: no background term to add
    for case in ("a", "b", "void"):
        rgb = torch.from_numpy(g[f"{case}_rgb"]).cuda(); sem = torch.from_numpy(g[f"{case}_sem"]).cuda()
        R, n_cls = sem.shape
        Ct = 6 + n_cls
        rend = torch.cat([rgb, torch.randn(R, 3, device="cuda"), sem], 1).contiguous()
        opacity = torch.from_numpy(g[f"{case}_opacity"]).cuda(); target = torch.from_numpy(g[f"{case}_target_rgb"]).cuda()
        labels = torch.from_numpy(g[f"{case}_labels"]).cuda()
        sums = torch.zeros(2, device="cuda"); ce = torch.zeros(2, device="cuda")
        d_rend = torch.full((R, Ct), 5.0, device="cuda"); d_op = torch.empty(R, device="cuda")
        check(L.ncn_photometric_loss(ptr(rend), ptr(opacity), ptr(target), R, Ct, bg, op_w, 1.0, None, ptr(sums), ptr(d_rend), ptr(d_op), stream()))
        check(L.ncn_semantic_ce_loss(ptr(rend), Ct, 6, n_cls, ptr(labels), R, sem_w, ptr(ce), ptr(d_rend), stream()))
        np.testing.assert_allclose(float(sums[0]) / (3 * R), float(g[f"{case}_loss_rgb"]), rtol=1e-5)
        np.testing.assert_allclose(op_w * float(sums[1]) / R, float(g[f"{case}_loss_opacity"]), rtol=1e-5)
        n_valid = int((labels > 0).sum())
        assert int(ce[1]) == n_valid
        loss_sem = sem_w * float(ce[0]) / n_valid if n_valid else 0.0
        np.testing.assert_allclose(loss_sem, float(g[f"{case}_loss_sem"]), rtol=1e-5, atol=1e-9)
        torch.testing.assert_close(d_rend[:, :3].cpu(), torch.from_numpy(g[f"{case}_grad_rgb"]), rtol=1e-4, atol=1e-9)
        torch.testing.assert_close(d_rend[:, 6:].cpu(), torch.from_numpy(g[f"{case}_grad_sem"]), rtol=1e-4, atol=1e-9)
        assert float(d_rend[:, 3:6].abs().max()) == 0.0


def test_kmeans_early_exit_is_exact(ncn):
    """the exact early exit (centroids bit-identical to the previous iteration's = a fixed point of the deterministic Lloyd map): on a
    cleanly separated planted frame the result of niter = 20 equals niter = 200 bit for bit, and a run that cannot converge
    (niter = 1 vs 2 on noise) still differs - the exit does not fire spuriously"""
    from ncn_b200 import clustering, synth
    x, _ = synth.manhattan_normals(6272, seed=3, noise=0.01, frac_axes=1.0, frac_zero=0.0)
    xt = torch.from_numpy(x).cuda()
    c20, a20, _ = clustering.kmeans_spherical(xt, 6, 40)
    c200, a200, _ = clustering.kmeans_spherical(xt, 6, 400)
    assert torch.equal(c20, c200) and torch.equal(a20, a200)
    g = torch.Generator(device="cuda").manual_seed(0)
    noise = torch.nn.functional.normalize(torch.randn(6272, 3, device="cuda", generator=g), dim=-1)
    c1, _, _ = clustering.kmeans_spherical(noise, 20, 1)
    c2, _, _ = clustering.kmeans_spherical(noise, 20, 2)
    assert not torch.equal(c1, c2)
