"""CPU: ncn_b200.batches (host side of SURVEY.md section 8 row f4) against golden vectors produced by the reference's OWN
datasets/base.py (oracle/gen_golden_batches.py): BaseDataset.__getitem__ index sampling for every strategy incl.
--random_tr_poses / --triang_max_expand, and generate_random_poses; plus the host logic of the random_tr_poses record."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "batches_a.npz")


def test_sample_batch_indices_match_reference_golden():
    import ncn_b200  # noqa: F401
    from ncn_b200 import batches
    g = np.load(GOLD)
    H, W, P, Q, B, PATCH = (int(g[k]) for k in ("H", "W", "P", "Q", "B", "PATCH"))
    seen = set()
    for c in range(int(g["n_cases"])):
        strategy, rtp, expand, seed = g[f"case{c}_meta"]
        rtp, expand, seed = bool(int(rtp)), int(expand), int(seed)
        rng = np.random.RandomState(seed)            # the reference draws from the legacy global stream seeded the same way
        s = batches.sample_batch_indices(str(strategy), B, P, H, W, patch_size=PATCH, max_expand=expand, random_tr_poses=rtp,
                                         n_random_poses=Q, rng=rng)
        want_img, want_pix = g[f"case{c}_img"], g[f"case{c}_pix"]
        assert np.array_equal(s["pix_idxs"], want_pix), (strategy, rtp, expand)
        # the reference returns a scalar image index for the same_image strategies
        assert np.array_equal(s["img_idxs"], np.broadcast_to(want_img, s["img_idxs"].shape) if want_img.size == 1 else want_img)
        if rtp:
            assert np.array_equal(s["rnd_img_idxs"], g[f"case{c}_rnd"])
            assert len(s["pix_idxs"]) == len(s["rnd_img_idxs"]) and 2 * len(s["pix_idxs"]) <= B
        seen.add((str(strategy), rtp, expand > 0))
    assert len(seen) == 14


def test_generate_random_poses_match_reference_golden():
    import ncn_b200  # noqa: F401
    from ncn_b200 import batches
    g = np.load(GOLD)
    poses = torch.from_numpy(g["poses"])
    np.testing.assert_allclose(batches.focus_point(g["poses"]), g["focus_pt"], rtol=1e-9, atol=1e-12)
    for name, jitter in (("plain", False), ("jitter", True)):
        rp, avg = batches.generate_random_poses(poses, torch.from_numpy(g["xyz_min"]), torch.from_numpy(g["xyz_max"]), 64,
                                                focus_jitter=jitter, rng=np.random.RandomState(7))
        assert rp.dtype == torch.float32 and tuple(rp.shape) == (64, 3, 4)
        np.testing.assert_allclose(rp.numpy(), g["rp_" + name], rtol=0, atol=2e-7)      # float64 maths, rounded once to fp32
        np.testing.assert_allclose(avg, g["rp_avg"], rtol=1e-6, atol=1e-7)
        R = rp[:, :, :3].double()
        eye = torch.eye(3, dtype=torch.float64).expand(64, 3, 3)
        assert torch.allclose(R.transpose(1, 2) @ R, eye, atol=1e-6)                     # proper camera frames
        pos = rp[:, :, 3].numpy()
        lo, hi = g["xyz_min"], g["xyz_max"]
        assert (pos >= lo + 0.1 * (hi - lo) - 1e-6).all() and (pos <= lo + 0.9 * (hi - lo) + 1e-6).all()
    # default rng = the global numpy stream, like the reference
    np.random.seed(7)
    rp2, _ = batches.generate_random_poses(poses, g["xyz_min"], g["xyz_max"], 64)
    np.testing.assert_allclose(rp2.numpy(), g["rp_plain"], rtol=0, atol=2e-7)


def test_random_tr_poses_pixel_record():
    """FusedStep.pack_pixel_batch with rnd_img_idx: [training-view half | the same pixels with pose rows P + rnd] (train_nerf.py:169-172)"""
    import ncn_b200  # noqa: F401
    from ncn_b200.fused import FusedStep
    n, P = 128, 7
    img = torch.randint(0, P, (n,)); pix = torch.randint(0, 999, (n,)); rgb = torch.rand(n, 3); rnd = torch.randint(0, 50, (n,))
    rec = FusedStep.pack_pixel_batch(img, pix, rgb, pin=False, rnd_img_idx=rnd, n_train_poses=P)
    R = 2 * n
    assert rec.dtype == torch.uint8 and rec.numel() == R * 28
    b_img = rec[:8 * R].view(torch.int64); b_pix = rec[8 * R:16 * R].view(torch.int64); b_rgb = rec[16 * R:].view(torch.float32).view(R, 3)
    assert torch.equal(b_img[:n], img) and torch.equal(b_img[n:], rnd + P)
    assert torch.equal(b_pix[:n], pix) and torch.equal(b_pix[n:], pix)
    assert torch.equal(b_rgb[:n], rgb) and float(b_rgb[n:].abs().max()) == 0.0


def test_batch_triangles_of_the_unsupervised_half():
    """losses.py:316-331: with random_tr_poses the triangle indices are built for the n_unsup = R/2 rays of the generated poses"""
    import ncn_b200  # noqa: F401
    from ncn_b200.fused import FusedStep
    t = FusedStep.batch_triangles(4096, "all_images_triang_patch", 8)
    assert t.shape == (3, 64 * 49) and int(t.max()) == 4095 and int(t.min()) == 1      # ray 0 (a patch corner) is no triangle vertex


def test_host_batcher_records():
    """HostBatcher = the DataLoader-worker half of a step (BaseDataset.__getitem__): the packed record equals pack_pixel_batch of the
    reference-pinned index draw, targets are gathered from the images, the ring rotates, random_tr_poses appends the generated half"""
    import ncn_b200  # noqa: F401
    from ncn_b200 import batches
    from ncn_b200.fused import FusedStep
    H, W, P, Q, B = 24, 32, 5, 37, 768
    g = torch.Generator().manual_seed(0)
    images = torch.rand(P, H * W, 3, generator=g)
    labels = torch.randint(0, 4, (P, H * W), generator=g)
    for strategy, rtp in (("all_images_triang_patch", False), ("all_images_triang_patch", True), ("same_image_triang", True),
                          ("all_images", False)):
        hb = batches.HostBatcher(images, H, W, strategy, B, patch_size=8, random_tr_poses=rtp, n_random_poses=Q, labels=labels, ring=2,
                                 rng=np.random.RandomState(11), pin=False)
        want = batches.sample_batch_indices(strategy, B, P, H, W, patch_size=8, random_tr_poses=rtp, n_random_poses=Q,
                                            rng=np.random.RandomState(11))
        rec, lab = hb.next()
        n = len(want["pix_idxs"])
        assert hb.n_gt == n and hb.n_rays == (2 * n if rtp else n) and rec.numel() == hb.n_rays * 28
        img = torch.from_numpy(np.broadcast_to(want["img_idxs"], (n,)).astype(np.int64).copy()); pix = torch.from_numpy(want["pix_idxs"].astype(np.int64))
        ref = FusedStep.pack_pixel_batch(img, pix, images[img, pix], pin=False,
                                         rnd_img_idx=torch.from_numpy(want["rnd_img_idxs"].astype(np.int64)) if rtp else None, n_train_poses=P)
        assert torch.equal(rec, ref)
        assert torch.equal(lab, labels[img, pix])
        rec2, _ = hb.next()
        assert rec2.data_ptr() != rec.data_ptr() and not torch.equal(rec2, ref)      # next slot of the ring, a fresh draw
        rec3, _ = hb.next()
        assert rec3.data_ptr() == rec.data_ptr()                                     # ring of 2
    for bad in (dict(strategy="nope"), dict(strategy="all_images", random_tr_poses=True, n_random_poses=3)):
        try:
            batches.HostBatcher(images, H, W, batch_size=B, pin=False, **bad)
            raise AssertionError("expected ValueError")
        except ValueError:
            pass
