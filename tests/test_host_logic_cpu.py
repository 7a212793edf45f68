"""CPU: host-side logic of the fused step that needs no device - triangle topologies and the packed pixel record."""
import numpy as np
import torch


def _ref_patch_triangles(n_rays, p):
    """datasets/base.py:51-58 (local offsets x1 = patch[1:,1:], x2 = patch[:-1,1:], x3 = patch[1:,:-1]) combined as
    losses.py:301-313 (get_patch_triang_idx: pix_idx.view(n, patch_area)[:, offsets].reshape(-1))"""
    import einops
    area = p * p
    loc = einops.rearrange(np.arange(area, dtype=np.int64), "(h w) -> h w", h=p)
    offs = {"x1": loc[1:, 1:].reshape(-1), "x2": loc[:-1, 1:].reshape(-1), "x3": loc[1:, :-1].reshape(-1)}
    pix = einops.rearrange(torch.arange(n_rays), "(n s) -> n s", s=area)
    return torch.stack([einops.rearrange(pix[:, torch.from_numpy(offs[k])], "n s -> (n s)") for k in ("x1", "x2", "x3")])


def test_batch_triangles_follow_the_reference_topology():
    import ncn_b200  # noqa: F401
    from ncn_b200.fused import FusedStep
    for n, p in ((8192, 8), (64, 8), (4 * 9, 3), (2 * 16, 4)):
        got = FusedStep.batch_triangles(n, "all_images_triang_patch", p)
        want = _ref_patch_triangles(n, p)
        assert got.shape == (3, (n // (p * p)) * (p - 1) ** 2) and torch.equal(got, want)
        assert torch.equal(FusedStep.batch_triangles(n, "same_image_triang_patch", p), want)
    # losses.py:294-299: consecutive triplets
    t = FusedStep.batch_triangles(8192, "all_images_triang")
    assert t.shape == (3, 2730) and torch.equal(t[:, 5], torch.tensor([15, 16, 17])) and int(t.max()) == 8189


def test_pixel_record_layout():
    """[img_idx i64 (R) | pix_idx i64 (R) | rgb f32 (R,3)] = 28 bytes per ray, little-endian views of the same bytes"""
    import ncn_b200  # noqa: F401
    from ncn_b200.fused import FusedStep
    R = 192
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 50, (R,), generator=g); pix = torch.randint(0, 786432, (R,), generator=g); rgb = torch.rand(R, 3, generator=g)
    rec = FusedStep.pack_pixel_batch(img, pix, rgb, pin=False)
    assert rec.dtype == torch.uint8 and rec.numel() == 28 * R
    assert torch.equal(rec[:8 * R].view(torch.int64), img) and torch.equal(rec[8 * R:16 * R].view(torch.int64), pix)
    assert torch.equal(rec[16 * R:].view(torch.float32).view(R, 3), rgb)
    rec32 = FusedStep.pack_pixel_batch(img.int(), pix.int(), rgb.double(), pin=False)       # other dtypes are converted, not reinterpreted
    assert torch.equal(rec32, rec)


def test_get_rays_golden_is_self_consistent():
    """tests/golden/get_rays_a.npz (the reference's own get_ray_directions + get_rays): rays_d = R d, rays_o = t, pixel-centre
    directions - the contract ncn_rays_from_pixels and ncn_b200.synth.pixel_directions implement (GPU side: test_fused_gpu.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "get_rays_a.npz"))
    c2w = g["poses"][g["img_idx"]]
    np.testing.assert_allclose(np.einsum("nij,nj->ni", c2w[:, :, :3], g["directions"][g["pix_idx"]]), g["rays_d"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(c2w[:, :, 3], g["rays_o"])
    K, H, W = g["K"], int(g["H"]), int(g["W"])
    v, u = 7, 11
    want = np.array([(u - K[0, 2] + 0.5) / K[0, 0], (v - K[1, 2] + 0.5) / K[1, 1], 1.0])
    np.testing.assert_allclose(g["directions"].reshape(H, W, 3)[v, u], want / np.linalg.norm(want), rtol=1e-6)
