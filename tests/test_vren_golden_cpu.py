"""CPU: the C / torch restatements in oracle/ against golden vectors produced on a B200 by the reference's OWN
csrc kernels (oracle/gen_golden_vren.py).  Bit-exact for AABB, marching (indices, counts and every fp32 value),
Morton codes; compositing within fp32 tolerance (__expf vs expf; summation order)."""
import glob
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "vren_ref_*.npz")))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)


@pytest.fixture(scope="module")
def bitfield():
    import ncn_b200
    from ncn_b200 import synth
    from oracle import march
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    bits = march.packbits(grid, 5.9)
    assert np.array_equal(bits, synth.packbits_np(grid, 5.9))
    return bits


@pytest.mark.parametrize("path", CASES)
def test_c_oracle_march_bit_exact(path, bitfield):
    from oracle import march
    g = np.load(path)
    hits_raw = march.aabb(g["rays_o"], g["rays_d"], [0, 0, 0], [0.5, 0.5, 0.5], -1.0)
    assert np.array_equal(_bits(hits_raw), _bits(g["hits_raw"]))
    hits = march.aabb(g["rays_o"], g["rays_d"], [0, 0, 0], [0.5, 0.5, 0.5], 0.01)
    assert np.array_equal(_bits(hits), _bits(g["hits_t"]))
    ra, xyzs, dirs, deltas, ts = march.march_train(g["rays_o"], g["rays_d"], hits, bitfield, 1, 0.5, float(g["esf"]),
                                                   g["noise"], 128, 1024)
    assert np.array_equal(ra, g["rays_a"])
    for a, b in ((ts, g["ts"]), (deltas, g["deltas"]), (xyzs, g["xyzs"]), (dirs, g["dirs"])):
        assert np.array_equal(_bits(a), _bits(b))
    assert np.array_equal(march.morton3d(g["coords"]), g["morton"])


@pytest.mark.parametrize("path", CASES)
def test_composite_oracle_vs_reference(path):
    from oracle import composite
    g = np.load(path)
    t = lambda k: torch.from_numpy(g[k])
    sig = t("sigmas").clone().requires_grad_(True)
    raws = t("raws").clone().requires_grad_(True)
    total, opacity, depth, rend, ws = composite.composite_train(sig, raws, t("deltas"), t("ts"), t("rays_a"), 1e-4)
    assert np.array_equal(total.numpy(), g["total_samples"])
    torch.testing.assert_close(ws, t("ws"), rtol=2e-5, atol=1e-7)
    torch.testing.assert_close(opacity, t("opacity"), rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(depth, t("depth"), rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(rend, t("rend"), rtol=2e-5, atol=1e-6)
    ((opacity * t("dO")).sum() + (depth * t("dD")).sum() + (rend * t("dR")).sum()).backward()
    torch.testing.assert_close(raws.grad, t("d_raws"), rtol=1e-4, atol=1e-6)
    scale = float(t("d_sigmas").abs().max())
    assert float((sig.grad - t("d_sigmas")).abs().max()) <= 1e-4 * scale
