"""CPU tests (no GPU): oracle self-checks, host logic, and the C-ABI surface."""
import ctypes as C
import os
import re

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(ncn):
    from ncn_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ncn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ncn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 30
    h = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(h, n)]
    assert not missing, f"declared in ncn.h but not exported: {missing}"
    unbound = [n for n in sorted(declared) if n not in _lib.SIGNATURES]
    assert not unbound, f"declared in ncn.h but not bound in _lib.SIGNATURES: {unbound}"
    assert _lib.lib().ncn_version() == 100


def test_grid_desc_matches_oracle_levels(ncn):
    from ncn_b200 import _lib
    from oracle import hashgrid
    for log2_T in (14, 19, 20, 21, 22):
        for scale in (0.5, 1.0, 8.0):
            b = float(np.exp(np.log(2048 * scale / 16) / 15))
            d = _lib.GridDesc()
            d.n_levels, d.n_features, d.log2_hashmap_size, d.base_resolution, d.per_level_scale = 16, 2, log2_T, 16, b
            n_params = _lib.lib().ncn_grid_desc_init(C.byref(d))
            levels, total = hashgrid.grid_levels(16, 2, log2_T, 16, b)
            assert n_params == total * 2
            for l, lv in enumerate(levels):
                assert d.level_res[l] == lv["res"] and d.level_size[l] == lv["size"] and d.level_offset[l] == lv["offset"]
                assert np.float32(d.level_scale[l]) == np.float32(lv["scale"])


def test_hashgrid_oracle_hand_vectors():
    """index rule by hand: dense level (res^3 <= size) and hashed level."""
    from oracle import hashgrid
    g = torch.tensor([[1, 2, 3]])
    assert int(hashgrid._index(g, 16, 4096)) == 1 + 2 * 16 + 3 * 256
    want = (1 ^ ((2 * 2654435761) & 0xFFFFFFFF) ^ ((3 * 805459861) & 0xFFFFFFFF)) % 524288
    assert int(hashgrid._index(g, 85, 524288)) == want
    # interpolation reproduces a table that is linear in the position on a dense level
    levels, total = hashgrid.grid_levels(1, 1, 19, 16, 2.0)
    lv = levels[0]
    idx = torch.arange(total)
    cx, cy, cz = idx % lv["res"], (idx // lv["res"]) % lv["res"], idx // lv["res"] ** 2
    table = (1.0 * cx + 2.0 * cy - 0.5 * cz).double().view(-1, 1)
    x = torch.rand(100, 3, dtype=torch.float64) * 0.9
    out = hashgrid.forward(x, table, levels, out_dtype=None)
    pos = x * lv["scale"] + 0.5
    torch.testing.assert_close(out[:, 0], pos[:, 0] + 2 * pos[:, 1] - 0.5 * pos[:, 2], rtol=1e-9, atol=1e-9)


def test_mlp_oracle_shapes():
    from oracle import mlp
    p = torch.randn(64 * 32 + 64 * 64 + 16 * 64)
    out = mlp.forward(torch.randn(5, 19), p, 19, 3, 2, "Sigmoid")
    assert out.shape == (5, 3) and out.dtype == torch.float16
    # the 13 pad columns are ones: they act as a first-layer bias
    mats = mlp.split_params(p, 19, 3, 2)
    x0 = torch.zeros(1, 19)
    h = torch.relu(mats[0].half().float()[:, 19:].sum(1))
    _, hidden = mlp.forward(x0, p, 19, 3, 2, "None", return_hidden=True)
    torch.testing.assert_close(hidden[0][0], h.half().float(), rtol=1e-3, atol=1e-3)
