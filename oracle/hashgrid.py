"""TEST INFRASTRUCTURE ONLY.  Pure-torch restatement of tiny-cuda-nn's multiresolution hash grid
(GridEncoding, type=Hash, interpolation=Linear, 3-D) as the reference uses it at
models/ngp_mt.py:70-82.

PARITY UNPINNED: tiny-cuda-nn is an un-vendored, un-pinned dependency of the reference
(README.md:17 `pip install git+https://github.com/NVlabs/tiny-cuda-nn/...`) and is absent from
/root/reference and from this image; the reference ships no tests or golden vectors for it.
This file restates tcnn's published algorithm:

    log2b   = log2f(per_level_scale)
    scale_l = exp2f(l*log2b)*base - 1          (fp32)
    res_l   = ceil(scale_l) + 1
    size_l  = min(next_multiple(res_l^3, 8), 2^log2_T)          offset_l = sum_{k<l} size_k
    pos     = fma(scale_l, x, 0.5);  cell = floor(pos);  w = pos - cell
    index   = dense x + y*res + z*res^2 while the stride fits, else (x*1)^(y*2654435761)^(z*805459861) (uint32); % size_l
    out[l*F+f] = sum_{8 corners} prod_d(w_d or 1-w_d) * table[offset_l + index][f]

Gradients (params, inputs, double-backward) come from torch autograd through this restatement.
"""
import numpy as np
import torch

PRIMES = (1, 2654435761, 805459861)


def grid_levels(n_levels=16, n_features=2, log2_hashmap_size=19, base_resolution=16, per_level_scale=2.0):
    """Per-level (scale, res, size, offset) in fp32 arithmetic; returns (levels, total_entries)."""
    # log2f / exp2f as correctly rounded fp32 (evaluated in double, rounded once) - host-libm independent
    b = np.float32(per_level_scale)
    log2b = np.float32(np.log2(np.float64(b)))
    T = 1 << log2_hashmap_size
    levels, off = [], 0
    for l in range(n_levels):
        p2 = np.float32(np.exp2(np.float64(np.float32(l) * log2b)))
        scale = np.float32(np.float32(p2 * np.float32(base_resolution)) - np.float32(1.0))
        res = int(np.ceil(scale)) + 1
        size = min(res ** 3, T)
        size = (size + 7) // 8 * 8
        size = min(size, T)
        levels.append(dict(scale=float(scale), res=res, size=size, offset=off))
        off += size
    return levels, off


def _index(g, res, size):
    """g: (..., 3) int64 grid coordinates -> entry index within the level (int64)."""
    stride, index, dense_dims = 1, torch.zeros_like(g[..., 0]), 0
    for d in range(3):
        if stride > size:
            break
        index = index + g[..., d] * stride
        stride *= res
        dense_dims += 1
    if size < stride:
        m = 0xFFFFFFFF
        index = ((g[..., 0] * PRIMES[0]) & m) ^ ((g[..., 1] * PRIMES[1]) & m) ^ ((g[..., 2] * PRIMES[2]) & m)
    return index % size


def forward(x, table, levels, out_dtype=torch.float16):
    """x (N,3) in [0,1] (fp32/fp64), table (entries, F) -> (N, L*F).  Differentiable w.r.t. x and table."""
    outs = []
    tab = table.to(x.dtype) if table.dtype != x.dtype else table
    for lv in levels:
        # fma(scale, x, 0.5) == single rounding of the exact product-sum: do it in float64
        pos64 = x.double() * float(np.float32(lv["scale"])) + 0.5
        pos = pos64.to(x.dtype) if x.dtype == torch.float32 else pos64
        cell = torch.floor(pos.detach())
        w = pos - cell
        g0 = cell.to(torch.int64)
        acc = 0
        for c in range(8):
            bits = torch.tensor([(c >> d) & 1 for d in range(3)], device=x.device)
            g = g0 + bits
            idx = _index(g, lv["res"], lv["size"]) + lv["offset"]
            wc = torch.where(bits.bool(), w, 1 - w).prod(-1, keepdim=True)
            acc = acc + wc * tab[idx]
        outs.append(acc)
    out = torch.cat(outs, -1)
    return out.to(out_dtype) if out_dtype is not None else out
