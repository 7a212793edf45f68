"""TEST INFRASTRUCTURE ONLY.  Pure-torch restatement of tiny-cuda-nn's FullyFusedMLP / Network as
used at models/ngp_mt.py:83-155 (no biases; ReLU hidden, width 64; output None / Sigmoid; the
wrapping Identity encoding pads the input width up to a multiple of 16 with the constant 1.0 and
the output width up to a multiple of 16; weight matrices are consecutive row-major (out, in)).

PARITY UNPINNED: tiny-cuda-nn is absent from /root/reference and this image (see oracle/hashgrid.py).
Numerics adopted here: fp16 weights and activations, fp32 accumulation (tcnn's wmma path accumulates
in fp16; ours is closer to exact - tolerance stated in tests/).
"""
import torch


def pad16(v):
    return (v + 15) // 16 * 16


def split_params(params, n_in, n_out, n_hidden):
    ip, op = pad16(n_in), pad16(n_out)
    shapes = [(64, ip)] + [(64, 64)] * (n_hidden - 1) + [(op, 64)]
    mats, o = [], 0
    for (r, c) in shapes:
        mats.append(params[o:o + r * c].reshape(r, c))
        o += r * c
    assert o == params.numel()
    return mats


def forward(x, params, n_in, n_out, n_hidden, out_act="None", emulate_half=True, return_hidden=False):
    """x (N, n_in) -> (N, n_out).  With emulate_half the weights / activations are rounded to fp16
    between layers exactly where the kernels round them; otherwise plain fp32/fp64 math (for autograd)."""
    ip = pad16(n_in)
    dt = torch.float32 if emulate_half else x.dtype
    h = torch.ones(x.shape[0], ip, dtype=dt, device=x.device)
    h[:, :n_in] = x.to(dt)
    if emulate_half:
        h = h.half().float()
    mats = split_params(params, n_in, n_out, n_hidden)
    hidden = []
    for i, W in enumerate(mats):
        Wf = W.half().float() if emulate_half else W.to(dt)
        h = h @ Wf.t()
        if i < len(mats) - 1:
            h = torch.relu(h)
            if emulate_half:
                h = h.half().float()
            hidden.append(h)
    if out_act == "Sigmoid":
        h = torch.sigmoid(h)
    elif out_act == "Exponential":
        h = torch.exp(h)
    out = h[:, :n_out]
    if emulate_half:
        out = out.half()
    return (out, hidden) if return_hidden else out
