"""Drop-in for the two tiny-cuda-nn torch modules the reference uses (models/ngp_mt.py:70-155):

    tcnn.Encoding(n_input_dims, encoding_config)                 Grid/Hash/Linear  (+ SphericalHarmonics, built but never called)
    tcnn.Network(n_input_dims, n_output_dims, network_config)    FullyFusedMLP, width 64, ReLU, output None/Sigmoid

Same module surface as tcnn's torch binding: ONE flat fp32 ``params``
nn.Parameter per module (state-dict keys ``<name>.params`` stay compatible with the reference's
optimizer split on 'xyz_encoder', train_nerf.py:264-274), ``n_input_dims`` / ``n_output_dims``,
half outputs, fp16 compute with fp32 accumulation, loss_scale = 128 on the backward pass,
gradients w.r.t. params and inputs (and the Encoding's double backward w.r.t. inputs).
All arithmetic runs in hand-written sm_100a kernels behind include/ncn.h; there is no torch path.
tiny-cuda-nn itself is absent from the reference tree and this image (parity unpinned).
"""
import ctypes as C
import math

import torch
from torch import nn

from . import _lib
from ._lib import check, ptr, stream

LOSS_SCALE = 128.0
SEED = 1337


def _half_copy(module):
    """fp16 copy of the fp32 master params, refreshed only when the parameter changed."""
    if module._half_key == "flat":        # a FlatAdam owns the fp16 copy and refreshes it inside its Adam kernel
        v = module.params._version          # ... which does not bump the version; a user-side in-place write does
        if getattr(module, "_flat_version", None) != v:
            if getattr(module, "_flat_version", None) is not None:
                module._half.copy_(module.params.detach())
            module._flat_version = v
        return module._half
    p = module.params
    key = (p.data_ptr(), p._version, p.device)
    if module._half_key != key:
        module._half = p.detach().to(torch.float16).contiguous()
        module._half_key = key
    return module._half


class _Workspace:
    buf = {}

    @classmethod
    def get(cls, device, nbytes):
        key = (device.type, device.index)
        b = cls.buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
            cls.buf[key] = b
        return b


# ============================================================================= Encoding
class _GridBackwardFn(torch.autograd.Function):
    """(dL/dparams, dL/dx) of the grid; differentiable once more w.r.t. dy and params (tcnn bwd_bwd_input)."""

    @staticmethod
    def forward(ctx, mod, x, params, dy, need_dx):
        table = _half_copy(mod)
        n = x.shape[0]
        L = _lib.lib()
        dy_s = (dy.float() * LOSS_SCALE).to(torch.float16).contiguous()
        grad = torch.zeros_like(params, dtype=torch.float32)
        check(L.ncn_grid_bwd(C.byref(mod.desc), ptr(x), ptr(dy_s), n, ptr(grad), 1.0 / LOSS_SCALE, None, None, stream()), "grid_bwd")
        dx = None
        if need_dx:
            dx = torch.empty(n, 3, dtype=torch.float32, device=x.device)
            check(L.ncn_grid_bwd_input(C.byref(mod.desc), ptr(x), ptr(table), ptr(dy_s), n, ptr(dx), stream()),
                  "grid_bwd_input")
            dx = dx / LOSS_SCALE
        ctx.mod = mod
        ctx.save_for_backward(x, dy)
        return grad, dx

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_grad, g_dx):
        # only the path through dL/dx is second-order relevant (the param gradient is linear in dy)
        x, dy = ctx.saved_tensors
        mod = ctx.mod
        if g_dx is None:
            return None, None, None, None, None
        table = _half_copy(mod)
        n = x.shape[0]
        v = g_dx.float().contiguous()
        dy_h = dy.to(torch.float16).contiguous()
        grad = torch.zeros_like(mod.params, dtype=torch.float32)
        ddy = torch.empty(n, mod.n_output_dims, dtype=torch.float16, device=x.device)
        check(_lib.lib().ncn_grid_bwd_bwd_input(C.byref(mod.desc), ptr(x), ptr(table), ptr(v), ptr(dy_h), n, ptr(grad),
                                                ptr(ddy), stream()), "grid_bwd_bwd_input")
        return None, None, grad, ddy.to(dy.dtype), None


class _GridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, params):
        table = _half_copy(mod)
        n = x.shape[0]
        out = torch.empty(n, mod.n_output_dims, dtype=torch.float16, device=x.device)
        check(_lib.lib().ncn_grid_fwd(C.byref(mod.desc), ptr(x), ptr(table), n, ptr(out), None, None, stream()), "grid_fwd")
        ctx.mod = mod
        ctx.save_for_backward(x, params)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, params = ctx.saved_tensors
        grad, dx = _GridBackwardFn.apply(ctx.mod, x, params, dy.contiguous(), ctx.needs_input_grad[1])
        return None, dx, grad


class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config, seed=SEED, dtype=None):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.encoding_config = dict(encoding_config)
        self.seed = seed
        self.loss_scale = LOSS_SCALE
        self._half, self._half_key = None, None
        otype = encoding_config.get("otype", "Grid")
        self.otype = otype
        if otype in ("Grid", "HashGrid"):
            if n_input_dims != 3:
                raise RuntimeError("ncn Encoding: only 3-D grid inputs are supported")
            if encoding_config.get("type", "Hash") != "Hash" or encoding_config.get("interpolation", "Linear") != "Linear":
                raise RuntimeError("ncn Encoding: only type=Hash, interpolation=Linear is supported")
            d = _lib.GridDesc()
            d.n_levels = int(encoding_config.get("n_levels", 16))
            d.n_features = int(encoding_config.get("n_features_per_level", 2))
            d.log2_hashmap_size = int(encoding_config.get("log2_hashmap_size", 19))
            d.base_resolution = int(encoding_config.get("base_resolution", 16))
            d.per_level_scale = float(encoding_config.get("per_level_scale", 2.0))
            n_params = _lib.lib().ncn_grid_desc_init(C.byref(d))
            if n_params < 0:
                raise RuntimeError("ncn Encoding: unsupported grid configuration")
            self.desc = d
            self.n_output_dims = d.n_levels * d.n_features
            g = torch.Generator().manual_seed(seed)
            init = (torch.rand(n_params, generator=g) * 2 - 1) * 1e-4     # tcnn: U(-1e-4, 1e-4)
            self.params = nn.Parameter(init)
        elif otype == "SphericalHarmonics":
            # built by NGPMT (models/ngp_mt.py:94-101) but never called (line 208 is commented out)
            self.degree = int(encoding_config.get("degree", 4))
            self.n_output_dims = (self.degree ** 2 + 15) // 16 * 16
            self.params = nn.Parameter(torch.zeros(0))
            self.desc = None
        else:
            raise RuntimeError(f"ncn Encoding: otype {otype!r} is not supported")

    def level_table(self):
        d = self.desc
        return [dict(scale=d.level_scale[l], res=d.level_res[l], size=d.level_size[l], offset=d.level_offset[l])
                for l in range(d.n_levels)]

    def forward(self, x):
        if self.desc is None:
            raise NotImplementedError("SphericalHarmonics encoding is constructed by the reference but never evaluated")
        if not x.is_cuda:
            raise RuntimeError("ncn Encoding: input must be a CUDA tensor (there is no CPU path)")
        x = x.to(torch.float32).contiguous()
        return _GridFn.apply(self, x, self.params)

    def extra_repr(self):
        return f"n_input_dims={self.n_input_dims}, n_output_dims={self.n_output_dims}, seed={self.seed}, cfg={self.encoding_config}"


# ============================================================================= Network
class _MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, params):
        w = _half_copy(mod)
        n = x.shape[0]
        dev = x.device
        xp = torch.ones(n, mod.in_pad, dtype=torch.float16, device=dev) if mod.in_pad != mod.n_input_dims else None
        if xp is None:
            xp = x.to(torch.float16).contiguous()
        else:
            xp[:, :mod.n_input_dims] = x                       # Identity encoding pads with 1.0
        out = torch.empty(n, mod.out_pad, dtype=torch.float16, device=dev)
        keep = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        # opaque saved activations in the backward kernel's tiled panel layout (include/ncn.h ncn_mlp_fwd)
        acts = torch.empty(_lib.lib().ncn_mlp_acts_bytes(C.byref(mod.desc), n) // 2, dtype=torch.float16, device=dev) if keep else None
        check(_lib.lib().ncn_mlp_fwd(C.byref(mod.desc), ptr(xp), ptr(w), n, ptr(out), ptr(acts), None, stream()), "mlp_fwd")
        ctx.mod = mod
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(xp, out, acts, w)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        xp, out, acts, w = ctx.saved_tensors
        mod = ctx.mod
        n = xp.shape[0]
        dev = xp.device
        L = _lib.lib()
        d = (dout.float() * LOSS_SCALE).to(torch.float16).contiguous()
        need_dx = ctx.needs_input_grad[1]
        dx = torch.empty(n, mod.in_pad, dtype=torch.float16, device=dev) if need_dx else None
        grad = torch.zeros(mod.params.numel(), dtype=torch.float32, device=dev)
        nbytes = L.ncn_mlp_bwd_workspace_bytes(C.byref(mod.desc), n)
        ws = _Workspace.get(dev, nbytes)
        check(L.ncn_mlp_bwd(C.byref(mod.desc), ptr(xp), ptr(w), ptr(out), ptr(acts), ptr(d), n, ptr(grad), ptr(dx),
                            1.0 / LOSS_SCALE, ptr(ws), ws.numel(), None, stream()), "mlp_bwd")
        gx = None
        if need_dx:
            gx = (dx[:, :mod.n_input_dims].float() / LOSS_SCALE).to(ctx.x_dtype)
        return None, gx, grad


class Network(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=SEED):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        self.network_config = dict(network_config)
        self.seed = seed
        self.loss_scale = LOSS_SCALE
        self._half, self._half_key = None, None
        if network_config.get("otype", "FullyFusedMLP") not in ("FullyFusedMLP", "CutlassMLP"):
            raise RuntimeError("ncn Network: unsupported otype")
        if int(network_config.get("n_neurons", 64)) != 64 or network_config.get("activation", "ReLU") != "ReLU":
            raise RuntimeError("ncn Network: only 64 neurons with ReLU are supported")
        d = _lib.MlpDesc()
        d.n_in, d.n_out = int(n_input_dims), int(n_output_dims)
        d.n_hidden = int(network_config.get("n_hidden_layers", 1))
        d.width = 64
        d.activation = _lib.ACT["ReLU"]
        d.out_activation = _lib.ACT[str(network_config.get("output_activation", "None"))]
        n_params = _lib.lib().ncn_mlp_n_params(C.byref(d))
        if n_params < 0:
            raise RuntimeError("ncn Network: unsupported configuration")
        self.desc = d
        self.n_hidden = d.n_hidden
        self.in_pad = (n_input_dims + 15) // 16 * 16
        self.out_pad = (n_output_dims + 15) // 16 * 16
        # Xavier-uniform per matrix on the padded shapes (tcnn), consecutive (out,in) row-major matrices
        g = torch.Generator().manual_seed(seed)
        shapes = [(64, self.in_pad)] + [(64, 64)] * (d.n_hidden - 1) + [(self.out_pad, 64)]
        mats = []
        for (o, i) in shapes:
            bound = math.sqrt(6.0 / (i + o))
            mats.append(((torch.rand(o, i, generator=g) * 2 - 1) * bound).reshape(-1))
        self.layer_shapes = shapes
        self.params = nn.Parameter(torch.cat(mats))
        assert self.params.numel() == n_params

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("ncn Network: input must be a CUDA tensor (there is no CPU path)")
        out = _MlpFn.apply(self, x.contiguous(), self.params)
        return out[:, :self.n_output_dims]

    def extra_repr(self):
        return (f"n_input_dims={self.n_input_dims}, n_output_dims={self.n_output_dims}, seed={self.seed}, "
                f"cfg={self.network_config}")


class NetworkWithInputEncoding(nn.Module):
    """tcnn.NetworkWithInputEncoding (commented-out alternative at models/ngp_mt.py:48-69)."""

    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=SEED):
        super().__init__()
        self.encoding = Encoding(n_input_dims, encoding_config, seed=seed)
        self.network = Network(self.encoding.n_output_dims, n_output_dims, network_config, seed=seed)
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims

    def forward(self, x):
        return self.network(self.encoding(x))
