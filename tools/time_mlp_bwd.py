"""Developer tool: time ncn_mlp_bwd alone (back-to-back launches between two CUDA events) for each implementation.
Not part of the product."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ncn_b200
from ncn_b200 import _lib, tinycudann as tcnn
from ncn_b200.tinycudann import ptr, stream, check, _half_copy, _Workspace

n = int(sys.argv[1]) if len(sys.argv) > 1 else 269000
impls = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
L = _lib.lib()
big = torch.empty(64 << 20, dtype=torch.float32, device="cuda")      # 256 MB: L2 flush between launches
for (n_in, n_out, nh, act) in ((32, 3, 2, "Sigmoid"), (32, 16, 1, "None")):
    net = tcnn.Network(n_in, n_out, dict(otype="FullyFusedMLP", activation="ReLU", output_activation=act, n_neurons=64, n_hidden_layers=nh)).cuda()
    w = _half_copy(net)
    x = torch.randn(n, net.in_pad, device="cuda").half()
    out = torch.empty(n, net.out_pad, dtype=torch.float16, device="cuda")
    acts = torch.empty(L.ncn_mlp_acts_bytes(C.byref(net.desc), n) // 2, dtype=torch.float16, device="cuda")
    check(L.ncn_mlp_fwd(C.byref(net.desc), ptr(x), ptr(w), n, ptr(out), ptr(acts), None, stream()), "fwd")
    d = torch.randn(n, net.out_pad, device="cuda").half()
    dx = torch.empty(n, net.in_pad, dtype=torch.float16, device="cuda")
    grad = torch.zeros(net.params.numel(), dtype=torch.float32, device="cuda")
    ws = _Workspace.get(x.device, L.ncn_mlp_bwd_workspace_bytes(C.byref(net.desc), n))
    for impl in impls:
        old = L.ncn_set_mlp_bwd_impl(impl)
        ts = []
        for rep in range(12):
            big.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            check(L.ncn_mlp_bwd(C.byref(net.desc), ptr(x), ptr(w), ptr(out), ptr(acts), ptr(d), n, ptr(grad), ptr(dx), 1.0 / 128, ptr(ws), ws.numel(), None, stream()), "bwd")
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        L.ncn_set_mlp_bwd_impl(old)
        ts = sorted(ts[2:])
        print(f"net {n_in}->{n_out} nh={nh} n={n} impl={impl}: median {ts[len(ts)//2]:.1f} us  min {ts[0]:.1f} us (cold L2, memset + kernel)")
