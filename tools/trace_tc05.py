"""Developer tool: per-tile phase clocks of the tcgen05 MLP backward (needs a libncn.so built with
NCN_NVCC_EXTRA=-DNCN_TC05_TRACE).  Not part of the product."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ncn_b200
from ncn_b200 import _lib, tinycudann as tcnn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 269000
nh = int(sys.argv[2]) if len(sys.argv) > 2 else 2
net = tcnn.Network(32, 3 if nh == 2 else 16, dict(otype="FullyFusedMLP", activation="ReLU", output_activation="Sigmoid" if nh == 2 else "None",
                                  n_neurons=64, n_hidden_layers=nh)).cuda()
x = torch.randn(n, 32, device="cuda").half().float().requires_grad_(True)
dy = torch.randn(n, 3 if nh == 2 else 16, device="cuda")
for _ in range(3):
    net.params.grad = None; x.grad = None
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    out = net(x)
    torch.cuda.synchronize()
    e0.record()
    (out.float() * dy).sum().backward()
    e1.record(); torch.cuda.synchronize()
print("backward total ms (incl. autograd glue)", e0.elapsed_time(e1))
raw = ctypes.CDLL(_lib.LIB_PATH)
buf = np.zeros(16 * 16 * 16, dtype=np.int64)
fn = raw.ncn_debug_tc05_trace
fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
assert fn(buf.ctypes.data) == 0
t = buf.reshape(16, 16, 16)
for b in (0, 1, 5):
    base = t[b, 15, 12]
    print(f"grid {t[b,15,15]} CTA {b}: kernel start 0, wgrad-epilogue start {t[b,15,13]-base}, end {t[b,15,14]-base} (clocks)")
    for it in range(10):
        r = t[b, it]
        if r[0] == 0:
            break
        print("  tile", it, " ".join(f"{int(v - base) if v else -1:7d}" for v in r[:15]))
