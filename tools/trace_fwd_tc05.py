"""Developer tool: in-kernel phase clocks of the tcgen05 field forward (build with NCN_NVCC_EXTRA=-DNCN_TC05_TRACE) + timing of the
two implementations of ncn_field_mlp_fwd at the bench's sample count.  Not part of the product."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ncn_b200
from ncn_b200 import _lib, tinycudann as tcnn
from ncn_b200._lib import check, ptr, stream
L = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 270000
g = torch.Generator(device="cuda").manual_seed(0)
sig = tcnn.Network(32, 16, dict(otype="FullyFusedMLP", activation="ReLU", output_activation="None", n_neurons=64, n_hidden_layers=1)).cuda()
rgb = tcnn.Network(19, 3, dict(otype="FullyFusedMLP", activation="ReLU", output_activation="Sigmoid", n_neurons=64, n_hidden_layers=2)).cuda()
w_sig = sig.params.detach().half().contiguous(); w_rgb = rgb.params.detach().half().contiguous()
feat = (torch.randn(n, 32, device="cuda", generator=g) * 0.5).half(); dirs = torch.randn(n, 3, device="cuda", generator=g)
cap_t = (n + 127) // 128 * 128
o = dict(sigmas=torch.empty(n, device="cuda"), raws=torch.empty(n, 3, device="cuda"), h=torch.empty(n, 16, dtype=torch.float16, device="cuda"),
         sig_acts=torch.empty(1, cap_t, 64, dtype=torch.float16, device="cuda"), x_rgb=torch.empty(n, 32, dtype=torch.float16, device="cuda"),
         rgb_acts=torch.empty(2, cap_t, 64, dtype=torch.float16, device="cuda"), rgb_out=torch.empty(n, 16, dtype=torch.float16, device="cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run():
    check(L.ncn_field_mlp_fwd(ptr(feat), ptr(dirs), ptr(w_sig), ptr(w_rgb), n, None, ptr(o["sigmas"]), ptr(o["raws"]), 3, ptr(o["h"]), ptr(o["sig_acts"]),
                              ptr(o["x_rgb"]), ptr(o["rgb_acts"]), ptr(o["rgb_out"]), stream()))
for impl in (0, 1):
    L.ncn_set_field_fwd_impl(impl)
    ts = []
    for rep in range(12):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    print(f"ncn_field_mlp_fwd impl {impl} n={n}: median {ts[len(ts)//2]:.1f} us min {ts[0]:.1f} us (cold L2, incl. ctypes launch)")
raw = ctypes.CDLL(_lib.LIB_PATH)
if hasattr(raw, "ncn_debug_fwd_trace"):
    buf = np.zeros(256, dtype=np.int64)
    raw.ncn_debug_fwd_trace.argtypes = [ctypes.c_void_p]
    assert raw.ncn_debug_fwd_trace(buf.ctypes.data) == 0
    names = ["tile start", "staged+sync", "L0 issued", "prefetch issued", "L0 done", "epi0 written", "sync", "L1 issued(+bulk)", "L1 done", "epi1 written+sync",
             "R0 done", "epi2+sync", "R1 done", "epi3+sync", "R2 done", "tile end"]
    for it in range(1, 5):
        r = buf[it * 16:(it + 1) * 16]
        print(f" tile {it}: total {r[15] - r[0]} clk | " + " ".join(f"{names[q]} @{r[q] - r[0]}" for q in range(1, 16)))
