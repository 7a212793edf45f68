"""Developer tool: time the spherical k-means kernel alone (CUDA events, back to back). Not part of the product."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ncn_b200
from ncn_b200 import synth, clustering

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6272
x, q = synth.manhattan_normals(n, seed=0)
xt = torch.from_numpy(x).cuda()
for niter in (1, 20):
    ts = []
    for rep in range(12):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        cent, assign, nv = clustering.kmeans_spherical(xt, 20, niter)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    print(f"kmeans n={n} K=20 niter={niter}: median {ts[len(ts)//2]:.1f} us min {ts[0]:.1f} us (incl. host wrapper launches)")

import ctypes
import numpy as np
from ncn_b200 import _lib
raw = ctypes.CDLL(_lib.LIB_PATH)
if hasattr(raw, "ncn_debug_km_trace"):
    buf = np.zeros(64, dtype=np.int64)
    raw.ncn_debug_km_trace.argtypes = [ctypes.c_void_p]
    assert raw.ncn_debug_km_trace(buf.ctypes.data) == 0
    b = buf[0]
    print("compaction done", buf[1] - b, "cluster sync", buf[2] - b, "iterations start", buf[3] - b, "iterations end", buf[4] - b,
          "final sync", buf[5] - b, "end", buf[6] - b)
    for it in range(4):
        r = buf[8 + 8 * it: 16 + 8 * it] - b
        print("  it", it, "start", r[0], "tiles done", r[1], "block sums", r[2], "cluster sync", r[3], "fold", r[4], "sync", r[5], "tail", r[6])
