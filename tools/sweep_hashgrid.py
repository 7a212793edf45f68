#!/usr/bin/env python
"""BASELINE config 5: hash-grid sweep, 16 levels, F=2, table size T = 2^19 .. 2^22, ScanNet-shaped 640x480 rays.
Samples = the occupancy march of `--rays` rays drawn from the 640x480 synthetic camera through the synthetic room
(ray-coherent sample order, as in training).  Per T: forward (ncn_grid_fwd) and parameter backward (ncn_grid_bwd) launch
time by CUDA events, with the L2 flushed between launches (a 256 MB write) so the table is read from where it lives after
the optimizer pass, and achieved GB/s of the ALGORITHMIC bytes (fwd 12 + 512 + 64 = 588 B/sample,
bwd 12 + 64 + 1024 B/sample with fp32 gradients) against the measured HBM peak.  `bwd_f16` = ncn_grid_bwd_f16, the selectable
fp16-gradient mode (12 + 64 + 512 B/sample; its GB/s is quoted on ITS bytes, its speed-up over fp32 is the ratio of the times).

    python tools/sweep_hashgrid.py [--rays 65536] [--out profiles/r1_hashgrid_sweep.json]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_hashgrid.py   (replicas: aggregate samples/s)
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch.distributed as dist
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib, synth, vren
    from ncn_b200 import tinycudann as tcnn
    from ncn_b200._lib import check, ptr, stream
    from ncn_b200.rendering import ray_aabb_near
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    scale, G = 0.5, 128
    occ = synth.room_occupancy(G, scale, seed=0)
    grid = torch.from_numpy(synth.density_grid_from_occupancy(occ)).to(dev)
    bits = torch.zeros(G ** 3 // 8, dtype=torch.uint8, device=dev)
    vren.packbits(grid, 5.9, bits)
    b = synth.random_batch(a.rays, cam="scannet", seed=rank)
    ro = torch.from_numpy(b["rays_o"]).to(dev); rd = torch.from_numpy(b["rays_d"]).to(dev)
    center = torch.zeros(1, 3, device=dev); half = torch.full((1, 3), scale, device=dev)
    hits_t = ray_aabb_near(ro, rd, center, half, 0.01)
    noise = torch.rand(a.rays, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    rays_a, xyzs, dirs, deltas, ts, _ = vren.raymarching_train(ro, rd, hits_t[:, 0], bits, 1, scale, 0.0, noise, G, 1024)
    n = xyzs.shape[0]
    x01 = ((xyzs + scale) / (2 * scale)).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    rows = []
    for log2_T in (19, 20, 21, 22):
        bsc = float(np.exp(np.log(2048 * scale / 16) / 15))
        enc = tcnn.Encoding(3, dict(otype="Grid", type="Hash", n_levels=16, n_features_per_level=2, log2_hashmap_size=log2_T,
                                    base_resolution=16, per_level_scale=bsc, interpolation="Linear")).to(dev)
        table = (torch.randn(enc.params.numel(), device=dev) * 0.1).to(torch.float16)
        feat = torch.empty(n, 32, dtype=torch.float16, device=dev)
        dfeat = (torch.randn(n, 32, device=dev) * 1e-2).to(torch.float16)
        grad = torch.zeros(enc.params.numel(), dtype=torch.float32, device=dev)
        grad16 = torch.zeros(enc.params.numel(), dtype=torch.float16, device=dev)
        st = stream()

        def fwd():
            check(L.ncn_grid_fwd(C.byref(enc.desc), ptr(x01), ptr(table), n, ptr(feat), None, None, st), "grid_fwd")

        def bwd():
            check(L.ncn_grid_bwd(C.byref(enc.desc), ptr(x01), ptr(dfeat), n, ptr(grad), 1.0, None, None, st), "grid_bwd")

        def bwd_f16():
            check(L.ncn_grid_bwd_f16(C.byref(enc.desc), ptr(x01), ptr(dfeat), n, ptr(grad16), 1.0, None, None, st), "grid_bwd_f16")

        def timed(fn, cold):
            tot = 0.0
            for _ in range(3):
                fn()
            for _ in range(a.reps):
                if cold:
                    flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            return tot / a.reps

        row = {"log2_T": log2_T, "params": enc.params.numel(), "table_fp16_MiB": enc.params.numel() * 2 / 2 ** 20, "samples": n}
        for name, fn, per in (("fwd", fwd, 588), ("bwd", bwd, 1100), ("bwd_f16", bwd_f16, 588)):
            for cold in (True, False):
                ms = timed(fn, cold)
                t = torch.tensor([ms], device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
                k = f"{name}_{'l2_flushed' if cold else 'warm'}"
                row[k + "_us"] = 1e3 * ms
                row[k + "_GBps_algorithmic"] = per * n / (ms * 1e-3) / 1e9
                row[k + "_frac_hbm_peak"] = row[k + "_GBps_algorithmic"] / peak
                row[k + "_Msamples_per_s_all_gpus"] = world * n / (ms * 1e-3) / 1e6
        rows.append(row)
        del enc, table, grad, grad16
    res = {"config": "hash-grid sweep L=16 F=2, T=2^19..2^22, ScanNet-shaped 640x480 camera, ray-coherent samples of the synthetic room",
           "n_gpus": world, "rays": a.rays, "hbm_peak_GBps": peak,
           "algorithmic_bytes_per_sample": {"fwd": 588, "bwd": 1100, "bwd_f16": 588}, "rows": rows}
    if rank == 0:
        print(json.dumps(res))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(res, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
