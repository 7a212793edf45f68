#!/usr/bin/env python
"""BASELINE config 4: full-image evaluation render 1024x768 (786 432 rays), image tiles sharded across the ranks
(parallel.shard_tiles), gathered on rank 0.  Times render_fast (one march / field pass / composite per tile) and - on one
rank, once - the reference-shaped round loop render(test_time=True) on the same tile, and checks they agree.

    python tools/eval_render.py [--heads sem,norm] [--reps 5] [--loop]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/eval_render.py
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--heads", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--loop", action="store_true", help="also time the reference-shaped test-time loop on this rank's tile")
    ap.add_argument("--sigma-scale", type=float, default=0.5, help="std of the random table (larger = more opaque field)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch.distributed as dist
    import ncn_b200  # noqa: F401
    from ncn_b200 import synth, vren
    from ncn_b200.parallel import shard_tiles
    from ncn_b200.rendering import render, render_fast
    from ncn_b200.trainer import NeRFTrainer
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    heads = [h for h in a.heads.split(",") if h]
    n_cls = 3 if "sem" in heads else 0
    torch.manual_seed(0)
    tr = NeRFTrainer(dict(batch_size=8192, pred_sem="sem" in heads, pred_norm_nn="norm" in heads), device=dev, rank=rank, world_size=world,
                     n_sem_cls=n_cls)
    m = tr.model
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    m.density_grid.copy_(torch.from_numpy(grid).to(dev))
    vren.packbits(m.density_grid, 5.9, m.density_bitfield)
    g = torch.Generator(device=dev).manual_seed(1)
    n = m.xyz_encoder.params.numel()
    tr.opt.flat[:n].copy_(torch.randn(n, device=dev, generator=g) * a.sigma_scale)
    tr.opt.flat16.copy_(tr.opt.flat)
    poses = torch.from_numpy(synth.camera_poses(50, 0)).to(dev); dirs = torch.from_numpy(synth.pixel_directions("hypersim")).to(dev)
    tr.set_cameras(poses, dirs)
    HW = dirs.shape[0]
    s, e = shard_tiles(HW, rank, world)
    pix = torch.arange(s, e, device=dev)
    ro, rd = tr.rays_from_batch(torch.zeros_like(pix), pix)
    kw = dict(near_distance=0.01, max_samples=1024, exp_step_factor=0.0, T_threshold=1e-4, n_sem_cls=n_cls)
    C = 8 + 3 * ("norm" in heads) + n_cls          # rgb 3, depth, opacity, + heads ... packed row for the gather

    def once():
        r = render_fast(m, ro, rd, **kw)
        cols = [r["rgb"], r["depth"][:, None], r["opacity"][:, None]] + ([r["norm_nn"]] if "norm" in heads else []) + ([r["sem"]] if n_cls else [])
        row = torch.cat(cols, 1).contiguous()
        if world > 1:
            per = [torch.empty(shard_tiles(HW, q, world)[1] - shard_tiles(HW, q, world)[0], row.shape[1], device=dev) for q in range(world)] if rank == 0 else None
            dist.gather(row, per, dst=0)
        return r

    for _ in range(2):
        r = once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(a.reps):
        r = once()
    ev[1].record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev[0].elapsed_time(ev[1]) / a.reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res = {"config": "full-image eval render 1024x768" + (" with " + "+".join(heads) + " heads" if heads else ""), "n_gpus": world,
           "rays": HW, "rays_per_gpu": e - s, "render_fast_ms_per_image": float(ms), "rays_per_s": HW / (float(ms) * 1e-3),
           "samples_composited_rank0": int(r["total_samples"]), "hit_fraction_rank0": float((r["opacity"] > 0).float().mean())}
    if a.loop and rank == 0:
        for _ in range(1):
            ref = render(m, ro, rd, test_time=True, **kw)
        torch.cuda.synchronize()
        ev[0].record()
        ref = render(m, ro, rd, test_time=True, **kw)
        ev[1].record()
        torch.cuda.synchronize()
        res["loop_ms_this_tile"] = ev[0].elapsed_time(ev[1])
        res["loop_total_samples"] = int(ref["total_samples"])
        res["max_abs_diff_rgb"] = float((ref["rgb"] - r["rgb"]).abs().max())
        res["max_abs_diff_depth"] = float((ref["depth"] - r["depth"]).abs().max())
    if rank == 0:
        print(json.dumps(res))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(res, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
