#!/usr/bin/env python
"""Multi-GPU check of the sharded peer-memory optimizer (ncn_peer_step) against ncclAllReduce + replicated Adam:

    timeout 300 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_peer.py

Two trainers per rank with identical initial state, one per exchange; each rank feeds BOTH the same rank-specific batches for
a few fused steps.  Checks: (1) the peer path's fp16 parameters are bit-identical on every rank and no wait timed out;
(2) they match the NCCL path's to fp16 rounding (the W gradient terms are summed in a different order); (3) the fp32 master
inside each rank's own slice matches the replicated master.  Prints per-step time of both exchanges (graph replay)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ncn_b200  # noqa: F401
    from ncn_b200 import synth, vren
    from ncn_b200.trainer import NeRFTrainer
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    R = 4096
    trs = {}
    for name, shard in (("peer", True), ("nccl", False)):
        torch.manual_seed(0)
        tr = NeRFTrainer(dict(batch_size=R), device=dev, rank=rank, world_size=world, shard_optimizer=shard)
        grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
        tr.model.density_grid.copy_(torch.from_numpy(grid).to(dev))
        vren.packbits(tr.model.density_grid, 5.9, tr.model.density_bitfield)
        g = torch.Generator(device=dev).manual_seed(1)
        n = tr.model.xyz_encoder.params.numel()
        tr.opt.flat[:n].copy_(torch.randn(n, device=dev, generator=g) * 0.3)
        tr.opt.flat16.copy_(tr.opt.flat)
        tr.global_step = 3000
        trs[name] = tr
    assert trs["peer"].peer is not None, "peer access unavailable"
    b = synth.patch_batch(R, seed=100 + rank)
    ro = torch.from_numpy(b["rays_o"]).to(dev); rd = torch.from_numpy(b["rays_d"]).to(dev); tri = torch.from_numpy(b["tri"]).to(dev)
    rgb = torch.rand(R, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
    noise = torch.rand(R, device=dev, generator=torch.Generator(device=dev).manual_seed(50 + rank))
    out = {}
    snap = {}
    for name, tr in trs.items():
        fs = tr.fused_step(use_graph=True); fs.set_triangles(tri)
        for _ in range(6):
            fs.step(ro, rd, rgb, noise=noise)
        fs.flush()
        torch.cuda.synchronize(); dist.barrier()
        snap[name] = (tr.opt.flat16.clone(), tr.opt.flat.clone())
    for name, tr in trs.items():           # timing (state keeps evolving; the comparison uses the 6-step snapshots)
        fs = tr.fused
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(100):
            fs.step(ro, rd, rgb, noise=noise)
        e1.record()
        fs.flush()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 100], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name + "_ms_per_step"] = float(t)
        dist.barrier()
    pe, nc = trs["peer"], trs["nccl"]
    p16, p32 = snap["peer"]; n16, n32 = snap["nccl"]
    ref = p16.clone(); dist.broadcast(ref, 0)
    same = torch.tensor([float(torch.equal(ref, p16))], device=dev); dist.all_reduce(same, op=dist.ReduceOp.MIN)
    ref2 = pe.opt.flat16.clone(); dist.broadcast(ref2, 0)
    same2 = torch.tensor([float(torch.equal(ref2, pe.opt.flat16))], device=dev); dist.all_reduce(same2, op=dist.ReduceOp.MIN)
    out["peer_fp16_bit_identical_across_ranks"] = bool(same.item()) and bool(same2.item())
    out["peer_error_word"] = pe.peer.error()
    out["fp16_rel_diff_peer_vs_nccl"] = float((p16.float() - n16.float()).norm() / n16.float().norm())
    out["fp16_max_abs_diff_peer_vs_nccl"] = float((p16.float() - n16.float()).abs().max())
    lo, hi = pe.peer.shard
    out["fp32_master_rel_diff_own_slice"] = float((p32[lo:hi] - n32[lo:hi]).norm() / n32[lo:hi].norm())
    pe.gather_master_params()
    ref3 = pe.opt.flat.clone(); dist.broadcast(ref3, 0)
    out["fp32_master_identical_after_gather"] = bool(torch.equal(ref3, pe.opt.flat))
    out["fp32_vs_fp16_after_gather_max_abs"] = float((pe.opt.flat - pe.opt.flat16.float()).abs().max())
    out["world"] = world
    ok = out["peer_fp16_bit_identical_across_ranks"] and out["peer_error_word"] == 0 and out["fp16_rel_diff_peer_vs_nccl"] < 2e-3 \
        and out["fp32_master_rel_diff_own_slice"] < 1e-3 and out["fp32_master_identical_after_gather"]
    out["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(out))
    for tr in trs.values():
        tr.comm.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
