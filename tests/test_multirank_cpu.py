"""world_size-2 gloo tests (CPU): the N>1 host logic - ray/tile sharding, NCCL-id plumbing, and that
'sum all-reduce then divide by world_size' of rank-local mean losses reproduces the single-process gradient of the
combined batch (using the CPU oracles as the model)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_shard_helpers():
    import ncn_b200
    from ncn_b200.parallel import shard_rays, shard_tiles
    spans = [shard_rays(65536, r, 8) for r in range(8)]
    assert spans[0] == (0, 8192) and spans[-1] == (57344, 65536)
    assert all(b[0] == a[1] for a, b in zip(spans, spans[1:]))
    spans = [shard_rays(64 * 10, r, 3) for r in range(3)]          # ragged: 4 + 3 + 3 patches
    assert [b - a for a, b in spans] == [256, 192, 192] and spans[-1][1] == 640
    assert all((b - a) % 64 == 0 for a, b in spans)
    with pytest.raises(ValueError):
        shard_rays(100, 0, 2)
    tiles = [shard_tiles(1024 * 768, r, 8) for r in range(8)]
    assert tiles[0] == (0, 98304) and tiles[-1][1] == 786432
    assert shard_tiles(10, 2, 3) == (7, 10) and shard_tiles(0, 0, 2) == (0, 0)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    import ncn_b200  # noqa: F401
    from ncn_b200 import synth
    from ncn_b200.parallel import shard_rays, dp_mean_gradient
    from oracle import composite, march
    torch.manual_seed(0)
    torch.set_num_threads(2)
    R = 256
    b = synth.patch_batch(R, seed=0)
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    bits = march.packbits(grid, 5.9)
    hits = march.aabb(b["rays_o"], b["rays_d"], [0, 0, 0], [0.5] * 3, 0.01)
    noise = np.random.RandomState(0).rand(R).astype(np.float32)
    # a tiny differentiable "field": sigma, rgb = f(theta, xyz); theta replicated on every rank
    theta = torch.linspace(-1, 1, 16).requires_grad_(True)
    target = torch.rand(R, 3, generator=torch.Generator().manual_seed(1))

    def loss_of(lo, hi):
        ra, xyzs, dirs, deltas, ts = march.march_train(b["rays_o"][lo:hi], b["rays_d"][lo:hi], hits[lo:hi], bits, 1, 0.5, 0.0,
                                                       noise[lo:hi], 128, 1024)
        x = torch.from_numpy(xyzs)
        sig = torch.exp(x @ theta[:3] + theta[3]) * 20
        rgb = torch.sigmoid(x @ theta[4:13].view(3, 3) + theta[13:16])
        _, opacity, depth, rend, _ = composite.composite_train(sig, rgb, torch.from_numpy(deltas), torch.from_numpy(ts),
                                                               torch.from_numpy(ra), 1e-4)
        return ((rend + (1 - opacity)[:, None] - target[lo:hi]) ** 2).mean()

    lo, hi = shard_rays(R, rank, world)
    loss_of(lo, hi).backward()
    g = theta.grad.clone()
    dist.all_reduce(g, op=dist.ReduceOp.SUM)               # what ncn_comm_allreduce_sum_f32 does on the flat gradient
    g_dp = dp_mean_gradient(g, world)
    theta.grad = None
    loss_of(0, R).backward()
    # NCCL id plumbing: 128 bytes from rank 0 reach every rank unchanged
    idt = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(idt, 0)
    if rank == 0:
        torch.save({"dp": g_dp, "single": theta.grad.clone(), "id_ok": bool((idt == torch.arange(128, dtype=torch.uint8)).all())}, out)
    else:
        assert bool((idt == torch.arange(128, dtype=torch.uint8)).all())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_gradient(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["id_ok"]
    torch.testing.assert_close(r["dp"], r["single"], rtol=1e-5, atol=1e-7)


def test_peer_shards_partition_the_parameter_vector():
    """ncn_peer_shard (host arithmetic of the sharded peer-memory optimizer): the W slices tile [0, n) exactly, in rank order,
    on float4 boundaries, balanced to within one float4"""
    import ctypes as C
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib
    L = _lib.lib()
    for n in (4, 64, 11468464, 70224656):
        for world in (1, 2, 3, 4, 8):
            prev, sizes = 0, []
            for r in range(world):
                lo, hi = C.c_int64(), C.c_int64()
                L.ncn_peer_shard(n, r, world, C.byref(lo), C.byref(hi))
                assert lo.value == prev and hi.value >= lo.value and lo.value % 4 == 0 and hi.value % 4 == 0
                prev = hi.value
                sizes.append(hi.value - lo.value)
            assert prev == n and max(sizes) - min(sizes) <= 4


def test_peer_segments_partition_both_ranges():
    """ncn_peer_segments_of (host arithmetic of the exchange with an early range, ncn_peer_set_cut): for every cut the W late
    slices tile [0, cut) and the W early slices tile [cut, n), in rank order, on float4 boundaries, each balanced to within one
    float4; cut == n reproduces ncn_peer_shard with an empty early range"""
    import ctypes as C
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib
    L = _lib.lib()
    for n in (64, 11468464):
        for cut in (0, 4, n // 2 // 4 * 4, 3090032 if n > 3090032 else n - 4, n):
            for world in (1, 2, 3, 8):
                pl, pe, sl, se = 0, cut, [], []
                for r in range(world):
                    seg = (C.c_int64 * 4)()
                    L.ncn_peer_segments_of(n, cut, r, world, seg)
                    assert seg[0] == pl and seg[2] == pe and all(v % 4 == 0 for v in seg)
                    assert seg[1] >= seg[0] and seg[3] >= seg[2]
                    pl, pe = seg[1], seg[3]
                    sl.append(seg[1] - seg[0]); se.append(seg[3] - seg[2])
                    if cut == n:
                        lo, hi = C.c_int64(), C.c_int64()
                        L.ncn_peer_shard(n, r, world, C.byref(lo), C.byref(hi))
                        assert (seg[0], seg[1]) == (lo.value, hi.value) and seg[2] == seg[3] == n
                assert pl == cut and pe == n
                assert max(sl) - min(sl) <= 4 and max(se) - min(se) <= 4
