"""Developer tool: timeline of the captured fused step, on 1 GPU or under torchrun on N GPUs.

Every libncn call of the step is followed by a one-thread kernel that stores %globaltimer (ncn_debug_stamp) on the same stream;
the stamps are captured into the CUDA graph with the step, and after a few replays the per-call completion times are printed per
stream.  With N > 1 ranks (sharded peer-memory exchange) every rank additionally reports the in-kernel stamps of the two exchange
kernels (ncn_peer_debug_times): flag waits vs reduction vs Adam + publish.  Not part of the product.

    python tools/timeline.py [--json out.json]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/timeline.py --json profiles/r2_timeline_8gpu.json
"""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import ncn_b200
from ncn_b200 import _lib, synth, vren
import ncn_b200.fused as F
from ncn_b200.trainer import NeRFTrainer

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(rank)
R = 8192
tr = NeRFTrainer(dict(batch_size=R), device=dev, rank=rank, world_size=world, shard_optimizer=True if os.environ.get("NCN_TIMELINE_SHARD") == "1" else None)
grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
tr.model.density_grid.copy_(torch.from_numpy(grid).to(dev))
vren.packbits(tr.model.density_grid, 5.9, tr.model.density_bitfield)
tr.global_step = 3009
tr.hp["update_interval"] = 1 << 30
NB = 4
batches = []
for i in range(NB):
    b = synth.patch_batch(R, seed=1000 * rank + i)
    batches.append((torch.from_numpy(b["rays_o"]).to(dev), torch.from_numpy(b["rays_d"]).to(dev), torch.rand(R, 3, device=dev)))

slots = torch.zeros(256, dtype=torch.int64, device=dev)
names = []
L = _lib.lib()
orig_check = F.check
streams = {}


def check_and_stamp(rc, what=""):
    orig_check(rc, what)
    s = torch.cuda.current_stream()
    sid = streams.setdefault(s.cuda_stream, len(streams))
    if torch.cuda.is_current_stream_capturing() and len(names) < 255:
        names.append((what, sid))
        orig_check(L.ncn_debug_stamp(slots.data_ptr(), len(names), s.cuda_stream), "stamp")


F.check = check_and_stamp
if tr.peer is not None:      # the peer step is issued through trainer.PeerLink (its own `check`): stamp it too
    import ncn_b200.trainer as T
    T.check = check_and_stamp
fs = tr.fused_step(use_graph=True)
fs.set_triangles(torch.from_numpy(b["tri"]).to(dev))
acc, peer_acc, early_acc = None, None, None
n_rep = 0
for i in range(24):
    ro, rd, rgb = batches[i % NB]
    tr.train_step_fused(ro, rd, rgb, update_grid=False)
    torch.cuda.synchronize()
    if i >= 8 and names:
        t = slots[1:len(names) + 1].cpu().numpy().astype(np.float64)
        t0 = t.min()
        t = (t - t0) / 1e3
        acc = t if acc is None else acc + t
        if tr.peer is not None:
            buf = (C.c_ulonglong * 8)()
            orig_check(L.ncn_peer_debug_times(tr.peer.handle, buf), "peer_debug_times")
            p = (np.array(list(buf)[:7], dtype=np.float64) - t0) / 1e3
            peer_acc = p if peer_acc is None else peer_acc + p
            if getattr(fs, "peer_early", 0):
                b4 = (C.c_ulonglong * 4)()
                orig_check(L.ncn_peer_debug_times_early(tr.peer.handle, b4), "peer_debug_times_early")
                q = (np.array(list(b4)[:3], dtype=np.float64) - t0) / 1e3
                early_acc = q if early_acc is None else early_acc + q
        n_rep += 1
fs.flush()
t = acc / n_rep
order = np.argsort(t)
rows = []
last = {}
for i in order:
    what, sid = names[i]
    d = t[i] - last.get(sid, 0.0)
    last[sid] = t[i]
    rows.append({"stream": sid, "done_us": float(t[i]), "since_prev_on_stream_us": float(d), "call": what})
res = {"rank": rank, "world": world, "replays": n_rep, "step_span_us": float(t.max() - t.min()), "calls": rows}
if peer_acc is not None:
    p = peer_acc / n_rep
    lab = ["K1 start", "K1 wait0 over (every rank's backward done)", "K1 reduced (last block)", "K2 start", "K2 wait1 over (norms in, peers done reading)",
           "K2 Adam + publish done (last block)", "K2 wait2 over (every shard of my fp16 copy written)"]
    res["peer_exchange_us (relative to the first stamp of the replay; the optimizer branch leads the replay)"] = {l: float(v) for l, v in zip(lab, p)}
    res["peer_split_us"] = {"K1 wait for peers' backward": float(p[1] - p[0]), "K1 reduce 7/8 of the shard over NVLink": float(p[2] - p[1]),
                            "K1 end -> K2 start": float(p[3] - p[2]), "K2 wait for norms": float(p[4] - p[3]),
                            "K2 Adam + publish": float(p[5] - p[4]), "K2 wait for all publishes": float(p[6] - p[5]), "total": float(p[6] - p[0])}
if early_acc is not None:
    q = early_acc / n_rep
    res["peer_early_us (same time base; this replay's backward)"] = {"start": float(q[0]), "every peer's early range complete": float(q[1]), "a block finished its pull": float(q[2])}
    res["peer_split_us"]["EARLY: wait for peers' fine levels"] = float(q[1] - q[0])
    res["peer_split_us"]["EARLY: pull of the early range (one block's end)"] = float(q[2] - q[1])
if world > 1:
    every = [None] * world
    dist.all_gather_object(every, res)
else:
    every = [res]
if rank == 0:
    for r in every:
        print(f"== rank {r['rank']}/{r['world']}: {len(r['calls'])} stamped calls, mean of {r['replays']} replays, step span {r['step_span_us']:.1f} us")
        if r["rank"] == 0 or world <= 2:
            for c in r["calls"]:
                print(f"  stream {c['stream']}  {c['done_us']:8.1f}  (+{c['since_prev_on_stream_us']:6.1f})  {c['call']}")
        if "peer_split_us" in r:
            print("  peer:", {k: round(v, 1) for k, v in r["peer_split_us"].items()})
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(every, f, indent=1)
if world > 1:
    dist.barrier()
    tr.comm.close()
    dist.destroy_process_group()
