"""Developer tool: timeline of the captured fused step.  Every libncn call of the step is followed by a one-thread kernel
that stores %globaltimer (ncn_debug_stamp) on the same stream, the stamps are captured into the CUDA graph with the step,
and after a few replays the per-call completion times are printed per stream.  Not part of the product."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ncn_b200
from ncn_b200 import _lib, synth, vren
import ncn_b200.fused as F
from ncn_b200.trainer import NeRFTrainer

dev = torch.device("cuda:0")
torch.manual_seed(0)
R = 8192
tr = NeRFTrainer(dict(batch_size=R), device=dev)
grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
tr.model.density_grid.copy_(torch.from_numpy(grid).to(dev))
vren.packbits(tr.model.density_grid, 5.9, tr.model.density_bitfield)
tr.global_step = 3009
tr.hp["update_interval"] = 1 << 30
b = synth.patch_batch(R, seed=0)
ro = torch.from_numpy(b["rays_o"]).to(dev); rd = torch.from_numpy(b["rays_d"]).to(dev)
rgb = torch.rand(R, 3, device=dev)

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
fs = tr.fused_step(use_graph=True)
fs.set_triangles(torch.from_numpy(b["tri"]).to(dev))
acc = None
n_rep = 0
for i in range(12):
    tr.train_step_fused(ro, rd, rgb, update_grid=False)
    torch.cuda.synchronize()
    if i >= 4 and names:
        t = slots[1:len(names) + 1].cpu().numpy().astype(np.float64)
        t = (t - t.min()) / 1e3
        acc = t if acc is None else acc + t
        n_rep += 1
fs.flush()
t = acc / n_rep
print(f"{len(names)} stamped calls, mean of {n_rep} replays; completion time (us) relative to the first stamp")
order = np.argsort(t)
last = {}
for i in order:
    what, sid = names[i]
    d = t[i] - last.get(sid, 0.0)
    last[sid] = t[i]
    print(f"  stream {sid}  {t[i]:8.1f}  (+{d:6.1f} since the previous call on this stream)  {what}")
print("step span", t.max() - t.min())
