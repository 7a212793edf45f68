"""One eager fused training step for ncu captures (no CUDA graph). Not part of the product."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ncn_b200
from ncn_b200 import synth, vren
from ncn_b200.trainer import NeRFTrainer

dev = torch.device("cuda:0")
torch.manual_seed(0)
R = 8192
heads = os.environ.get("NCN_PROF_HEADS", "")          # e.g. "sem,norm" = BASELINE config 3
hp = dict(batch_size=R, pred_sem="sem" in heads, pred_norm_nn="norm" in heads, loss_sem_w=4e-2 if "sem" in heads else 0)
tr = NeRFTrainer(hp, device=dev, n_sem_cls=3 if "sem" in heads else 0, shard_optimizer=bool(os.environ.get("NCN_PROF_PEER")))
grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
tr.model.density_grid.copy_(torch.from_numpy(grid).to(dev))
vren.packbits(tr.model.density_grid, 5.9, tr.model.density_bitfield)
tr.global_step = 3009
b = synth.patch_batch(R, seed=0)
ro = torch.from_numpy(b["rays_o"]).to(dev); rd = torch.from_numpy(b["rays_d"]).to(dev)
rgb = torch.rand(R, 3, device=dev)
fs = tr.fused_step(use_graph=False)
fs.set_triangles(torch.from_numpy(b["tri"]).to(dev))
if "sem" in heads:
    fs.sem_target.copy_(torch.randint(0, 4, (R,), device=dev))
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(n_steps):
    if i == n_steps - 1:          # ncu --nvtx --nvtx-include "prof/" captures only the last (warm) step
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("prof")
    tr.train_step_fused(ro, rd, rgb, update_grid=False)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("ok", fs.stats_host())
