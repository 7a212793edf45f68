"""Developer tool: time the occupancy-grid update (graph replay and per-call breakdown). Not part of the product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ncn_b200
from ncn_b200 import synth, vren, _lib
from ncn_b200.trainer import NeRFTrainer
dev = torch.device("cuda:0")
tr = NeRFTrainer(dict(batch_size=8192), device=dev)
grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
tr.model.density_grid.copy_(torch.from_numpy(grid).to(dev))
vren.packbits(tr.model.density_grid, 5.9, tr.model.density_bitfield)
fs = tr.fused_step(use_graph=True)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("update graph replay us:", t(lambda: fs.update_grid()))
_lib.Profiler.reset(); _lib.Profiler.counting = True; _lib.Profiler.timing = {"*"}
fs.use_graph = False
for _ in range(4): fs.update_grid()
torch.cuda.synchronize()
for k, (c, ms) in sorted(_lib.Profiler.summary().items(), key=lambda kv: -kv[1][1]): print(k, c, round(1e3 * ms / 4, 1), "us per update")
thr = 5.0
print("cumsum us:", t(lambda: torch.cumsum(tr.model.density_grid[0] > thr, 0, dtype=torch.int32)))
