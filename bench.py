#!/usr/bin/env python
"""bench.py - training rays/s of the ngp_mt hot path on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W     # the CPU arm: oracle/step.py on the host cores

A step = one full training pass over one batch of 8192 synthetic Hypersim-shaped rays per GPU:
AABB + occupancy march -> hash-grid encode -> sigma/rgb MLPs -> compositing -> photometric + opacity +
Manhattan normal-clustering loss -> backward -> NCCL all-reduce (N>1) -> clip + Adam (+ occupancy-grid update
every 16 steps).  `value` times K steps with the ray batches already resident in HBM; `e2e` times the same step
through the public API from pinned HOST buffers (H2D of the batch + D2H of the loss inside the timed region).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC, UNIT = "training rays/sec", "rays/s"
WORKLOAD = ("ngp_mt single synthetic Hypersim-shaped scene 1024x768, 8192 rays/step/GPU (128 patches of 8x8), "
            "fp16 params/activations fp32 accumulate, RGB+depth heads + opacity + normal-clustering loss, "
            "L=16 F=2 T=2^19, grid 128^3, max_samples 1024")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", "--no-baselines", dest="no_cpu_baseline", action="store_true",
                    help="skip the baseline legs that run after the timed regions (CPU reference losses.py / CPU port / reference csrc kernels and step on the GPU)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 = the headline RGB+depth step (default), 3 = + semantic and normal heads, "
                         "4 = full-image evaluation render (tools/eval_render.py), 5 = hash-grid sweep T=2^19..2^22 (tools/sweep_hashgrid.py)")
    ap.add_argument("--breakdown", default=None, help="write a per-call CUDA-event breakdown (json) to this path")
    ap.add_argument("--update-interval", type=int, default=None, help="occupancy-grid update period in steps (default: the reference's 16)")
    ap.add_argument("--fuse-fwd", default="mlp", choices=["none", "mlp", "all"], help="forward fusion of the fused step (A/B)")
    ap.add_argument("--heads", default="", help="extra heads of BASELINE config 3, e.g. 'sem,norm' (semantic head with 3 classes + "
                    "cross-entropy, normal head); default: the headline RGB+depth configuration")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"], help="N>1 gradient exchange: sharded optimizer over NVLink "
                    "peer memory (default when every rank can map every peer) or ncclAllReduce + replicated Adam")
    ap.add_argument("--no-graph", action="store_true", help="run the fused step eagerly instead of replaying its CUDA graph")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """samples SM clock + throttle reasons during the timed region (pynvml; nvidia-smi fallback)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            while not self._stop.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_steps(n_rays, steps, warmup, seed=0):
    """oracle/step.py on the host cores; returns (rays/s, ms/step, threads, samples/ray)."""
    import numpy as np
    import torch
    import ncn_b200  # noqa: F401
    from ncn_b200 import synth
    from oracle import march, step as ostep
    torch.set_num_threads(os.cpu_count() or 1)
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    bits = march.packbits(grid, 5.9)
    field = ostep.CpuField()
    st = {}
    batches = [synth.patch_batch(n_rays, seed=seed + i) for i in range(2)]
    tgt = [torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(i)) for i in range(2)]
    n_samp = 0
    for i in range(warmup):
        ostep.train_step(field, bits, batches[i % 2]["rays_o"], batches[i % 2]["rays_d"], tgt[i % 2], batches[i % 2]["tri"], opt_state=st)
    t0 = time.perf_counter()
    for i in range(steps):
        _, n = ostep.train_step(field, bits, batches[i % 2]["rays_o"], batches[i % 2]["rays_d"], tgt[i % 2], batches[i % 2]["tri"], opt_state=st)
        n_samp += n
    dt = time.perf_counter() - t0
    return n_rays * steps / dt, dt / steps * 1e3, torch.get_num_threads(), n_samp / max(1, steps * n_rays)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    n_rays = 512 if total <= 40 else (256 if total <= 120 else 128)
    v, ms, threads, spr = cpu_steps(n_rays, args.steps, args.warmup)
    sample = f"{n_rays}-ray sub-batch of the 8192-ray step (same scene, same step: march+encode+MLP+composite+loss+backward+Adam), torch CPU"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample, "samples_per_ray": spr},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    line["config"]["same_config_as_gpu_arm"] = False      # a sub-batch of the 8192-ray step, CPU port of the step
    try:        # BASELINE.md 3a / BASELINE.json config 1: the reference's OWN losses.py on the same host cores, in the same run
        from tools import baselines
        rl = baselines.ref_losses_cpu(reps=5)
        if rl is not None:
            line["reference_losses_py"] = {k: rl[k] for k in ("what", "config", "cores", "ms_fwd_bwd", "sub_ms", "rays_per_s", "normals_per_s")}
    except Exception as e:  # noqa: BLE001
        line["reference_losses_py"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- our arm
ALGO = {  # algorithmic bytes (or flops) per unit, SURVEY.md section 8d / DESIGN.md
    "ncn_grid_fwd": ("hbm", "sample", 588.0), "ncn_grid_bwd": ("hbm", "sample", 12 + 64 + 1024.0),
    "ncn_adam_step": ("hbm", "param", 34.0), "ncn_adam_step_groups": ("hbm", "param", 34.0), "ncn_grad_sumsq": ("hbm", "param", 4.0),
    "ncn_composite_train_fw": ("hbm", "sample", 28.0), "ncn_composite_train_bw": ("hbm", "sample", 48.0),
    "ncn_composite_train_fw_photometric": ("hbm", "sample", 28.0),      # + 56 B/ray of loss terms (negligible against 28 B x 33 samples)
    "ncn_composite_train_fw_photometric_gt": ("hbm", "sample", 28.0),   # the random_tr_poses form of the same launch
    "ncn_march_train_expand": ("hbm", "sample", 36.0),
    # tcgen05 MLP backward, two launches per step (colour head 448 B/sample, density trunk 288 B/sample): average per launch
    "ncn_mlp_bwd_src_fused": ("hbm", "sample", 368.0), "ncn_mlp_bwd": ("hbm", "sample", 368.0),
    "ncn_mlp_fwd": ("hbm", "sample", 288.0),
    "ncn_field_mlp_fwd": ("hbm", "sample", 64 + 12 + 128 + 32 + 64 + 256 + 32 + 12 + 4.0),     # feat, dirs in; sig_acts, h, x_rgb, rgb_acts, rgb_out, raws, sigmas out
}


# kernels BASELINE.json's metric names explicitly ("composite/hashgrid GB/s vs HBM peak") + the other step kernels with an
# algorithmic byte count: reported per launch in the line's "kernels" object
KERNEL_REPORT = ("ncn_composite_train_fw", "ncn_composite_train_fw_photometric", "ncn_composite_train_bw", "ncn_grid_fwd", "ncn_grid_bwd", "ncn_field_mlp_fwd",
                 "ncn_mlp_bwd_src_fused", "ncn_adam_step_groups", "ncn_grad_sumsq")


NOTES = {
    "ncn_grid_bwd": "algorithmic bytes = x (12) + dL/dfeat (64) + 16 levels x 8 corners x 8 B of red.global.add (1024) per sample; the "
                    "reductions are absorbed by the L2 atomic units (the 46 MB fp32 gradient stays L2-resident until Adam reads it), so "
                    "DRAM traffic is far BELOW the algorithmic bytes and the limiter is L2 reduction transactions, not HBM",
    "ncn_mlp_bwd_src_fused": "two launches per step (colour head 448 B/sample, density trunk 288 B/sample); average per launch",
}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import ncn_b200  # noqa: F401
    from ncn_b200 import _lib, synth, vren
    from ncn_b200.trainer import NeRFTrainer

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)
    R = args.rays
    heads = [h for h in args.heads.split(",") if h]
    hp_extra = dict(pred_sem="sem" in heads, pred_norm_nn="norm" in heads, loss_sem_w=4e-2 if "sem" in heads else 0)
    tr = NeRFTrainer(dict(batch_size=R, **hp_extra), device=dev, rank=rank, world_size=world, n_sem_cls=3 if "sem" in heads else 0,
                     shard_optimizer={"auto": None, "peer": True, "nccl": False}[args.exchange] if world > 1 else False)
    grid = synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))
    tr.model.density_grid.copy_(torch.from_numpy(grid).to(dev))
    vren.packbits(tr.model.density_grid, 5.9, tr.model.density_bitfield)
    poses = synth.camera_poses(50, 0); dirs = synth.pixel_directions("hypersim")
    tr.set_cameras(poses, dirs)
    if args.update_interval:
        tr.hp["update_interval"] = args.update_interval
    # steady state: clustering weights on, past the occupancy warm-up.  The timed region STARTS on a grid-update step (3008 = 16 * 188),
    # so it always contains ceil(K / 16) occupancy-grid updates whatever --steps is (never fewer than the reference's 1-in-16)
    W_eff = max(args.warmup, 3)
    tr.global_step = 3008 - W_eff
    # The occupancy-grid update (1 M-point density query + decay/max + packbits, every 16 steps) runs in full, but
    # with a RANDOM-INIT field it would flood the grid within a few updates and the samples/ray would drift with the
    # number of steps run.  To keep the synthetic room stationary the pre-update grid / bitfield are restored after
    # each update (two extra device copies; no work is skipped).
    grid0, bits0 = tr.model.density_grid.clone(), tr.model.density_bitfield.clone()
    restore = (grid0, bits0)
    NB = 8
    host = []
    for i in range(NB):
        b = synth.patch_batch(R, seed=1000 * rank + i)
        host.append(dict(img=torch.from_numpy(b["img_idx"]).pin_memory(), pix=torch.from_numpy(b["pix_idx"]).pin_memory(),
                         rgb=torch.rand(R, 3, generator=torch.Generator().manual_seed(i)).pin_memory(), tri=b["tri"]))
    tri_local = {k: torch.from_numpy(host[0]["tri"][j][:49] % 64).to(dev) for j, k in enumerate(("x1", "x2", "x3"))}

    def target_of(rgb):
        return {"rgb": rgb, "patch_area": 64, "x1_offsets_local": tri_local["x1"], "x2_offsets_local": tri_local["x2"],
                "x3_offsets_local": tri_local["x3"]}

    resident = []
    for h in host:
        ro, rd = tr.rays_from_batch(h["img"].to(dev), h["pix"].to(dev))
        resident.append(torch.stack([ro, rd, h["rgb"].to(dev)]).contiguous())      # (3,R,3) [rays_o | rays_d | rgb]
    h2d_bytes = sum(host[0][k].numel() * host[0][k].element_size() for k in ("img", "pix", "rgb"))

    tri_dev = torch.from_numpy(host[0]["tri"]).to(dev)      # identical triangle topology for every patch batch
    fs = tr.fused_step(use_graph=not args.no_graph, fuse_fwd={"none": False, "mlp": "mlp", "all": True}[args.fuse_fwd])
    fs.set_triangles(tri_dev)
    if "sem" in heads:          # semantics_WF-style labels (0 = void, 1..3), resident: 64 KB per step if they were copied
        fs.sem_target.copy_(torch.randint(0, 4, (R,), generator=torch.Generator().manual_seed(7 + rank)).to(dev))
    loss_pin = torch.zeros(8, dtype=torch.float32).pin_memory()

    def step_resident(i):
        tr.train_step_fused(packed=resident[i % NB], grid_restore=restore)


    # End to end: every step copies its batch from pinned host memory, runs, and copies its loss sums back to pinned host
    # memory; the host READS the loss of step i-1 while step i is already enqueued (double-buffered slots + events), the
    # way a training loop logs - so the device never idles on the read-back.  All losses are read by the end of the region.
    loss_slots = [torch.zeros(8, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_events = [None, None]
    e2e_losses = []

    def read_loss(slot):
        if loss_events[slot] is not None:
            loss_events[slot].synchronize()
            e2e_losses.append(float(loss_slots[slot][0]) / (3 * R))
            loss_events[slot] = None

    for h in host:      # the batch as ONE pinned record [img_idx i64 | pix_idx i64 | rgb f32] = a single H2D copy per step
        h["packed"] = fs.pack_pixel_batch(h["img"], h["pix"], h["rgb"])
    assert host[0]["packed"].numel() == h2d_bytes

    def step_e2e(i):
        h = host[i % NB]
        fs.batch.copy_(h["packed"], non_blocking=True)            # H2D of the batch; rays are generated inside the step graph
        tr.train_step_fused(grid_restore=restore)
        slot = i & 1
        loss_slots[slot].copy_(fs.zeros, non_blocking=True)      # D2H of this step's loss sums (32 B)
        ev = torch.cuda.Event(); ev.record(); loss_events[slot] = ev
        read_loss(slot ^ 1)                                       # host reads the PREVIOUS step's loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        fs.flush()                      # the last step's (deferred) optimizer pass belongs to the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    fs.census = {}                      # count the nodes of the captured graphs (gpu_launches)
    fs.update_grid(restore=restore)     # untimed: captures the grid-update graph (the warm-up steps below may not hit an update step)
    for i in range(W_eff):
        step_resident(i)
    assert tr.global_step == 3008

    if tr.peer is not None and tr.peer.error() != 0:
        raise RuntimeError("peer-memory exchange: a cross-GPU wait timed out during warm-up (ncn_peer_error)")

    # ---- timed region 1: inputs resident in HBM (CUDA-graph replay of the fused step)
    clocks = ClockSampler(local); clocks.start()
    ms = timed(step_resident, args.steps)
    clk = clocks.stop()
    value = world * R * args.steps / (ms * 1e-3)
    census = dict(fs.census)            # node counts of the graphs replayed in region 1 (region 2 re-captures with ray generation in front)
    n_updates = sum(1 for q in range(args.steps) if (3008 + q) % tr.hp["update_interval"] == 0)

    # ---- timed region 2: end to end from pinned host buffers, loss read back every step
    fs.use_pixel_batches(True)
    tr.global_step = 3008 - 2
    for i in range(2):
        step_e2e(i)

    def timed_e2e(steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step_e2e(i)
        fs.flush()
        e1.record()
        read_loss(0); read_loss(1)          # the last losses are read inside the region's wall clock too (e1 is already recorded)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    ms_e2e = timed_e2e(args.steps)
    assert len(e2e_losses) >= args.steps and all(np.isfinite(e2e_losses))
    e2e = world * R * args.steps / (ms_e2e * 1e-3)
    _, n_samples = fs.stats_host()
    fs.use_pixel_batches(False)

    # ---- instrumented pass: the SAME step, same buffers, eager (no graph) with CUDA events around every libncn call
    #      -> launch count and the dominant kernel's average launch time (events cannot be read inside a replayed graph)
    fs.use_graph = False
    fs.serial = True                    # no concurrent branches: every call is timed alone
    _lib.Profiler.reset(); _lib.Profiler.counting = True; _lib.Profiler.timing = {"*"}
    nprof = 8
    saved_interval = tr.hp["update_interval"]; tr.hp["update_interval"] = 1 << 30      # time the step itself
    for i in range(nprof):
        step_resident(i)
    summ = _lib.Profiler.summary()
    launches_per_step = _lib.Profiler.launches / nprof
    _lib.Profiler.counting = False; _lib.Profiler.timing = None
    tr.hp["update_interval"] = saved_interval
    fs.use_graph = not args.no_graph
    fs.serial = False
    # launches inside the timed region: from the replayed graphs' own kernel-node census when available (libncn kernels AND the few
    # torch nodes - jitter, accumulator resets - that are captured with them), else from the eager pass's count of libncn kernels
    own_per_step = launches_per_step
    step_nodes = census.get("step")
    if isinstance(step_nodes, list):
        launches_per_step = float(step_nodes[0])
        upd = census.get("grid_update")
        launches = int(step_nodes[0] * args.steps + (upd[0] * n_updates if isinstance(upd, list) else 0))
    else:
        launches = int(round(launches_per_step * args.steps))
    # the dominant KERNEL = the longest single launch (a call name that launches twice per step is compared per launch)
    top = max((k for k in summ if k in ALGO), key=lambda k: summ[k][1] / max(summ[k][0], 1), default=None)
    top_stats = summ.get(top) if top else None
    if args.breakdown and rank == 0:
        with open(args.breakdown, "w") as f:
            json.dump({k: {"calls_per_step": c / nprof, "us_per_step": 1e3 * t / nprof} for k, (c, t) in sorted(summ.items(), key=lambda kv: -kv[1][1])}, f, indent=1)

    ranks_in_sync = None
    if world > 1:      # replicated parameters must be bit-identical on every rank after the run (same all-reduced gradient, same Adam)
        fs.flush()
        # sharded peer-memory optimizer: the replicated state is the fp16 working copy (each slice written by its owner into
        # every rank's buffer); the fp32 master is only current inside a rank's own slice
        mine = tr.opt.flat16.float() if tr.peer is not None else tr.opt.flat
        ref_p = mine.clone()
        dist.broadcast(ref_p, 0)
        diff = (mine - ref_p).abs().max()
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        ranks_in_sync = bool(diff.item() == 0.0) and (tr.peer is None or tr.peer.error() == 0)

    roofline = None
    if top and top_stats:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        bound, unit_of, per_unit = ALGO[top]
        n_params = tr.opt.grad.numel()
        units = n_samples if unit_of == "sample" else n_params
        calls, tot_ms = top_stats
        avg_ms = tot_ms / calls
        if top == "ncn_adam_step":          # two parameter groups -> two launches per step; time per launch, bytes per launch
            units = n_params / 2.0
        achieved = per_unit * units / (avg_ms * 1e-3) / 1e9
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        traffic_src = None
        for tf in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):      # DRAM bytes per launch of this kernel from the committed ncu --set full capture
            try:
                traffic = float(json.load(open(os.path.join(ROOT, "profiles", tf)))[top]["bytes_per_launch"])
                traffic_src = "profiles/" + tf
                break
            except Exception:  # noqa: BLE001
                pass
        roofline = {"kernel": top, "bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": (traffic_src + " (ncu --set full, dram read+write per launch)") if traffic else None,
                    "note": NOTES.get(top), "avg_launch_ms": avg_ms, "launches_timed": calls,
                    "timed_in": "instrumented eager pass of the same step right after the timed region (CUDA events around the launch)",
                    "algorithmic_bytes_per_launch": per_unit * units,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}

    # per-kernel achieved bandwidth of the kernels BASELINE.json's metric names (composite, hash grid) and of the other step kernels
    # with an algorithmic byte count: bytes per launch / CUDA-event time per launch (instrumented eager pass), against the measured peak
    peaks_k = {}
    try:
        peaks_k = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_k = float(peaks_k.get("hbm_gbs", 6650.0))
    kernels = {}
    for k in KERNEL_REPORT:
        if k not in summ or k not in ALGO:
            continue
        calls, tot_ms = summ[k]
        _, unit_of, per_unit = ALGO[k]
        units = n_samples if unit_of == "sample" else tr.opt.grad.numel()
        us = 1e3 * tot_ms / calls
        gbs = per_unit * units / (us * 1e-6) / 1e9
        kernels[k] = {"us_per_launch": us, "launches_per_step": calls / nprof, "algorithmic_bytes_per_launch": per_unit * units,
                      "GB/s": gbs, "frac_of_hbm_peak": gbs / peak_k}
    cpu = None
    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # --- baseline legs (after the timed regions; the oracle / reference here is the thing MEASURED AGAINST, never the product)
        v, cms, threads, spr = cpu_steps(512, 4, 1)
        port = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": cms,
                "sample": "512-ray sub-batch of the same step (oracle/step.py: C march + torch hash-grid/MLP/composite/loss/Adam), 4 steps after 1 warm-up"}
        cpu = port
        try:
            from tools import baselines
            rl = baselines.ref_losses_cpu()
            if rl is not None:      # the reference's own losses.py on config 1, same host cores, same run
                cpu = {"value": rl["rays_per_s"], "unit": "rays/s through losses.py (config 1)", "cores": rl["cores"], "kind": "reference",
                       "sample": rl["config"] + "; median of %d fwd+bwd" % rl["reps"], "ms_fwd_bwd": rl["ms_fwd_bwd"], "sub_ms": rl["sub_ms"],
                       "what": rl["what"], "port_full_step": port}
            gpu_ref = {"step_reference_csrc": baselines.ref_step_gpu(R, 10, 3, "ref"),
                       "step_reference_on_shims": baselines.ref_step_gpu(R, 10, 3, "shim")}
            for k in ("step_reference_csrc", "step_reference_on_shims"):
                if gpu_ref[k]:
                    gpu_ref[k]["ours_over_it"] = (value / world) / gpu_ref[k]["rays_per_s"]
        except Exception as e:  # noqa: BLE001  (a baseline leg must never take the product's line down)
            gpu_ref = {"error": f"{type(e).__name__}: {e}"}
        try:                    # last: its graph-capture attempts of the reference's functions are the most fragile part
            gpu_ref["csrc_kernels_us"] = baselines.ref_kernels_gpu(R)
        except Exception as e:  # noqa: BLE001
            gpu_ref["csrc_kernels_us"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16", "data": "synthetic",
                "config": {"workload": WORKLOAD if not heads else WORKLOAD.replace("RGB+depth heads", "RGB+depth+" + "+".join(heads) + " heads (config 3: n_sem_cls 3, cross-entropy w 4e-2)"),
                           "rays_per_step_per_gpu": R, "samples_per_step_per_gpu": n_samples,
                           "parallelism": f"dp{world}", "ranks_in_sync": ranks_in_sync,
                           "exchange": None if world == 1 else ("sharded reduce + Adam + fp16 publish over NVLink peer memory (ncn_peer_step)" if tr.peer is not None else "ncclAllReduce(fp32 flat gradient) + replicated Adam"), "path": "fused CUDA-graph step (ncn_b200.fused.FusedStep)" if not args.no_graph else "fused eager step",
                           "occupancy": "synthetic room (13.6 % of 128^3 cells); grid update every 16 steps runs in full (the timed region starts on an update step: ceil(K/16) updates inside), its result is reverted to keep samples/ray stationary",
                           "init": "random (tcnn-style U(-1e-4,1e-4) table, Xavier MLPs)",
                           "l2": "no explicit flush: per-step working set (fp32 params+grads+Adam m,v = 183 MB, + 22 MB fp16 table) exceeds the 126 MB L2"},
                "clocks": clk, "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                                       "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 32,
                                       "readback": "loss sums copied to pinned host memory every step; the host reads step i-1's loss while step i runs"},
                "gpu_launches": launches, "gpu_launches_per_step": launches_per_step,
                "gpu_launch_census": {"graphs [kernel, memcpy, memset, other nodes]": census, "grid_updates_in_region": n_updates,
                                      "libncn_kernels_per_step": own_per_step},
                "kernels": kernels, "roofline": roofline, "cpu_baseline": cpu, "gpu_reference": gpu_ref}
        print(json.dumps(line))
    tr.comm.close()
    if world > 1:
        dist.destroy_process_group()


def run_tool(args):
    """configs 4 / 5: the evaluation render and the hash-grid sweep have their own drivers (same launch contract, one JSON line)"""
    import runpy
    tool, argv = {4: ("eval_render.py", ["--heads", "sem,norm", "--reps", str(max(3, min(args.steps, 10)))]),
                  5: ("sweep_hashgrid.py", ["--reps", str(max(3, min(args.steps, 10)))])}[args.config]
    sys.argv = [os.path.join(ROOT, "tools", tool)] + argv
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    a = parse()
    if a.config == 3 and not a.heads:
        a.heads = "sem,norm"
    if a.impl == "reference":
        run_reference(a)
    elif a.config in (4, 5):
        run_tool(a)
    else:
        run_ours(a)
