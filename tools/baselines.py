"""Reference baselines timed beside the product.  Measurement infrastructure: imported by bench.py's
baseline legs (after its timed regions) and runnable on its own; nothing here is on the product path.

  ref_losses_cpu    section 3a, BASELINE.json config 1: the reference's OWN losses.py (NeRFMTLoss forward + backward, imported
                    unchanged via oracle/ref_py.py) on CPU tensors, 8192 synthetic room rays -> 6272 depth-derived normals, with
                    k-means / selection / tail sub-timings.  faiss is absent: faiss.Kmeans is the numpy stand-in of SURVEY App. C.
  ref_kernels_gpu   section 3b: the 15 functions of the reference's own csrc (oracle/_ref/vren_ref.so, compiled from the reference's
                    sources for sm_100) timed with CUDA events next to the libncn entry points on the same inputs.
  ref_step_gpu      the closest runnable form of "the reference's CUDA step": the reference's own render() + NeRFMTLoss + autograd
                    (models/rendering.py, ngp_mt.py, custom_functions.py, losses.py, unchanged) bound to the reference's own csrc
                    kernels, CPU k-means (the reference calls faiss with gpu=False), GradScaler + clip + Adam as Lightning
                    would drive them - with tiny-cuda-nn replaced by the libncn hash grid / MLPs, because tcnn's source is not
                    available.  So it under-states the reference's step time by whatever tcnn is slower than libncn.

    python tools/baselines.py [losses|kernels|step] [--json out.json]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HP = dict(loss_opacity_w=1e-3, loss_norm_can_tres=0.01, loss_norm_D_C_ort_dot_w=2e-3, loss_norm_D_C_centr_dot_w=2e-3,
          loss_norm_D_C_centr_L1_w=2e-3, loss_norm_can_start=500, loss_norm_can_grow=2500, loss_norm_can_end=-1,
          ray_sampling_strategy="all_images_triang_patch", random_tr_poses=False, pred_norm_nn=False, pred_norm_depth=True)


def _median(v):
    v = sorted(v)
    return v[len(v) // 2]


# ------------------------------------------------------------------------------------------- 3a: losses.py on the host cores
def ref_losses_cpu(n_rays=8192, reps=7, warmup=2, seed=5):
    import numpy as np
    import torch
    import ncn_b200  # noqa: F401
    from ncn_b200 import synth
    from oracle import ref_py
    ns = ref_py.load(faiss="cpu", want_models=False)
    if ns is None:
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    L = ns.losses
    b = synth.patch_batch(n_rays, seed=seed)
    g = torch.Generator().manual_seed(seed)
    rays_d = torch.from_numpy(b["rays_d"]); rays_o = torch.from_numpy(b["rays_o"])
    t_wall = torch.where(rays_d > 0, (0.4 - rays_o) / rays_d, (-0.4 - rays_o) / rays_d).min(-1)[0]
    depth0 = (t_wall + 0.002 * torch.randn(n_rays, generator=g)).clamp_min(0.02)
    opacity = torch.rand(n_rays, generator=g).clamp(0.05, 0.99)
    rgb0 = torch.rand(n_rays, 3, generator=g)
    tri = b["tri"][:, :49] % 64
    target = {"rgb": torch.rand(n_rays, 3, generator=g), "patch_area": 64, "x1_offsets_local": torch.from_numpy(tri[0]),
              "x2_offsets_local": torch.from_numpy(tri[1]), "x3_offsets_local": torch.from_numpy(tri[2])}
    loss_fn = L.NeRFMTLoss(dict(HP))
    # sub-timers: wrap the reference's own functions (not edits - attribute rebinding on the imported module)
    acc = {"kmeans": 0.0, "clustering": 0.0}
    Km = ns.faiss.Kmeans
    orig_train = Km.train
    orig_clu = L._normals_clustering

    def train(self, x):
        t0 = time.perf_counter(); r = orig_train(self, x); acc["kmeans"] += time.perf_counter() - t0
        return r

    def clu(*a, **k):
        t0 = time.perf_counter(); r = orig_clu(*a, **k); acc["clustering"] += time.perf_counter() - t0
        return r

    Km.train = train
    L._normals_clustering = clu
    rows = []
    n_normals = None
    try:
        for i in range(warmup + reps):
            depth = depth0.clone().requires_grad_(True); rgb = rgb0.clone().requires_grad_(True)
            pred = {"rgb": rgb, "depth": depth, "opacity": opacity, "rays_o": rays_d, "rays_d": rays_d,      # rays_o := rays_d (rendering.py:227)
                    "deltas": torch.zeros(1), "ts": torch.zeros(1), "rays_a": torch.zeros(1, 3, dtype=torch.int64)}
            acc["kmeans"] = acc["clustering"] = 0.0
            t0 = time.perf_counter()
            loss_d = loss_fn(pred, target, global_step=3000)
            t1 = time.perf_counter()
            loss_d["total"].backward()
            t2 = time.perf_counter()
            n_normals = int(Km.last["x"].shape[0])
            if i >= warmup:
                rows.append((t1 - t0, t2 - t1, acc["kmeans"], acc["clustering"]))
    finally:
        Km.train = orig_train
        L._normals_clustering = orig_clu
    fw = _median([r[0] for r in rows]); bw = _median([r[1] for r in rows])
    km = _median([r[2] for r in rows]); cl = _median([r[3] for r in rows])
    tot = _median([r[0] + r[1] for r in rows])
    return {"what": "the reference's own losses.py::NeRFMTLoss forward+backward (imported unchanged) on CPU tensors; faiss.Kmeans = numpy "
                    "stand-in (faiss absent, SURVEY App. C)",
            "config": f"BASELINE.json config 1: {n_rays} synthetic room rays -> {n_normals} depth-derived normals, paper weights, global_step 3000",
            "cores": torch.get_num_threads(), "reps": reps, "ms_fwd_bwd": tot * 1e3, "ms_fwd": fw * 1e3, "ms_bwd": bw * 1e3,
            "sub_ms": {"kmeans_train (losses.py:86-88)": km * 1e3, "search+selection (losses.py:89-166)": (cl - km) * 1e3,
                       "normals + loss tail (losses.py:331-509)": (fw - cl) * 1e3, "backward": bw * 1e3},
            "rays_per_s": n_rays / tot, "normals_per_s": n_normals / tot,
            "losses": {k: float(v) for k, v in loss_d.items()}}


# ------------------------------------------------------------------------------------------- 3b: csrc kernels on the same GPU
def _time_gpu(fn, reps=20, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return _median(ts)


def _time_graph(fn, calls=10, replays=5):
    """device time per call with the host out of the picture: `calls` invocations captured into one CUDA graph, the replay timed
    with events.  None when the function cannot be captured (host synchronisation / host-side sizing inside it)."""
    import torch
    try:
        fn(); torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(calls):
                    fn()
        torch.cuda.current_stream().wait_stream(s)
        g.replay(); torch.cuda.synchronize()
        ts = []
        for _ in range(replays):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / calls)
        del g
        return _median(ts)
    except Exception:  # noqa: BLE001
        try:
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass
        return None


def ref_kernels_gpu(n_rays=8192, reps=20):
    """us per call of each of the 15 vren functions (binding.cpp:330-350): reference csrc (vren_ref) vs libncn (ncn_b200.vren), same
    inputs.  `ours_us` / `ref_us`: CUDA events around ONE python-level call on an idle GPU (both sides pay their own output
    allocations and their own host-side call path - ctypes here, pybind there: the calls of a few microseconds of device work are
    host bound on both sides).  `ours_device_us`: the libncn call captured 10x into a CUDA graph and replayed = its device time
    alone (null where the call sizes its outputs on the host).  The reference's functions launch on the legacy default stream
    (`<<<blocks, threads>>>` without a stream argument), which a stream capture does not record, so there is no such figure for
    them - a graph of them contains only their output allocations."""
    import numpy as np
    import torch
    import ncn_b200  # noqa: F401
    from ncn_b200 import synth, vren as ours
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    ref = build_ref.load()
    if ref is None:
        return None
    dev = torch.device("cuda")
    grid = torch.from_numpy(synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))).to(dev)
    bits = torch.from_numpy(synth.packbits_np(grid.cpu().numpy(), 5.9)).to(dev)
    b = synth.patch_batch(n_rays, seed=1000)
    rays_o = torch.from_numpy(b["rays_o"]).to(dev); rays_d = torch.from_numpy(b["rays_d"]).to(dev)
    center = torch.zeros(1, 3, device=dev); half = torch.full((1, 3), 0.5, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    noise = torch.rand(n_rays, device=dev, generator=g)
    coords = torch.randint(0, 128, (128 ** 3 // 4, 3), dtype=torch.int32, device=dev, generator=g)
    _, hits_t, _ = ours.ray_aabb_intersect(rays_o, rays_d, center, half, 1)
    hits_t[(hits_t[:, 0, 0] >= 0) & (hits_t[:, 0, 0] < 0.01), 0, 0] = 0.01
    ht = hits_t[:, 0].contiguous()
    rays_a, xyzs, dirs, deltas, ts, cnt = ours.raymarching_train(rays_o, rays_d, ht, bits, 1, 0.5, 0.0, noise, 128, 1024)
    N = xyzs.shape[0]
    sig = torch.rand(N, device=dev, generator=g) * 20
    out = {}
    radii = torch.full((1,), 0.5, device=dev)
    idx = ours.morton3D(coords)
    obits = torch.zeros_like(bits); rbits = torch.zeros_like(bits)
    alive = torch.arange(n_rays, device=dev)

    def both(name, f_ours, f_ref, work=None):
        t_o = _time_gpu(f_ours, reps); t_r = _time_gpu(f_ref, reps)
        out[name] = {"ours_us": t_o, "ref_us": t_r, "speedup": t_r / t_o}
        out[name]["ours_device_us"] = _time_graph(f_ours)
        if work:
            out[name]["work"] = work

    both("ray_aabb_intersect", lambda: ours.ray_aabb_intersect(rays_o, rays_d, center, half, 1),
         lambda: ref.ray_aabb_intersect(rays_o, rays_d, center, half, 1), f"{n_rays} rays x 1 box")
    both("ray_sphere_intersect", lambda: ours.ray_sphere_intersect(rays_o, rays_d, center, radii, 1),
         lambda: ref.ray_sphere_intersect(rays_o, rays_d, center, radii, 1), f"{n_rays} rays x 1 sphere")
    both("morton3D", lambda: ours.morton3D(coords), lambda: ref.morton3D(coords), f"{coords.shape[0]} cells")
    both("morton3D_invert", lambda: ours.morton3D_invert(idx), lambda: ref.morton3D_invert(idx), f"{coords.shape[0]} cells")
    both("packbits", lambda: ours.packbits(grid, 5.9, obits), lambda: ref.packbits(grid, 5.9, rbits), "128^3 cells")
    both("raymarching_train", lambda: ours.raymarching_train(rays_o, rays_d, ht, bits, 1, 0.5, 0.0, noise, 128, 1024),
         lambda: ref.raymarching_train(rays_o, rays_d, ht, bits, 1, 0.5, 0.0, noise, 128, 1024), f"{n_rays} rays -> {N} samples")
    both("raymarching_test", lambda: ours.raymarching_test(rays_o, rays_d, ht.clone(), alive, bits, 1, 0.5, 0.0, 128, 1024, 4),
         lambda: ref.raymarching_test(rays_o, rays_d, ht.clone(), alive, bits, 1, 0.5, 0.0, 128, 1024, 4), f"{n_rays} rays x 4 samples")
    for C, tag in ((3, ""), (9, "_multi")):
        raws = torch.rand(N, C, device=dev, generator=g)
        fw_o = getattr(ours, f"composite_train{tag}_fw"); fw_r = getattr(ref, f"composite_train{tag}_fw")
        bw_o = getattr(ours, f"composite_train{tag}_bw"); bw_r = getattr(ref, f"composite_train{tag}_bw")
        both(f"composite_train{tag}_fw", lambda: fw_o(sig, raws, deltas, ts, rays_a, 1e-4), lambda: fw_r(sig, raws, deltas, ts, rays_a, 1e-4),
             f"{N} samples x {C} channels")
        tot, opa, dep, rend, ws = fw_o(sig, raws, deltas, ts, rays_a, 1e-4)
        dO = torch.rand_like(opa); dD = torch.rand_like(dep); dR = torch.rand_like(rend); dW = torch.zeros_like(ws)
        both(f"composite_train{tag}_bw", lambda: bw_o(dO, dD, dR, dW, sig, raws, ws, deltas, ts, rays_a, opa, dep, rend, 1e-4),
             lambda: bw_r(dO, dD, dR, dW, sig, raws, ws, deltas, ts, rays_a, opa, dep, rend, 1e-4), f"{N} samples x {C} channels")
        # test-time compositing: one round of 4 samples per ray
        S = 4
        sg = torch.rand(n_rays, S, device=dev, generator=g) * 20; rw = torch.rand(n_rays, S, C, device=dev, generator=g)
        dl = torch.full((n_rays, S), 1.7e-3, device=dev); tt = torch.rand(n_rays, S, device=dev, generator=g).sort(1)[0]
        ne = torch.full((n_rays,), S, dtype=torch.int32, device=dev)
        st = {k: (torch.zeros(n_rays, device=dev), torch.zeros(n_rays, device=dev), torch.zeros(n_rays, C, device=dev)) for k in "or"}
        ht2 = {k: ht.clone() for k in "or"}; al = {k: alive.clone() for k in "or"}
        name = "composite_test_fw" if C == 3 else "composite_test_multi_fw"

        def run_test(mod, k, name=name, sg=sg, rw=rw, dl=dl, tt=tt, ne=ne, st=st, ht2=ht2, al=al):
            al[k].copy_(alive)                      # finished rays are marked -1 in place: start every call from the full set
            for t_ in st[k]:
                t_.zero_()
            getattr(mod, name)(sg, rw, dl, tt, ht2[k], al[k], 1e-4, ne, *st[k])

        both(name, lambda: run_test(ours, "o"), lambda: run_test(ref, "r"),
             f"{n_rays} rays x {S} samples x {C} channels (incl. 4 state resets on both sides)")
    wsd = torch.rand(N, device=dev, generator=g)
    both("distortion_loss_fw", lambda: ours.distortion_loss_fw(wsd, deltas, ts, rays_a), lambda: ref.distortion_loss_fw(wsd, deltas, ts, rays_a), f"{N} samples")
    loss, wi, wti = ours.distortion_loss_fw(wsd, deltas, ts, rays_a)
    dL = torch.rand_like(loss)
    both("distortion_loss_bw", lambda: ours.distortion_loss_bw(dL, wi, wti, wsd, deltas, ts, rays_a),
         lambda: ref.distortion_loss_bw(dL, wi, wti, wsd, deltas, ts, rays_a), f"{N} samples")
    return out


# ------------------------------------------------------------------------------------------- the reference's step on the GPU
def ref_step_gpu(n_rays=8192, steps=10, warmup=3, vren="ref"):
    """ms per step of the reference's own Python step (render + NeRFMTLoss + backward + GradScaler/clip/Adam) on this GPU.
    vren="ref": bound to the reference's csrc kernels; vren="shim": bound to libncn (the drop-in path a user of the shims gets)."""
    import contextlib
    import io
    import numpy as np
    import torch
    import ncn_b200  # noqa: F401
    from ncn_b200 import synth
    from oracle import ref_py
    ns = ref_py.load(faiss="cpu", want_models=True, vren=vren)
    if ns is None:
        return None
    dev = torch.device("cuda")
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = ns.ngp_mt.NGPMT(scale=0.5, grid_size=128, rgb_act="Sigmoid").to(dev)
    grid = torch.from_numpy(synth.density_grid_from_occupancy(synth.room_occupancy(128, 0.5, seed=0))).to(dev)
    model.register_buffer("density_grid", grid)
    model.density_bitfield.copy_(torch.from_numpy(synth.packbits_np(grid.cpu().numpy(), 5.9)).to(dev))
    loss_fn = ns.losses.NeRFMTLoss(dict(HP))
    # configure_optimizers (train_nerf.py:262-291): two groups, eps 1e-15; apex FusedAdam is absent -> torch's fused Adam
    enc = [p for n, p in model.named_parameters() if "xyz_encoder" in n and p.numel()]
    net = [p for n, p in model.named_parameters() if "xyz_encoder" not in n and p.numel()]
    opt = torch.optim.Adam([{"params": enc, "weight_decay": 0.0}, {"params": net, "weight_decay": 1e-6}], lr=1e-2, eps=1e-15, fused=True)
    scaler = torch.amp.GradScaler("cuda")
    batches = []
    for i in range(4):
        b = synth.patch_batch(n_rays, seed=1000 + i)
        tri = torch.from_numpy(b["tri"]).to(dev)
        batches.append((torch.from_numpy(b["rays_o"]).to(dev), torch.from_numpy(b["rays_d"]).to(dev),
                        {"rgb": torch.rand(n_rays, 3, device=dev), "patch_area": 64, "x1_offsets_local": tri[0][:49] % 64,
                         "x2_offsets_local": tri[1][:49] % 64, "x3_offsets_local": tri[2][:49] % 64}))
    kw = dict(near_distance=0.01, max_samples=1024, exp_step_factor=0.0, n_sem_cls=0, pred_norm_nn_norm=False)
    n_samples = 0

    def one(i):
        nonlocal n_samples
        ro, rd, tgt = batches[i % 4]
        with torch.autocast("cuda", dtype=torch.float16), contextlib.redirect_stdout(io.StringIO()):
            res = ns.rendering.render(model, ro, rd, global_step=3000 + i, **kw)
            loss_d = loss_fn(res, tgt, global_step=3000 + i)
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss_d["total"]).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(enc + net, 0.05)
        scaler.step(opt); scaler.update()
        n_samples = int(res["rm_samples"])
        return loss_d

    for i in range(warmup):
        one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        loss_d = one(warmup + i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return {"what": "the reference's own render() + NeRFMTLoss + autograd (unchanged files) + GradScaler / clip_grad_norm_ / fused Adam, bound to "
                    + ("the reference's own csrc CUDA kernels (vren_ref, compiled for sm_100)" if vren == "ref" else "libncn through the drop-in shims")
                    + "; tiny-cuda-nn = the libncn shim (tcnn's source is not available), k-means on the CPU as the reference does (gpu=False), "
                      "occupancy-grid update excluded",
            "ms_per_step": dt * 1e3, "rays_per_s": n_rays / dt, "steps": steps, "rays": n_rays, "samples_per_step": n_samples,
            "loss_total": float(loss_d["total"])}


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["losses", "kernels", "step"]
    res = {}
    if "losses" in which:
        res["ref_losses_cpu"] = ref_losses_cpu()
    if "kernels" in which:
        res["ref_kernels_gpu"] = ref_kernels_gpu()
    if "step" in which:
        res["ref_step_gpu_csrc"] = ref_step_gpu(vren="ref")
        res["ref_step_gpu_shims"] = ref_step_gpu(vren="shim")
    txt = json.dumps(res, indent=1)
    if "--json" in sys.argv:
        open(sys.argv[sys.argv.index("--json") + 1], "w").write(txt)
    print(txt)
