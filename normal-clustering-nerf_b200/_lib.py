"""ctypes binding of libncn.so - the C-ABI declared in include/ncn.h.

There is no fallback: if the shared library is missing this module raises at import of
the first symbol (``lib()``), and every non-zero return code becomes a RuntimeError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NCN_LIB_PATH") or os.path.join(_HERE, "libncn.so")      # override: developer A/B builds only
_lib = None

c_i64, c_i32, c_f32, c_vp, c_sz = C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); must list every function include/ncn.h declares.
SIGNATURES = {
    "ncn_version": (c_i32, []),
    "ncn_error_string": (C.c_char_p, [c_i32]),
    "ncn_device_info": (c_i32, [c_vp, c_vp, c_vp]),
    "ncn_ray_aabb_intersect": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "ncn_ray_sphere_intersect": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "ncn_ray_aabb_near": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_f32, c_i64, c_vp, c_vp]),
    "ncn_morton3d": (c_i32, [c_vp, c_i64, c_vp, c_vp]),
    "ncn_morton3d_invert": (c_i32, [c_vp, c_i64, c_vp, c_vp]),
    "ncn_packbits": (c_i32, [c_vp, c_i64, c_f32, c_vp, c_vp]),
    "ncn_density_grid_update": (c_i32, [c_vp, c_vp, c_i64, c_f32, c_vp, c_vp]),
    "ncn_packbits_auto": (c_i32, [c_vp, c_i64, c_vp, c_f32, c_vp, c_vp]),
    "ncn_march_train_workspace_bytes": (c_sz, [c_i64, c_i32]),
    "ncn_march_train": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_f32, c_f32, c_vp, c_i32, c_i32, c_i64, c_i64,
                                c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "ncn_march_train_count": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_f32, c_f32, c_vp, c_i32, c_i32, c_i64,
                                      c_vp, c_vp, c_vp, c_sz, c_vp]),
    "ncn_march_train_expand": (c_i32, [c_vp, c_vp, c_vp, c_f32, c_f32, c_i32, c_i32, c_i64, c_i64,
                                       c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "ncn_march_test": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_f32, c_f32, c_i32, c_i32, c_i32, c_i64,
                               c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_composite_train_fw": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_i64, c_i64, c_i32,
                                       c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_composite_train_fw_photometric": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                                                   c_vp, c_vp, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_composite_train_fw_photometric_gt": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                                                      c_vp, c_i64, c_vp, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_composite_train_bw": (c_i32, [c_vp] * 13 + [c_f32, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "ncn_composite_test_fw": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_i64, c_i32, c_i32,
                                      c_vp, c_vp, c_vp, c_vp]),
    "ncn_distortion_fw": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "ncn_distortion_bw": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "ncn_segment_csr_sum": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "ncn_set_composite_width": (c_i32, [c_i32]),
    "ncn_set_field_fwd_impl": (c_i32, [c_i32]),
    "ncn_set_pdl": (c_i32, [c_i32]),
    "ncn_set_grid_bwd_occupancy": (c_i32, [c_i32]),
}


class Profiler:
    """Optional launch counter / per-call CUDA-event timer (used by bench.py; off by default).
    launches: number of libncn kernels launched (every C call is counted with the kernels it enqueues)."""
    counting = False
    timing = None          # None, or a set of call names to time ("*" = all)
    launches = 0
    events = []            # (name, start_event, end_event, n_items)
    KERNELS_PER_CALL = {"ncn_march_train": 3, "ncn_march_train_count": 3, "ncn_cluster_tail": 2, "ncn_kmeans_workspace_bytes": 0,
                        "ncn_march_train_workspace_bytes": 0, "ncn_mlp_bwd_workspace_bytes": 0, "ncn_mlp_acts_bytes": 0, "ncn_mlp_n_params": 0,
                        "ncn_grid_desc_init": 0, "ncn_version": 0, "ncn_set_mlp_bwd_impl": 0, "ncn_set_march_segments": 0, "ncn_set_composite_width": 0, "ncn_set_field_fwd_impl": 0, "ncn_set_pdl": 0, "ncn_set_grid_bwd_occupancy": 0, "ncn_debug_stamp": 1, "ncn_set_grid_bwd_merge": 0, "ncn_set_grid_fwd_coherent": 0, "ncn_error_string": 0, "ncn_device_info": 0,
                        "ncn_comm_unique_id": 0, "ncn_comm_init": 0, "ncn_comm_destroy": 0, "ncn_comm_last_error": 0,
                        "ncn_sample_ray_batch": 2, "ncn_sample_ray_batch_ex": 2, "ncn_peer_create": 0, "ncn_peer_grad": 0, "ncn_peer_p16": 0, "ncn_peer_handles": 0, "ncn_peer_connect": 0,
                        "ncn_peer_shard": 0, "ncn_peer_step": 2, "ncn_peer_error": 0, "ncn_peer_destroy": 0,
                        "ncn_peer_poll": 0, "ncn_peer_set_timeout": 0, "ncn_graph_node_counts": 0, "ncn_peer_debug_times": 0, "ncn_peer_set_external_zero": 0, "ncn_peer_set_loads": 0, "ncn_peer_set_shape": 0,
                        "ncn_peer_set_cut": 0, "ncn_peer_set_early_loads": 0, "ncn_peer_segments": 0, "ncn_peer_segments_of": 0, "ncn_peer_debug_times_early": 0}

    @classmethod
    def reset(cls):
        cls.launches = 0
        cls.events = []

    @classmethod
    def summary(cls):
        """name -> (calls, total_ms) from the recorded events (synchronises)."""
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in cls.events:
            c, t = out.get(name, (0, 0.0))
            out[name] = (c + 1, t + e0.elapsed_time(e1))
        return out


class _Proxy:
    """Attribute access returns the ctypes function, wrapped with counting / event timing when enabled."""

    def __init__(self, h):
        self._h = h

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if not Profiler.counting and Profiler.timing is None:
            return fn
        k = Profiler.KERNELS_PER_CALL.get(name, 1)

        def wrapped(*args):
            import torch
            timed = Profiler.timing is not None and ("*" in Profiler.timing or name in Profiler.timing) and k > 0
            if timed:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = fn(*args)
            if timed:
                e1.record()
                Profiler.events.append((name, e0, e1))
            if Profiler.counting:
                kk = k
                if name == "ncn_mlp_bwd":      # dgrad + one wgrad per layer (when a weight gradient is requested)
                    kk = 1 + (args[0]._obj.n_hidden + 1 if args[7] else 0)
                Profiler.launches += kk
            return rc
        return wrapped


def lib():
    """Load libncn.so once; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"libncn.so not found at {LIB_PATH}: build it with "
                "`python normal-clustering-nerf_b200/csrc/build.py` (there is no CPU / torch fallback)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)   # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = _Proxy(h)
        if os.environ.get("NCN_PDL", "1") == "0":        # developer A/B knobs
            h.ncn_set_pdl(0)
        if os.environ.get("NCN_GRID_BWD_OCC"):
            h.ncn_set_grid_bwd_occupancy(int(os.environ["NCN_GRID_BWD_OCC"]))
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ncn_error_string(int(rc)).decode()
        raise RuntimeError(f"{what}: {msg} (code {rc})" if what else f"{msg} (code {rc})")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    """raw handle of torch's current CUDA stream (the C-level getter: the python Stream object costs ~2 us per call)"""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


# ----------------------------------------------------------------------------- structs (include/ncn.h)
NCN_GRID_MAX_LEVELS = 32


class GridDesc(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("n_features", C.c_int32), ("log2_hashmap_size", C.c_int32),
                ("base_resolution", C.c_int32), ("per_level_scale", C.c_float),
                ("level_scale", C.c_float * NCN_GRID_MAX_LEVELS), ("level_res", C.c_uint32 * NCN_GRID_MAX_LEVELS),
                ("level_size", C.c_uint32 * NCN_GRID_MAX_LEVELS),
                ("level_offset", C.c_uint32 * (NCN_GRID_MAX_LEVELS + 1))]


class MlpDesc(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_out", C.c_int32), ("n_hidden", C.c_int32), ("width", C.c_int32),
                ("activation", C.c_int32), ("out_activation", C.c_int32)]


class MlpBwdSrc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("d_raws", C.c_void_p), ("c_total", C.c_int32), ("c_off", C.c_int32), ("n_ch", C.c_int32),
                ("dx_rgb", C.c_void_p), ("d_sigmas", C.c_void_p), ("h", C.c_void_p), ("scale", C.c_float), ("perm", C.c_int32),
                ("dx_extra", C.c_void_p)]


class AdamGroups(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("reserved", C.c_int32), ("start", C.c_int64 * 4), ("weight_decay", C.c_float * 4),
                ("max_norm", C.c_float), ("reserved2", C.c_float)]


ACT = {"None": 0, "ReLU": 1, "Sigmoid": 2, "Exponential": 3}

SIGNATURES.update({
    "ncn_grid_desc_init": (c_i64, [C.POINTER(GridDesc)]),
    "ncn_grid_fwd": (c_i32, [C.POINTER(GridDesc), c_vp, c_vp, c_i64, c_vp, C.POINTER(c_f32), c_vp, c_vp]),
    "ncn_grid_bwd": (c_i32, [C.POINTER(GridDesc), c_vp, c_vp, c_i64, c_vp, c_f32, C.POINTER(c_f32), c_vp, c_vp]),
    "ncn_grid_bwd_f16": (c_i32, [C.POINTER(GridDesc), c_vp, c_vp, c_i64, c_vp, c_f32, C.POINTER(c_f32), c_vp, c_vp]),
    "ncn_grid_bwd_levels": (c_i32, [C.POINTER(GridDesc), c_vp, c_vp, c_i64, c_vp, c_f32, C.POINTER(c_f32), c_vp, c_i32, c_i32, c_i32, c_vp]),
    "ncn_grid_bwd_input": (c_i32, [C.POINTER(GridDesc), c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "ncn_grid_bwd_bwd_input": (c_i32, [C.POINTER(GridDesc), c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "ncn_mlp_n_params": (c_i64, [C.POINTER(MlpDesc)]),
    "ncn_mlp_bwd_workspace_bytes": (c_sz, [C.POINTER(MlpDesc), c_i64]),
    "ncn_mlp_acts_bytes": (c_sz, [C.POINTER(MlpDesc), c_i64]),
    "ncn_mlp_fwd": (c_i32, [C.POINTER(MlpDesc), c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "ncn_mlp_bwd": (c_i32, [C.POINTER(MlpDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_f32, c_vp, c_sz, c_vp, c_vp]),
})


class KmeansParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("niter", C.c_int32), ("seed", C.c_int32),
                ("max_points_per_centroid", C.c_int32), ("spherical", C.c_int32)]


SIGNATURES.update({
    "ncn_normals_from_depth_fw": (c_i32, [c_vp] * 6 + [c_i64, c_vp, c_vp]),
    "ncn_normals_from_depth_bw": (c_i32, [c_vp] * 7 + [c_i64, c_vp, c_vp]),
    "ncn_kmeans_workspace_bytes": (c_sz, [c_i64, c_i32]),
    "ncn_kmeans_spherical": (c_i32, [c_vp, c_i64, C.POINTER(KmeansParams), c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "ncn_cluster_bw_depth": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_cluster_chain": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, C.POINTER(KmeansParams), c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                  c_vp, c_sz, c_vp]),
    "ncn_cluster_select": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_f32, c_vp, c_vp, c_vp]),
    "ncn_cluster_loss_fw": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "ncn_cluster_loss_bw": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "ncn_cluster_tail": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_photometric_loss": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, C.POINTER(c_f32), c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_photometric_loss_gt": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, C.POINTER(c_f32), c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_sample_ray_batch": (c_i32, [c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "ncn_sample_ray_batch_ex": (c_i32, [c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "ncn_sample_random_pose_half": (c_i32, [c_i32, c_vp, c_i32, c_i32, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "ncn_gather_pixels": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp]),
    "ncn_normals_from_depth_image": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "ncn_semantic_ce_loss": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_i64, c_f32, c_vp, c_vp, c_vp]),
    "ncn_adam_step": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_adam_step_groups": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, C.POINTER(AdamGroups), c_f32, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_debug_stamp": (c_i32, [c_vp, c_i32, c_vp]),
    "ncn_step_guard": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "ncn_graph_node_counts": (c_i32, [c_vp, c_vp]),
    "ncn_grid_sample_cells": (c_i32, [c_vp, c_i32, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "ncn_grid_scatter_density": (c_i32, [c_vp, c_i32, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "ncn_field_mlp_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_set_march_segments": (c_i32, [c_i32]),
    "ncn_set_mlp_bwd_impl": (c_i32, [c_i32]),
    "ncn_mlp_bwd_src_fused": (c_i32, [C.POINTER(MlpDesc), C.POINTER(MlpBwdSrc), c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_f32, c_vp, c_sz, c_vp, c_vp]),
    "ncn_set_grid_bwd_merge": (c_i32, [c_i32]),
    "ncn_set_grid_fwd_coherent": (c_i32, [c_i32]),
    "ncn_field_fwd": (c_i32, [C.POINTER(GridDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, C.POINTER(c_f32), c_vp, c_vp, c_i32,
                              c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_rays_from_pixels": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "ncn_field_prepare_rgb": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "ncn_field_heads_fwd": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "ncn_field_head_out": (c_i32, [c_vp, c_i32, c_i64, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "ncn_field_head_dout": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_f32, c_i64, c_vp, c_vp, c_i32, c_vp]),
    "ncn_field_bwd_h": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_i64, c_vp, c_vp, c_vp]),
    "ncn_grad_sumsq": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "ncn_clip_coef": (c_i32, [c_vp, c_f32, c_vp, c_vp]),
    "ncn_comm_unique_id": (c_i32, [c_vp]),
    "ncn_comm_init": (c_i32, [C.POINTER(c_vp), c_vp, c_i32, c_i32]),
    "ncn_comm_allreduce_sum_f32": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "ncn_comm_destroy": (c_i32, [c_vp]),
    "ncn_comm_last_error": (C.c_char_p, []),
    "ncn_peer_create": (c_i32, [C.POINTER(c_vp), c_i32, c_i32, c_i64]),
    "ncn_peer_grad": (c_vp, [c_vp]),
    "ncn_peer_p16": (c_vp, [c_vp]),
    "ncn_peer_handles": (c_i32, [c_vp, c_vp]),
    "ncn_peer_connect": (c_i32, [c_vp, c_vp]),
    "ncn_peer_set_cut": (c_i32, [c_vp, c_i64]),
    "ncn_peer_set_early_loads": (c_i32, [c_i32]),
    "ncn_peer_segments": (c_i32, [c_vp, c_i32, C.POINTER(c_i64)]),
    "ncn_peer_segments_of": (None, [c_i64, c_i64, c_i32, c_i32, C.POINTER(c_i64)]),
    "ncn_peer_early": (c_i32, [c_vp, c_vp, c_vp]),
    "ncn_peer_debug_times_early": (c_i32, [c_vp, C.POINTER(C.c_ulonglong)]),
    "ncn_peer_shard": (None, [c_i64, c_i32, c_i32, C.POINTER(c_i64), C.POINTER(c_i64)]),
    "ncn_peer_step": (c_i32, [c_vp, c_vp, c_vp, c_vp, C.POINTER(AdamGroups), c_f32, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ncn_peer_error": (c_i32, [c_vp, C.POINTER(C.c_uint32)]),
    "ncn_peer_poll": (C.c_uint32, [c_vp]),
    "ncn_peer_set_loads": (c_i32, [c_i32]),
    "ncn_peer_set_shape": (c_i32, [c_i32, c_i32]),
    "ncn_peer_set_external_zero": (c_i32, [c_vp, c_i32]),
    "ncn_peer_debug_times": (c_i32, [c_vp, c_vp]),
    "ncn_peer_set_timeout": (c_i32, [C.c_double]),
    "ncn_peer_destroy": (c_i32, [c_vp]),
})
