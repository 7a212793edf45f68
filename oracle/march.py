"""TEST INFRASTRUCTURE ONLY.  ctypes front-end of oracle/march_ref.c (numpy in / numpy out)."""
import ctypes as C

import numpy as np

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build())
        _lib.ncn_oracle_march_train.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def aabb(rays_o, rays_d, center, half, near=-1.0):
    R = len(rays_o)
    out = np.empty((R, 2), np.float32)
    lib().ncn_oracle_aabb(_p(rays_o), _p(rays_d), _p(np.ascontiguousarray(center, np.float32)),
                          _p(np.ascontiguousarray(half, np.float32)), C.c_float(near), C.c_int64(R), _p(out))
    return out


def march_train(rays_o, rays_d, hits_t, bitfield, cascades, scale, esf, noise, grid_size, max_samples):
    R = len(rays_o)
    rays_a = np.empty((R, 3), np.int64)
    args = (_p(rays_o), _p(rays_d), _p(hits_t), _p(bitfield), C.c_int(cascades), C.c_float(scale), C.c_float(esf), _p(noise),
            C.c_int(grid_size), C.c_int(max_samples), C.c_int64(R))
    n = lib().ncn_oracle_march_train(*args, C.c_int64(0), _p(rays_a), None, None, None, None)
    xyzs = np.empty((n, 3), np.float32); dirs = np.empty((n, 3), np.float32)
    deltas = np.empty(n, np.float32); ts = np.empty(n, np.float32)
    lib().ncn_oracle_march_train(*args, C.c_int64(n), _p(rays_a), _p(xyzs), _p(dirs), _p(deltas), _p(ts))
    return rays_a, xyzs, dirs, deltas, ts


def packbits(grid, thr):
    grid = np.ascontiguousarray(grid.reshape(-1), np.float32)
    out = np.empty(len(grid) // 8, np.uint8)
    lib().ncn_oracle_packbits(_p(grid), C.c_int64(len(out)), C.c_float(thr), _p(out))
    return out


def morton3d(coords):
    coords = np.ascontiguousarray(coords, np.int32)
    out = np.empty(len(coords), np.int32)
    lib().ncn_oracle_morton3d(_p(coords), C.c_int64(len(coords)), _p(out))
    return out
