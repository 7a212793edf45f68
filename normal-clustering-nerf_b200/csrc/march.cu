// Occupancy-grid ray marching.  Replaces raymarching.cu:166-453 of the reference
// (vren.raymarching_train / vren.raymarching_test).
//
// Bit-exactness.  The DDA below reproduces the reference's fp32 rounding sequence as it
// is actually compiled (nvcc -O2, default -fmad=true; SASS of the reference build
// inspected, see DESIGN.md "marching arithmetic"): every operation is an explicit
// round-to-nearest intrinsic so neither nvcc nor ptxas may re-associate or (un)fuse it:
//   x      = fma(d, t, o)
//   dt     = max(dt_min, min(t*exp_step_factor, dt_max)),  dt_min = sqrt3/max_samples (IEEE div),
//            dt_max = (scale*2sqrt3)/grid_size (IEEE div)
//   cell   = trunc(max(0, min(((fma(x, 1/mip_bound, 1))*0.5)*G, G-1)))
//   t_face = (fma(mip_bound, fma(((cell+0.5) + 0.5*sign(d))*(1/G), 2, -1), -x)) * (1/d)
//   skip   : t_target = t + max(0, min3(tx,ty,tz));  do t += dt(t) while (t < t_target)
//
// B200 design.  The reference marches every ray twice (count pass, then an atomicAdd for
// the output slot, then a write pass) into 268 MB of zero-filled worst-case buffers.
// Here a ray is marched ONCE: the thread records the t of each emitted sample in an
// L2-resident scratch slab, a scan turns the per-ray counts into start offsets in RAY
// ORDER (deterministic; one of the layouts the reference's atomics can produce), and a
// warp-per-ray expansion writes the exact-size sample arrays with coalesced stores.
// The DDA is an inherently serial, divergent fp32 chain (latency bound); the expansion is
// a pure streaming pass (32 B/sample).
#include "ncn_common.cuh"
#include "morton.cuh"

namespace ncn {

struct MarchCfg {
  int cascades, grid_size, max_samples;
  int const_dt;            // exp_step_factor == 0: calc_dt(t) == dt_min for every finite t (t*0 = +-0)
  uint32_t grid_size3;
  float scale;             // mip_bound cap (train: scale; test: scale)
  float exp_step_factor;
  float dt_min, dt_max;    // calc_dt clamp bounds
  float gs_f, gs_inv, gs_m1;
};

// calc_dt(t) - raymarching.cu:11-13 : clamp(t*f, lo, hi) = fmaxf(lo, fminf(t*f, hi))
__device__ __forceinline__ float calc_dt(float t, const MarchCfg& c) {
  if (c.const_dt) return c.dt_min;   // bit-identical: max(dt_min, min(+-0, dt_max)) with dt_max >= dt_min > 0
  return fmaxf(c.dt_min, fminf(__fmul_rn(t, c.exp_step_factor), c.dt_max));
}

// raymarching.cu:19-23 / 29-32
__device__ __forceinline__ int mip_from_pos(float x, float y, float z, int cascades) {
  const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
  int e; frexpf(mx, &e);
  return min(cascades - 1, max(0, e + 1));
}
__device__ __forceinline__ int mip_from_dt(float dt, float gs_f, int cascades) {
  int e; frexpf(__fmul_rn(dt, gs_f), &e);
  return min(cascades - 1, max(0, e));
}

struct Cell { int nx, ny, nz; float mip_bound; bool occ; };

// kSimple = (cascades == 1 && exp_step_factor == 0): dt, the mip bound and its reciprocal are per-launch constants, so the
// serial per-ray chain sheds the calc_dt / frexp / IEEE-reciprocal work (the values are bit-identical, see const_dt)
struct SimpleCfg { float dt, mip_bound, mb_inv; };

__device__ __forceinline__ int cell_coord(float x, float mb_inv, const MarchCfg& c) {
  const float v = __fmul_rn(__fmul_rn(__fmaf_rn(x, mb_inv, 1.0f), 0.5f), c.gs_f);
  return (int)fmaxf(0.0f, fminf(v, c.gs_m1));   // clamp(v, 0, G-1) then float->int truncation
}

__device__ __forceinline__ Cell locate(float x, float y, float z, float dt, const MarchCfg& c,
                                       const uint8_t* __restrict__ bitfield) {
  Cell k;
  int mip = 0;
  if (c.cascades > 1) mip = max(mip_from_pos(x, y, z, c.cascades), mip_from_dt(dt, c.gs_f, c.cascades));
  // scalbnf(1, mip-1): exact power of two
  const float p2 = __int_as_float((126 + mip) << 23);
  k.mip_bound = fminf(p2, c.scale);
  // IEEE reciprocal (same value every iteration when cascades == 1: the compiler hoists it out of the loop)
  const float mb_inv = __frcp_rn(k.mip_bound);
  k.nx = cell_coord(x, mb_inv, c); k.ny = cell_coord(y, mb_inv, c); k.nz = cell_coord(z, mb_inv, c);
  const uint32_t idx = (uint32_t)mip * c.grid_size3 + morton3d((uint32_t)k.nx, (uint32_t)k.ny, (uint32_t)k.nz);
  k.occ = (__ldg(bitfield + (idx >> 3)) >> (idx & 7u)) & 1u;
  return k;
}

__device__ __forceinline__ float face_t(int n, float sgn_half, float mip_bound, float x, float d_inv,
                                        const MarchCfg& c) {
  // (((n+0.5f+0.5f*sign)*grid_size_inv*2-1)*mip_bound-x)*d_inv
  const float a = __fadd_rn(__fadd_rn((float)n, 0.5f), sgn_half);
  const float b = __fmaf_rn(__fmul_rn(a, c.gs_inv), 2.0f, -1.0f);
  return __fmul_rn(__fmaf_rn(mip_bound, b, -x), d_inv);
}

// advance t past the current (empty) cell in whole dt steps
__device__ __forceinline__ float skip_cell(float t, const Cell& k, float x, float y, float z,
                                           float sx, float sy, float sz, float ix, float iy, float iz,
                                           const MarchCfg& c) {
  const float tx = face_t(k.nx, sx, k.mip_bound, x, ix, c);
  const float ty = face_t(k.ny, sy, k.mip_bound, y, iy, c);
  const float tz = face_t(k.nz, sz, k.mip_bound, z, iz, c);
  const float t_target = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
  do { t = __fadd_rn(t, calc_dt(t, c)); } while (t < t_target);
  return t;
}

// ---- train: pass 1 (one thread per ray): march, record sample ts, count -------------------
// The DDA is a serial, divergent fp32 chain: a warp costs the UNION of its lanes' paths and there is nothing to hide
// latency with.  Only kRaysPerWarp lanes of each warp carry a ray, which shrinks the union and gives every SM
// scheduler several warps to interleave (8192 rays -> 1024 warps instead of 256).
constexpr int kRaysPerWarp = 8;

template <bool kSimple>
__global__ void __launch_bounds__(64)
march_train_count_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                         const float* __restrict__ hits_t, const uint8_t* __restrict__ bitfield,
                         const float* __restrict__ noise, MarchCfg c, int64_t n_rays,
                         int32_t* __restrict__ counts, int32_t* __restrict__ segcnt, float* __restrict__ ts_scratch, int slab_stride) {
  const int lane = threadIdx.x & 31;
  if (lane >= kRaysPerWarp) return;
  const int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kRaysPerWarp + lane;
  if (r >= n_rays) return;
  SimpleCfg sc;
  sc.dt = c.dt_min; sc.mip_bound = fminf(0.5f, c.scale); sc.mb_inv = __frcp_rn(sc.mip_bound);
  const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
  const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
  const float ix = __frcp_rn(dx), iy = __frcp_rn(dy), iz = __frcp_rn(dz);
  const float sx = copysignf(0.5f, dx), sy = copysignf(0.5f, dy), sz = copysignf(0.5f, dz);
  const float2 h = reinterpret_cast<const float2*>(hits_t)[r];
  float t = h.x;
  const float t2 = h.y;
  if (t >= 0.0f && noise != nullptr) t = __fmaf_rn(calc_dt(t, c), noise[r], t);   // raymarching.cu:195-198
  float* out = ts_scratch + r * (int64_t)slab_stride;
  int n = 0;
  if (kSimple) {
    const int max_samples = c.max_samples;
    const float gs_f = c.gs_f, gs_m1 = c.gs_m1, gs_inv = c.gs_inv;
    while (0.0f <= t && t < t2 && n < max_samples) {
      const float x = __fmaf_rn(dx, t, ox), y = __fmaf_rn(dy, t, oy), z = __fmaf_rn(dz, t, oz);
      const int nx = (int)fmaxf(0.0f, fminf(__fmul_rn(__fmul_rn(__fmaf_rn(x, sc.mb_inv, 1.0f), 0.5f), gs_f), gs_m1));
      const int ny = (int)fmaxf(0.0f, fminf(__fmul_rn(__fmul_rn(__fmaf_rn(y, sc.mb_inv, 1.0f), 0.5f), gs_f), gs_m1));
      const int nz = (int)fmaxf(0.0f, fminf(__fmul_rn(__fmul_rn(__fmaf_rn(z, sc.mb_inv, 1.0f), 0.5f), gs_f), gs_m1));
      const uint32_t idx = morton3d((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
      const bool occ = (__ldg(bitfield + (idx >> 3)) >> (idx & 7u)) & 1u;
      if (occ) {
        out[n++] = t;
        t = __fadd_rn(t, sc.dt);
      } else {
        const float tx = __fmul_rn(__fmaf_rn(sc.mip_bound, __fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)nx, 0.5f), sx), gs_inv), 2.0f, -1.0f), -x), ix);
        const float ty = __fmul_rn(__fmaf_rn(sc.mip_bound, __fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)ny, 0.5f), sy), gs_inv), 2.0f, -1.0f), -y), iy);
        const float tz = __fmul_rn(__fmaf_rn(sc.mip_bound, __fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)nz, 0.5f), sz), gs_inv), 2.0f, -1.0f), -z), iz);
        const float t_target = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
        do { t = __fadd_rn(t, sc.dt); } while (t < t_target);
      }
    }
  } else {
    while (0.0f <= t && t < t2 && n < c.max_samples) {
      const float x = __fmaf_rn(dx, t, ox), y = __fmaf_rn(dy, t, oy), z = __fmaf_rn(dz, t, oz);
      const float dt = calc_dt(t, c);
      const Cell k = locate(x, y, z, dt, c, bitfield);
      if (k.occ) {
        out[n++] = t;
        t = __fadd_rn(t, dt);
      } else {
        t = skip_cell(t, k, x, y, z, sx, sy, sz, ix, iy, iz, c);
      }
    }
  }
  counts[r] = n;
  reinterpret_cast<int4*>(segcnt)[r] = make_int4(n, 0, 0, 0);
}

// ---- train: pass 1, constant-dt fast path: FOUR lanes per ray, each marching a quarter of the candidate sequence -----
// With exp_step_factor == 0 every march visits a subset of ONE fixed sequence of candidate times T_0 = t_start,
// T_{k+1} = fl(T_k + dt) - an occupied cell emits the candidate and moves to the next one, an empty cell skips forward to
// the first candidate at or beyond the cell's exit - so the sequence itself does not depend on the occupancy grid.  Lane s
// of a ray starts at candidate s*Q (Q = max_samples/4; it gets there with s*Q dependent additions, a few clocks each) and
// marches until it passes candidate (s+1)*Q, writing its samples to its own region of the ray's slab.  The speculation
// "candidate s*Q is visited by the real march" is then checked segment by segment: the previous segment must have landed
// exactly on s*Q, or - the common case in empty space - on the candidate this segment's first (empty-cell) skip reached;
// otherwise the lane re-marches its segment from the true landing point.  Every emitted t is the value the serial march
// produces (same chain, same cell tests, same skip targets), so the result stays bit-identical to the reference while the
// serial dependent chain per lane is 4x shorter.
constexpr int kSegs = 4;
constexpr int kSegPad = 16;        // slack behind each segment region (a segment holds at most Q (+1 for the last) samples)

struct SegWalk { float t; int idx; int n; float v1; bool first_occ; bool any; };

__device__ __forceinline__ void march_segment(SegWalk& w, bool go, int idx_end, int cap, float t2, float ox, float oy, float oz,
                                              float dx, float dy, float dz, float ix, float iy, float iz, float sx, float sy,
                                              float sz, float dt, float mip_bound, float mb_inv, float gs_f, float gs_m1,
                                              float gs_inv, const uint8_t* __restrict__ bitfield, float* __restrict__ out) {
  float t = w.t;
  int idx = w.idx, n = 0;
  bool first = true;
  w.first_occ = false; w.any = false; w.v1 = t;
  while (go && t < t2 && idx < idx_end && n < cap) {
    const float x = __fmaf_rn(dx, t, ox), y = __fmaf_rn(dy, t, oy), z = __fmaf_rn(dz, t, oz);
    const int nx = (int)fmaxf(0.0f, fminf(__fmul_rn(__fmul_rn(__fmaf_rn(x, mb_inv, 1.0f), 0.5f), gs_f), gs_m1));
    const int ny = (int)fmaxf(0.0f, fminf(__fmul_rn(__fmul_rn(__fmaf_rn(y, mb_inv, 1.0f), 0.5f), gs_f), gs_m1));
    const int nz = (int)fmaxf(0.0f, fminf(__fmul_rn(__fmul_rn(__fmaf_rn(z, mb_inv, 1.0f), 0.5f), gs_f), gs_m1));
    const uint32_t cell = morton3d((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
    const bool occ = (__ldg(bitfield + (cell >> 3)) >> (cell & 7u)) & 1u;
    if (occ) {
      out[n++] = t;
      t = __fadd_rn(t, dt);
      ++idx;
    } else {
      const float tx = __fmul_rn(__fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)nx, 0.5f), sx), gs_inv), 2.0f, -1.0f), -x), ix);
      const float ty = __fmul_rn(__fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)ny, 0.5f), sy), gs_inv), 2.0f, -1.0f), -y), iy);
      const float tz = __fmul_rn(__fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)nz, 0.5f), sz), gs_inv), 2.0f, -1.0f), -z), iz);
      const float t_target = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
      do { t = __fadd_rn(t, dt); ++idx; } while (t < t_target);
    }
    if (first) { first = false; w.first_occ = occ; w.v1 = t; w.any = true; }
  }
  w.t = t; w.idx = idx; w.n = n;
}

__global__ void __launch_bounds__(64)
march_train_count_seg_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                             const float* __restrict__ hits_t, const uint8_t* __restrict__ bitfield,
                             const float* __restrict__ noise, MarchCfg c, int64_t n_rays, int32_t* __restrict__ counts,
                             int32_t* __restrict__ segcnt, float* __restrict__ slabs, int slab_stride, int seg_stride, int force_redo) {
  const int lane = threadIdx.x & 31;
  const int seg = lane & (kSegs - 1);
  const int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / kSegs) + (lane >> 2);
  const bool have_ray = r < n_rays;
  const int64_t rr = have_ray ? r : 0;
  const float dt = c.dt_min, mip_bound = fminf(0.5f, c.scale), mb_inv = __frcp_rn(mip_bound);
  const float gs_f = c.gs_f, gs_m1 = c.gs_m1, gs_inv = c.gs_inv;
  const float ox = rays_o[3 * rr], oy = rays_o[3 * rr + 1], oz = rays_o[3 * rr + 2];
  const float dx = rays_d[3 * rr], dy = rays_d[3 * rr + 1], dz = rays_d[3 * rr + 2];
  const float ix = __frcp_rn(dx), iy = __frcp_rn(dy), iz = __frcp_rn(dz);
  const float sx = copysignf(0.5f, dx), sy = copysignf(0.5f, dy), sz = copysignf(0.5f, dz);
  const float2 h = reinterpret_cast<const float2*>(hits_t)[rr];
  float t0 = h.x;
  const float t2 = h.y;
  if (t0 >= 0.0f && noise != nullptr) t0 = __fmaf_rn(dt, noise[rr], t0);            // raymarching.cu:195-198
  const bool ray_live = have_ray && 0.0f <= t0;
  const int Q = c.max_samples / kSegs;
  const int idx_begin = seg * Q, idx_end = seg == kSegs - 1 ? 0x7fffffff : (seg + 1) * Q;
  const int cap = seg == kSegs - 1 ? Q + kSegPad : Q;
  float* out = slabs + rr * (int64_t)slab_stride + seg * seg_stride;
  // candidate time at the start of this lane's segment: idx_begin dependent additions
  float ts = t0;
#pragma unroll 8
  for (int i = 0; i < idx_begin; ++i) ts = __fadd_rn(ts, dt);
  const float start_t = ts;
  SegWalk w;
  w.t = start_t; w.idx = idx_begin;
  march_segment(w, ray_live, idx_end, cap, t2, ox, oy, oz, dx, dy, dz, ix, iy, iz, sx, sy, sz, dt, mip_bound, mb_inv, gs_f, gs_m1, gs_inv,
                bitfield, out);
  // ---- stitch: validate (or re-march) segments 1..3 in order.  All lanes run the loop; shuffles stay inside the 4-lane group
  const int grp = lane & ~(kSegs - 1);
#pragma unroll
  for (int s = 1; s < kSegs; ++s) {
    const float L_t = __shfl_sync(0xffffffffu, w.t, grp + s - 1);
    const int L_idx = __shfl_sync(0xffffffffu, w.idx, grp + s - 1);
    const bool mine = seg == s;
    const bool over = !(L_t < t2);                                  // the ray ended inside an earlier segment
    bool redo = false;
    if (mine && ray_live) {
      if (over) { w.n = 0; w.t = L_t; w.idx = L_idx; }
      else if (force_redo) { redo = true; w.t = L_t; w.idx = L_idx; }                         // test mode: always repair
      else if (L_idx == idx_begin) { /* landed exactly on this segment's first candidate */ }
      else if (w.any && !w.first_occ && L_t == w.v1) { /* both skipped to the same candidate out of the shared empty cell */ }
      else { redo = true; w.t = L_t; w.idx = L_idx; }
    }
    if (__any_sync(0xffffffffu, redo)) {
      SegWalk w2 = w;
      march_segment(w2, redo, idx_end, cap, t2, ox, oy, oz, dx, dy, dz, ix, iy, iz, sx, sy, sz, dt, mip_bound, mb_inv, gs_f, gs_m1,
                    gs_inv, bitfield, out);
      if (redo) w = w2;
    }
  }
  // ---- counts (the reference stops at max_samples samples: trim from the back)
  int n = ray_live ? w.n : 0;
  int n0 = __shfl_sync(0xffffffffu, n, grp), n1 = __shfl_sync(0xffffffffu, n, grp + 1), n2 = __shfl_sync(0xffffffffu, n, grp + 2),
      n3 = __shfl_sync(0xffffffffu, n, grp + 3);
  const int m = c.max_samples;
  n0 = min(n0, m); n1 = min(n1, m - n0); n2 = min(n2, m - n0 - n1); n3 = min(n3, m - n0 - n1 - n2);
  if (have_ray && seg == 0) {
    counts[r] = n0 + n1 + n2 + n3;
    reinterpret_cast<int4*>(segcnt)[r] = make_int4(n0, n1, n2, n3);
  }
}

// ---- train: pass 2 (single CTA): exclusive scan of counts -> rays_a, counter ---------------
// A thread owns 8 CONSECUTIVE rays (two 16-byte loads, scanned in registers), so 8192 rays are one pass with one
// memory round trip and two block barriers (the per-1024 loop paid a load latency and four barriers per pass).
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 8;
__global__ void __launch_bounds__(kScanThreads)
march_scan_kernel(const int32_t* __restrict__ counts, int64_t n_rays, int32_t* __restrict__ starts,
                  int32_t* __restrict__ counter) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_rays; base += kScanThreads * kScanItems) {
    const int64_t r0 = base + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    if (r0 + kScanItems <= n_rays) {
      const int4 a = *reinterpret_cast<const int4*>(counts + r0), b = *reinterpret_cast<const int4*>(counts + r0 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int q = 0; q < kScanItems; ++q) v[q] = r0 + q < n_rays ? counts[r0 + q] : 0;
    }
    int tot = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) tot += v[q];
    const int incl = warp_scan_incl_i(tot, lane);
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int w = s_warp[lane];
      const int wi = warp_scan_incl_i(w, lane);
      s_warp[lane] = wi - w;   // exclusive offset of each warp
    }
    __syncthreads();
    int excl = s_carry + s_warp[wid] + incl - tot;
    int st[kScanItems];
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) { st[q] = excl; excl += v[q]; }
    if (r0 + kScanItems <= n_rays) {      // two 16-byte stores; the (R,3) i64 rays_a rows are written by a parallel kernel
      *reinterpret_cast<int4*>(starts + r0) = make_int4(st[0], st[1], st[2], st[3]);
      *reinterpret_cast<int4*>(starts + r0 + 4) = make_int4(st[4], st[5], st[6], st[7]);
    } else {
#pragma unroll
      for (int q = 0; q < kScanItems; ++q) if (r0 + q < n_rays) starts[r0 + q] = st[q];
    }
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) s_carry = excl;
    __syncthreads();
  }
  if (threadIdx.x == 0) { counter[0] = s_carry; counter[1] = (int32_t)n_rays; }
}

// rays_a (R,3) i64 = [ray, start, n] from the compact count / start arrays (all SMs; one CTA writing 196 KB of 8-byte
// rows was 12 us of the old scan kernel)
__global__ void __launch_bounds__(256)
march_rays_a_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ starts, int64_t n_rays,
                    int64_t* __restrict__ rays_a) {
  const int64_t total = n_rays * 3, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / 3;
    const int j = (int)(i - 3 * r);
    rays_a[i] = j == 0 ? r : (j == 1 ? (int64_t)starts[r] : (int64_t)counts[r]);
  }
}

// ---- train: pass 3 (one warp per ray): expand recorded ts into the sample arrays -----------
__global__ void __launch_bounds__(256)
march_train_expand_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                          const int32_t* __restrict__ counts, const int32_t* __restrict__ starts,
                          int64_t* __restrict__ rays_a_out, const float* __restrict__ ts_scratch,
                          const int32_t* __restrict__ segcnt, int slab_stride, int seg_stride,
                          MarchCfg c, int64_t n_rays, int64_t capacity,
                          float* __restrict__ xyzs, float* __restrict__ dirs,
                          float* __restrict__ deltas, float* __restrict__ ts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rays; r += n_warps) {
    const int64_t start = starts[r];
    const int n = counts[r];
    if (rays_a_out != nullptr && lane < 3) rays_a_out[3 * r + lane] = lane == 0 ? r : (lane == 1 ? start : (int64_t)n);
    if (n == 0) continue;
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float* src = ts_scratch + r * (int64_t)slab_stride;
    const int4 sc = reinterpret_cast<const int4*>(segcnt)[r];      // samples per segment region of the ray's slab
    const int p1 = sc.x, p2 = p1 + sc.y, p3 = p2 + sc.z;
    for (int k = lane; k < n; k += 32) {
      const int64_t s = start + k;
      if (s >= capacity) break;
      const int sg = k < p1 ? 0 : (k < p2 ? 1 : (k < p3 ? 2 : 3));
      const int off = sg * seg_stride + (k - (sg == 0 ? 0 : (sg == 1 ? p1 : (sg == 2 ? p2 : p3))));
      const float t = src[off];
      float* p = xyzs + 3 * s;
      p[0] = __fmaf_rn(dx, t, ox); p[1] = __fmaf_rn(dy, t, oy); p[2] = __fmaf_rn(dz, t, oz);
      float* q = dirs + 3 * s;
      q[0] = dx; q[1] = dy; q[2] = dz;
      ts[s] = t;
      deltas[s] = calc_dt(t, c);
    }
  }
}

// ---- test-time march (one thread per alive ray) --------------------------------------------
__global__ void __launch_bounds__(128)
march_test_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                  float* __restrict__ hits_t, const int64_t* __restrict__ alive,
                  const uint8_t* __restrict__ bitfield, MarchCfg c, int n_samples, int64_t n_alive,
                  float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
                  float* __restrict__ ts, int32_t* __restrict__ n_eff) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_alive) return;
  const int64_t r = alive[n];
  const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
  const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
  const float ix = __frcp_rn(dx), iy = __frcp_rn(dy), iz = __frcp_rn(dz);
  const float sx = copysignf(0.5f, dx), sy = copysignf(0.5f, dy), sz = copysignf(0.5f, dz);
  float t = hits_t[2 * r];
  const float t2 = hits_t[2 * r + 1];
  float* px = xyzs + n * (int64_t)n_samples * 3;
  float* pd = dirs + n * (int64_t)n_samples * 3;
  float* pt = ts + n * (int64_t)n_samples;
  float* pdt = deltas + n * (int64_t)n_samples;
  int s = 0;
  float t_resume = t;
  while (t < t2 && s < n_samples) {
    const float x = __fmaf_rn(dx, t, ox), y = __fmaf_rn(dy, t, oy), z = __fmaf_rn(dz, t, oz);
    const float dt = calc_dt(t, c);
    const Cell k = locate(x, y, z, dt, c, bitfield);
    if (k.occ) {
      px[3 * s] = x; px[3 * s + 1] = y; px[3 * s + 2] = z;
      pd[3 * s] = dx; pd[3 * s + 1] = dy; pd[3 * s + 2] = dz;
      pt[s] = t; pdt[s] = dt;
      t = __fadd_rn(t, dt);
      t_resume = t;
      ++s;
    } else {
      t = skip_cell(t, k, x, y, z, sx, sy, sz, ix, iy, iz, c);
    }
  }
  if (s > 0) hits_t[2 * r] = t_resume;  // start of the next march (raymarching.cu:390)
  n_eff[n] = s;
  for (int j = s; j < n_samples; ++j) {   // padding slots read as zeros (torch::zeros in the reference)
    px[3 * j] = 0.f; px[3 * j + 1] = 0.f; px[3 * j + 2] = 0.f;
    pd[3 * j] = 0.f; pd[3 * j + 1] = 0.f; pd[3 * j + 2] = 0.f;
    pt[j] = 0.f; pdt[j] = 0.f;
  }
}

}  // namespace ncn

using namespace ncn;

static int make_cfg(MarchCfg* c, int cascades, float scale, float dt_scale, float exp_step_factor,
                    int grid_size, int max_samples) {
  if (cascades < 1 || cascades > 8) return NCN_E_CONFIG;
  if (grid_size < 2 || grid_size > 1024) return NCN_E_CONFIG;
  if (max_samples < 1) return NCN_E_CONFIG;
  c->cascades = cascades; c->grid_size = grid_size; c->max_samples = max_samples;
  c->grid_size3 = (uint32_t)grid_size * grid_size * grid_size;
  c->scale = scale; c->exp_step_factor = exp_step_factor;
  // host IEEE single division == div.rn.f32
  c->dt_min = 1.73205080757f / (float)max_samples;
  volatile float two_sqrt3_scale = dt_scale * 3.4641015529632568359f;   // (SQRT3*2 folded)*scale
  c->dt_max = two_sqrt3_scale / (float)grid_size;
  c->const_dt = (exp_step_factor == 0.0f && c->dt_max >= c->dt_min && c->dt_min > 0.0f) ? 1 : 0;
  c->gs_f = (float)grid_size; c->gs_inv = 1.0f / (float)grid_size; c->gs_m1 = (float)grid_size - 1.0f;
  return NCN_OK;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// workspace = [counts (R) i32 | starts (R) i32 | samples per segment region (R,4) i32 | per-ray slabs of sample times (R, slab_stride) f32];
// a slab is 4 segment regions of seg_stride floats (the general kernel uses it as one region of >= max_samples floats)
static inline int march_seg_stride(int max_samples) { return (max_samples + kSegs - 1) / kSegs + kSegPad; }
static inline int march_slab_stride(int max_samples) { return kSegs * march_seg_stride(max_samples); }
static inline size_t march_off_starts(int64_t n_rays) { return align256((size_t)n_rays * sizeof(int32_t)); }
static inline size_t march_off_segcnt(int64_t n_rays) { return 2 * align256((size_t)n_rays * sizeof(int32_t)); }
static inline size_t march_off_slabs(int64_t n_rays) { return march_off_segcnt(n_rays) + align256((size_t)n_rays * kSegs * sizeof(int32_t)); }

extern "C" size_t ncn_march_train_workspace_bytes(int64_t n_rays, int max_samples) {
  if (n_rays < 0 || max_samples < 1) return 0;
  return march_off_slabs(n_rays) + align256((size_t)n_rays * (size_t)march_slab_stride(max_samples) * sizeof(float));
}

static int g_march_segments = 1;   // 1 = four lanes per ray on the constant-dt path (default), 0 = one lane per ray, 2 = four lanes and
                                   // every segment re-marched from its predecessor's landing point (exercises the repair path)
extern "C" int ncn_set_march_segments(int mode) { const int old = g_march_segments; g_march_segments = mode; return old; }

static int march_count_impl(const float* rays_o, const float* rays_d, const float* hits_t,
                            const uint8_t* density_bitfield, int cascades, float scale,
                            float exp_step_factor, const float* noise, int grid_size, int max_samples,
                            int64_t n_rays, int64_t* rays_a, bool write_rays_a, int32_t* counter, void* workspace,
                            size_t workspace_bytes, ncn_stream_t stream) {
  MarchCfg c;
  int rc = make_cfg(&c, cascades, scale, scale, exp_step_factor, grid_size, max_samples);
  if (rc) return rc;
  NCN_CHECK_SIZE(n_rays >= 0 && n_rays < (int64_t)1 << 31);
  NCN_CHECK_PTR(counter);
  if (n_rays > 0) {
    NCN_CHECK_PTR(rays_o); NCN_CHECK_PTR(rays_d); NCN_CHECK_PTR(hits_t); NCN_CHECK_PTR(density_bitfield);
    NCN_CHECK_PTR(rays_a); NCN_CHECK_PTR(workspace);
    if (workspace_bytes < ncn_march_train_workspace_bytes(n_rays, max_samples)) return NCN_E_SIZE;
    if (((uintptr_t)hits_t & 7) || ((uintptr_t)workspace & 255)) return NCN_E_ALIGN;
    int32_t* counts = (int32_t*)workspace;
    int32_t* starts = (int32_t*)((char*)workspace + march_off_starts(n_rays));
    int32_t* segcnt = (int32_t*)((char*)workspace + march_off_segcnt(n_rays));
    float* ts_scratch = (float*)((char*)workspace + march_off_slabs(n_rays));
    const int slab_stride = march_slab_stride(max_samples), seg_stride = march_seg_stride(max_samples);
    const int threads = 64;                                   // 2 warps x kRaysPerWarp rays
    const unsigned blocks = (unsigned)ceil_div(n_rays, (threads / 32) * kRaysPerWarp);
    if (c.cascades == 1 && c.const_dt && g_march_segments != 0 && max_samples % kSegs == 0 && max_samples >= 64)
      march_train_count_seg_kernel<<<blocks, threads, 0, as_stream(stream)>>>(rays_o, rays_d, hits_t, density_bitfield, noise, c, n_rays,
                                                                              counts, segcnt, ts_scratch, slab_stride, seg_stride,
                                                                              g_march_segments == 2 ? 1 : 0);
    else if (c.cascades == 1 && c.const_dt)
      march_train_count_kernel<true><<<blocks, threads, 0, as_stream(stream)>>>(rays_o, rays_d, hits_t, density_bitfield, noise, c,
                                                                             n_rays, counts, segcnt, ts_scratch, slab_stride);
    else
      march_train_count_kernel<false><<<blocks, threads, 0, as_stream(stream)>>>(rays_o, rays_d, hits_t, density_bitfield, noise, c,
                                                                              n_rays, counts, segcnt, ts_scratch, slab_stride);
    NCN_LAUNCH_OK();
    march_scan_kernel<<<1, kScanThreads, 0, as_stream(stream)>>>(counts, n_rays, starts, counter);
    NCN_LAUNCH_OK();
    if (write_rays_a)
      march_rays_a_kernel<<<persistent_grid(n_rays * 3, 256, 4), 256, 0, as_stream(stream)>>>(counts, starts, n_rays, rays_a);
  } else {
    march_scan_kernel<<<1, kScanThreads, 0, as_stream(stream)>>>(nullptr, 0, nullptr, counter);
  }
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_march_train_count(const float* rays_o, const float* rays_d, const float* hits_t,
                                     const uint8_t* density_bitfield, int cascades, float scale,
                                     float exp_step_factor, const float* noise, int grid_size, int max_samples,
                                     int64_t n_rays, int64_t* rays_a, int32_t* counter, void* workspace,
                                     size_t workspace_bytes, ncn_stream_t stream) {
  return march_count_impl(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise, grid_size, max_samples,
                          n_rays, rays_a, true, counter, workspace, workspace_bytes, stream);
}

static int march_expand_impl(const float* rays_o, const float* rays_d, const int64_t* rays_a, int64_t* rays_a_out,
                             float exp_step_factor, float scale, int grid_size, int max_samples,
                             int64_t n_rays, int64_t capacity, float* xyzs, float* dirs, float* deltas,
                             float* ts, const void* workspace, size_t workspace_bytes,
                             ncn_stream_t stream) {
  MarchCfg c;
  int rc = make_cfg(&c, 1, scale, scale, exp_step_factor, grid_size, max_samples);
  if (rc) return rc;
  NCN_CHECK_SIZE(n_rays >= 0 && capacity >= 0);
  if (n_rays == 0 || capacity == 0) return NCN_OK;
  NCN_CHECK_PTR(rays_o); NCN_CHECK_PTR(rays_d); NCN_CHECK_PTR(rays_a); NCN_CHECK_PTR(workspace);
  NCN_CHECK_PTR(xyzs); NCN_CHECK_PTR(dirs); NCN_CHECK_PTR(deltas); NCN_CHECK_PTR(ts);
  if (workspace_bytes < ncn_march_train_workspace_bytes(n_rays, max_samples)) return NCN_E_SIZE;
  const int32_t* counts = (const int32_t*)workspace;
  const int32_t* starts = (const int32_t*)((const char*)workspace + march_off_starts(n_rays));
  const int32_t* segcnt = (const int32_t*)((const char*)workspace + march_off_segcnt(n_rays));
  const float* ts_scratch = (const float*)((const char*)workspace + march_off_slabs(n_rays));
  const int grid = persistent_grid(n_rays * 32, 256, 8);
  march_train_expand_kernel<<<grid, 256, 0, as_stream(stream)>>>(rays_o, rays_d, counts, starts, rays_a_out, ts_scratch, segcnt,
                                                                 march_slab_stride(max_samples), march_seg_stride(max_samples), c, n_rays,
                                                                 capacity, xyzs, dirs, deltas, ts);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

// rays_a must be the array ncn_march_train_count filled (the per-ray counts / starts are re-read from the workspace)
extern "C" int ncn_march_train_expand(const float* rays_o, const float* rays_d, const int64_t* rays_a,
                                      float exp_step_factor, float scale, int grid_size, int max_samples,
                                      int64_t n_rays, int64_t capacity, float* xyzs, float* dirs, float* deltas,
                                      float* ts, const void* workspace, size_t workspace_bytes,
                                      ncn_stream_t stream) {
  return march_expand_impl(rays_o, rays_d, rays_a, nullptr, exp_step_factor, scale, grid_size, max_samples, n_rays, capacity, xyzs, dirs,
                           deltas, ts, workspace, workspace_bytes, stream);
}

extern "C" int ncn_march_train(const float* rays_o, const float* rays_d, const float* hits_t,
                               const uint8_t* density_bitfield, int cascades, float scale, float exp_step_factor,
                               const float* noise, int grid_size, int max_samples, int64_t n_rays, int64_t capacity,
                               int64_t* rays_a, float* xyzs, float* dirs, float* deltas, float* ts,
                               int32_t* counter, void* workspace, size_t workspace_bytes, ncn_stream_t stream) {
  // count -> scan -> expand; the (R,3) rays_a rows are written by the (all-SM) expansion kernel instead of a separate pass
  const bool expand_writes = n_rays > 0 && capacity > 0;
  int rc = march_count_impl(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise, grid_size, max_samples,
                            n_rays, rays_a, !expand_writes, counter, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return march_expand_impl(rays_o, rays_d, rays_a, expand_writes ? rays_a : nullptr, exp_step_factor, scale, grid_size, max_samples,
                           n_rays, capacity, xyzs, dirs, deltas, ts, workspace, workspace_bytes, stream);
}

extern "C" int ncn_march_test(const float* rays_o, const float* rays_d, float* hits_t,
                              const int64_t* alive_indices, const uint8_t* density_bitfield, int cascades,
                              float scale, float exp_step_factor, int grid_size, int max_samples, int n_samples,
                              int64_t n_alive, float* xyzs, float* dirs, float* deltas, float* ts, int32_t* n_eff,
                              ncn_stream_t stream) {
  MarchCfg c;
  // the reference passes `cascades` where calc_dt expects `scale` (raymarching.cu:370,399)
  int rc = make_cfg(&c, cascades, scale, (float)cascades, exp_step_factor, grid_size, max_samples);
  if (rc) return rc;
  NCN_CHECK_SIZE(n_alive >= 0 && n_samples >= 1);
  if (n_alive == 0) return NCN_OK;
  NCN_CHECK_PTR(rays_o); NCN_CHECK_PTR(rays_d); NCN_CHECK_PTR(hits_t); NCN_CHECK_PTR(alive_indices);
  NCN_CHECK_PTR(density_bitfield); NCN_CHECK_PTR(xyzs); NCN_CHECK_PTR(dirs); NCN_CHECK_PTR(deltas);
  NCN_CHECK_PTR(ts); NCN_CHECK_PTR(n_eff);
  const int threads = 128;
  march_test_kernel<<<(unsigned)ceil_div(n_alive, threads), threads, 0, as_stream(stream)>>>(
      rays_o, rays_d, hits_t, alive_indices, density_bitfield, c, n_samples, n_alive, xyzs, dirs, deltas, ts, n_eff);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
