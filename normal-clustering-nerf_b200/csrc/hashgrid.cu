// Multiresolution hash-grid encoding (tcnn "Grid"/"Hash"/"Linear", 3-D inputs).
// Replaces tcnn.Encoding as used by models/ngp_mt.py:70-82 (semantics: SURVEY.md Appendix B;
// tiny-cuda-nn itself is NOT in the reference tree - parity is pinned against oracle/hashgrid.py).
//
// Thread mapping: one thread per (sample, level) with the level index fastest, so a warp
// covers 2 samples x 16 levels and writes 2 x 64 B = one contiguous 128 B line of the
// (N, L*F) fp16 output; the 8 corner gathers per thread are independent 4 B (half2) loads
// (MLP = 8) served by L1/L2 (the 21.8 MiB fp16 table at T=2^19 is L2 resident on B200).
// Backward scatters fp32 with vector red.global.add.v2.f32 (one per corner).
#include "ncn_common.cuh"
#include <math.h>

namespace ncn {

struct GridXform {         // optional input normalisation x' = (x - lo) / size  (models/ngp_mt.py:166)
  float lo[3], size[3];
  int on;
};

struct GridMeta {          // per-level constants, copied to shared memory by each CTA
  float scale[NCN_GRID_MAX_LEVELS];
  uint32_t res[NCN_GRID_MAX_LEVELS];
  uint32_t size[NCN_GRID_MAX_LEVELS];
  uint32_t offset[NCN_GRID_MAX_LEVELS];
  int32_t n_levels;
};

__device__ __forceinline__ uint32_t grid_index(uint32_t gx, uint32_t gy, uint32_t gz, uint32_t res, uint32_t size) {
  // dense while the running stride fits in the level, else the coherent prime hash
  uint32_t stride = 1, index = 0;
  index += gx * stride; stride *= res;
  if (stride <= size) { index += gy * stride; stride *= res;
    if (stride <= size) { index += gz * stride; stride *= res; } }
  if (size < stride) index = gx ^ (gy * 2654435761u) ^ (gz * 805459861u);
  // hashed levels hold exactly T = 2^log2_T entries (mask); on a dense level index < res^3 <= size for every in-range corner,
  // so the general modulo is only reached by out-of-range inputs
  if ((size & (size - 1u)) == 0u) return index & (size - 1u);
  if (index < size) return index;
  return index % size;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void red_add_h2(__half2* addr, __half2 v) {
  asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(addr), "r"(*reinterpret_cast<const uint32_t*>(&v)) : "memory");
}
__device__ __forceinline__ void red_add_v2_h2(__half2* addr, __half2 a, __half2 b) {
  asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(addr), "r"(*reinterpret_cast<const uint32_t*>(&a)),
               "r"(*reinterpret_cast<const uint32_t*>(&b)) : "memory");
}

struct Cell8 {
  uint32_t idx[8];
  float w[3];       // fractional position
};

__device__ __forceinline__ void locate8(const float* __restrict__ x, int64_t n, float scale, uint32_t res, uint32_t size,
                                        Cell8& c, const GridXform& xf, uint32_t* cell = nullptr) {
  uint32_t g[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float xi = x[3 * n + d];
    if (xf.on) xi = __fdiv_rn(__fsub_rn(xi, xf.lo[d]), xf.size[d]);   // (x - xyz_min) / (xyz_max - xyz_min), IEEE like torch
    const float pos = fmaf(scale, xi, 0.5f);
    const float fl = floorf(pos);
    g[d] = (uint32_t)(int)fl;
    c.w[d] = pos - fl;
  }
  // The 8 corner indices share their per-axis terms (same values as 8 calls of grid_index, ~4x fewer instructions):
  //   dense level  (res^3 <= size): base + dx + dy*res + dz*res^2, valid while every corner is inside the level;
  //   hashed level (size = 2^k)   : (x+dx) ^ (y*P1 + dy*P1) ^ (z*P2 + dz*P2), masked.
  const bool dense = (uint64_t)res * res * res <= (uint64_t)size;
  if (dense && g[0] + 1u < res && g[1] + 1u < res && g[2] + 1u < res) {
    const uint32_t r2 = res * res;
    const uint32_t base = g[0] + g[1] * res + g[2] * r2;
#pragma unroll
    for (int k = 0; k < 8; ++k) c.idx[k] = base + (k & 1) + ((k >> 1) & 1) * res + ((k >> 2) & 1) * r2;
  } else if (!dense && (size & (size - 1u)) == 0u) {
    const uint32_t hy0 = g[1] * 2654435761u, hz0 = g[2] * 805459861u;
    const uint32_t hx[2] = {g[0], g[0] + 1u}, hy[2] = {hy0, hy0 + 2654435761u}, hz[2] = {hz0, hz0 + 805459861u};
    const uint32_t mask = size - 1u;
#pragma unroll
    for (int k = 0; k < 8; ++k) c.idx[k] = (hx[k & 1] ^ hy[(k >> 1) & 1] ^ hz[(k >> 2) & 1]) & mask;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      c.idx[k] = grid_index(g[0] + (k & 1), g[1] + ((k >> 1) & 1), g[2] + ((k >> 2) & 1), res, size);
  }
  if (cell) { cell[0] = g[0] | (g[1] << 16); cell[1] = g[2]; }
}

__device__ __forceinline__ float corner_w(const float* w, int k) {
  return ((k & 1) ? w[0] : 1.f - w[0]) * ((k & 2) ? w[1] : 1.f - w[1]) * ((k & 4) ? w[2] : 1.f - w[2]);
}
// d(corner weight)/d(w[d])
__device__ __forceinline__ float corner_dw(const float* w, int k, int d) {
  float r = ((k >> d) & 1) ? 1.f : -1.f;
#pragma unroll
  for (int e = 0; e < 3; ++e) if (e != d) r *= ((k >> e) & 1) ? w[e] : 1.f - w[e];
  return r;
}

#define NCN_GRID_PROLOGUE                                                          \
  __shared__ GridMeta sm;                                                          \
  for (int i = threadIdx.x; i < (int)(sizeof(GridMeta) / 4); i += blockDim.x)     \
    reinterpret_cast<uint32_t*>(&sm)[i] = reinterpret_cast<const uint32_t*>(&meta)[i]; \
  __syncthreads();                                                                 \
  const int L = sm.n_levels;                                                       \
  int64_t n = n_cap;                                                               \
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }         \
  const int64_t total = n * L;                                                     \
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;

template <int F>
__global__ void __launch_bounds__(256)
grid_fwd_kernel(const __grid_constant__ GridMeta meta, const float* __restrict__ x, const __half* __restrict__ table,
                int64_t n_cap, const int32_t* __restrict__ n_dev, GridXform xf, __half* __restrict__ out) {
  NCN_GRID_PROLOGUE
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t s = i / L;
    const int l = (int)(i - s * L);
    Cell8 c;
    locate8(x, s, sm.scale[l], sm.res[l], sm.size[l], c, xf);
    const __half* tl = table + (size_t)sm.offset[l] * F;
    float acc[F];
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = 0.f;
    if (F == 2) {
      __half2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const __half2*>(tl) + c.idx[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float wk = corner_w(c.w, k);
        const float2 f2 = __half22float2(v[k]);
        acc[0] = fmaf(wk, f2.x, acc[0]); acc[1] = fmaf(wk, f2.y, acc[1]);
      }
      reinterpret_cast<__half2*>(out)[i] = __floats2half2_rn(acc[0], acc[1]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float wk = corner_w(c.w, k);
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = fmaf(wk, __half2float(__ldg(tl + (size_t)c.idx[k] * F + f)), acc[f]);
      }
#pragma unroll
      for (int f = 0; f < F; ++f) out[i * F + f] = __float2half_rn(acc[f]);
    }
  }
}

// Forward with a sample-coherent mapping (F = 2): a warp evaluates ONE level for 32 CONSECUTIVE samples, so at the coarse
// levels its 8 corner gathers touch a handful of cache lines instead of 32 different ones (the (sample, level)-interleaved
// mapping above is bound by L1 wavefronts: every lane reads a different level's table).  A CTA of 8 warps covers 32
// samples x 16 levels (2 levels per warp) and transposes the 32 x 32 fp16 tile through shared memory so the output rows
// are still written as full 64 B segments.
__global__ void __launch_bounds__(256)
grid_fwd_coherent_kernel(const __grid_constant__ GridMeta meta, const float* __restrict__ x, const __half* __restrict__ table,
                         int64_t n_cap, const int32_t* __restrict__ n_dev, GridXform xf, __half* __restrict__ out) {
  __shared__ GridMeta sm;
  __shared__ __align__(16) __half2 tile[32][17];      // [sample][level] (+1 pad)
  for (int i = threadIdx.x; i < (int)(sizeof(GridMeta) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&sm)[i] = reinterpret_cast<const uint32_t*>(&meta)[i];
  __syncthreads();
  pdl_wait(); pdl_trigger();
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t n_chunks = (n + 31) >> 5;
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int64_t s = (chunk << 5) + lane;
    const bool valid = s < n;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int l = wid + 8 * q;
      __half2 r = __floats2half2_rn(0.f, 0.f);
      if (valid) {
        Cell8 c;
        locate8(x, s, sm.scale[l], sm.res[l], sm.size[l], c, xf);
        const __half2* tl = reinterpret_cast<const __half2*>(table) + sm.offset[l];
        __half2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(tl + c.idx[k]);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float wk = corner_w(c.w, k);
          const float2 f2 = __half22float2(v[k]);
          a0 = fmaf(wk, f2.x, a0); a1 = fmaf(wk, f2.y, a1);
        }
        r = __floats2half2_rn(a0, a1);
      }
      tile[lane][l] = r;
    }
    __syncthreads();
    // 32 samples x 16 half2: thread -> (sample = tid / 8, two half2 = 8 B) : rows of 64 B, fully coalesced
    {
      const int sr = threadIdx.x >> 3, c2 = (threadIdx.x & 7) * 2;
      const int64_t so = (chunk << 5) + sr;
      if (so < n) {
        const __half2 a = tile[sr][c2], b = tile[sr][c2 + 1];
        uint2 pk; pk.x = *reinterpret_cast<const uint32_t*>(&a); pk.y = *reinterpret_cast<const uint32_t*>(&b);
        reinterpret_cast<uint2*>(out + so * 32)[c2 >> 1] = pk;
      }
    }
    __syncthreads();
  }
}

template <int F>
__global__ void __launch_bounds__(256)
grid_bwd_kernel(const __grid_constant__ GridMeta meta, const float* __restrict__ x, const __half* __restrict__ dy,
                int64_t n_cap, const int32_t* __restrict__ n_dev, GridXform xf, float grad_scale, float* __restrict__ grad) {
  NCN_GRID_PROLOGUE
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t s = i / L;
    const int l = (int)(i - s * L);
    float g[F];
    bool any = false;
#pragma unroll
    for (int f = 0; f < F; ++f) { g[f] = __half2float(dy[i * F + f]) * grad_scale; any |= (g[f] != 0.f); }
    if (!any) continue;    // samples cut by early ray termination carry exact-zero gradients
    Cell8 c;
    locate8(x, s, sm.scale[l], sm.res[l], sm.size[l], c, xf);
    float* gl = grad + (size_t)sm.offset[l] * F;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float wk = corner_w(c.w, k);
      if (F == 2) {
        atomicAdd(reinterpret_cast<float2*>(gl) + c.idx[k], make_float2(wk * g[0], wk * g[1]));
      } else {
#pragma unroll
        for (int f = 0; f < F; ++f) atomicAdd(gl + (size_t)c.idx[k] * F + f, wk * g[f]);
      }
    }
  }
}

// Parameter-gradient scatter with warp-level run merging (F = 2).
// Samples arrive in ray order, 1.7e-3 apart, so at the coarse levels dozens of consecutive samples fall in the same
// cell and would issue dozens of same-address reductions (which the L2 atomic unit serialises).  Here a warp takes 32
// CONSECUTIVE samples of ONE level; lanes in the same grid cell as their left neighbour form a run (found once per
// sample, not per corner), each corner's contributions are summed over the run with a windowed segmented warp scan whose
// depth follows the average run length, and only the run's last lane issues the reduction.  Warps whose samples
// are not coherent at this level (fine / hashed levels) detect that with one ballot and take the direct path.
// segmented sums of the 8 corner contributions over runs of lanes in the same cell; runs are cut at W-lane windows so
// the prefix scan needs log2(W) steps.  `heads` already contains the window starts.
// the contributions (ax, ay) to entry i0 and (bx, by) to entry i1 of one level, i0 / i1 = the two x-neighbours of a cell.
// For an even cell x the two entries are neighbours in the dense layout AND in the hashed one (x enters the hash with
// multiplier 1, so x ^ h and (x+1) ^ h differ in bit 0 only): one 16-byte reduction instead of two 8-byte ones.
// HALF: the gradient buffer holds __half2 entries (what tiny-cuda-nn accumulates into for F = 2): 4 / 8 byte packed reductions.
template <bool HALF>
__device__ __forceinline__ void emit_pair(void* __restrict__ gl, uint32_t i0, uint32_t i1, float ax, float ay, float bx, float by) {
  const bool z0 = ax == 0.f && ay == 0.f, z1 = bx == 0.f && by == 0.f;
  if (z0 && z1) return;
  if ((i0 ^ i1) == 1u) {
    const uint32_t lo = i0 & ~1u;
    const bool sw = (i0 & 1u) != 0;                             // hashed: the pair may come out swapped
    if (HALF) {
      const __half2 h0 = __floats2half2_rn(sw ? bx : ax, sw ? by : ay), h1 = __floats2half2_rn(sw ? ax : bx, sw ? ay : by);
      red_add_v2_h2(reinterpret_cast<__half2*>(gl) + lo, h0, h1);
    } else {
      red_add_v4(reinterpret_cast<float*>(reinterpret_cast<float2*>(gl) + lo), sw ? bx : ax, sw ? by : ay, sw ? ax : bx, sw ? ay : by);
    }
  } else if (HALF) {
    if (!z0) red_add_h2(reinterpret_cast<__half2*>(gl) + i0, __floats2half2_rn(ax, ay));
    if (!z1) red_add_h2(reinterpret_cast<__half2*>(gl) + i1, __floats2half2_rn(bx, by));
  } else {
    if (!z0) atomicAdd(reinterpret_cast<float2*>(gl) + i0, make_float2(ax, ay));
    if (!z1) atomicAdd(reinterpret_cast<float2*>(gl) + i1, make_float2(bx, by));
  }
}

template <int W, bool HALF>
__device__ __forceinline__ void merge_corners(const Cell8& c, bool valid, float g0, float g1, unsigned heads, int lane, void* __restrict__ gl) {
  // segment = run of lanes in one cell, cut at W-lane windows (`heads` has the window starts set).  Segmented Hillis-Steele
  // scan: at distance o a lane adds its left neighbour's partial sum only if that neighbour belongs to the same segment, so the
  // LAST lane of a segment ends up with the segment's sum and issues the reduction itself (same cell = same 8 entries) - no
  // prefix differences, no broadcast back to the head: 2 log2(W) shuffles per corner instead of 2 log2(W) + 4
  const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));   // bit 0 of `heads` is always set
  const int dist = lane - start;
  const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    float s[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float wk = valid ? corner_w(c.w, k + q) : 0.f;
      float vx = wk * g0, vy = wk * g1;
#pragma unroll
      for (int o = 1; o < W; o <<= 1) {
        const float tx = __shfl_up_sync(0xffffffffu, vx, o), ty = __shfl_up_sync(0xffffffffu, vy, o);
        if (dist >= o) { vx += tx; vy += ty; }
      }
      s[q][0] = vx; s[q][1] = vy;
    }
    if (tail && valid) emit_pair<HALF>(gl, c.idx[k], c.idx[k + 1], s[0][0], s[0][1], s[1][0], s[1][1]);
  }
}

template <bool HALF, int OCC>
__global__ void __launch_bounds__(256, OCC)
grid_bwd_merge_kernel(const __grid_constant__ GridMeta meta, const float* __restrict__ x, const __half* __restrict__ dy,
                      int64_t n_cap, const int32_t* __restrict__ n_dev, GridXform xf, float grad_scale, void* __restrict__ grad,
                      int level_begin, int level_end) {
  __shared__ GridMeta sm;
  for (int i = threadIdx.x; i < (int)(sizeof(GridMeta) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&sm)[i] = reinterpret_cast<const uint32_t*>(&meta)[i];
  __syncthreads();
  pdl_wait(); pdl_trigger();
  const int L = sm.n_levels;
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  const int lane = threadIdx.x & 31;
  const int LR = level_end - level_begin;                      // levels handled by this launch
  const int64_t n_items = ((n + 31) >> 5) * LR;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int64_t trip = 0;
  const bool rotate = (n_warps % LR) == 0;
  for (int64_t item = warp0; item < n_items; item += n_warps, ++trip) {
    const int64_t chunk = item / LR;
    // rotate the level with the trip count: with a warp count that is a multiple of LR a warp would stay on ONE level for all
    // of its items, and CTAs of coarse (merging) levels and of fine (direct) levels would run for different times (the LR
    // items of a chunk then share one trip count, so the rotation permutes them; other warp counts rotate by themselves)
    int lr = (int)(item - chunk * LR) + (rotate ? (int)(trip % LR) : 0);
    if (lr >= LR) lr -= LR;
    const int l = level_begin + lr;
    const int64_t s = (chunk << 5) + lane;
    const bool valid = s < n;
    float g0 = 0.f, g1 = 0.f;
    if (valid) {
      const float2 gg = __half22float2(*reinterpret_cast<const __half2*>(dy + (s * L + l) * 2));
      g0 = gg.x * grad_scale; g1 = gg.y * grad_scale;
    }
    const bool live = valid && (g0 != 0.f || g1 != 0.f);
    if (__ballot_sync(0xffffffffu, live) == 0u) continue;      // e.g. samples behind an early-terminated ray
    Cell8 c;
    uint32_t cell[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
    if (valid) locate8(x, s, sm.scale[l], sm.res[l], sm.size[l], c, xf, cell);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) c.idx[k] = 0xFFFFFFFFu;
      c.w[0] = c.w[1] = c.w[2] = 0.f;
    }
    void* gl = HALF ? static_cast<void*>(reinterpret_cast<__half2*>(grad) + sm.offset[l])
                    : static_cast<void*>(reinterpret_cast<float2*>(grad) + sm.offset[l]);
    // runs of consecutive samples in the SAME cell share all 8 table entries (exact: compared on the integer cell)
    const uint32_t pa = __shfl_up_sync(0xffffffffu, cell[0], 1), pb = __shfl_up_sync(0xffffffffu, cell[1], 1);
    const unsigned heads0 = __ballot_sync(0xffffffffu, lane == 0 || cell[0] != pa || cell[1] != pb || !valid);
    const int nh = __popc(heads0);
    if (nh > 20) {                                              // incoherent at this level (fine / hashed): direct path
      if (live) {
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          const float w0 = corner_w(c.w, k), w1 = corner_w(c.w, k + 1);
          emit_pair<HALF>(gl, c.idx[k], c.idx[k + 1], w0 * g0, w0 * g1, w1 * g0, w1 * g1);
        }
      }
      continue;
    }
    // window = about twice the average run: short scans where runs are short, whole-warp sums on the coarse levels
    if (nh <= 2) merge_corners<32, HALF>(c, valid, g0, g1, heads0 | 0x00000001u, lane, gl);
    else if (nh <= 4) merge_corners<16, HALF>(c, valid, g0, g1, heads0 | 0x00010001u, lane, gl);
    else if (nh <= 10) merge_corners<8, HALF>(c, valid, g0, g1, heads0 | 0x01010101u, lane, gl);
    else merge_corners<4, HALF>(c, valid, g0, g1, heads0 | 0x11111111u, lane, gl);
  }
}

// dL/dx (atomically accumulated over levels into a zeroed (N,3) buffer)
template <int F>
__global__ void __launch_bounds__(256)
grid_bwd_input_kernel(const __grid_constant__ GridMeta meta, const float* __restrict__ x,
                      const __half* __restrict__ table, const __half* __restrict__ dy, int64_t n_cap,
                      const int32_t* __restrict__ n_dev, GridXform xf, float* __restrict__ dx) {
  NCN_GRID_PROLOGUE
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t s = i / L;
    const int l = (int)(i - s * L);
    Cell8 c;
    const float scale = sm.scale[l];
    locate8(x, s, scale, sm.res[l], sm.size[l], c, xf);
    const __half* tl = table + (size_t)sm.offset[l] * F;
    float g[F];
#pragma unroll
    for (int f = 0; f < F; ++f) g[f] = __half2float(dy[i * F + f]);
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float dot = 0.f;
#pragma unroll
      for (int f = 0; f < F; ++f) dot = fmaf(__half2float(__ldg(tl + (size_t)c.idx[k] * F + f)), g[f], dot);
#pragma unroll
      for (int d = 0; d < 3; ++d) acc[d] = fmaf(corner_dw(c.w, k, d), dot, acc[d]);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) atomicAdd(dx + 3 * s + d, scale * acc[d]);
  }
}

// double backward of the input gradient: given v = dL/d(dL/dx) (N,3)
//   grad_table[corner] += sum_d v_d * scale * dw_corner/dw_d * dy_f
//   dL/d(dy)_f          = sum_d v_d * scale * sum_corner dw_corner/dw_d * table[corner]_f
template <int F>
__global__ void __launch_bounds__(256)
grid_bwd_bwd_input_kernel(const __grid_constant__ GridMeta meta, const float* __restrict__ x,
                          const __half* __restrict__ table, const float* __restrict__ v,
                          const __half* __restrict__ dy, int64_t n_cap, const int32_t* __restrict__ n_dev,
                          GridXform xf, float* __restrict__ grad, __half* __restrict__ ddy) {
  NCN_GRID_PROLOGUE
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t s = i / L;
    const int l = (int)(i - s * L);
    Cell8 c;
    const float scale = sm.scale[l];
    locate8(x, s, scale, sm.res[l], sm.size[l], c, xf);
    const float v0 = v[3 * s] * scale, v1 = v[3 * s + 1] * scale, v2 = v[3 * s + 2] * scale;
    float g[F], o[F];
#pragma unroll
    for (int f = 0; f < F; ++f) { g[f] = dy ? __half2float(dy[i * F + f]) : 0.f; o[f] = 0.f; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float coef = v0 * corner_dw(c.w, k, 0) + v1 * corner_dw(c.w, k, 1) + v2 * corner_dw(c.w, k, 2);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        if (grad) atomicAdd(grad + ((size_t)sm.offset[l] + c.idx[k]) * F + f, coef * g[f]);
        if (ddy) o[f] = fmaf(coef, __half2float(__ldg(table + ((size_t)sm.offset[l] + c.idx[k]) * F + f)), o[f]);
      }
    }
    if (ddy) {
#pragma unroll
      for (int f = 0; f < F; ++f) ddy[i * F + f] = __float2half_rn(o[f]);
    }
  }
}

}  // namespace ncn

using namespace ncn;

extern "C" int64_t ncn_grid_desc_init(ncn_grid_desc* d) {
  if (!d || d->n_levels < 1 || d->n_levels > NCN_GRID_MAX_LEVELS) return -1;
  if (d->n_features != 1 && d->n_features != 2 && d->n_features != 4 && d->n_features != 8) return -1;
  if (d->log2_hashmap_size < 1 || d->log2_hashmap_size > 28 || d->base_resolution < 1) return -1;
  // tcnn evaluates log2f / exp2f in fp32; libm implementations may differ in the last ulp between
  // hosts, so both are evaluated in double and rounded once (== the correctly rounded fp32 result).
  const float log2b = (float)log2((double)d->per_level_scale);
  uint32_t off = 0;
  const uint32_t T = 1u << d->log2_hashmap_size;
  for (int l = 0; l < d->n_levels; ++l) {
    const float p2 = (float)exp2((double)((float)l * log2b));
    volatile float scaled = p2 * (float)d->base_resolution;
    const float scale = scaled - 1.0f;
    const uint32_t res = (uint32_t)ceilf(scale) + 1u;
    // entries of the level: res^3 rounded up to a multiple of 8, capped at T
    const double cube = (double)res * res * res;
    uint32_t size = cube > (double)T ? T : (uint32_t)cube;
    size = (size + 7u) / 8u * 8u;
    if (size > T) size = T;
    d->level_scale[l] = scale; d->level_res[l] = res; d->level_size[l] = size; d->level_offset[l] = off;
    off += size;
  }
  d->level_offset[d->n_levels] = off;
  for (int l = d->n_levels; l < NCN_GRID_MAX_LEVELS; ++l) { d->level_scale[l] = 0; d->level_res[l] = 0; d->level_size[l] = 0; if (l > d->n_levels) d->level_offset[l] = off; }
  return (int64_t)off * d->n_features;
}

static int to_meta(const ncn_grid_desc* d, GridMeta* m) {
  if (!d) return NCN_E_NULL;
  if (d->n_levels < 1 || d->n_levels > NCN_GRID_MAX_LEVELS) return NCN_E_CONFIG;
  for (int l = 0; l < NCN_GRID_MAX_LEVELS; ++l) {
    m->scale[l] = d->level_scale[l]; m->res[l] = d->level_res[l]; m->size[l] = d->level_size[l];
    m->offset[l] = d->level_offset[l];
    if (l < d->n_levels && d->level_size[l] == 0) return NCN_E_CONFIG;
  }
  m->n_levels = d->n_levels;
  return NCN_OK;
}

static GridXform make_xform(const float* xform_host) {
  GridXform xf;
  for (int d = 0; d < 3; ++d) { xf.lo[d] = xform_host ? xform_host[d] : 0.f; xf.size[d] = xform_host ? xform_host[3 + d] : 1.f; }
  xf.on = xform_host != nullptr;
  return xf;
}

static int g_grid_fwd_coherent = 1; // 1 = warp-per-(32 samples, level) forward with smem transposition (default), 0 = thread-per-(sample, level)
extern "C" int ncn_set_grid_fwd_coherent(int on) { const int old = g_grid_fwd_coherent; g_grid_fwd_coherent = on; return old; }
static int g_grid_bwd_merge = 1;    // 1 = warp-level run merging before the scatter (default), 0 = one reduction per corner
extern "C" int ncn_set_grid_bwd_merge(int on) { const int old = g_grid_bwd_merge; g_grid_bwd_merge = on; return old; }

#define NCN_GRID_DISPATCH(F, CALL)                 \
  switch (F) {                                     \
    case 1: { constexpr int kF = 1; CALL; } break; \
    case 2: { constexpr int kF = 2; CALL; } break; \
    case 4: { constexpr int kF = 4; CALL; } break; \
    case 8: { constexpr int kF = 8; CALL; } break; \
    default: return NCN_E_CONFIG;                  \
  }

extern "C" int ncn_grid_fwd(const ncn_grid_desc* desc, const float* x, const void* table, int64_t n, void* out,
                            const float* xform_host, const int32_t* n_dev, ncn_stream_t stream) {
  GridMeta m; int rc = to_meta(desc, &m); if (rc) return rc;
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(x); NCN_CHECK_PTR(table); NCN_CHECK_PTR(out);
  if (((uintptr_t)table | (uintptr_t)out) & 3) return NCN_E_ALIGN;
  if (desc->n_features == 2 && m.n_levels == 16 && g_grid_fwd_coherent) {
    static int resident = 0;
    const int cg = resident_grid(grid_fwd_coherent_kernel, 256, 0, &resident, (n + 31) / 32);
    NCN_CUDA(launch_pdl(grid_fwd_coherent_kernel, dim3(cg), dim3(256), 0, as_stream(stream), m, x, (const __half*)table, n, n_dev,
                        make_xform(xform_host), (__half*)out));
    NCN_LAUNCH_OK();
    return NCN_OK;
  }
  const int grid = persistent_grid(n * m.n_levels, 256, 8);
  NCN_GRID_DISPATCH(desc->n_features, (grid_fwd_kernel<kF><<<grid, 256, 0, as_stream(stream)>>>(
      m, x, (const __half*)table, n, n_dev, make_xform(xform_host), (__half*)out)));
  NCN_LAUNCH_OK();
  return NCN_OK;
}

// level_begin/level_end restrict the launch to a range of levels (their gradient regions are contiguous in the table), so a
// data-parallel caller can all-reduce the first range while the second is still being computed; ctas_per_sm < 8 leaves
// room on the SMs for the collective's kernels
static int g_grid_bwd_occ = 5;   // developer A/B knob: CTAs per SM the merge kernel is compiled for (5 or 6)
extern "C" int ncn_set_grid_bwd_occupancy(int ctas_per_sm) {
  const int old = g_grid_bwd_occ;
  if (ctas_per_sm == 5 || ctas_per_sm == 6) g_grid_bwd_occ = ctas_per_sm;
  return old;
}

static int grid_bwd_impl(const ncn_grid_desc* desc, const float* x, const void* dy, int64_t n, void* grad, bool half_grad,
                         float grad_scale, const float* xform_host, const int32_t* n_dev, int level_begin, int level_end,
                         int ctas_per_sm, ncn_stream_t stream) {
  GridMeta m; int rc = to_meta(desc, &m); if (rc) return rc;
  NCN_CHECK_SIZE(n >= 0);
  if (level_begin < 0 || level_end > m.n_levels || level_begin >= level_end || ctas_per_sm < 1 || ctas_per_sm > 8) return NCN_E_CONFIG;
  if (half_grad && desc->n_features != 2) return NCN_E_UNSUPPORTED;
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(x); NCN_CHECK_PTR(dy); NCN_CHECK_PTR(grad);
  if ((uintptr_t)grad & 7) return NCN_E_ALIGN;
  const bool whole = level_begin == 0 && level_end == m.n_levels;
  const int64_t need = ceil_div(n * (level_end - level_begin), (int64_t)256);
  if (half_grad || (desc->n_features == 2 && (g_grid_bwd_merge || !whole) && ((uintptr_t)grad & 15) == 0)) {
    // fp32 table: 16-byte paired reductions; fp16 table: 8-byte ones.  CTAs per SM: 5 (46 registers) or 6 (40, 12 B spilled)
    static int resident[2][2] = {{0, 0}, {0, 0}};
    const int o6 = g_grid_bwd_occ == 6 ? 1 : 0;
#define NCN_GBM(H, O)                                                                                                              \
  do {                                                                                                                             \
    int grid = resident_grid(grid_bwd_merge_kernel<H, O>, 256, 0, &resident[H ? 1 : 0][o6], need);                                 \
    if (ctas_per_sm < 8 && grid > sm_count() * ctas_per_sm) grid = sm_count() * ctas_per_sm;                                       \
    NCN_CUDA(launch_pdl(grid_bwd_merge_kernel<H, O>, dim3(grid), dim3(256), 0, as_stream(stream), m, x, (const __half*)dy, n, n_dev, \
                        make_xform(xform_host), grad_scale, grad, level_begin, level_end));                                        \
  } while (0)
    if (half_grad) { if (o6) NCN_GBM(true, 6); else NCN_GBM(true, 5); }
    else { if (o6) NCN_GBM(false, 6); else NCN_GBM(false, 5); }
#undef NCN_GBM
    return NCN_OK;
  }
  if (!whole) return NCN_E_UNSUPPORTED;
  const int grid = persistent_grid(n * m.n_levels, 256, 8);
  NCN_GRID_DISPATCH(desc->n_features, (grid_bwd_kernel<kF><<<grid, 256, 0, as_stream(stream)>>>(
      m, x, (const __half*)dy, n, n_dev, make_xform(xform_host), grad_scale, (float*)grad)));
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_grid_bwd(const ncn_grid_desc* desc, const float* x, const void* dy, int64_t n, float* grad,
                            float grad_scale, const float* xform_host, const int32_t* n_dev, ncn_stream_t stream) {
  if (!desc) return NCN_E_NULL;
  return grid_bwd_impl(desc, x, dy, n, grad, false, grad_scale, xform_host, n_dev, 0, desc->n_levels, 8, stream);
}

extern "C" int ncn_grid_bwd_levels(const ncn_grid_desc* desc, const float* x, const void* dy, int64_t n, float* grad,
                                   float grad_scale, const float* xform_host, const int32_t* n_dev, int level_begin, int level_end,
                                   int ctas_per_sm, ncn_stream_t stream) {
  if (!desc) return NCN_E_NULL;
  return grid_bwd_impl(desc, x, dy, n, grad, false, grad_scale, xform_host, n_dev, level_begin, level_end, ctas_per_sm, stream);
}

extern "C" int ncn_grid_bwd_f16(const ncn_grid_desc* desc, const float* x, const void* dy, int64_t n, void* grad_f16,
                                float grad_scale, const float* xform_host, const int32_t* n_dev, ncn_stream_t stream) {
  if (!desc) return NCN_E_NULL;
  return grid_bwd_impl(desc, x, dy, n, grad_f16, true, grad_scale, xform_host, n_dev, 0, desc->n_levels, 8, stream);
}

extern "C" int ncn_grid_bwd_input(const ncn_grid_desc* desc, const float* x, const void* table, const void* dy,
                                  int64_t n, float* dx, ncn_stream_t stream) {
  GridMeta m; int rc = to_meta(desc, &m); if (rc) return rc;
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(x); NCN_CHECK_PTR(table); NCN_CHECK_PTR(dy); NCN_CHECK_PTR(dx);
  NCN_CUDA(cudaMemsetAsync(dx, 0, (size_t)n * 3 * sizeof(float), as_stream(stream)));
  const int grid = persistent_grid(n * m.n_levels, 256, 8);
  NCN_GRID_DISPATCH(desc->n_features, (grid_bwd_input_kernel<kF><<<grid, 256, 0, as_stream(stream)>>>(
      m, x, (const __half*)table, (const __half*)dy, n, nullptr, make_xform(nullptr), dx)));
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_grid_bwd_bwd_input(const ncn_grid_desc* desc, const float* x, const void* table,
                                      const float* dL_ddLdx, const void* dy, int64_t n, float* grad, void* ddy,
                                      ncn_stream_t stream) {
  GridMeta m; int rc = to_meta(desc, &m); if (rc) return rc;
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0 || (!grad && !ddy)) return NCN_OK;
  NCN_CHECK_PTR(x); NCN_CHECK_PTR(table); NCN_CHECK_PTR(dL_ddLdx);
  if (grad) NCN_CHECK_PTR(dy);
  const int grid = persistent_grid(n * m.n_levels, 256, 8);
  NCN_GRID_DISPATCH(desc->n_features, (grid_bwd_bwd_input_kernel<kF><<<grid, 256, 0, as_stream(stream)>>>(
      m, x, (const __half*)table, dL_ddLdx, (const __half*)dy, n, nullptr, make_xform(nullptr), grad, (__half*)ddy)));
  NCN_LAUNCH_OK();
  return NCN_OK;
}
