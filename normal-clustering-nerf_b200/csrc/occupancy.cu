// Occupancy-grid integer kernels: Morton encode / decode, bit packing and the fused
// decay/max grid update.  Replaces raymarching.cu:35-161 of the reference (bit-exact).
//
// B200 notes: all three are pure streaming passes (HBM/L2 bound), so the work is
// 128-bit vectorised and launched as a grid-stride loop over <= 8 CTAs/SM.
//   morton3d        : 12 B in / 4 B out per cell  (3x int4 -> 1x int4 per 4 cells)
//   morton3d_invert :  4 B in / 12 B out
//   packbits        : 32 B in / 1 B out per byte; one thread packs 32 cells -> 1 u32
#include "ncn_common.cuh"
#include "morton.cuh"

namespace ncn {

__global__ void __launch_bounds__(256)
morton3d_kernel(const int32_t* __restrict__ coords, int64_t n, int32_t* __restrict__ indices) {
  // 4 cells per thread-iteration: three int4 loads (48 B) -> one int4 store (16 B)
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int4* p = reinterpret_cast<const int4*>(coords) + 3 * i;
    const int4 a = __ldcs(p), b = __ldcs(p + 1), c = __ldcs(p + 2);
    int4 o;
    o.x = (int32_t)morton3d((uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z);
    o.y = (int32_t)morton3d((uint32_t)a.w, (uint32_t)b.x, (uint32_t)b.y);
    o.z = (int32_t)morton3d((uint32_t)b.z, (uint32_t)b.w, (uint32_t)c.x);
    o.w = (int32_t)morton3d((uint32_t)c.y, (uint32_t)c.z, (uint32_t)c.w);
    __stcs(reinterpret_cast<int4*>(indices) + i, o);
  }
  // tail (< 4 cells)
  const int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) indices[t] = (int32_t)morton3d((uint32_t)coords[3 * t], (uint32_t)coords[3 * t + 1],
                                            (uint32_t)coords[3 * t + 2]);
}

__global__ void __launch_bounds__(256)
morton3d_invert_kernel(const int32_t* __restrict__ indices, int64_t n, int32_t* __restrict__ coords) {
  // NOTE: the reference shifts the SIGNED index (`const int ind; ind >> k`, raymarching.cu:97-100),
  // i.e. an arithmetic shift; for negative inputs that differs from a logical shift in the
  // masked bit 30, so the signed shift is kept for bit-exactness on any input.
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int4 v = __ldcs(reinterpret_cast<const int4*>(indices) + i);
    int4 a, b, c;
    a.x = (int32_t)morton3d_invert((uint32_t)(v.x >> 0)); a.y = (int32_t)morton3d_invert((uint32_t)(v.x >> 1));
    a.z = (int32_t)morton3d_invert((uint32_t)(v.x >> 2)); a.w = (int32_t)morton3d_invert((uint32_t)(v.y >> 0));
    b.x = (int32_t)morton3d_invert((uint32_t)(v.y >> 1)); b.y = (int32_t)morton3d_invert((uint32_t)(v.y >> 2));
    b.z = (int32_t)morton3d_invert((uint32_t)(v.z >> 0)); b.w = (int32_t)morton3d_invert((uint32_t)(v.z >> 1));
    c.x = (int32_t)morton3d_invert((uint32_t)(v.z >> 2)); c.y = (int32_t)morton3d_invert((uint32_t)(v.w >> 0));
    c.z = (int32_t)morton3d_invert((uint32_t)(v.w >> 1)); c.w = (int32_t)morton3d_invert((uint32_t)(v.w >> 2));
    int4* q = reinterpret_cast<int4*>(coords) + 3 * i;
    __stcs(q, a); __stcs(q + 1, b); __stcs(q + 2, c);
  }
  const int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    const int32_t ind = indices[t];
    coords[3 * t + 0] = (int32_t)morton3d_invert((uint32_t)(ind >> 0));
    coords[3 * t + 1] = (int32_t)morton3d_invert((uint32_t)(ind >> 1));
    coords[3 * t + 2] = (int32_t)morton3d_invert((uint32_t)(ind >> 2));
  }
}

// One thread packs 32 consecutive cells (8x float4 = 128 B) into one u32 (4 bytes).
__device__ __forceinline__ uint32_t pack32(const float* __restrict__ g, float thr) {
  uint32_t bits = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(g) + k);
    bits |= (v.x > thr ? 1u : 0u) << (4 * k + 0);
    bits |= (v.y > thr ? 1u : 0u) << (4 * k + 1);
    bits |= (v.z > thr ? 1u : 0u) << (4 * k + 2);
    bits |= (v.w > thr ? 1u : 0u) << (4 * k + 3);
  }
  return bits;  // little-endian: byte b holds cells 8b..8b+7, LSB first (raymarching.cu:136-138)
}

__device__ __forceinline__ void packbits_body(const float* __restrict__ grid, int64_t n_bytes,
                                              float thr, uint8_t* __restrict__ bitfield) {
  const int64_t n_words = n_bytes >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride)
    reinterpret_cast<uint32_t*>(bitfield)[w] = pack32(grid + 32 * w, thr);
  // tail bytes (n_bytes not a multiple of 4)
  const int64_t b = (n_words << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n_bytes) {
    uint8_t bits = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) bits |= (grid[8 * b + i] > thr) ? (uint8_t)(1u << i) : (uint8_t)0;
    bitfield[b] = bits;
  }
}

__global__ void __launch_bounds__(256)
packbits_kernel(const float* __restrict__ grid, int64_t n_bytes, float thr, uint8_t* __restrict__ bitfield) {
  packbits_body(grid, n_bytes, thr, bitfield);
}

// threshold = min(mean(grid[grid>0]), density_threshold), the mean read from device stats
// (ngp_mt.py:365-367 without the .item()).  mean of an empty set is NaN in the reference and
// python's min(nan, thr) returns nan -> nothing is > nan -> all bits 0; reproduce that.
__global__ void __launch_bounds__(256)
packbits_auto_kernel(const float* __restrict__ grid, int64_t n_bytes, const float* __restrict__ stats,
                     float density_threshold, uint8_t* __restrict__ bitfield) {
  const float mean = stats[0] / stats[1];  // 0/0 -> NaN like torch's empty mean
  // python: min(mean, thr) == thr if thr < mean else mean  (NaN mean -> NaN)
  const float thr = (density_threshold < mean) ? density_threshold : mean;
  packbits_body(grid, n_bytes, thr, bitfield);
}

__global__ void __launch_bounds__(256)
density_grid_update_kernel(float* __restrict__ grid, const float* __restrict__ tmp, int64_t n,
                           float decay, float* __restrict__ stats) {
  float sum = 0.f, cnt = 0.f;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 g = reinterpret_cast<float4*>(grid)[i];
    const float4 t = __ldcs(reinterpret_cast<const float4*>(tmp) + i);
    g.x = g.x < 0.f ? g.x : fmaxf(__fmul_rn(g.x, decay), t.x);
    g.y = g.y < 0.f ? g.y : fmaxf(__fmul_rn(g.y, decay), t.y);
    g.z = g.z < 0.f ? g.z : fmaxf(__fmul_rn(g.z, decay), t.z);
    g.w = g.w < 0.f ? g.w : fmaxf(__fmul_rn(g.w, decay), t.w);
    reinterpret_cast<float4*>(grid)[i] = g;
    if (g.x > 0.f) { sum += g.x; cnt += 1.f; }
    if (g.y > 0.f) { sum += g.y; cnt += 1.f; }
    if (g.z > 0.f) { sum += g.z; cnt += 1.f; }
    if (g.w > 0.f) { sum += g.w; cnt += 1.f; }
  }
  const int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    float g = grid[t];
    g = g < 0.f ? g : fmaxf(__fmul_rn(g, decay), tmp[t]);
    grid[t] = g;
    if (g > 0.f) { sum += g; cnt += 1.f; }
  }
  sum = warp_sum(sum); cnt = warp_sum(cnt);
  __shared__ float s_sum[8], s_cnt[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_sum[wid] = sum; s_cnt[wid] = cnt; }
  __syncthreads();
  if (wid == 0) {
    sum = lane < (blockDim.x >> 5) ? s_sum[lane] : 0.f;
    cnt = lane < (blockDim.x >> 5) ? s_cnt[lane] : 0.f;
    sum = warp_sum(sum); cnt = warp_sum(cnt);
    if (lane == 0) { atomicAdd(stats, sum); atomicAdd(stats + 1, cnt); }
  }
}


// ---- occupancy-grid upkeep, sampling half (models/ngp_mt.py:254-271, 339-357) ---------------------------------------
// One launch replaces randint + morton3D + cumsum lookup (searchsorted) + morton3D_invert + two cats + the jittered
// cell-centre arithmetic + the (M,3) temporaries of the reference's sample_uniform_and_occupied_cells /
// update_density_grid: thread i < M draws a uniform cell, thread M <= i < 2M draws the k-th occupied cell for a uniform k
// (upper_bound over the inclusive cumsum of the occupied flags = torch.searchsorted(right=True), clamped like the
// reference), and both emit the Morton index and a uniformly jittered point inside the cell.
__device__ __forceinline__ uint32_t rng_hash(uint64_t seed, uint64_t i, uint32_t k) {      // splitmix64 finaliser
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (i * 8ull + k + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}
__device__ __forceinline__ float rng_unit(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }   // [0, 1)

__global__ void __launch_bounds__(256)
grid_sample_cells_kernel(const int32_t* __restrict__ occ_csum, int G, int64_t M, float s, const int64_t* __restrict__ seed_dev,
                         int32_t* __restrict__ indices, float* __restrict__ xyz) {
  const int64_t G3 = (int64_t)G * G * G;
  const uint64_t seed = (uint64_t)*seed_dev;
  const float half_cell = s / (float)G;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * M; i += stride) {
    uint32_t cx, cy, cz, idx;
    if (i < M) {
      cx = rng_hash(seed, i, 0) % (uint32_t)G; cy = rng_hash(seed, i, 1) % (uint32_t)G; cz = rng_hash(seed, i, 2) % (uint32_t)G;
      idx = morton3d(cx, cy, cz);
    } else {
      const int32_t total = occ_csum[G3 - 1];
      const int32_t r = (int32_t)(rng_unit(rng_hash(seed, i, 0)) * (float)total);
      int64_t lo = 0, hi = G3;                               // first position with csum > r
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (occ_csum[mid] > r) hi = mid; else lo = mid + 1;
      }
      idx = (uint32_t)(lo < G3 - 1 ? lo : G3 - 1);
      cx = morton3d_invert(idx); cy = morton3d_invert(idx >> 1); cz = morton3d_invert(idx >> 2);
    }
    indices[i] = (int32_t)idx;
    const uint32_t c3[3] = {cx, cy, cz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float centre = ((float)c3[d] / (float)(G - 1) * 2.0f - 1.0f) * (s - half_cell);
      xyz[3 * i + d] = centre + (rng_unit(rng_hash(seed, i, 3 + d)) * 2.0f - 1.0f) * half_cell;
    }
  }
}

// density_tmp[indices[i]] = exp(h[i, 0])  (TruncExp forward of the density head; duplicate cells: any one of the draws,
// like the reference's index_put); also advances the sampling seed once per launch
__global__ void __launch_bounds__(256)
grid_scatter_density_kernel(const __half* __restrict__ h, int h_stride, const int32_t* __restrict__ indices, int64_t n,
                            float* __restrict__ density_tmp, int64_t* __restrict__ seed_dev) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    density_tmp[indices[i]] = expf(__half2float(h[i * h_stride]));
  if (seed_dev != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *seed_dev += 1;
}

}  // namespace ncn

using namespace ncn;

extern "C" int ncn_morton3d(const int32_t* coords, int64_t n, int32_t* indices, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(coords); NCN_CHECK_PTR(indices);
  if (((uintptr_t)coords | (uintptr_t)indices) & 15) return NCN_E_ALIGN;
  const int grid = persistent_grid((n + 3) / 4, 256, 8);
  morton3d_kernel<<<grid, 256, 0, as_stream(stream)>>>(coords, n, indices);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_morton3d_invert(const int32_t* indices, int64_t n, int32_t* coords, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(coords); NCN_CHECK_PTR(indices);
  if (((uintptr_t)coords | (uintptr_t)indices) & 15) return NCN_E_ALIGN;
  const int grid = persistent_grid((n + 3) / 4, 256, 8);
  morton3d_invert_kernel<<<grid, 256, 0, as_stream(stream)>>>(indices, n, coords);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_packbits(const float* density_grid, int64_t n_bytes, float threshold,
                            uint8_t* density_bitfield, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_bytes >= 0);
  if (n_bytes == 0) return NCN_OK;
  NCN_CHECK_PTR(density_grid); NCN_CHECK_PTR(density_bitfield);
  if (((uintptr_t)density_grid & 15) || ((uintptr_t)density_bitfield & 3)) return NCN_E_ALIGN;
  const int grid = persistent_grid((n_bytes + 3) / 4, 256, 8);
  packbits_kernel<<<grid, 256, 0, as_stream(stream)>>>(density_grid, n_bytes, threshold, density_bitfield);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_packbits_auto(const float* density_grid, int64_t n_bytes, const float* stats,
                                 float density_threshold, uint8_t* density_bitfield, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_bytes >= 0);
  if (n_bytes == 0) return NCN_OK;
  NCN_CHECK_PTR(density_grid); NCN_CHECK_PTR(density_bitfield); NCN_CHECK_PTR(stats);
  if (((uintptr_t)density_grid & 15) || ((uintptr_t)density_bitfield & 3)) return NCN_E_ALIGN;
  const int grid = persistent_grid((n_bytes + 3) / 4, 256, 8);
  packbits_auto_kernel<<<grid, 256, 0, as_stream(stream)>>>(density_grid, n_bytes, stats,
                                                            density_threshold, density_bitfield);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_density_grid_update(float* density_grid, const float* density_tmp, int64_t n_cells,
                                       float decay, float* stats, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_cells >= 0);
  if (n_cells == 0) return NCN_OK;
  NCN_CHECK_PTR(density_grid); NCN_CHECK_PTR(density_tmp); NCN_CHECK_PTR(stats);
  if (((uintptr_t)density_grid | (uintptr_t)density_tmp) & 15) return NCN_E_ALIGN;
  const int grid = persistent_grid((n_cells + 3) / 4, 256, 4);
  density_grid_update_kernel<<<grid, 256, 0, as_stream(stream)>>>(density_grid, density_tmp, n_cells,
                                                                  decay, stats);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_grid_sample_cells(const int32_t* occ_csum, int grid_size, int64_t m, float s, const int64_t* seed_dev,
                                     int32_t* indices, float* xyz, ncn_stream_t stream) {
  NCN_CHECK_SIZE(m >= 0 && grid_size >= 2 && grid_size <= 1024);
  if (m == 0) return NCN_OK;
  NCN_CHECK_PTR(occ_csum); NCN_CHECK_PTR(seed_dev); NCN_CHECK_PTR(indices); NCN_CHECK_PTR(xyz);
  grid_sample_cells_kernel<<<persistent_grid(2 * m, 256, 8), 256, 0, as_stream(stream)>>>(occ_csum, grid_size, m, s, seed_dev, indices, xyz);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_grid_scatter_density(const void* h_f16, int h_stride, const int32_t* indices, int64_t n, float* density_tmp,
                                        int64_t* seed_dev, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0 && h_stride >= 1);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(h_f16); NCN_CHECK_PTR(indices); NCN_CHECK_PTR(density_tmp);
  grid_scatter_density_kernel<<<persistent_grid(n, 256, 8), 256, 0, as_stream(stream)>>>((const __half*)h_f16, h_stride, indices, n,
                                                                                      density_tmp, seed_dev);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
