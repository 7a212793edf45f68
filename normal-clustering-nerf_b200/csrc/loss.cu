// Normals from rendered depth + Manhattan normal-clustering loss, on the GPU with no host
// round trip.  Replaces
//   datasets/hypersim_src/utils.py:505-541  (_extract_normals_from_ray_batch)
//   losses.py:86-92     faiss.Kmeans(3, K, niter, spherical=True) on the CPU + D2H/H2D + syncs
//   losses.py:97-166    orthogonal-triple selection, merge, opposite labelling (4 .item() syncs)
//   losses.py:441-478   per-cluster means, L_ort / L_dot / L_L1 and (through autograd) their gradient
// plus the photometric terms of losses.py:347-361 fused with the background composite of
// models/rendering.py:231-241.
//
// The clustering problem is tiny (M <= 6272 3-vectors, K = 20, 20 iterations): it is latency
// bound, so it runs as ONE persistent CTA (1024 threads) that keeps the centroids in shared
// memory and iterates without ever leaving the SM.  Cluster sums are accumulated in 64-bit
// fixed point (2^-30 resolution) so the result is independent of the atomic ordering
// (bit-reproducible run to run, unlike float atomics).
#include "ncn_common.cuh"

namespace ncn {

// ---------------------------------------------------------------- normals from depth
struct Tri { float p[3][3]; };

__global__ void __launch_bounds__(256)
normals_fw_kernel(const float* __restrict__ origin, const float* __restrict__ dir, const float* __restrict__ depth,
                  const int64_t* __restrict__ i1, const int64_t* __restrict__ i2, const int64_t* __restrict__ i3,
                  int64_t n_tri, float* __restrict__ normals) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < n_tri; m += stride) {
    const int64_t idx[3] = {i1[m], i2[m], i3[m]};
    float P[3][3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const float dep = depth[idx[v]];
#pragma unroll
      for (int c = 0; c < 3; ++c) P[v][c] = __fadd_rn(origin[3 * idx[v] + c], __fmul_rn(dir[3 * idx[v] + c], dep));
    }
    float a[3], b[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { a[c] = P[1][c] - P[0][c]; b[c] = P[2][c] - P[0][c]; }
    const float cx = a[1] * b[2] - a[2] * b[1];
    const float cy = a[2] * b[0] - a[0] * b[2];
    const float cz = a[0] * b[1] - a[1] * b[0];
    const float nrm = fmaxf(sqrtf(cx * cx + cy * cy + cz * cz), 1e-12f);   // F.normalize eps
    normals[3 * m] = cx / nrm; normals[3 * m + 1] = cy / nrm; normals[3 * m + 2] = cz / nrm;
  }
}

__global__ void __launch_bounds__(256)
normals_bw_kernel(const float* __restrict__ origin, const float* __restrict__ dir, const float* __restrict__ depth,
                  const int64_t* __restrict__ i1, const int64_t* __restrict__ i2, const int64_t* __restrict__ i3,
                  const float* __restrict__ dn, int64_t n_tri, float* __restrict__ ddepth) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < n_tri; m += stride) {
    const float g[3] = {dn[3 * m], dn[3 * m + 1], dn[3 * m + 2]};
    if (g[0] == 0.f && g[1] == 0.f && g[2] == 0.f) continue;
    const int64_t idx[3] = {i1[m], i2[m], i3[m]};
    float P[3][3], D[3][3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const float dep = depth[idx[v]];
#pragma unroll
      for (int c = 0; c < 3; ++c) { D[v][c] = dir[3 * idx[v] + c]; P[v][c] = __fadd_rn(origin[3 * idx[v] + c], __fmul_rn(D[v][c], dep)); }
    }
    float a[3], b[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { a[c] = P[1][c] - P[0][c]; b[c] = P[2][c] - P[0][c]; }
    const float cr[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const float len = sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
    float gc[3];     // dL/dcross
    if (len > 1e-12f) {
      const float inv = 1.0f / len;
      const float n[3] = {cr[0] * inv, cr[1] * inv, cr[2] * inv};
      const float gn = g[0] * n[0] + g[1] * n[1] + g[2] * n[2];
#pragma unroll
      for (int c = 0; c < 3; ++c) gc[c] = (g[c] - n[c] * gn) * inv;
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) gc[c] = g[c] * 1e12f;
    }
    // cross = a x b:  dL/da = b x gc,  dL/db = gc x a
    const float ga[3] = {b[1] * gc[2] - b[2] * gc[1], b[2] * gc[0] - b[0] * gc[2], b[0] * gc[1] - b[1] * gc[0]};
    const float gb[3] = {gc[1] * a[2] - gc[2] * a[1], gc[2] * a[0] - gc[0] * a[2], gc[0] * a[1] - gc[1] * a[0]};
    // a = P2 - P1, b = P3 - P1, dP_v/ddepth_v = dir_v
    float g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) { g1 -= (ga[c] + gb[c]) * D[0][c]; g2 += ga[c] * D[1][c]; g3 += gb[c] * D[2][c]; }
    atomicAdd(ddepth + idx[0], g1); atomicAdd(ddepth + idx[1], g2); atomicAdd(ddepth + idx[2], g3);
  }
}

// ---------------------------------------------------------------- spherical k-means (one CTA)
constexpr int kKmThreads = 1024;
constexpr int kKmMaxK = 64;
constexpr double kFix = 1073741824.0;   // 2^30 fixed point for order-independent (reproducible) sums

__device__ __forceinline__ bool valid_normal(float x, float y, float z) {
  // losses.py:427-429: drop rows that are all zero / contain NaN / contain Inf
  const float s = fabsf(x) + fabsf(y) + fabsf(z);
  return (s != 0.f) && isfinite(x) && isfinite(y) && isfinite(z);
}

__device__ __forceinline__ int best_centroid(float x, float y, float z, const float* __restrict__ c, int k) {
  int best = 0; float bs = -INFINITY;
  for (int j = 0; j < k; ++j) {
    const float s = x * c[3 * j] + y * c[3 * j + 1] + z * c[3 * j + 2];
    if (s > bs) { bs = s; best = j; }     // ties -> lowest index
  }
  return best;
}

// Shared-memory plan (dynamic): training points xs[nt][3] (<= 256*K points, 60 KB at K=20).  Every Lloyd
// iteration is (1) thread-per-point assignment (K dot products against the centroids in smem), (2) a WARP-
// AGGREGATED integer reduction: the warp walks the distinct clusters present among its 32 points and folds
// each cluster's members with redux.sync (fixed point 2^20, so the sums are order independent and the result is
// bit-reproducible), one lane adds the warp total to the warp's PRIVATE accumulator row - no atomics, no
// contention - (3) K*4 threads fold the 32 warp rows, and one warp updates / re-seeds / renormalises.
constexpr float kKmFix = 1048576.0f;     // 2^20
constexpr int kAccStride = 33;

// adds (x,y,z,1) of every valid lane to acc[key*4 + {0,1,2,3}], a WARP-PRIVATE row of integer accumulators:
// integer atomics commute, so the result does not depend on the order in which the lanes are served
__device__ __forceinline__ void warp_accumulate_by_key(int key, bool valid, float x, float y, float z, int* __restrict__ acc,
                                                       int lane) {
  (void)lane;
  // acc points at column `warp` of a [4*K][kAccStride] matrix (stride 33 words: conflict-free both for these atomics,
  // whose addresses differ in the row, and for the fold below, which reads along a row)
  if (valid) {
    atomicAdd(acc + (4 * key) * kAccStride, __float2int_rn(x * kKmFix));
    atomicAdd(acc + (4 * key + 1) * kAccStride, __float2int_rn(y * kKmFix));
    atomicAdd(acc + (4 * key + 2) * kAccStride, __float2int_rn(z * kKmFix));
    atomicAdd(acc + (4 * key + 3) * kAccStride, 1);
  }
}

__global__ void __launch_bounds__(kKmThreads, 1)
kmeans_kernel(const float* __restrict__ x, int64_t n, ncn_kmeans_params p, float* __restrict__ centroids,
              int32_t* __restrict__ assign, int32_t* __restrict__ n_valid_out, int32_t* __restrict__ valid_idx,
              int nt_cap) {
  extern __shared__ __align__(16) unsigned char km_smem[];
  float* xs = reinterpret_cast<float*>(km_smem);                 // [nt_cap][3]
  __shared__ float s_c[kKmMaxK * 3];
  __shared__ int s_wacc[kKmMaxK * 4 * kAccStride];   // [accumulator][warp] (stride 33), fixed point
  __shared__ float s_acc[kKmMaxK * 4];
  __shared__ int s_nvalid, s_warp_tot[32], s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int K = p.k;
  // 1) compact the valid rows (stable order) into valid_idx
  if (tid == 0) { s_base = 0; }
  __syncthreads();
  for (int64_t b0 = 0; b0 < n; b0 += kKmThreads) {
    const int64_t i = b0 + tid;
    bool v = false;
    if (i < n) v = valid_normal(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    if (lane == 0) s_warp_tot[wid] = __popc(bal);
    __syncthreads();
    if (wid == 0) {
      const int w = s_warp_tot[lane];
      const int inc = warp_scan_incl_i(w, lane);
      s_warp_tot[lane] = inc - w;
      if (lane == 31) s_nvalid = inc;
    }
    __syncthreads();
    if (v) valid_idx[s_base + s_warp_tot[wid] + __popc(bal & ((1u << lane) - 1))] = (int32_t)i;
    if (i < n && !v) assign[i] = -1;
    __syncthreads();
    if (tid == 0) s_base += s_nvalid;
    __syncthreads();
  }
  const int nv = s_base;
  if (tid == 0) *n_valid_out = nv;
  if (nv == 0) {
    for (int j = tid; j < K * 3; j += kKmThreads) centroids[j] = 0.f;
    return;
  }
  // 2) training subset: at most max_points_per_centroid*K points, taken at a uniform stride over the valid rows
  //    (faiss draws a random subset; equality with faiss is not a parity criterion), staged in shared memory
  const int nt = nv > nt_cap ? nt_cap : nv;
  for (int j = tid; j < nt; j += kKmThreads) {
    const int r = valid_idx[(int)(((int64_t)j * nv) / nt)];
    xs[3 * j] = x[3 * r]; xs[3 * j + 1] = x[3 * r + 1]; xs[3 * j + 2] = x[3 * r + 2];
  }
  __syncthreads();
  // 3) init: K training points spread over the subset with a seeded offset
  if (tid < K) {
    const uint32_t h = (uint32_t)p.seed * 2654435761u + 12345u;
    const int span = nt / K > 0 ? nt / K : 1;
    const int j = (int)((((int64_t)tid * nt) / K + (h % (uint32_t)span)) % nt);
    float cx = xs[3 * j], cy = xs[3 * j + 1], cz = xs[3 * j + 2];
    if (p.spherical) { const float l = sqrtf(cx * cx + cy * cy + cz * cz); if (l > 0.f) { cx /= l; cy /= l; cz /= l; } }
    s_c[3 * tid] = cx; s_c[3 * tid + 1] = cy; s_c[3 * tid + 2] = cz;
  }
  __syncthreads();
  // 4) Lloyd iterations
  const int nt_round = (nt + 31) & ~31;
  for (int it = 0; it < p.niter; ++it) {
    for (int a = tid; a < K * 4 * kAccStride; a += kKmThreads) s_wacc[a] = 0;
    __syncthreads();
    if (K <= 32) {
      // centroids in registers: lane j holds centroid j, fetched by shuffle -> no shared-memory traffic in the dot loop
      const float cx = lane < K ? s_c[3 * lane] : 0.f, cy = lane < K ? s_c[3 * lane + 1] : 0.f, cz = lane < K ? s_c[3 * lane + 2] : 0.f;
      for (int j = tid; j < nt_round; j += kKmThreads) {
        const bool v = j < nt;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (v) { px = xs[3 * j]; py = xs[3 * j + 1]; pz = xs[3 * j + 2]; }
        int b = 0; float bs = -INFINITY;
        for (int q = 0; q < K; ++q) {
          const float sc = px * __shfl_sync(0xffffffffu, cx, q) + py * __shfl_sync(0xffffffffu, cy, q) + pz * __shfl_sync(0xffffffffu, cz, q);
          if (sc > bs) { bs = sc; b = q; }
        }
        warp_accumulate_by_key(b, v, px, py, pz, s_wacc + wid, lane);
      }
    } else {
      for (int j = tid; j < nt_round; j += kKmThreads) {
        const bool v = j < nt;
        float px = 0.f, py = 0.f, pz = 0.f;
        int b = 0;
        if (v) { px = xs[3 * j]; py = xs[3 * j + 1]; pz = xs[3 * j + 2]; b = best_centroid(px, py, pz, s_c, K); }
        warp_accumulate_by_key(b, v, px, py, pz, s_wacc + wid, lane);
      }
    }
    __syncthreads();
    for (int a = wid; a < K * 4; a += 32) {       // warp `wid` folds accumulator rows wid, wid+32, ...
      int part = s_wacc[a * kAccStride + lane];
      const int tot = __reduce_add_sync(0xffffffffu, part);
      if (lane == 0) s_acc[a] = (a & 3) == 3 ? (float)tot : (float)((double)tot / (double)kKmFix);
    }
    __syncthreads();
    // new centroids = member means (thread per cluster); empty clusters split the currently largest one
    // (faiss-style +-eps, rare -> one thread); then renormalise (thread per cluster)
    __shared__ int s_any_empty;
    if (tid == 0) s_any_empty = 0;
    __syncthreads();
    if (tid < K) {
      const float c = s_acc[4 * tid + 3];
      if (c > 0.f) { s_c[3 * tid] = s_acc[4 * tid] / c; s_c[3 * tid + 1] = s_acc[4 * tid + 1] / c; s_c[3 * tid + 2] = s_acc[4 * tid + 2] / c; }
      else s_any_empty = 1;
    }
    __syncthreads();
    if (tid == 0 && s_any_empty) {
      for (int j = 0; j < K; ++j) {
        if (s_acc[4 * j + 3] == 0.f) {
          int big = 0;
          for (int q = 1; q < K; ++q) if (s_acc[4 * q + 3] > s_acc[4 * big + 3]) big = q;
          const float eps = 1.0f / 1024.0f;
          for (int d = 0; d < 3; ++d) {
            const float v = s_c[3 * big + d];
            const float sgn = (d & 1) ? -1.f : 1.f;
            s_c[3 * j + d] = v * (1.f + sgn * eps);
            s_c[3 * big + d] = v * (1.f - sgn * eps);
          }
          const float half = floorf(s_acc[4 * big + 3] * 0.5f);
          s_acc[4 * j + 3] = half; s_acc[4 * big + 3] -= half;
        }
      }
    }
    __syncthreads();
    if (tid < K && p.spherical) {
      const float l = sqrtf(s_c[3 * tid] * s_c[3 * tid] + s_c[3 * tid + 1] * s_c[3 * tid + 1] + s_c[3 * tid + 2] * s_c[3 * tid + 2]);
      if (l > 0.f) { s_c[3 * tid] /= l; s_c[3 * tid + 1] /= l; s_c[3 * tid + 2] /= l; }
    }
    __syncthreads();
  }
  // 5) final assignment of every valid row (kmeans.index.search, losses.py:89) + centroids out
  for (int j = tid; j < nv; j += kKmThreads) {
    const int r = valid_idx[j];
    assign[r] = best_centroid(x[3 * r], x[3 * r + 1], x[3 * r + 2], s_c, K);
  }
  for (int j = tid; j < K * 3; j += kKmThreads) centroids[j] = s_c[j];
}

// ---------------------------------------------------------------- orthogonal-triple selection (one CTA)
__global__ void __launch_bounds__(1024, 1)
cluster_select_kernel(const float* __restrict__ centroids, const int32_t* __restrict__ assign, int64_t n, int K,
                      float t_similar, int32_t* __restrict__ labels, int32_t* __restrict__ sel) {
  __shared__ int s_size[kKmMaxK];
  __shared__ int s_lab[kKmMaxK];
  __shared__ float s_sim[kKmMaxK * kKmMaxK];
  const int tid = threadIdx.x;
  for (int j = tid; j < K; j += blockDim.x) s_size[j] = 0;
  __syncthreads();
  for (int64_t i = tid; i < n; i += blockDim.x) { const int a = assign[i]; if (a >= 0) atomicAdd(&s_size[a], 1); }
  for (int e = tid; e < K * K; e += blockDim.x) {
    const int i = e / K, j = e % K;
    // sim = centrs @ centrs.T (fp32)
    s_sim[e] = centroids[3 * i] * centroids[3 * j] + centroids[3 * i + 1] * centroids[3 * j + 1] +
               centroids[3 * i + 2] * centroids[3 * j + 2];
  }
  __syncthreads();
  if (tid == 0) {
    int c1 = 0;
    for (int j = 1; j < K; ++j) if (s_size[j] > s_size[c1]) c1 = j;            // biggest cluster (losses.py:104-107)
    // criteria[i][j] = |sim[i,c1]| + |sim[c1,j]| + |sim[i,j]|; mins over i, then argmin over j (losses.py:117-120)
    int c2 = 0, c3 = 0; float best = INFINITY;
    for (int j = 0; j < K; ++j) {
      float mn = INFINITY; int arg = 0;
      for (int i = 0; i < K; ++i) {
        const float v = fabsf(s_sim[i * K + c1]) + fabsf(s_sim[c1 * K + j]) + fabsf(s_sim[i * K + j]);
        if (v < mn) { mn = v; arg = i; }
      }
      if (mn < best) { best = mn; c2 = j; c3 = arg; }
    }
    for (int j = 0; j < K; ++j) s_lab[j] = 0;
    const int cs[3] = {c1, c2, c3};
    for (int q = 0; q < 3; ++q)                                                // merge similar (losses.py:47-54)
      for (int j = 0; j < K; ++j) if (s_sim[cs[q] * K + j] > t_similar) s_lab[j] = q + 1;
    for (int q = 0; q < 3; ++q) {                                              // opposite clusters (losses.py:57-72)
      int o = 0;
      for (int j = 1; j < K; ++j) if (s_sim[cs[q] * K + j] < s_sim[cs[q] * K + o]) o = j;
      if (-s_sim[cs[q] * K + o] > t_similar)
        for (int j = 0; j < K; ++j) if (s_sim[o * K + j] > t_similar) s_lab[j] = -(q + 1);
    }
    sel[0] = c1; sel[1] = c2; sel[2] = c3;
  }
  __syncthreads();
  for (int64_t i = tid; i < n; i += blockDim.x) { const int a = assign[i]; labels[i] = a >= 0 ? s_lab[a] : 0; }
}

// ---------------------------------------------------------------- cluster statistics and loss (one CTA)
// stats layout (floats): per cluster k in 0..2 at stats[8k..]: [count, cx, cy, cz, |mu|, gl1x, gl1y, gl1z]
// where gl1 = sum_i sign(n'_i - c_k); stats[24..26] = losses (ort, dot, L1); stats[27] = valid flag
constexpr int kStats = 32;

__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(1024, 1)
cluster_loss_fw_kernel(const float* __restrict__ nrm, const int32_t* __restrict__ labels, int64_t n,
                       float* __restrict__ losses, float* __restrict__ stats) {
  __shared__ int s_wacc[12 * kAccStride];
  __shared__ long long s_sum[9];
  __shared__ int s_cnt[3];
  __shared__ float s_c[9], s_mu[3];
  __shared__ float s_red[32][8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int a = tid; a < 12 * kAccStride; a += blockDim.x) s_wacc[a] = 0;
  __syncthreads();
  const int64_t n_round = (n + 31) & ~(int64_t)31;
  for (int64_t i = tid; i < n_round; i += blockDim.x) {
    int l = 0;
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < n) l = labels[i];
    const bool v = l != 0;
    int k = 0;
    if (v) {
      k = (l > 0 ? l : -l) - 1;
      const float sg = l > 0 ? 1.f : -1.f;
      x = sg * nrm[3 * i]; y = sg * nrm[3 * i + 1]; z = sg * nrm[3 * i + 2];
    }
    warp_accumulate_by_key(k, v, x, y, z, s_wacc + wid, lane);
  }
  __syncthreads();
  if (tid < 12) {
    long long t = 0;
    for (int w = 0; w < 32; ++w) t += s_wacc[tid * kAccStride + w];
    const int k = tid >> 2, d = tid & 3;
    if (d == 3) s_cnt[k] = (int)t; else s_sum[3 * k + d] = t;
  }
  __syncthreads();
  if (tid < 3) {
    const int k = tid;
    if (s_cnt[k] > 0) {
      const double inv = 1.0 / ((double)s_cnt[k] * (double)kKmFix);
      const float mx = (float)((double)s_sum[3 * k] * inv), my = (float)((double)s_sum[3 * k + 1] * inv), mz = (float)((double)s_sum[3 * k + 2] * inv);
      const float len = sqrtf(mx * mx + my * my + mz * mz);
      const float d = fmaxf(len, 1e-12f);
      s_c[3 * k] = mx / d; s_c[3 * k + 1] = my / d; s_c[3 * k + 2] = mz / d; s_mu[k] = len;
    } else { s_c[3 * k] = s_c[3 * k + 1] = s_c[3 * k + 2] = 0.f; s_mu[k] = 0.f; }
  }
  __syncthreads();
  // second pass: per cluster sum of dots, sum of L1 distances and sum of sign(n' - c)
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // generic slots reused per cluster below
  float dotv[3] = {0, 0, 0}, l1v[3] = {0, 0, 0}, sg[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = tid; i < n; i += blockDim.x) {
    const int l = labels[i];
    if (l == 0) continue;
    const int k = (l > 0 ? l : -l) - 1;
    const float s = l > 0 ? 1.f : -1.f;
    const float x = s * nrm[3 * i], y = s * nrm[3 * i + 1], z = s * nrm[3 * i + 2];
    const float dx = x - s_c[3 * k], dy = y - s_c[3 * k + 1], dz = z - s_c[3 * k + 2];
#pragma unroll
    for (int q = 0; q < 3; ++q) if (q == k) {
      dotv[q] += x * s_c[3 * q] + y * s_c[3 * q + 1] + z * s_c[3 * q + 2];
      l1v[q] += fabsf(dx) + fabsf(dy) + fabsf(dz);
      sg[3 * q] += sgnf(dx); sg[3 * q + 1] += sgnf(dy); sg[3 * q + 2] += sgnf(dz);
    }
  }
  (void)acc;
  // block reduction of 15 values (3 dot, 3 l1, 9 sign sums) in two rounds of 8
  float vals[16] = {dotv[0], dotv[1], dotv[2], l1v[0], l1v[1], l1v[2], sg[0], sg[1], sg[2], sg[3], sg[4], sg[5], sg[6], sg[7], sg[8], 0.f};
  __shared__ float s_tot[16];
  for (int round = 0; round < 2; ++round) {
#pragma unroll
    for (int q = 0; q < 8; ++q) { const float v = warp_sum(vals[round * 8 + q]); if (lane == 0) s_red[wid][q] = v; }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float v = warp_sum(s_red[lane][q]); if (lane == 0) s_tot[round * 8 + q] = v; }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const bool ok = s_cnt[0] > 0 && s_cnt[1] > 0 && s_cnt[2] > 0;
    float l_ort = NAN, l_dot = NAN, l_l1 = NAN;
    if (ok) {
      auto dot3 = [&](int a, int b) { return s_c[3 * a] * s_c[3 * b] + s_c[3 * a + 1] * s_c[3 * b + 1] + s_c[3 * a + 2] * s_c[3 * b + 2]; };
      l_ort = (fabsf(dot3(0, 1)) + fabsf(dot3(0, 2)) + fabsf(dot3(1, 2))) / 3.0f;
      l_dot = ((1.f - s_tot[0] / s_cnt[0]) + (1.f - s_tot[1] / s_cnt[1]) + (1.f - s_tot[2] / s_cnt[2])) / 3.0f;
      l_l1 = (s_tot[3] / s_cnt[0] + s_tot[4] / s_cnt[1] + s_tot[5] / s_cnt[2]) / 3.0f;
    }
    losses[0] = l_ort; losses[1] = l_dot; losses[2] = l_l1;
    for (int k = 0; k < 3; ++k) {
      stats[8 * k] = (float)s_cnt[k];
      stats[8 * k + 1] = s_c[3 * k]; stats[8 * k + 2] = s_c[3 * k + 1]; stats[8 * k + 3] = s_c[3 * k + 2];
      stats[8 * k + 4] = s_mu[k];
      stats[8 * k + 5] = s_tot[6 + 3 * k]; stats[8 * k + 6] = s_tot[7 + 3 * k]; stats[8 * k + 7] = s_tot[8 + 3 * k];
    }
    stats[24] = l_ort; stats[25] = l_dot; stats[26] = l_l1; stats[27] = ok ? 1.f : 0.f;
  }
}

// dL/dn for  L = w[0] L_ort + w[1] L_dot + w[2] L_L1   (w read from device memory)
__global__ void __launch_bounds__(256)
cluster_loss_bw_kernel(const float* __restrict__ nrm, const int32_t* __restrict__ labels, int64_t n,
                       const float* __restrict__ stats, const float* __restrict__ w, float* __restrict__ dn) {
  __shared__ float s_A[9], s_c[9], s_m[3];
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    s_ok = stats[27] > 0.5f;
    if (s_ok) {
      float c[3][3], m[3], mu[3], g1[3][3];
      for (int k = 0; k < 3; ++k) {
        m[k] = stats[8 * k]; mu[k] = fmaxf(stats[8 * k + 4], 1e-12f);
        for (int d = 0; d < 3; ++d) { c[k][d] = stats[8 * k + 1 + d]; g1[k][d] = stats[8 * k + 5 + d]; }
      }
      auto dot3 = [&](int a, int b) { return c[a][0] * c[b][0] + c[a][1] * c[b][1] + c[a][2] * c[b][2]; };
      const float s01 = sgnf(dot3(0, 1)), s02 = sgnf(dot3(0, 2)), s12 = sgnf(dot3(1, 2));
      for (int k = 0; k < 3; ++k) {
        float gc[3];   // dL/dc_k from the ort term and the L1 term
        for (int d = 0; d < 3; ++d) {
          float go = 0.f;
          if (k == 0) go = s01 * c[1][d] + s02 * c[2][d];
          if (k == 1) go = s01 * c[0][d] + s12 * c[2][d];
          if (k == 2) go = s02 * c[0][d] + s12 * c[1][d];
          gc[d] = w[0] * go / 3.0f - w[2] * g1[k][d] / (3.0f * m[k]);
        }
        // through c = mu/|mu|:  (I - c c^T)/|mu| gc ; then mean: 1/m_k
        const float cg = c[k][0] * gc[0] + c[k][1] * gc[1] + c[k][2] * gc[2];
        for (int d = 0; d < 3; ++d)
          s_A[3 * k + d] = ((gc[d] - c[k][d] * cg) / mu[k]) / m[k] - w[1] * c[k][d] / (3.0f * m[k]);
        for (int d = 0; d < 3; ++d) s_c[3 * k + d] = c[k][d];
        s_m[k] = w[2] / (3.0f * m[k]);
      }
    }
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int l = labels[i];
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (l != 0 && s_ok) {
      const int k = (l > 0 ? l : -l) - 1;
      const float s = l > 0 ? 1.f : -1.f;
      const float x = s * nrm[3 * i], y = s * nrm[3 * i + 1], z = s * nrm[3 * i + 2];
      gx = s * (s_A[3 * k] + s_m[k] * sgnf(x - s_c[3 * k]));
      gy = s * (s_A[3 * k + 1] + s_m[k] * sgnf(y - s_c[3 * k + 1]));
      gz = s * (s_A[3 * k + 2] + s_m[k] * sgnf(z - s_c[3 * k + 2]));
    }
    dn[3 * i] = gx; dn[3 * i + 1] = gy; dn[3 * i + 2] = gz;
  }
}

// ---------------------------------------------------------------- photometric terms (fused fwd + grad)
// rgb = rend[:, :3] + bg*(1-opacity)  (rendering.py:231-241);  L_rgb = mean((rgb-target)^2) (losses.py:353);
// L_op = w_op * mean(-(o+1e-10) log(o+1e-10)) (losses.py:357-361).  sums[0] += sum sq err, sums[1] += sum entropy.
// Gradients are multiplied by `gscale` (loss scale / grad divisor) and written (not accumulated).
__global__ void __launch_bounds__(256)
photometric_kernel(const float* __restrict__ rend, const float* __restrict__ opacity, const float* __restrict__ target,
                   int64_t n_rays, int C, float bg0, float bg1, float bg2, float opacity_w, float gscale,
                   float* __restrict__ rgb_out, float* __restrict__ sums, float* __restrict__ d_rend,
                   float* __restrict__ d_opacity) {
  float se = 0.f, ent = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float inv3n = 1.0f / (3.0f * (float)n_rays), invn = 1.0f / (float)n_rays;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += stride) {
    const float o = opacity[r];
    const float bg[3] = {bg0, bg1, bg2};
    float go = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = rend[r * C + c] + bg[c] * (1.f - o);
      if (rgb_out) rgb_out[3 * r + c] = v;
      const float e = v - target[3 * r + c];
      se += e * e;
      const float g = 2.f * e * inv3n * gscale;
      if (d_rend) d_rend[r * C + c] = g;
      go -= bg[c] * g;
    }
    if (d_rend) for (int c = 3; c < C; ++c) d_rend[r * C + c] = 0.f;
    if (opacity_w > 0.f) {
      const float oe = o + 1e-10f;
      const float lg = logf(oe);
      ent += -oe * lg;
      go += opacity_w * (-(lg + 1.f)) * invn * gscale;
    }
    if (d_opacity) d_opacity[r] = go;
  }
  se = warp_sum(se); ent = warp_sum(ent);
  __shared__ float s_a[8], s_b[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_a[wid] = se; s_b[wid] = ent; }
  __syncthreads();
  if (wid == 0) {
    se = lane < 8 ? s_a[lane] : 0.f; ent = lane < 8 ? s_b[lane] : 0.f;
    se = warp_sum(se); ent = warp_sum(ent);
    if (lane == 0) { atomicAdd(sums, se); atomicAdd(sums + 1, ent); }
  }
}

}  // namespace ncn

using namespace ncn;

extern "C" int ncn_normals_from_depth_fw(const float* origin, const float* dir, const float* depth, const int64_t* idx1,
                                         const int64_t* idx2, const int64_t* idx3, int64_t n_tri, float* normals,
                                         ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_tri >= 0);
  if (n_tri == 0) return NCN_OK;
  NCN_CHECK_PTR(origin); NCN_CHECK_PTR(dir); NCN_CHECK_PTR(depth); NCN_CHECK_PTR(idx1); NCN_CHECK_PTR(idx2);
  NCN_CHECK_PTR(idx3); NCN_CHECK_PTR(normals);
  normals_fw_kernel<<<persistent_grid(n_tri, 256, 8), 256, 0, as_stream(stream)>>>(origin, dir, depth, idx1, idx2, idx3, n_tri, normals);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_normals_from_depth_bw(const float* origin, const float* dir, const float* depth, const int64_t* idx1,
                                         const int64_t* idx2, const int64_t* idx3, const float* dL_dnormals,
                                         int64_t n_tri, float* dL_ddepth, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_tri >= 0);
  if (n_tri == 0) return NCN_OK;
  NCN_CHECK_PTR(origin); NCN_CHECK_PTR(dir); NCN_CHECK_PTR(depth); NCN_CHECK_PTR(idx1); NCN_CHECK_PTR(idx2);
  NCN_CHECK_PTR(idx3); NCN_CHECK_PTR(dL_dnormals); NCN_CHECK_PTR(dL_ddepth);
  normals_bw_kernel<<<persistent_grid(n_tri, 256, 8), 256, 0, as_stream(stream)>>>(origin, dir, depth, idx1, idx2, idx3,
                                                                                  dL_dnormals, n_tri, dL_ddepth);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" size_t ncn_kmeans_workspace_bytes(int64_t n_points_max, int k) {
  (void)k;
  return n_points_max < 0 ? 0 : (size_t)n_points_max * sizeof(int32_t) + 256;
}

extern "C" int ncn_kmeans_spherical(const float* x, int64_t n_points, const ncn_kmeans_params* p, float* centroids,
                                    int32_t* assign, int32_t* n_valid, void* workspace, size_t workspace_bytes,
                                    ncn_stream_t stream) {
  NCN_CHECK_PTR(p); NCN_CHECK_PTR(centroids); NCN_CHECK_PTR(n_valid);
  if (p->k < 1 || p->k > kKmMaxK || p->niter < 0 || p->max_points_per_centroid < 1) return NCN_E_CONFIG;
  NCN_CHECK_SIZE(n_points >= 0 && n_points < ((int64_t)1 << 31));
  if (n_points > 0) { NCN_CHECK_PTR(x); NCN_CHECK_PTR(assign); NCN_CHECK_PTR(workspace); }
  if (workspace_bytes < ncn_kmeans_workspace_bytes(n_points, p->k)) return NCN_E_SIZE;
  int64_t cap = (int64_t)p->max_points_per_centroid * p->k;
  if (cap > 14336) cap = 14336;            // 14336 * 12 B = 168 KB of dynamic shared memory (+ 41 KB static)
  if (cap > n_points) cap = n_points > 0 ? n_points : 1;
  const size_t smem = (size_t)cap * 12 + 16;
  // static (41 KB) + dynamic shared memory exceed the 48 KB default: always opt in
  NCN_CUDA(cudaFuncSetAttribute(kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 172 * 1024));
  kmeans_kernel<<<1, kKmThreads, smem, as_stream(stream)>>>(x, n_points, *p, centroids, assign, n_valid, (int32_t*)workspace, (int)cap);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_cluster_select(const float* centroids, const int32_t* assign, int64_t n_points, int k, float t_similar,
                                  int32_t* labels, int32_t* sel, ncn_stream_t stream) {
  NCN_CHECK_PTR(centroids); NCN_CHECK_PTR(sel);
  if (k < 3 || k > kKmMaxK) return NCN_E_CONFIG;
  NCN_CHECK_SIZE(n_points >= 0);
  if (n_points > 0) { NCN_CHECK_PTR(assign); NCN_CHECK_PTR(labels); }
  cluster_select_kernel<<<1, 1024, 0, as_stream(stream)>>>(centroids, assign, n_points, k, t_similar, labels, sel);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_cluster_loss_fw(const float* normals, const int32_t* labels, int64_t n_points, float* losses,
                                   float* stats, ncn_stream_t stream) {
  NCN_CHECK_PTR(losses); NCN_CHECK_PTR(stats);
  NCN_CHECK_SIZE(n_points >= 0);
  if (n_points > 0) { NCN_CHECK_PTR(normals); NCN_CHECK_PTR(labels); }
  cluster_loss_fw_kernel<<<1, 1024, 0, as_stream(stream)>>>(normals, labels, n_points, losses, stats);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_cluster_loss_bw(const float* normals, const int32_t* labels, int64_t n_points, const float* stats,
                                   const float* weights_dev, float* dL_dnormals, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_points >= 0);
  if (n_points == 0) return NCN_OK;
  NCN_CHECK_PTR(normals); NCN_CHECK_PTR(labels); NCN_CHECK_PTR(stats); NCN_CHECK_PTR(weights_dev); NCN_CHECK_PTR(dL_dnormals);
  cluster_loss_bw_kernel<<<persistent_grid(n_points, 256, 8), 256, 0, as_stream(stream)>>>(normals, labels, n_points, stats,
                                                                                          weights_dev, dL_dnormals);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_photometric_loss(const float* rend, const float* opacity, const float* target_rgb, int64_t n_rays,
                                    int n_channels, const float* bg_rgb_host, float opacity_w, float grad_scale,
                                    float* rgb_out, float* sums, float* dL_drend, float* dL_dopacity,
                                    ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_rays >= 0 && n_channels >= 3);
  if (n_rays == 0) return NCN_OK;
  NCN_CHECK_PTR(rend); NCN_CHECK_PTR(opacity); NCN_CHECK_PTR(target_rgb); NCN_CHECK_PTR(bg_rgb_host); NCN_CHECK_PTR(sums);
  photometric_kernel<<<persistent_grid(n_rays, 256, 4), 256, 0, as_stream(stream)>>>(
      rend, opacity, target_rgb, n_rays, n_channels, bg_rgb_host[0], bg_rgb_host[1], bg_rgb_host[2], opacity_w, grad_scale,
      rgb_out, sums, dL_drend, dL_dopacity);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
