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
#include "mma.cuh"

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

__device__ __forceinline__ void normals_bw_body(const float* __restrict__ origin, const float* __restrict__ dir, const float* __restrict__ depth,
                                                const int64_t* __restrict__ i1, const int64_t* __restrict__ i2, const int64_t* __restrict__ i3,
                                                const float* __restrict__ dn, int64_t n_tri, float* __restrict__ ddepth,
                                                int64_t first, int64_t stride) {
  for (int64_t m = first; m < n_tri; m += stride) {
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
__global__ void __launch_bounds__(256)
normals_bw_kernel(const float* __restrict__ origin, const float* __restrict__ dir, const float* __restrict__ depth,
                  const int64_t* __restrict__ i1, const int64_t* __restrict__ i2, const int64_t* __restrict__ i3,
                  const float* __restrict__ dn, int64_t n_tri, float* __restrict__ ddepth) {
  normals_bw_body(origin, dir, depth, i1, i2, i3, dn, n_tri, ddepth, (int64_t)blockIdx.x * blockDim.x + threadIdx.x,
                  (int64_t)gridDim.x * blockDim.x);
}

// ---------------------------------------------------------------- spherical k-means (one thread-block cluster)
// The problem is tiny (<= 5120 training points x K = 20 x 20 iterations) but a single SM can issue only 4 warp
// instructions per clock, and one Lloyd iteration is ~10^5 (point, centroid) pairs: on one CTA the kernel is
// instruction-issue bound (measured: 1.4 M warp instructions, 300 us).  It therefore runs on a CLUSTER of 8 CTAs
// (8 SMs): each CTA owns 1/8 of the training points in its shared memory, does assignment + a warp-aggregated
// integer reduction (redux.sync; fixed point 2^20 => order independent => bit-reproducible), publishes its K x 4
// partial sums in its own shared memory, and after one barrier.cluster every CTA folds all 8 partials through
// distributed shared memory and updates the (replicated) centroids.  No atomics, no host round trip.
#ifndef NCN_KM_THREADS
#define NCN_KM_THREADS 512           // 16 warps: with a 16-CTA cluster a warp owns 1-2 sixteen-point tiles per Lloyd iteration
#endif                               // (-DNCN_KM_THREADS=640: 20 warps = one tile per warp at 5120 training points; A/B build)
constexpr int kKmThreads = NCN_KM_THREADS;
constexpr int kKmCluster = 8;        // portable cluster size (fallback)
constexpr int kKmClusterMax = 16;    // non-portable size tried first: half the tiles per CTA and Lloyd iteration
constexpr int kKmMaxK = 64;
constexpr float kKmFix = 1048576.0f;     // 2^20
constexpr int kAccStride = 33;

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

// adds (x,y,z,1) of every valid lane to acc[key*4 + {0,1,2,3}] (a warp-private row of int accumulators): the warp walks
// the distinct keys present and folds each key's members with redux.sync
__device__ __forceinline__ void warp_accumulate_by_key(int key, bool valid, float x, float y, float z, int* __restrict__ acc,
                                                       int lane) {
  const int fx = __float2int_rn(x * kKmFix), fy = __float2int_rn(y * kKmFix), fz = __float2int_rn(z * kKmFix);
  unsigned remaining = __ballot_sync(0xffffffffu, valid);
  while (remaining) {
    const int leader = __ffs(remaining) - 1;
    const int k = __shfl_sync(0xffffffffu, key, leader);
    const bool mine = valid && key == k;
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    const int sx = __reduce_add_sync(0xffffffffu, mine ? fx : 0);
    const int sy = __reduce_add_sync(0xffffffffu, mine ? fy : 0);
    const int sz = __reduce_add_sync(0xffffffffu, mine ? fz : 0);
    if (lane == leader) { acc[4 * k] += sx; acc[4 * k + 1] += sy; acc[4 * k + 2] += sz; acc[4 * k + 3] += __popc(m); }
    remaining &= ~m;
  }
  __syncwarp();
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// execution barrier only (no memory ordering: no MEMBAR, used where nothing read afterwards depends on the peers' writes)
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// read an int from the same shared-memory variable of CTA `rank` of the cluster (DSMEM)
__device__ __forceinline__ int dsmem_ld_int(const int* local_ptr, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local_ptr), ra;
  int v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(ra) : "memory");
  return v;
}

__device__ __forceinline__ float dsmem_ld_float(const float* local_ptr, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local_ptr), ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}


// ---- cluster exchange without barrier.cluster ------------------------------------------------------------------------------
// barrier.cluster.arrive.release lowers to MEMBAR.ALL.GPU + the hardware cluster barrier + an L1 invalidate (SASS), ~2 us per use
// on a 16-CTA cluster - two thirds of a Lloyd iteration.  The exchange below needs none of it: every CTA pushes its partial sums
// into an inbox in every peer's shared memory with st.async, whose completion is counted (bytes) by an mbarrier in the RECEIVER's
// shared memory; the receiver waits on its own mbarrier and then reads its inbox with ordinary shared-memory loads.  Inboxes and
// mbarriers are double buffered by round parity (a peer can be at most one round ahead: it cannot finish round r+1 before it has
// received this CTA's round r+1 contribution, which is sent only after round r's inbox has been consumed).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(addr), "r"(rank));
  return ra;
}
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint4 v, uint32_t remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(mbar), "r"(parity) : "memory");
}
constexpr int kKmInboxLanes = 32;       // one 16-byte slot per lane of the exchanging warp
// one round of the all-to-all: lanes [0, n_lanes) of ONE warp each contribute 16 bytes; returns when all CL contributions of the
// round have landed in this CTA's inbox[round & 1][source rank][lane]
template <int CL>
__device__ __forceinline__ const uint4* km_exchange(uint4* inbox, uint64_t* mbar, int round, uint32_t rank, int lane, int n_lanes, uint4 v) {
  const int buf = round & 1;
  const uint32_t mb = smem_u32(mbar + buf);
  if (lane < n_lanes) {
    const uint32_t dst = smem_u32(inbox + ((size_t)(buf * CL + (int)rank) * kKmInboxLanes + lane));
#pragma unroll
    for (uint32_t r = 0; r < CL; ++r) st_async_v4(mapa_u32(dst, r), v, mapa_u32(mb, r));
  }
  if (lane == 0) mbar_arrive_expect_tx(mb, (uint32_t)(CL * n_lanes * 16));
  mbar_wait(mb, (uint32_t)((round >> 1) & 1));
  __syncwarp();
  return inbox + (size_t)buf * CL * kKmInboxLanes;
}

// ---- tensor-core Lloyd iteration (K <= 32) ---------------------------------------------------------------
// Both halves of an iteration are tiny GEMMs over 16-point tiles, so they run on the warp MMA units:
//   assignment : S[16 points x 8 clusters] = P[16 x 16] * C[16 x 8],  P row = (xh,yh,zh, xh,yh,zh, xl,yl,zl, 0..),
//                C col = (ch, cl, ch, 0..): fp16 hi/lo splits, exact fp32 products, ~2^-22 relative accuracy
//   accumulate : SUM[16 clusters x 8] += ONEHOT[16 clusters x 16 points] * Q[16 points x 8], Q row = (xh,yh,zh,xl,yl,zl,1,0)
// ~120 instructions per 16-point tile instead of ~8 per (point, centroid) pair + a serial redux chain.
__device__ __forceinline__ void split_hl(float v, float& hi, float& lo) {
  hi = __half2float(__float2half_rn(v));
  lo = v - hi;
}
__device__ __forceinline__ float centroid_row(const float* __restrict__ c, int row) {
  if (row >= 9) return 0.f;
  float hi, lo; split_hl(c[row % 3], hi, lo);
  return (row >= 3 && row < 6) ? lo : hi;
}

#ifdef NCN_KM_TRACE      // developer build only
__device__ long long g_km_trace[64];
#define KM_TRACE(slot) do { if (threadIdx.x == 0 && cluster_ctarank() == 0) g_km_trace[slot] = clock64(); } while (0)
#else
#define KM_TRACE(slot) do { } while (0)
#endif

// empty clusters split the currently largest one (faiss-style +-eps); rare, one thread
__device__ __forceinline__ void split_empty_clusters(float* __restrict__ s_c, float* __restrict__ s_acc, int K) {
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


// warp arg-min / arg-max over (value, index) pairs with ties to the LOWEST index (the reference's argmin / first-max loops)
__device__ __forceinline__ void warp_argmin(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

// Orthogonal-triple selection, merge and opposite labelling (losses.py:97-166) for K <= 32 by ONE warp, one lane per cluster:
// s_sim = centroids @ centroids^T (K x K), s_size = members per cluster -> s_lab[j] in {-3..3}, (c1, c2, c3)
__device__ __forceinline__ void select_triple_warp(const float* __restrict__ s_sim, const int* __restrict__ s_size, int K, float t_similar,
                                                   int* __restrict__ s_lab, int& c1_out, int& c2_out, int& c3_out, int lane) {
  const int j = lane;
  const bool on = j < K;
  // biggest cluster (losses.py:104-107): first maximum
  float negsz = on ? -(float)s_size[j] : INFINITY; int c1 = j;
  warp_argmin(negsz, c1);
  // criteria[i][j] = |sim[i,c1]| + |sim[c1,j]| + |sim[i,j]|; column j: min / argmin over i (losses.py:117-118)
  float mn = INFINITY; int arg = 0;
  if (on)
    for (int i = 0; i < K; ++i) {
      const float v = fabsf(s_sim[i * K + c1]) + fabsf(s_sim[c1 * K + j]) + fabsf(s_sim[i * K + j]);
      if (v < mn) { mn = v; arg = i; }
    }
  float best = mn; int c2 = j;                                            // losses.py:119-120
  warp_argmin(best, c2);
  const int c3 = __shfl_sync(0xffffffffu, arg, c2);
  int lab = 0;
  const int cs[3] = {c1, c2, c3};
#pragma unroll
  for (int q = 0; q < 3; ++q)                                             // merge similar (losses.py:47-54)
    if (on && s_sim[cs[q] * K + j] > t_similar) lab = q + 1;
#pragma unroll
  for (int q = 0; q < 3; ++q) {                                           // opposite clusters (losses.py:57-72)
    float v = on ? s_sim[cs[q] * K + j] : INFINITY; int o = j;
    warp_argmin(v, o);
    if (-v > t_similar && on && s_sim[o * K + j] > t_similar) lab = -(q + 1);
  }
  if (on) s_lab[j] = lab;
  c1_out = c1; c2_out = c2; c3_out = c3;
}

__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
constexpr int kStats = 32;

// Optional stages the k-means launch can carry in front of and behind the clustering itself (ncn_cluster_chain): the normals from
// the rendered depth (datasets/hypersim_src/utils.py:505-541) as a prologue, and - with the points and centroids still resident
// in the cluster's shared memory - the triple selection (losses.py:97-166) and the cluster statistics / three loss terms
// (losses.py:441-478) as an epilogue.  All-null = plain k-means.
struct ChainArgs {
  const float* origin; const float* dir; const float* depth;          // prologue inputs (origin == nullptr: x is given)
  const int64_t* i1; const int64_t* i2; const int64_t* i3;
  float* normals_out;
  float t_similar;                                                    // epilogue (labels == nullptr: none)
  int32_t* labels; int32_t* sel; float* losses; float* stats;
};
__device__ __forceinline__ long long dsmem_ld_i64(const long long* local_ptr, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local_ptr), ra;
  long long v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.s64 %0, [%1];" : "=l"(v) : "r"(ra) : "memory");
  return v;
}

template <int CL>
__global__ void __launch_bounds__(kKmThreads, 1)
kmeans_kernel(const float* __restrict__ x_in, int64_t n, ncn_kmeans_params p, float* __restrict__ centroids,
              int32_t* __restrict__ assign, int32_t* __restrict__ n_valid_out, int32_t* __restrict__ valid_idx,
              int nt_cap, const ChainArgs ch) {
  extern __shared__ __align__(16) unsigned char km_smem[];
  float* xs = reinterpret_cast<float*>(km_smem);                 // this CTA's training points [my_n][3]
  __shared__ float s_c[kKmMaxK * 3];
  // per-warp private accumulators (integer Lloyd path, epilogue) and the tensor-core path's per-warp (32 clusters x 8) partial
  // sums are never live at the same time (block barriers in between): one array
  constexpr int kWaccInts = (kKmThreads / 32) * kKmMaxK * 4, kFwFloats = (kKmThreads / 32) * 32 * 8;
  __shared__ __align__(16) int s_wacc[kWaccInts > kFwFloats ? kWaccInts : kFwFloats];
  __shared__ int s_part[2][kKmMaxK * 4];                         // this CTA's partial sums (double buffered), read by peers
  __shared__ float s_acc[kKmMaxK * 4];
  float* s_fw = reinterpret_cast<float*>(s_wacc);
  __shared__ float s_fpart[2][32 * 8];                           // this CTA's partial sums (double buffered), read by peers
  __shared__ int s_nvalid, s_warp_tot[32], s_base, s_any_empty;
  __shared__ __align__(8) uint64_t s_mbar[2];                    // completion of the two exchange inboxes (double buffered)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    mbar_init(smem_u32(&s_mbar[0]), 1); mbar_init(smem_u32(&s_mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int K = p.k;
  const uint32_t rank = cluster_ctarank();
  pdl_wait(); pdl_trigger();
  KM_TRACE(0);
  // 0) chain prologue: the normals of the rendered depth (one triangle per thread, the cluster's CTAs split them), published
  //    to global memory (the backward pass and the caller read them) and made visible to the whole cluster by one barrier
  const float* x = x_in;
  if (ch.origin != nullptr) {
    for (int64_t m = tid + (int64_t)rank * kKmThreads; m < n; m += (int64_t)kKmThreads * CL) {
      const int64_t idx[3] = {ch.i1[m], ch.i2[m], ch.i3[m]};
      float P[3][3];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const float dep = ch.depth[idx[v]];
#pragma unroll
        for (int c = 0; c < 3; ++c) P[v][c] = __fadd_rn(ch.origin[3 * idx[v] + c], __fmul_rn(ch.dir[3 * idx[v] + c], dep));
      }
      float a[3], b[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) { a[c] = P[1][c] - P[0][c]; b[c] = P[2][c] - P[0][c]; }
      const float cx = a[1] * b[2] - a[2] * b[1];
      const float cy = a[2] * b[0] - a[0] * b[2];
      const float cz = a[0] * b[1] - a[1] * b[0];
      const float nrm = fmaxf(sqrtf(cx * cx + cy * cy + cz * cz), 1e-12f);   // F.normalize eps (same arithmetic as normals_fw_kernel)
      ch.normals_out[3 * m] = cx / nrm; ch.normals_out[3 * m + 1] = cy / nrm; ch.normals_out[3 * m + 2] = cz / nrm;
    }
    __threadfence();
    cluster_sync_all();
    x = ch.normals_out;
  }
  // 1) every CTA compacts the valid rows (stable order) - identical results, CTA 0 publishes them.  Warp w owns a
  //    contiguous slice of rows, walked in chunks of 32 x 32 rows whose validity bits a lane gathers with independent
  //    (pipelined) loads: count, one block-wide exclusive scan of the 8 warp totals, then write.
  {
    const int64_t per_warp = (((n + (kKmThreads / 32) - 1) / (kKmThreads / 32)) + 31) & ~(int64_t)31;
    const int64_t r_begin = (int64_t)wid * per_warp, r_end = r_begin + per_warp < n ? r_begin + per_warp : n;
    auto chunk_mask = [&](int64_t c0) -> uint32_t {            // bit b: row c0 + 32*b + lane is a valid normal
      uint32_t vm = 0u;
#pragma unroll
      for (int b = 0; b < 32; ++b) {
        const int64_t i = c0 + 32 * b + lane;
        if (i < r_end && valid_normal(x[3 * i], x[3 * i + 1], x[3 * i + 2])) vm |= 1u << b;
      }
      return vm;
    };
    const uint32_t vm0 = r_begin < r_end ? chunk_mask(r_begin) : 0u;      // kept for the write pass (n <= 8192: the only chunk)
    int cnt = __popc(vm0);
    for (int64_t c0 = r_begin + 1024; c0 < r_end; c0 += 1024) cnt += __popc(chunk_mask(c0));
    cnt = warp_sum_i(cnt);
    if (lane == 0) s_warp_tot[wid] = cnt;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kKmThreads / 32; ++w) { const int c = s_warp_tot[w]; if (w < wid) base += c; total += c; }
    if (rank == 0) {
      for (int64_t c0 = r_begin; c0 < r_end; c0 += 1024) {
        const uint32_t vm = c0 == r_begin ? vm0 : chunk_mask(c0);
        for (int b = 0; b < 32 && c0 + 32 * b < r_end; ++b) {
          const int64_t i = c0 + 32 * b + lane;
          const bool v = (vm >> b) & 1u;
          const unsigned bal = __ballot_sync(0xffffffffu, v);
          if (v) valid_idx[base + __popc(bal & ((1u << lane) - 1))] = (int32_t)i;
          else if (i < r_end) { assign[i] = -1; if (ch.labels != nullptr) ch.labels[i] = 0; }
          base += __popc(bal);
        }
      }
    }
    if (tid == 0) s_base = total;
    __syncthreads();
  }
  KM_TRACE(1);
  const int nv = s_base;
  if (rank == 0 && tid == 0) *n_valid_out = nv;
  if (nv == 0) {
    if (rank == 0) for (int j = tid; j < K * 3; j += kKmThreads) centroids[j] = 0.f;
    if (rank == 0 && tid == 0 && ch.labels != nullptr) {      // no valid normal: every cluster is empty -> NaN terms, flag 0
      ch.sel[0] = ch.sel[1] = ch.sel[2] = 0;
      for (int q = 0; q < kStats; ++q) ch.stats[q] = 0.f;
      ch.losses[0] = ch.losses[1] = ch.losses[2] = NAN; ch.stats[24] = ch.stats[25] = ch.stats[26] = NAN;
    }
    return;                         // uniform across the cluster: nobody reaches a cluster barrier
  }
  __threadfence();
  cluster_sync_all();               // valid_idx (written by CTA 0) is visible to the whole cluster
  KM_TRACE(2);
  // 2) training subset: at most max_points_per_centroid*K points at a uniform stride over the valid rows (faiss draws a
  //    random subset; equality with faiss is not a parity criterion); CTA r owns training points r, r+8, r+16, ...
  const int nt = nv > nt_cap ? nt_cap : nv;
  const int my_n = (nt - (int)rank + CL - 1) / CL;
  for (int j = tid; j < my_n; j += kKmThreads) {
    const int gj = j * CL + (int)rank;
    const int r = valid_idx[(int)(((int64_t)gj * nv) / nt)];
    xs[3 * j] = x[3 * r]; xs[3 * j + 1] = x[3 * r + 1]; xs[3 * j + 2] = x[3 * r + 2];
  }
  // 3) init (replicated): K training points spread over the subset with a seeded offset
  if (tid < K) {
    const uint32_t h = (uint32_t)p.seed * 2654435761u + 12345u;
    const int span = nt / K > 0 ? nt / K : 1;
    const int j = (int)((((int64_t)tid * nt) / K + (h % (uint32_t)span)) % nt);
    const int r = valid_idx[(int)(((int64_t)j * nv) / nt)];
    float cx = x[3 * r], cy = x[3 * r + 1], cz = x[3 * r + 2];
    if (p.spherical) { const float l = sqrtf(cx * cx + cy * cy + cz * cz); if (l > 0.f) { cx /= l; cy /= l; cz /= l; } }
    s_c[3 * tid] = cx; s_c[3 * tid + 1] = cy; s_c[3 * tid + 2] = cz;
  }
  __syncthreads();
  // 4) Lloyd iterations
  const int my_round = (my_n + 31) & ~31;
  const bool use_tc = K <= 32;
  const int g = lane >> 2, t = lane & 3;
  const int n_tiles = (my_n + 15) >> 4;
  __half* hq = reinterpret_cast<__half*>(km_smem + (((size_t)((nt_cap + CL - 1) / CL) * 12 + 15) & ~(size_t)15));
  // exchange inboxes [2][CL][32 lanes] x 16 B behind the fp16 rows (only ever written by the peers' st.async)
  uint4* inbox = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(hq) + (size_t)(((nt_cap + CL - 1) / CL + 16) * 16));
  int xround = 0;                                                // exchange rounds done (uniform across the cluster)
  if (use_tc) {
    // iteration-invariant fp16 point rows: hq[point] = (xh,yh,zh,xl,yl,zl,1,0); all-zero rows pad the last tile
    for (int j = tid; j < n_tiles * 16; j += kKmThreads) {
      uint4 row = make_uint4(0u, 0u, 0u, 0u);
      if (j < my_n) {
        float xh, xl, yh, yl, zh, zl;
        split_hl(xs[3 * j], xh, xl); split_hl(xs[3 * j + 1], yh, yl); split_hl(xs[3 * j + 2], zh, zl);
        row.x = pack_half2(xh, yh); row.y = pack_half2(zh, xl); row.z = pack_half2(yl, zl); row.w = pack_half2(1.f, 0.f);
      }
      reinterpret_cast<uint4*>(hq)[j] = row;
    }
    __syncthreads();
  }
  KM_TRACE(3);
  float prev_c[3] = {0.f, 0.f, 0.f};                               // warp 0, lane j: centroid j before the current iteration
  if (wid == 0 && lane < K) { prev_c[0] = s_c[3 * lane]; prev_c[1] = s_c[3 * lane + 1]; prev_c[2] = s_c[3 * lane + 2]; }
  for (int it = 0; it < p.niter; ++it) {
    if (it < 4) KM_TRACE(8 + 8 * it);
    if (use_tc) {
      uint32_t cb[4][2];        // centroid fragments: clusters 8j+g, rows 2t,2t+1 / 2t+8,2t+9 of the expanded column
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 8 * j + g;
        const float zero3[3] = {0.f, 0.f, 0.f};
        const float* cp = c < K ? s_c + 3 * c : zero3;
        cb[j][0] = pack_half2(centroid_row(cp, 2 * t), centroid_row(cp, 2 * t + 1));
        cb[j][1] = pack_half2(centroid_row(cp, 2 * t + 8), centroid_row(cp, 2 * t + 9));
      }
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      for (int tile = wid; tile < n_tiles; tile += kKmThreads / 32) {
        const int p0 = tile * 16;
        uint32_t pa[4];
        {
          const uint4 r0 = reinterpret_cast<const uint4*>(hq)[p0 + g], r1 = reinterpret_cast<const uint4*>(hq)[p0 + g + 8];
          // words: x=(xh,yh) y=(zh,xl) z=(yl,zl);  pairs needed: t0:(xh,yh) t1:(zh,xh) t2:(yh,zh) t3:(xl,yl) | t0:(zl,0)
          auto sel = [&](const uint4& r) -> uint32_t {
            if (t == 0) return r.x;
            if (t == 1) return (r.y & 0xFFFFu) | (r.x << 16);
            if (t == 2) return (r.x >> 16) | (r.y << 16);
            return (r.y >> 16) | (r.z << 16);
          };
          pa[0] = sel(r0); pa[1] = sel(r1);
          pa[2] = t == 0 ? (r0.z >> 16) : 0u; pa[3] = t == 0 ? (r1.z >> 16) : 0u;
        }
        uint32_t pq[2];       // accumulation B fragment: column g of points 2t,2t+1 / 2t+8,2t+9
        ldmatrix_x2_trans(pq, hq + (size_t)(p0 + (lane & 15)) * 8);
        float best0 = -INFINITY, best1 = -INFINITY;
        int bi0 = 0, bi1 = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (8 * j >= K) break;
          float sc[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(sc, pa, cb[j][0], cb[j][1]);
          const int c0 = 8 * j + 2 * t;
          if (c0 < K && sc[0] > best0) { best0 = sc[0]; bi0 = c0; }
          if (c0 + 1 < K && sc[1] > best0) { best0 = sc[1]; bi0 = c0 + 1; }
          if (c0 < K && sc[2] > best1) { best1 = sc[2]; bi1 = c0; }
          if (c0 + 1 < K && sc[3] > best1) { best1 = sc[3]; bi1 = c0 + 1; }
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {      // reduce over the quad (ties -> lowest cluster index)
          const float ob0 = __shfl_xor_sync(0xffffffffu, best0, o), ob1 = __shfl_xor_sync(0xffffffffu, best1, o);
          const int oi0 = __shfl_xor_sync(0xffffffffu, bi0, o), oi1 = __shfl_xor_sync(0xffffffffu, bi1, o);
          if (ob0 > best0 || (ob0 == best0 && oi0 < bi0)) { best0 = ob0; bi0 = oi0; }
          if (ob1 > best1 || (ob1 == best1 && oi1 < bi1)) { best1 = ob1; bi1 = oi1; }
        }
        // one-hot A fragments: assignments of points 2t, 2t+1 (quads 2t, 2t+1: "point g") and 2t+8, 2t+9 ("point g+8")
        const int a_p0 = __shfl_sync(0xffffffffu, bi0, 8 * t), a_p1 = __shfl_sync(0xffffffffu, bi0, 8 * t + 4);
        const int a_p8 = __shfl_sync(0xffffffffu, bi1, 8 * t), a_p9 = __shfl_sync(0xffffffffu, bi1, 8 * t + 4);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (16 * mt >= K) break;
          const int c_lo = g + 16 * mt, c_hi = c_lo + 8;
          uint32_t oh[4];
          oh[0] = pack_half2(a_p0 == c_lo ? 1.f : 0.f, a_p1 == c_lo ? 1.f : 0.f);
          oh[1] = pack_half2(a_p0 == c_hi ? 1.f : 0.f, a_p1 == c_hi ? 1.f : 0.f);
          oh[2] = pack_half2(a_p8 == c_lo ? 1.f : 0.f, a_p9 == c_lo ? 1.f : 0.f);
          oh[3] = pack_half2(a_p8 == c_hi ? 1.f : 0.f, a_p9 == c_hi ? 1.f : 0.f);
          mma16816(acc[mt], oh, pq[0], pq[1]);
        }
      }
      if (it < 4) KM_TRACE(9 + 8 * it);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float* row0 = s_fw + ((size_t)wid * 32 + g + 16 * mt) * 8 + 2 * t;
        row0[0] = acc[mt][0]; row0[1] = acc[mt][1];
        row0[64] = acc[mt][2]; row0[65] = acc[mt][3];            // cluster +8 -> 8 rows * 8 floats further
      }
      __syncthreads();
      if (tid < 256) {
        float tot = 0.f;                                          // 256 threads <-> 32 clusters x 8 columns, fixed order
#pragma unroll
        for (int w = 0; w < kKmThreads / 32; ++w) tot += s_fw[(size_t)w * 256 + tid];
        s_fpart[0][tid] = tot;                                    // this CTA's sums (local scratch)
      }
      if (it < 4) KM_TRACE(10 + 8 * it);
      __syncthreads();
      // exchange + centroid update by warp 0 alone (one lane per cluster); the other warps wait at the block barrier below
      if (wid == 0) {
        uint4 mine = make_uint4(0u, 0u, 0u, 0u);
        if (lane < K) {
          const float* r = s_fpart[0] + lane * 8;               // hi + lo halves of (x, y, z), member count
          mine.x = __float_as_uint(r[0] + r[3]); mine.y = __float_as_uint(r[1] + r[4]); mine.z = __float_as_uint(r[2] + r[5]);
          mine.w = __float_as_uint(r[6]);
        }
        const uint4* in = km_exchange<CL>(inbox, s_mbar, xround, rank, lane, K, mine);
        if (it < 4) KM_TRACE(11 + 8 * it);
        if (lane < K) {
          float sx = 0.f, sy = 0.f, sz = 0.f, sc = 0.f;           // fold in rank order: identical in every CTA
#pragma unroll
          for (int r = 0; r < CL; ++r) {
            const uint4 q = in[r * kKmInboxLanes + lane];
            sx += __uint_as_float(q.x); sy += __uint_as_float(q.y); sz += __uint_as_float(q.z); sc += __uint_as_float(q.w);
          }
          s_acc[4 * lane] = sx; s_acc[4 * lane + 1] = sy; s_acc[4 * lane + 2] = sz; s_acc[4 * lane + 3] = sc;
        }
        if (it < 4) KM_TRACE(12 + 8 * it);
        __syncwarp();
        bool empty = false;
        if (lane < K) {
          const float c = s_acc[4 * lane + 3];
          if (c > 0.f) {
            // spherical: normalize(sum / c) == normalize(sum) - the member count only scales the vector, skip the divisions
            const float ic = p.spherical ? 1.0f : 1.0f / c;
            s_c[3 * lane] = s_acc[4 * lane] * ic; s_c[3 * lane + 1] = s_acc[4 * lane + 1] * ic; s_c[3 * lane + 2] = s_acc[4 * lane + 2] * ic;
          } else empty = true;
        }
        const unsigned any_empty = __ballot_sync(0xffffffffu, empty);
        if (any_empty) {
          __syncwarp();
          if (lane == 0) split_empty_clusters(s_c, s_acc, K);
          __syncwarp();
        }
        if (lane < K && p.spherical) {
          const float l = sqrtf(s_c[3 * lane] * s_c[3 * lane] + s_c[3 * lane + 1] * s_c[3 * lane + 1] + s_c[3 * lane + 2] * s_c[3 * lane + 2]);
          if (l > 0.f) { const float il = 1.0f / l; s_c[3 * lane] *= il; s_c[3 * lane + 1] *= il; s_c[3 * lane + 2] *= il; }
        }
        // exact early exit: centroids bit-identical to the previous iteration's are a fixed point of the (deterministic) Lloyd map -
        // every further iteration would reproduce them, so the result equals running all niter iterations.  Every CTA holds the
        // same centroids, so the decision is uniform across the cluster without any extra exchange.
        __syncwarp();
        bool same = true;
        if (lane < K)
          same = __float_as_uint(s_c[3 * lane]) == __float_as_uint(prev_c[0]) && __float_as_uint(s_c[3 * lane + 1]) == __float_as_uint(prev_c[1]) &&
                 __float_as_uint(s_c[3 * lane + 2]) == __float_as_uint(prev_c[2]);
        if (lane < K) { prev_c[0] = s_c[3 * lane]; prev_c[1] = s_c[3 * lane + 1]; prev_c[2] = s_c[3 * lane + 2]; }
        const bool fixed = __all_sync(0xffffffffu, same);
        if (lane == 0) s_any_empty = fixed ? 1 : 0;               // (s_any_empty doubles as the block-wide "converged" flag on this path)
      }
      if (it < 4) KM_TRACE(14 + 8 * it);
      ++xround;
      __syncthreads();
      if (s_any_empty) break;                                     // (xround counts the exchanges actually done: parities stay consistent)
      continue;
    } else {
    int* wacc = s_wacc + wid * K * 4;
    for (int a = lane; a < K * 4; a += 32) wacc[a] = 0;
    __syncwarp();
    for (int j = tid; j < my_round; j += kKmThreads) {
      const bool v = j < my_n;
      float px = 0.f, py = 0.f, pz = 0.f;
      int b = 0;
      if (v) { px = xs[3 * j]; py = xs[3 * j + 1]; pz = xs[3 * j + 2]; b = best_centroid(px, py, pz, s_c, K); }
      warp_accumulate_by_key(b, v, px, py, pz, wacc, lane);
    }
    __syncthreads();
    int* part = s_part[it & 1];
    for (int a = tid; a < K * 4; a += kKmThreads) {
      int t = 0;
#pragma unroll
      for (int w = 0; w < kKmThreads / 32; ++w) t += s_wacc[w * K * 4 + a];
      part[a] = t;
    }
    cluster_sync_all();             // every CTA's partials are published
    for (int a = tid; a < K * 4; a += kKmThreads) {
      long long t = 0;
#pragma unroll
      for (uint32_t r = 0; r < CL; ++r) t += dsmem_ld_int(part + a, r);
      s_acc[a] = (a & 3) == 3 ? (float)t : (float)((double)t / (double)kKmFix);
    }
    if (tid == 0) s_any_empty = 0;
    __syncthreads();
    }   // scalar path (K > 32)
    // new centroids = member means (thread per cluster); empty clusters split the currently largest one
    // (faiss-style +-eps, rare -> one thread); then renormalise.  Replicated identically in every CTA.
    if (tid < K) {
      const float c = s_acc[4 * tid + 3];
      if (c > 0.f) { s_c[3 * tid] = s_acc[4 * tid] / c; s_c[3 * tid + 1] = s_acc[4 * tid + 1] / c; s_c[3 * tid + 2] = s_acc[4 * tid + 2] / c; }
      else s_any_empty = 1;
    }
    __syncthreads();
    if (tid == 0 && s_any_empty) split_empty_clusters(s_c, s_acc, K);
    __syncthreads();
    if (tid < K && p.spherical) {
      const float l = sqrtf(s_c[3 * tid] * s_c[3 * tid] + s_c[3 * tid + 1] * s_c[3 * tid + 1] + s_c[3 * tid + 2] * s_c[3 * tid + 2]);
      if (l > 0.f) { s_c[3 * tid] /= l; s_c[3 * tid + 1] /= l; s_c[3 * tid + 2] /= l; }
    }
    __syncthreads();
    // the partial buffer of iteration `it` is only overwritten in iteration it+2, i.e. after the barrier of it+1,
    // which every CTA reaches only after it finished reading buffer `it`
  }
  KM_TRACE(4);
  if (!use_tc) cluster_sync_all();  // pull-based exchange: no CTA may run ahead / exit while a peer can still read its shared memory
  KM_TRACE(5);
  // 5) final assignment of every valid row (kmeans.index.search, losses.py:89), split over the cluster + centroids out
  __shared__ int s_sz[kKmMaxK];                                   // chain: this CTA's member counts, read by the peers
  if (ch.labels != nullptr) { for (int j = tid; j < K; j += kKmThreads) s_sz[j] = 0; __syncthreads(); }
  for (int j = tid + (int)rank * kKmThreads; j < nv; j += kKmThreads * CL) {
    const int r = valid_idx[j];
    const int a = best_centroid(x[3 * r], x[3 * r + 1], x[3 * r + 2], s_c, K);
    assign[r] = a;
    if (ch.labels != nullptr) atomicAdd(&s_sz[a], 1);
  }
  if (rank == 0) for (int j = tid; j < K * 3; j += kKmThreads) centroids[j] = s_c[j];
  KM_TRACE(6);
  if (ch.labels == nullptr) return;
  // ---- 6) chain epilogue (K <= 32): selection, labels, cluster statistics and the three loss terms, with every CTA working on
  //         the rows it has just assigned.  Integer / fixed-point partials make the folds order independent; the float partials
  //         of the second pass are folded in rank order - the result is bit-reproducible run to run.
  __shared__ float s_sim[32 * 32];
  __shared__ int s_size[32], s_lab[32], s_cnt3[3];
  __shared__ long long s_p1[12], s_t1[12];                       // this CTA's / the cluster's fixed-point sums (x, y, z, count) x 3 clusters
  __shared__ float s_p2[16], s_c3[9], s_mu3[3], s_w2[kKmThreads / 32][16];
  for (int e = tid; e < K * K; e += kKmThreads) {
    const int i = e / K, j = e % K;
    s_sim[e] = s_c[3 * i] * s_c[3 * j] + s_c[3 * i + 1] * s_c[3 * j + 1] + s_c[3 * i + 2] * s_c[3 * j + 2];      // centrs @ centrs.T
  }
  if (tid < 12) s_p1[tid] = 0;
  __syncthreads();
  if (wid == 0) {                                                 // X1: member counts of all CTAs, then the selection (replicated)
    uint4 mine = make_uint4(lane < K ? (uint32_t)s_sz[lane] : 0u, 0u, 0u, 0u);
    const uint4* in = km_exchange<CL>(inbox, s_mbar, xround, rank, lane, K, mine);
    if (lane < K) {
      int t = 0;
#pragma unroll
      for (int r = 0; r < CL; ++r) t += (int)in[r * kKmInboxLanes + lane].x;
      s_size[lane] = t;
    }
    __syncwarp();
    int c1, c2, c3;
    select_triple_warp(s_sim, s_size, K, ch.t_similar, s_lab, c1, c2, c3, lane);
    if (rank == 0 && lane == 0) { ch.sel[0] = c1; ch.sel[1] = c2; ch.sel[2] = c3; }
  }
  ++xround;
  __syncthreads();
  // labels + first pass: sign-flipped member sums per selected cluster in 2^-20 fixed point - warp-aggregated (redux.sync) into a
  // warp-private row of integer accumulators, no shared-memory atomics (64-bit shared atomics compile to CAS spin loops)
  {
    int* wacc = s_wacc + wid * 12;
    if (lane < 12) wacc[lane] = 0;
    __syncwarp();
    const int j_end = ((nv + kKmThreads * CL - 1) / (kKmThreads * CL)) * (kKmThreads * CL);      // whole warps take part in every round
    for (int j = tid + (int)rank * kKmThreads; j < j_end; j += kKmThreads * CL) {
      int l = 0, k = 0;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (j < nv) {
        const int r = valid_idx[j];
        l = s_lab[assign[r]];
        ch.labels[r] = l;
        if (l != 0) {
          k = (l > 0 ? l : -l) - 1;
          const float sg = l > 0 ? 1.f : -1.f;
          px = sg * x[3 * r]; py = sg * x[3 * r + 1]; pz = sg * x[3 * r + 2];
        }
      }
      warp_accumulate_by_key(k, l != 0, px, py, pz, wacc, lane);
    }
  }
  __syncthreads();
  if (tid < 12) {
    long long t = 0;
#pragma unroll
    for (int w = 0; w < kKmThreads / 32; ++w) t += s_wacc[w * 12 + tid];
    s_p1[tid] = t;
  }
  __syncthreads();
  if (wid == 0) {                                                 // X2: fixed-point first-pass sums of all CTAs (12 x i64 = 6 slots)
    uint4 mine = make_uint4(0u, 0u, 0u, 0u);
    if (lane < 6) {
      const unsigned long long a = (unsigned long long)s_p1[2 * lane], b = (unsigned long long)s_p1[2 * lane + 1];
      mine = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
    }
    const uint4* in = km_exchange<CL>(inbox, s_mbar, xround, rank, lane, 6, mine);
    if (lane < 6) {
      long long ta = 0, tb = 0;
#pragma unroll
      for (int r = 0; r < CL; ++r) {
        const uint4 q = in[r * kKmInboxLanes + lane];
        ta += (long long)(((unsigned long long)q.y << 32) | q.x); tb += (long long)(((unsigned long long)q.w << 32) | q.z);
      }
      s_t1[2 * lane] = ta; s_t1[2 * lane + 1] = tb;
    }
    __syncwarp();
    if (lane < 3) {
      const int k = lane;
      const long long sx = s_t1[4 * k], sy = s_t1[4 * k + 1], sz = s_t1[4 * k + 2], cnt = s_t1[4 * k + 3];
      s_cnt3[k] = (int)cnt;
      if (cnt > 0) {                                              // same arithmetic as cluster_loss_fw_body
        const double inv = 1.0 / ((double)cnt * (double)kKmFix);
        const float mx = (float)((double)sx * inv), my = (float)((double)sy * inv), mz = (float)((double)sz * inv);
        const float len = sqrtf(mx * mx + my * my + mz * mz);
        const float d = fmaxf(len, 1e-12f);
        s_c3[3 * k] = mx / d; s_c3[3 * k + 1] = my / d; s_c3[3 * k + 2] = mz / d; s_mu3[k] = len;
      } else { s_c3[3 * k] = s_c3[3 * k + 1] = s_c3[3 * k + 2] = 0.f; s_mu3[k] = 0.f; }
    }
  }
  ++xround;
  __syncthreads();
  // second pass: per cluster sum of dots, of L1 distances and of sign(n' - c)
  float v2[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) v2[q] = 0.f;
  for (int j = tid + (int)rank * kKmThreads; j < nv; j += kKmThreads * CL) {
    const int r = valid_idx[j];
    const int l = s_lab[assign[r]];
    if (l == 0) continue;
    const int k = (l > 0 ? l : -l) - 1;
    const float sg = l > 0 ? 1.f : -1.f;
    const float px = sg * x[3 * r], py = sg * x[3 * r + 1], pz = sg * x[3 * r + 2];
    const float dx = px - s_c3[3 * k], dy = py - s_c3[3 * k + 1], dz = pz - s_c3[3 * k + 2];
#pragma unroll
    for (int q = 0; q < 3; ++q) if (q == k) {
      v2[q] += px * s_c3[3 * q] + py * s_c3[3 * q + 1] + pz * s_c3[3 * q + 2];
      v2[3 + q] += fabsf(dx) + fabsf(dy) + fabsf(dz);
      v2[6 + 3 * q] += sgnf(dx); v2[7 + 3 * q] += sgnf(dy); v2[8 + 3 * q] += sgnf(dz);
    }
  }
#pragma unroll
  for (int q = 0; q < 15; ++q) { const float t = warp_sum(v2[q]); if (lane == 0) s_w2[wid][q] = t; }
  __syncthreads();
  if (tid < 15) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kKmThreads / 32; ++w) t += s_w2[w][tid];
    s_p2[tid] = t;
  }
  __syncthreads();
  if (wid == 0) {                                                 // X3: second-pass float partials (15 values = 4 slots), folded in rank order
    uint4 mine = make_uint4(0u, 0u, 0u, 0u);
    if (lane < 4) mine = make_uint4(__float_as_uint(s_p2[4 * lane]), __float_as_uint(s_p2[4 * lane + 1]), __float_as_uint(s_p2[4 * lane + 2]),
                                    __float_as_uint(s_p2[4 * lane + 3]));
    const uint4* in = km_exchange<CL>(inbox, s_mbar, xround, rank, lane, 4, mine);
    float tot = 0.f;
    if (lane < 15) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(in);
#pragma unroll
      for (int r = 0; r < CL; ++r) tot += __uint_as_float(w[(r * kKmInboxLanes + (lane >> 2)) * 4 + (lane & 3)]);
    }
    // lane q holds total q: 0-2 dots, 3-5 L1, 6-14 sign sums
    const float d0 = __shfl_sync(0xffffffffu, tot, 0), d1 = __shfl_sync(0xffffffffu, tot, 1), d2 = __shfl_sync(0xffffffffu, tot, 2);
    const float a0 = __shfl_sync(0xffffffffu, tot, 3), a1 = __shfl_sync(0xffffffffu, tot, 4), a2 = __shfl_sync(0xffffffffu, tot, 5);
    if (rank == 0 && lane >= 6 && lane < 15) { const int k = (lane - 6) / 3, d = (lane - 6) % 3; ch.stats[8 * k + 5 + d] = tot; }
    if (rank == 0 && lane == 0) {
      const bool ok = s_cnt3[0] > 0 && s_cnt3[1] > 0 && s_cnt3[2] > 0;
      float l_ort = NAN, l_dot = NAN, l_l1 = NAN;
      if (ok) {
        auto dot3 = [&](int a, int b) { return s_c3[3 * a] * s_c3[3 * b] + s_c3[3 * a + 1] * s_c3[3 * b + 1] + s_c3[3 * a + 2] * s_c3[3 * b + 2]; };
        l_ort = (fabsf(dot3(0, 1)) + fabsf(dot3(0, 2)) + fabsf(dot3(1, 2))) / 3.0f;
        l_dot = ((1.f - d0 / s_cnt3[0]) + (1.f - d1 / s_cnt3[1]) + (1.f - d2 / s_cnt3[2])) / 3.0f;
        l_l1 = (a0 / s_cnt3[0] + a1 / s_cnt3[1] + a2 / s_cnt3[2]) / 3.0f;
      }
      ch.losses[0] = l_ort; ch.losses[1] = l_dot; ch.losses[2] = l_l1;
      for (int k = 0; k < 3; ++k) {
        ch.stats[8 * k] = (float)s_cnt3[k];
        ch.stats[8 * k + 1] = s_c3[3 * k]; ch.stats[8 * k + 2] = s_c3[3 * k + 1]; ch.stats[8 * k + 3] = s_c3[3 * k + 2];
        ch.stats[8 * k + 4] = s_mu3[k];
      }
      ch.stats[24] = l_ort; ch.stats[25] = l_dot; ch.stats[26] = l_l1; ch.stats[27] = ok ? 1.f : 0.f;
    }
  }
  cluster_sync_relaxed();                                         // leave together: no CTA exits while pushes into a peer are in flight
}
#ifdef NCN_KM_TRACE
extern "C" int ncn_debug_km_trace(long long* host_dst) {
  return cudaMemcpyFromSymbol(host_dst, g_km_trace, sizeof(long long) * 64) == cudaSuccess ? 0 : 1;
}
#endif

// ---------------------------------------------------------------- orthogonal-triple selection (one CTA)
__device__ __forceinline__ void cluster_select_body(const float* __restrict__ centroids, const int32_t* __restrict__ assign, int64_t n, int K,
                                                    float t_similar, int32_t* __restrict__ labels, int32_t* __restrict__ sel) {
  __shared__ int s_size[kKmMaxK];
  __shared__ int s_lab[kKmMaxK];
  __shared__ float s_sim[kKmMaxK * kKmMaxK];
  const int tid = threadIdx.x, lane = tid & 31;
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
  if (K <= 32) {
    // one lane per cluster, warp 0 only: the selection is a handful of warp arg-min / arg-max reductions
    if (tid < 32) {
      int c1, c2, c3;
      select_triple_warp(s_sim, s_size, K, t_similar, s_lab, c1, c2, c3, lane);
      if (lane == 0) { sel[0] = c1; sel[1] = c2; sel[2] = c3; }
    }
    __syncthreads();
    for (int64_t i = tid; i < n; i += blockDim.x) { const int a = assign[i]; labels[i] = a >= 0 ? s_lab[a] : 0; }
    return;
  }
  __shared__ float s_mn[kKmMaxK];
  __shared__ int s_arg[kKmMaxK], s_c1;
  if (tid == 0) {
    int c1 = 0;
    for (int j = 1; j < K; ++j) if (s_size[j] > s_size[c1]) c1 = j;            // biggest cluster (losses.py:104-107)
    s_c1 = c1;
  }
  __syncthreads();
  if (tid < K) {
    // criteria[i][j] = |sim[i,c1]| + |sim[c1,j]| + |sim[i,j]|; column j = tid: min / argmin over i (losses.py:117-118)
    const int c1 = s_c1, j = tid;
    float mn = INFINITY; int arg = 0;
    for (int i = 0; i < K; ++i) {
      const float v = fabsf(s_sim[i * K + c1]) + fabsf(s_sim[c1 * K + j]) + fabsf(s_sim[i * K + j]);
      if (v < mn) { mn = v; arg = i; }
    }
    s_mn[j] = mn; s_arg[j] = arg;
  }
  __syncthreads();
  if (tid == 0) {
    const int c1 = s_c1;
    int c2 = 0, c3 = 0; float best = INFINITY;
    for (int j = 0; j < K; ++j) if (s_mn[j] < best) { best = s_mn[j]; c2 = j; c3 = s_arg[j]; }      // losses.py:119-120
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
__global__ void __launch_bounds__(1024, 1)
cluster_select_kernel(const float* __restrict__ centroids, const int32_t* __restrict__ assign, int64_t n, int K,
                      float t_similar, int32_t* __restrict__ labels, int32_t* __restrict__ sel) {
  cluster_select_body(centroids, assign, n, K, t_similar, labels, sel);
}

// ---------------------------------------------------------------- cluster statistics and loss (one CTA)
// stats layout (floats): per cluster k in 0..2 at stats[8k..]: [count, cx, cy, cz, |mu|, gl1x, gl1y, gl1z]
// where gl1 = sum_i sign(n'_i - c_k); stats[24..26] = losses (ort, dot, L1); stats[27] = valid flag
__device__ __forceinline__ void cluster_loss_fw_body(const float* __restrict__ nrm, const int32_t* __restrict__ labels, int64_t n,
                                                     float* __restrict__ losses, float* __restrict__ stats) {
  __shared__ int s_wacc[32 * 12];
  __shared__ long long s_sum[9];
  __shared__ int s_cnt[3];
  __shared__ float s_c[9], s_mu[3];
  __shared__ float s_red[32][8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int a = tid; a < 32 * 12; a += blockDim.x) s_wacc[a] = 0;
  __syncthreads();
  const int64_t n_round = (n + 31) & ~(int64_t)31;
  // up to kPer points per thread are fetched ONCE, in one batch of independent loads, and kept in registers for both passes
  constexpr int kPer = 8;
  const bool in_regs = n <= (int64_t)kPer * blockDim.x;
  int rl[kPer];
  float rx[kPer], ry[kPer], rz[kPer];
  if (in_regs) {
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int64_t i = tid + (int64_t)q * blockDim.x;
      rl[q] = i < n ? labels[i] : 0;
    }
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int64_t i = tid + (int64_t)q * blockDim.x;
      rx[q] = ry[q] = rz[q] = 0.f;
      if (rl[q] != 0) {
        const float sg = rl[q] > 0 ? 1.f : -1.f;
        rx[q] = sg * nrm[3 * i]; ry[q] = sg * nrm[3 * i + 1]; rz[q] = sg * nrm[3 * i + 2];
      }
    }
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int64_t i = tid + (int64_t)q * blockDim.x;
      if (i < n_round) {
        const int l = rl[q];
        warp_accumulate_by_key(l != 0 ? (l > 0 ? l : -l) - 1 : 0, l != 0, rx[q], ry[q], rz[q], s_wacc + wid * 12, lane);
      }
    }
  } else {
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
      warp_accumulate_by_key(k, v, x, y, z, s_wacc + wid * 12, lane);
    }
  }
  __syncthreads();
  if (tid < 12) {
    long long t = 0;
    for (int w = 0; w < 32; ++w) t += s_wacc[w * 12 + tid];
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
  auto second_pass = [&](int l, float x, float y, float z) {
    const int k = (l > 0 ? l : -l) - 1;
    const float dx = x - s_c[3 * k], dy = y - s_c[3 * k + 1], dz = z - s_c[3 * k + 2];
#pragma unroll
    for (int q = 0; q < 3; ++q) if (q == k) {
      dotv[q] += x * s_c[3 * q] + y * s_c[3 * q + 1] + z * s_c[3 * q + 2];
      l1v[q] += fabsf(dx) + fabsf(dy) + fabsf(dz);
      sg[3 * q] += sgnf(dx); sg[3 * q + 1] += sgnf(dy); sg[3 * q + 2] += sgnf(dz);
    }
  };
  if (in_regs) {
#pragma unroll
    for (int q = 0; q < kPer; ++q) if (rl[q] != 0) second_pass(rl[q], rx[q], ry[q], rz[q]);
  } else {
    for (int64_t i = tid; i < n; i += blockDim.x) {
      const int l = labels[i];
      if (l == 0) continue;
      const float s = l > 0 ? 1.f : -1.f;
      second_pass(l, s * nrm[3 * i], s * nrm[3 * i + 1], s * nrm[3 * i + 2]);
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
__global__ void __launch_bounds__(1024, 1)
cluster_loss_fw_kernel(const float* __restrict__ nrm, const int32_t* __restrict__ labels, int64_t n,
                       float* __restrict__ losses, float* __restrict__ stats) {
  cluster_loss_fw_body(nrm, labels, n, losses, stats);
}

// dL/dn for  L = w[0] L_ort + w[1] L_dot + w[2] L_L1   (w read from device memory)
__device__ __forceinline__ void cluster_loss_bw_body(const float* __restrict__ nrm, const int32_t* __restrict__ labels, int64_t n,
                                                     const float* __restrict__ stats, const float* __restrict__ w, float* __restrict__ dn,
                                                     int64_t first, int64_t stride) {
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
  for (int64_t i = first; i < n; i += stride) {
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
__global__ void __launch_bounds__(256)
cluster_loss_bw_kernel(const float* __restrict__ nrm, const int32_t* __restrict__ labels, int64_t n,
                       const float* __restrict__ stats, const float* __restrict__ w, float* __restrict__ dn) {
  cluster_loss_bw_body(nrm, labels, n, stats, w, dn, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// Everything between the k-means result and dL/ddepth in TWO launches instead of four: the two single-CTA stages
// (selection -> cluster statistics + losses) share one launch, the two per-triangle stages (dL/dnormals -> dL/ddepth) the
// other - a thread consumes the dL/dnormal it has just written, so no grid-wide ordering is needed.
__global__ void __launch_bounds__(1024, 1)
cluster_select_loss_kernel(const float* __restrict__ centroids, const int32_t* __restrict__ assign, int64_t n, int K, float t_similar,
                           int32_t* __restrict__ labels, int32_t* __restrict__ sel, const float* __restrict__ nrm,
                           float* __restrict__ losses, float* __restrict__ stats) {
  cluster_select_body(centroids, assign, n, K, t_similar, labels, sel);
  __syncthreads();
  cluster_loss_fw_body(nrm, labels, n, losses, stats);
}
__global__ void __launch_bounds__(256)
cluster_bw_depth_kernel(const float* __restrict__ nrm, const int32_t* __restrict__ labels, int64_t n, const float* __restrict__ stats,
                        const float* __restrict__ w, float* __restrict__ dn, const float* __restrict__ origin,
                        const float* __restrict__ dir, const float* __restrict__ depth, const int64_t* __restrict__ i1,
                        const int64_t* __restrict__ i2, const int64_t* __restrict__ i3, float* __restrict__ ddepth) {
  pdl_wait(); pdl_trigger();
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  cluster_loss_bw_body(nrm, labels, n, stats, w, dn, first, stride);
  normals_bw_body(origin, dir, depth, i1, i2, i3, dn, n, ddepth, first, stride);
}

// ---------------------------------------------------------------- photometric terms (fused fwd + grad)
// rgb = rend[:, :3] + bg*(1-opacity)  (rendering.py:231-241);  L_rgb = mean((rgb-target)^2) (losses.py:353);
// L_op = w_op * mean(-(o+1e-10) log(o+1e-10)) (losses.py:357-361).  sums[0] += sum sq err, sums[1] += sum entropy.
// Gradients are multiplied by `gscale` (loss scale / grad divisor) and written (not accumulated).
__global__ void __launch_bounds__(256)
photometric_kernel(const float* __restrict__ rend, const float* __restrict__ opacity, const float* __restrict__ target,
                   int64_t n_rays, int64_t n_gt, int C, float bg0, float bg1, float bg2, float opacity_w, float gscale,
                   float* __restrict__ rgb_out, float* __restrict__ sums, float* __restrict__ d_rend,
                   float* __restrict__ d_opacity) {
  float se = 0.f, ent = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float inv3n = n_gt > 0 ? 1.0f / (3.0f * (float)n_gt) : 0.f, invn = 1.0f / (float)n_rays;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += stride) {
    const float o = opacity[r];
    const float bg[3] = {bg0, bg1, bg2};
    float go = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = rend[r * C + c] + bg[c] * (1.f - o);
      if (rgb_out) rgb_out[3 * r + c] = v;
      const float e = r < n_gt ? v - target[3 * r + c] : 0.f;      // rays beyond n_gt have no target colour (random_tr_poses)
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

// ---------------------------------------------------------------- normals of a rendered depth IMAGE (evaluation)
// datasets/hypersim_src/utils.py:544-611 (_extract_normals_from_depth_batch): P = ray_dir_cc * depth (camera frame),
// n(y,x) = normalize(cross(P(y-1,x) - P(y,x), P(y,x-1) - P(y,x))) (eps 1e-12), rotated to the world frame by the pose's 3x3
// block; the one-pixel border and pixels whose own depth is 0 / NaN / Inf get (0,0,0).  One thread per pixel; the three
// points of a pixel are recomputed from depth + directions (28 B/pixel of reads that neighbours share through L1/L2, 12 B written).
__global__ void __launch_bounds__(256)
normals_image_kernel(const float* __restrict__ depth, const float* __restrict__ dirs, const float* __restrict__ poses,
                     int pose_stride, int B, int H, int W, float* __restrict__ out) {
  const int64_t hw = (int64_t)H * W;
  const int64_t total = (int64_t)B * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    const int64_t p = i - (int64_t)b * hw;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    const float d1 = depth[i];
    if (y >= 1 && y < H - 1 && x >= 1 && x < W - 1 && d1 != 0.f && isfinite(d1)) {
      const float d2 = depth[i - W], d3 = depth[i - 1];
      const float* r1 = dirs + 3 * p; const float* r2 = dirs + 3 * (p - W); const float* r3 = dirs + 3 * (p - 1);
      const float p1x = r1[0] * d1, p1y = r1[1] * d1, p1z = r1[2] * d1;
      const float ax = r2[0] * d2 - p1x, ay = r2[1] * d2 - p1y, az = r2[2] * d2 - p1z;
      const float bx = r3[0] * d3 - p1x, by = r3[1] * d3 - p1y, bz = r3[2] * d3 - p1z;
      float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
      const float inv = 1.f / fmaxf(sqrtf(cx * cx + cy * cy + cz * cz), 1e-12f);
      cx *= inv; cy *= inv; cz *= inv;
      const float* R = poses + (int64_t)b * pose_stride;       // row-major (3|4, 4): R[r][c] = R[4 r + c]
      n0 = R[0] * cx + R[1] * cy + R[2] * cz;
      n1 = R[4] * cx + R[5] * cy + R[6] * cz;
      n2 = R[8] * cx + R[9] * cy + R[10] * cz;
    }
    out[3 * i] = n0; out[3 * i + 1] = n1; out[3 * i + 2] = n2;
  }
}

// ---------------------------------------------------------------- semantic cross-entropy (fused fwd + grad)
// losses.py:226-242, 569-573: CrossEntropyLoss(ignore_index=-1)(sem_pred, target - 1), mean over non-void rays.
// Thread per ray; every CTA counts the valid rays itself (R labels from L2) so the mean's divisor needs no second launch.
constexpr int kCeMaxCls = 64;
__global__ void __launch_bounds__(256)
semantic_ce_kernel(const float* __restrict__ rend, int C, int c_off, int n_cls, const int64_t* __restrict__ labels,
                   int64_t n_rays, float gscale, float* __restrict__ sums, float* __restrict__ d_rend) {
  __shared__ int s_cnt[8];
  __shared__ float s_sum[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int cnt = 0;
  for (int64_t base = 0; base < n_rays; base += 8 * 256) {      // 8 independent loads in flight per thread
    int64_t v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int64_t r = base + u * 256 + threadIdx.x; v[u] = r < n_rays ? __ldg(labels + r) : 0; }
#pragma unroll
    for (int u = 0; u < 8; ++u) cnt += (v[u] >= 1 && v[u] <= n_cls) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_cnt[wid] = cnt;
  __syncthreads();
  int n_valid = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) n_valid += s_cnt[q];
  const float inv = n_valid > 0 ? gscale / (float)n_valid : 0.f;
  float acc = 0.f;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rays) {
    const float* row = rend + r * C + c_off;
    const int64_t t = labels[r] - 1;
    const bool valid = t >= 0 && t < n_cls;
    float mx = -INFINITY;
    for (int c = 0; c < n_cls; ++c) mx = fmaxf(mx, row[c]);
    float se = 0.f;
    for (int c = 0; c < n_cls; ++c) se += expf(row[c] - mx);
    const float lse = mx + logf(se);
    if (valid) acc = lse - row[t];
    if (d_rend) {
      float* g = d_rend + r * C + c_off;
      for (int c = 0; c < n_cls; ++c)
        g[c] = valid ? (expf(row[c] - lse) - (c == (int)t ? 1.f : 0.f)) * inv : 0.f;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) s_sum[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) tot += s_sum[q];
    atomicAdd(sums, tot);
    if (blockIdx.x == 0) sums[1] = (float)n_valid;
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

static int launch_kmeans(const float* x, int64_t n_points, const ncn_kmeans_params* p, float* centroids, int32_t* assign,
                         int32_t* n_valid, void* workspace, size_t workspace_bytes, const ChainArgs& ch, ncn_stream_t stream) {
  NCN_CHECK_PTR(p); NCN_CHECK_PTR(centroids); NCN_CHECK_PTR(n_valid);
  if (p->k < 1 || p->k > kKmMaxK || p->niter < 0 || p->max_points_per_centroid < 1) return NCN_E_CONFIG;
  NCN_CHECK_SIZE(n_points >= 0 && n_points < ((int64_t)1 << 31));
  if (n_points > 0) { NCN_CHECK_PTR(x); NCN_CHECK_PTR(assign); NCN_CHECK_PTR(workspace); }
  if (workspace_bytes < ncn_kmeans_workspace_bytes(n_points, p->k)) return NCN_E_SIZE;
  int64_t cap = (int64_t)p->max_points_per_centroid * p->k;
  if (cap > 65536) cap = 65536;
  if (cap > n_points) cap = n_points > 0 ? n_points : 1;
  // 16-CTA cluster (non-portable size, one GPC) when the device grants it, else the portable 8
  static int cluster = 0;
  if (cluster == 0) {
    cluster = kKmCluster;
    if (cudaFuncSetAttribute(kmeans_kernel<kKmClusterMax>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(kKmClusterMax); q.blockDim = dim3(kKmThreads); q.dynamicSmemBytes = 96 * 1024;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = kKmClusterMax; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int n_clusters = 0;
      cudaFuncSetAttribute(kmeans_kernel<kKmClusterMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (cudaOccupancyMaxActiveClusters(&n_clusters, kmeans_kernel<kKmClusterMax>, &q) == cudaSuccess && n_clusters >= 1) cluster = kKmClusterMax;
    }
    (void)cudaGetLastError();
  }
  const size_t per_cta = (size_t)(cap + cluster - 1) / cluster;
  // this CTA's points (fp32) + fp16 rows + the two exchange inboxes (cluster x 32 slots x 16 B each)
  const size_t smem = per_cta * 12 + 16 + (per_cta + 16) * 16 + 16 + (size_t)2 * cluster * 32 * 16;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster); cfg.blockDim = dim3(kKmThreads); cfg.dynamicSmemBytes = smem; cfg.stream = as_stream(stream);
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  int32_t* wsp = (int32_t*)workspace;
  const int cap_i = (int)cap;
  // static (~28 KB) + dynamic shared memory can cross the 48 KB default limit: always opt in
  if (cluster == kKmClusterMax) {
    NCN_CUDA(cudaFuncSetAttribute(kmeans_kernel<kKmClusterMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NCN_CUDA(cudaLaunchKernelEx(&cfg, kmeans_kernel<kKmClusterMax>, x, n_points, *p, centroids, assign, n_valid, wsp, cap_i, ch));
  } else {
    NCN_CUDA(cudaFuncSetAttribute(kmeans_kernel<kKmCluster>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NCN_CUDA(cudaLaunchKernelEx(&cfg, kmeans_kernel<kKmCluster>, x, n_points, *p, centroids, assign, n_valid, wsp, cap_i, ch));
  }
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_kmeans_spherical(const float* x, int64_t n_points, const ncn_kmeans_params* p, float* centroids,
                                    int32_t* assign, int32_t* n_valid, void* workspace, size_t workspace_bytes,
                                    ncn_stream_t stream) {
  ChainArgs ch = ChainArgs();
  return launch_kmeans(x, n_points, p, centroids, assign, n_valid, workspace, workspace_bytes, ch, stream);
}

// normals from the rendered depth -> spherical k-means -> triple selection -> cluster statistics and the three loss terms in ONE
// cluster launch (K <= 32): = ncn_normals_from_depth_fw + ncn_kmeans_spherical + ncn_cluster_select + ncn_cluster_loss_fw
extern "C" int ncn_cluster_chain(const float* origin, const float* dir, const float* depth, const int64_t* idx1, const int64_t* idx2,
                                 const int64_t* idx3, int64_t n_tri, const ncn_kmeans_params* p, float t_similar, float* normals,
                                 float* centroids, int32_t* assign, int32_t* n_valid, int32_t* labels, int32_t* sel, float* losses,
                                 float* stats, void* workspace, size_t workspace_bytes, ncn_stream_t stream) {
  NCN_CHECK_PTR(p);
  if (p->k < 3 || p->k > 32) return NCN_E_UNSUPPORTED;
  NCN_CHECK_SIZE(n_tri >= 0);
  if (n_tri > 262144) return NCN_E_UNSUPPORTED;      // the epilogue's warp-private 32-bit fixed-point sums cover 2047 rows per warp
  NCN_CHECK_PTR(sel); NCN_CHECK_PTR(losses); NCN_CHECK_PTR(stats);
  if (n_tri > 0) {
    NCN_CHECK_PTR(origin); NCN_CHECK_PTR(dir); NCN_CHECK_PTR(depth); NCN_CHECK_PTR(idx1); NCN_CHECK_PTR(idx2); NCN_CHECK_PTR(idx3);
    NCN_CHECK_PTR(normals); NCN_CHECK_PTR(labels);
  }
  ChainArgs ch;
  ch.origin = origin; ch.dir = dir; ch.depth = depth; ch.i1 = idx1; ch.i2 = idx2; ch.i3 = idx3; ch.normals_out = normals;
  ch.t_similar = t_similar; ch.labels = labels; ch.sel = sel; ch.losses = losses; ch.stats = stats;
  if (n_tri == 0) { ch.origin = nullptr; }
  // (the kernel reads its points from `normals` after the prologue; launch_kmeans only checks that the pointer is set)
  return launch_kmeans(normals, n_tri, p, centroids, assign, n_valid, workspace, workspace_bytes, ch, stream);
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

extern "C" int ncn_cluster_tail(const float* centroids, const int32_t* assign, int64_t n_points, int k, float t_similar,
                                int32_t* labels, int32_t* sel, const float* normals, float* losses, float* stats,
                                const float* weights_dev, float* dL_dnormals, const float* origin, const float* dir,
                                const float* depth, const int64_t* idx1, const int64_t* idx2, const int64_t* idx3,
                                float* dL_ddepth, ncn_stream_t stream) {
  NCN_CHECK_PTR(centroids); NCN_CHECK_PTR(sel); NCN_CHECK_PTR(losses); NCN_CHECK_PTR(stats); NCN_CHECK_PTR(weights_dev);
  if (k < 3 || k > kKmMaxK) return NCN_E_CONFIG;
  NCN_CHECK_SIZE(n_points >= 0);
  if (n_points > 0) {
    NCN_CHECK_PTR(assign); NCN_CHECK_PTR(labels); NCN_CHECK_PTR(normals); NCN_CHECK_PTR(dL_dnormals); NCN_CHECK_PTR(origin);
    NCN_CHECK_PTR(dir); NCN_CHECK_PTR(depth); NCN_CHECK_PTR(idx1); NCN_CHECK_PTR(idx2); NCN_CHECK_PTR(idx3); NCN_CHECK_PTR(dL_ddepth);
  }
  cluster_select_loss_kernel<<<1, 1024, 0, as_stream(stream)>>>(centroids, assign, n_points, k, t_similar, labels, sel, normals, losses, stats);
  NCN_LAUNCH_OK();
  if (n_points > 0)
    cluster_bw_depth_kernel<<<persistent_grid(n_points, 256, 8), 256, 0, as_stream(stream)>>>(normals, labels, n_points, stats, weights_dev,
                                                                                             dL_dnormals, origin, dir, depth, idx1, idx2, idx3,
                                                                                             dL_ddepth);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

// second half of ncn_cluster_tail on its own (after ncn_cluster_chain): dL/dnormals of the three cluster terms and, through the
// normals, dL/ddepth - one launch over the triangles
extern "C" int ncn_cluster_bw_depth(const float* normals, const int32_t* labels, int64_t n_points, const float* stats,
                                    const float* weights_dev, float* dL_dnormals, const float* origin, const float* dir,
                                    const float* depth, const int64_t* idx1, const int64_t* idx2, const int64_t* idx3,
                                    float* dL_ddepth, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_points >= 0);
  if (n_points == 0) return NCN_OK;
  NCN_CHECK_PTR(normals); NCN_CHECK_PTR(labels); NCN_CHECK_PTR(stats); NCN_CHECK_PTR(weights_dev); NCN_CHECK_PTR(dL_dnormals);
  NCN_CHECK_PTR(origin); NCN_CHECK_PTR(dir); NCN_CHECK_PTR(depth); NCN_CHECK_PTR(idx1); NCN_CHECK_PTR(idx2); NCN_CHECK_PTR(idx3);
  NCN_CHECK_PTR(dL_ddepth);
  NCN_CUDA(launch_pdl(cluster_bw_depth_kernel, dim3(persistent_grid(n_points, 256, 8)), dim3(256), 0, as_stream(stream), normals, labels, n_points,
                      stats, weights_dev, dL_dnormals, origin, dir, depth, idx1, idx2, idx3, dL_ddepth));
  return NCN_OK;
}

extern "C" int ncn_photometric_loss_gt(const float* rend, const float* opacity, const float* target_rgb, int64_t n_rays,
                                       int64_t n_gt_rays, int n_channels, const float* bg_rgb_host, float opacity_w, float grad_scale,
                                       float* rgb_out, float* sums, float* dL_drend, float* dL_dopacity, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_rays >= 0 && n_channels >= 3 && n_gt_rays >= 0 && n_gt_rays <= n_rays);
  if (n_rays == 0) return NCN_OK;
  NCN_CHECK_PTR(rend); NCN_CHECK_PTR(opacity); NCN_CHECK_PTR(bg_rgb_host); NCN_CHECK_PTR(sums);
  if (n_gt_rays > 0) NCN_CHECK_PTR(target_rgb);
  photometric_kernel<<<persistent_grid(n_rays, 256, 4), 256, 0, as_stream(stream)>>>(
      rend, opacity, target_rgb, n_rays, n_gt_rays, n_channels, bg_rgb_host[0], bg_rgb_host[1], bg_rgb_host[2], opacity_w, grad_scale,
      rgb_out, sums, dL_drend, dL_dopacity);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_photometric_loss(const float* rend, const float* opacity, const float* target_rgb, int64_t n_rays,
                                    int n_channels, const float* bg_rgb_host, float opacity_w, float grad_scale,
                                    float* rgb_out, float* sums, float* dL_drend, float* dL_dopacity,
                                    ncn_stream_t stream) {
  return ncn_photometric_loss_gt(rend, opacity, target_rgb, n_rays, n_rays, n_channels, bg_rgb_host, opacity_w, grad_scale, rgb_out, sums,
                                 dL_drend, dL_dopacity, stream);
}

extern "C" int ncn_semantic_ce_loss(const float* rend, int c_total, int c_off, int n_cls, const int64_t* labels, int64_t n_rays,
                                    float grad_scale, float* sums, float* dL_drend, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_rays >= 0 && n_cls >= 1 && n_cls <= kCeMaxCls && c_off >= 0 && c_off + n_cls <= c_total);
  if (n_rays == 0) return NCN_OK;
  NCN_CHECK_PTR(rend); NCN_CHECK_PTR(labels); NCN_CHECK_PTR(sums);
  semantic_ce_kernel<<<(unsigned)ceil_div(n_rays, 256), 256, 0, as_stream(stream)>>>(rend, c_total, c_off, n_cls, labels, n_rays,
                                                                                    grad_scale, sums, dL_drend);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_normals_from_depth_image(const float* depth, const float* ray_dirs_cc, const float* poses, int pose_rows,
                                            int n_images, int height, int width, float* normals, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_images >= 0 && height >= 0 && width >= 0 && (pose_rows == 3 || pose_rows == 4));
  const int64_t total = (int64_t)n_images * height * width;
  if (total == 0) return NCN_OK;
  NCN_CHECK_PTR(depth); NCN_CHECK_PTR(ray_dirs_cc); NCN_CHECK_PTR(poses); NCN_CHECK_PTR(normals);
  normals_image_kernel<<<persistent_grid(total, 256, 8), 256, 0, as_stream(stream)>>>(depth, ray_dirs_cc, poses, pose_rows * 4, n_images,
                                                                                     height, width, normals);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
