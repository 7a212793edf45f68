// Density trunk + colour head FORWARD on the 5th-generation tensor cores (tcgen05), accumulators in tensor memory.
// Replaces the warp-MMA (mma.sync) form of ncn_field_mlp_fwd for the training / evaluation step of models/ngp_mt.py:157-229:
//
//   feat (N,32) -> [64 ReLU] -> h (16) ; sigma = TruncExp(h0) ; x_rgb = [h | d/|d| | 1] -> [64 ReLU] -> [64 ReLU] -> sigmoid -> rgb
//
// One CTA = 128 threads = one 128-sample tile at a time (thread r <-> sample r <-> TMEM lane r), persistent over tiles.
// Every layer is ONE accumulator tile  D[128 x N] = A[128 x K] * W^T  issued by one elected thread:
//   A = the previous layer's activations as a K-major panel [K/8][128][8 halfs] in shared memory (no swizzle; the canonical
//       core-matrix layout), written by the 128 threads straight from their accumulator rows (tcgen05.ld -> ReLU -> fp16);
//   B = the layer's weight matrix (out x in, row-major = K-major) staged once per CTA as a panel [in/8][out][8 halfs];
//   D = 64 TMEM columns, reused layer after layer (the chain of a tile is serial by data dependence).
// The hidden-activation panels ARE the layout the tcgen05 backward reads (ncn_common.cuh act_offset: one contiguous 16 KB block
// per tile and layer), so each of them leaves with ONE bulk store (cp.async.bulk shared -> global) while the next layer's MMA
// is already running; the row-major outputs (h, x_rgb, rgb_out, sigmas, raws) are written by their owner threads, 32-64
// contiguous bytes each.  The next tile's feature rows are prefetched with cp.async during the current tile's chain.
// Column order of the colour head's input: [h (16) | d (3) | ones (13)] (see field_fused.cu) - the first layer's weight
// panel is permuted accordingly when it is staged.
#include "ncn_common.cuh"
#include "mma.cuh"
#include "tc05.cuh"

namespace ncn {

__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// K-major operand over a panel [K/8][rows][8 halfs]: rows 16 B apart, next 8 rows SBO = 128 B, next 8 K-elements LBO = rows*16 B
__device__ __forceinline__ uint64_t make_desc_kmajor(const void* panel, int rows) {
  const uint64_t addr = (uint64_t)(smem_u32(panel) >> 4) & 0x3FFF;
  const uint64_t lbo = (uint64_t)((rows * 16) >> 4), sbo = (128 >> 4);
  return addr | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// instruction descriptor: D = f32, A / B = f16, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_fwd(int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// weight matrix (N rows = outputs, K cols = inputs, row-major fp16) -> K-major panel [K/8][N][8]: 16-byte chunks
__device__ __forceinline__ void load_wk_panel_async(const __half* __restrict__ w, int N, int K, __half* __restrict__ P) {
  const int kc_count = K >> 3;
  for (int i = threadIdx.x; i < N * kc_count; i += blockDim.x) {
    const int nrow = i / kc_count, kc = i - nrow * kc_count;
    cp_async16(P + ((size_t)kc * N + nrow) * 8, w + (size_t)nrow * K + kc * 8);
  }
}
// same with the input columns permuted to the fused-forward order [h | d | 1]
__device__ __forceinline__ void load_wk_panel_perm(const __half* __restrict__ w, int N, int K, __half* __restrict__ P) {
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int nrow = i / K, k = i % K;
    P[((size_t)(k >> 3) * N + nrow) * 8 + (k & 7)] = w[(size_t)nrow * K + perm_col(k)];
  }
}

#ifdef NCN_TC05_TRACE      // developer build only: per-tile phase clocks of CTA 0
__device__ long long g_fwd_trace[16 * 16];
#define FWD_TRACE(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0 && it < 16) g_fwd_trace[it * 16 + (slot)] = clock64(); } while (0)
#else
#define FWD_TRACE(slot) do { } while (0)
#endif

struct FwdLayout {
  // weight panels (halfs)
  static constexpr int kS0 = 32 * 64, kS1 = 64 * 16, kR0 = 32 * 64, kR1 = 64 * 64, kR2 = 64 * 16;
  static constexpr int kWeights = kS0 + kS1 + kR0 + kR1 + kR2;                   // 10240 halfs = 20 KB
  // activation panels (halfs): feat x2 (prefetch), sigma hidden, x_rgb, rgb hidden 0, rgb hidden 1
  static constexpr int kPx = 32 * kTile, kPh = 64 * kTile;
  static constexpr int kPanels = 2 * kPx + kPh + kPx + kPh + kPh;                // 36864 halfs = 72 KB
  // row-major output staging: x_rgb 8 KB + h 4 KB + rgb_out 4 KB + sigmas 512 B + raws 1536 B
  static constexpr int kStageBytes = kTile * (32 + 16 + 16) * 2 + kTile * 4 + kTile * 3 * 4;
  static constexpr size_t kBytes = (size_t)(kWeights + kPanels) * 2 + kStageBytes + 64;
  static constexpr int kTmemCols = 64;
};

__device__ __forceinline__ void store_row_chunks(__half* __restrict__ panel, int tid, const uint32_t* v, int n_cols) {
  // fp32 accumulator columns [0, n_cols) of this thread's row -> fp16 chunks of the K-major panel
  for (int c = 0; c < n_cols / 8; ++c) {
    const uint4 q = make_uint4(pack_half2(__uint_as_float(v[8 * c]), __uint_as_float(v[8 * c + 1])), pack_half2(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3])),
                               pack_half2(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5])), pack_half2(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7])));
    *reinterpret_cast<uint4*>(panel + ((size_t)c * kTile + tid) * 8) = q;
  }
}

__global__ void __launch_bounds__(kTcThreads, 2)
field_mlp_fwd_tc05_kernel(const __half* __restrict__ feat, const float* __restrict__ dirs, const __half* __restrict__ w_sigma,
                          const __half* __restrict__ w_rgb, int64_t n_cap, const int32_t* __restrict__ n_dev, float* __restrict__ sigmas,
                          float* __restrict__ raws, int c_total, __half* __restrict__ h_out, __half* __restrict__ sig_acts,
                          __half* __restrict__ x_rgb, __half* __restrict__ rgb_acts, __half* __restrict__ rgb_out) {
  using LY = FwdLayout;
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  extern __shared__ __align__(128) unsigned char fw_smem[];
  __half* WS0 = reinterpret_cast<__half*>(fw_smem);          // sigma layer 0: N = 64, K = 32
  __half* WS1 = WS0 + LY::kS0;                               // sigma layer 1: N = 16, K = 64
  __half* WR0 = WS1 + LY::kS1;                               // rgb layer 0:   N = 64, K = 32 (columns permuted)
  __half* WR1 = WR0 + LY::kR0;                               // rgb layer 1:   N = 64, K = 64
  __half* WR2 = WR1 + LY::kR1;                               // rgb layer 2:   N = 16, K = 64
  __half* PX0 = WR2 + LY::kR2;                               // feature rows, two sets (prefetch)
  __half* PH = PX0 + 2 * LY::kPx;                            // sigma hidden
  __half* PX1 = PH + LY::kPh;                                // colour-head input
  __half* PA1 = PX1 + LY::kPx;                               // rgb hidden 0
  __half* PA2 = PA1 + LY::kPh;                               // rgb hidden 1
  // row-major staging of a tile's outputs: each leaves with one bulk store (a tile's rows are contiguous in global memory)
  __half* OX = PA2 + LY::kPh;                                // x_rgb   [128][32]
  __half* OH = OX + kTile * 32;                              // h       [128][16]
  __half* OR = OH + kTile * 16;                              // rgb_out [128][16]
  float* OS = reinterpret_cast<float*>(OR + kTile * 16);     // sigmas  [128]
  float* OW = OS + kTile;                                    // raws    [128][3]  (c_total == 3 only)
  uint64_t* mbar = reinterpret_cast<uint64_t*>(OW + kTile * 3);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tid = threadIdx.x, wid = tid >> 5;
  const bool mma_warp = __shfl_sync(0xffffffffu, wid, 0) == 0;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int64_t layer_stride = act_rows(n_cap) * 64;

  if (tid == 0) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  int64_t cur = blockIdx.x;
  // the weights were written by the optimizer at the head of the step: staged before the wait on the predecessor (the encoder)
  load_wk_panel_async(w_sigma, 64, 32, WS0);
  load_wk_panel_async(w_sigma + 64 * 32, 16, 64, WS1);
  load_wk_panel_perm(w_rgb, 64, 32, WR0);
  load_wk_panel_async(w_rgb + 64 * 32, 64, 64, WR1);
  load_wk_panel_async(w_rgb + 64 * 32 + 64 * 64, 16, 64, WR2);
  pdl_wait(); pdl_trigger();
  if (cur < n_tiles) stage_rows<32>(feat, cur * kTile, n, PX0);
  cp_async_commit();
  __syncwarp();
  if (wid == 0) tmem_alloc<LY::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t my_lane = tmem_d + ((uint32_t)(32 * wid) << 16);

  // one layer: [all rows written, visible to the async proxy] -> elected thread issues the k-steps and commits
  auto issue = [&](const __half* A, const __half* W, int K, int N) {
    if (mma_warp) {
      if (elect_one()) {
        tc_fence_after();
        const uint64_t a = make_desc_kmajor(A, kTile), b = make_desc_kmajor(W, N);
        const uint32_t idesc = make_idesc_fwd(N);
        for (int ks = 0; ks < K / 16; ++ks)
          tc_mma_f16(tmem_d, desc_add(a, (uint32_t)(ks * 2 * kTile * 16)), desc_add(b, (uint32_t)(ks * 2 * N * 16)), idesc, ks > 0 ? 1u : 0u);
        tc_commit(mbar);
      }
      __syncwarp();
    }
  };
  // Bulk-store groups leave in a fixed order, five per tile: G1 {sigma hidden} G2 {x_rgb, h, sigmas} G3 {rgb hidden 0} G4 {rgb hidden 1}
  // G5 {rgb_out, raws}.  A staging buffer is rewritten exactly five groups after the one that reads it was committed, and a group is
  // committed right after each of the tile's five block barriers - so "at most 3 groups still reading" in front of every barrier
  // (thread 0) is exactly the guarantee the writes behind that barrier need; nobody ever waits for the most recent stores.
  auto sync_point = [&]() {
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
  };
  auto relu_row_to_panel = [&](__half* P, bool live) {
    uint32_t v0[32], v1[32];
    tmem_ld32_nowait(my_lane, v0);
    tmem_ld32_nowait(my_lane + 32, v1);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      v0[q] = live ? __float_as_uint(fmaxf(__uint_as_float(v0[q]), 0.f)) : 0u;
      v1[q] = live ? __float_as_uint(fmaxf(__uint_as_float(v1[q]), 0.f)) : 0u;
    }
    store_row_chunks(P, tid, v0, 32);
    store_row_chunks(P + (size_t)4 * kTile * 8, tid, v1, 32);
  };

  uint32_t phase = 0;
  int it = 0;
  bool prev_full = false;           // the previous tile's G5 (issued behind this tile's first barrier) uses the bulk path
  int64_t prev = 0;
  for (; cur < n_tiles; cur += gridDim.x, ++it) {
    const int set = it & 1;
    __half* PX = PX0 + (size_t)set * LY::kPx;
    const int64_t row0 = cur * kTile;
    const int64_t row = row0 + tid;
    const bool live = row < n;
    const bool full = row0 + kTile <= n;          // whole tile live: row-major outputs leave as bulk stores
    const int64_t nxt = cur + gridDim.x;
    // this tile's view direction (normalised later); the features were staged one tile ago
    float dx = 0.f, dy = 0.f, dz = 1.f;
    if (live) { dx = dirs[3 * row]; dy = dirs[3 * row + 1]; dz = dirs[3 * row + 2]; }
    FWD_TRACE(0);
    cp_async_wait_all();
    sync_point();
    FWD_TRACE(1);
    // ---- density trunk, layer 0
    issue(PX, WS0, 32, 64);
    if (tid == 0) {                               // G5 of the previous tile
      if (it > 0 && prev_full) {
        if (rgb_out != nullptr) bulk_s2g(rgb_out + prev * kTile * 16, OR, kTile * 16 * 2);
        if (c_total == 3) bulk_s2g(raws + prev * kTile * 3, OW, kTile * 3 * 4);
      }
      bulk_commit();
    }
    FWD_TRACE(2);
    if (nxt < n_tiles) stage_rows<32>(feat, nxt * kTile, n, PX0 + (size_t)(set ^ 1) * LY::kPx);      // prefetch under the chain
    cp_async_commit();
    FWD_TRACE(3);
    mbar_wait(mbar, phase); phase ^= 1u;
    FWD_TRACE(4);
    tc_fence_after();
    relu_row_to_panel(PH, live);
    sync_point();
    // ---- density trunk, layer 1 (+ the hidden panel leaves for the backward pass)
    issue(PH, WS1, 64, 16);
    if (tid == 0) { if (sig_acts != nullptr) bulk_s2g(sig_acts + cur * (64 * kTile), PH, 64 * kTile * 2); bulk_commit(); }      // G1
    FWD_TRACE(7);
    mbar_wait(mbar, phase); phase ^= 1u;
    FWD_TRACE(8);
    tc_fence_after();
    {
      uint32_t v[16];
      tmem_ld16(my_lane, v);
      // h in fp16 (what the tcnn module returns) is an output, the source of sigma, and columns 0..15 of the colour head's input
      uint32_t hp[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) hp[q] = pack_half2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
      const float inr = __frcp_rn(sqrtf(dx * dx + dy * dy + dz * dz));     // d / |d| to fp16: one reciprocal instead of three divisions
      const uint32_t ones = pack_half2(1.f, 1.f);
      uint32_t xr[16];
#pragma unroll
      for (int q = 0; q < 8; ++q) xr[q] = live ? hp[q] : 0u;
      xr[8] = live ? pack_half2(dx * inr, dy * inr) : 0u;
      xr[9] = live ? pack_half2(dz * inr, 1.f) : 0u;
#pragma unroll
      for (int q = 10; q < 16; ++q) xr[q] = live ? ones : 0u;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(PX1 + ((size_t)c * kTile + tid) * 8) = make_uint4(xr[4 * c], xr[4 * c + 1], xr[4 * c + 2], xr[4 * c + 3]);
      const float sg = expf(__low2float(*reinterpret_cast<const __half2*>(&hp[0])));
      if (full) {                                  // row-major staging (16-byte chunks rotated by the row to spread the banks)
        OS[tid] = sg;
#pragma unroll
        for (int c = 0; c < 2; ++c) *reinterpret_cast<uint4*>(OH + tid * 16 + c * 8) = make_uint4(hp[4 * c], hp[4 * c + 1], hp[4 * c + 2], hp[4 * c + 3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(OX + tid * 32 + c * 8) = make_uint4(xr[4 * c], xr[4 * c + 1], xr[4 * c + 2], xr[4 * c + 3]);
      } else if (live) {
        sigmas[row] = sg;
        if (h_out != nullptr) {
          *reinterpret_cast<uint4*>(h_out + row * 16) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
          *reinterpret_cast<uint4*>(h_out + row * 16 + 8) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
        }
        if (x_rgb != nullptr) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(x_rgb + row * 32 + c * 8) = make_uint4(xr[4 * c], xr[4 * c + 1], xr[4 * c + 2], xr[4 * c + 3]);
        }
      }
    }
    sync_point();
    FWD_TRACE(9);
    // ---- colour head, layer 0
    issue(PX1, WR0, 32, 64);
    if (tid == 0) {                               // G2
      if (full) {
        bulk_s2g(sigmas + row0, OS, kTile * 4);
        if (h_out != nullptr) bulk_s2g(h_out + row0 * 16, OH, kTile * 16 * 2);
        if (x_rgb != nullptr) bulk_s2g(x_rgb + row0 * 32, OX, kTile * 32 * 2);
      }
      bulk_commit();
    }
    mbar_wait(mbar, phase); phase ^= 1u;
    FWD_TRACE(10);
    tc_fence_after();
    relu_row_to_panel(PA1, live);
    sync_point();
    FWD_TRACE(11);
    // ---- colour head, layer 1
    issue(PA1, WR1, 64, 64);
    if (tid == 0) { if (rgb_acts != nullptr) bulk_s2g(rgb_acts + cur * (64 * kTile), PA1, 64 * kTile * 2); bulk_commit(); }      // G3
    mbar_wait(mbar, phase); phase ^= 1u;
    FWD_TRACE(12);
    tc_fence_after();
    relu_row_to_panel(PA2, live);
    sync_point();
    FWD_TRACE(13);
    // ---- colour head, layer 2 (sigmoid)
    issue(PA2, WR2, 64, 16);
    if (tid == 0) { if (rgb_acts != nullptr) bulk_s2g(rgb_acts + layer_stride + cur * (64 * kTile), PA2, 64 * kTile * 2); bulk_commit(); }   // G4
    mbar_wait(mbar, phase); phase ^= 1u;
    FWD_TRACE(14);
    tc_fence_after();
    {
      uint32_t v[16];
      tmem_ld16(my_lane, v);
      FWD_TRACE(5);
      uint32_t o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        // (__fdividef: one MUFU.RCP + FMUL; the IEEE division subroutine cost 130 clocks per output here, 2000 per tile, for
        //  bits that the fp16 rounding of the network output discards anyway)
        o[q] = pack_half2(__fdividef(1.0f, 1.0f + __expf(-__uint_as_float(v[2 * q]))), __fdividef(1.0f, 1.0f + __expf(-__uint_as_float(v[2 * q + 1]))));
      FWD_TRACE(6);
      // raws[:, 0:3] = fp32 of the fp16 network output
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&o[0]));
      const float r2 = __low2float(*reinterpret_cast<const __half2*>(&o[1]));
      if (full) {
        *reinterpret_cast<uint4*>(OR + tid * 16) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(OR + tid * 16 + 8) = make_uint4(o[4], o[5], o[6], o[7]);
        if (c_total == 3) { OW[3 * tid] = a.x; OW[3 * tid + 1] = a.y; OW[3 * tid + 2] = r2; }
      }
      if (live && (!full || c_total != 3)) { raws[row * c_total] = a.x; raws[row * c_total + 1] = a.y; raws[row * c_total + 2] = r2; }
      if (live && !full && rgb_out != nullptr) {
        *reinterpret_cast<uint4*>(rgb_out + row * 16) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(rgb_out + row * 16 + 8) = make_uint4(o[4], o[5], o[6], o[7]);
      }
      tc_fence_before();      // the next tile's first MMA overwrites the accumulator columns this thread has just read
    }
    prev_full = full; prev = cur;
    FWD_TRACE(15);
  }
  cp_async_wait_all();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    if (it > 0 && prev_full) {                      // G5 of the last tile
      if (rgb_out != nullptr) bulk_s2g(rgb_out + prev * kTile * 16, OR, kTile * 16 * 2);
      if (c_total == 3) bulk_s2g(raws + prev * kTile * 3, OW, kTile * 3 * 4);
    }
    bulk_commit();
    bulk_wait_read_all();
  }
  __syncthreads();
  if (wid == 0) tmem_dealloc<LY::kTmemCols>(tmem_d);
}

}  // namespace ncn

using namespace ncn;

#ifdef NCN_TC05_TRACE
extern "C" int ncn_debug_fwd_trace(long long* host_dst) {
  return cudaMemcpyFromSymbol(host_dst, ncn::g_fwd_trace, sizeof(long long) * 16 * 16) == cudaSuccess ? 0 : 1;
}
#endif

// returns NCN_E_UNSUPPORTED when the tcgen05 form cannot run (missing alignment): the caller takes the warp-MMA kernel
int ncn_field_mlp_fwd_tc05_try(const void* feat_f16, const float* dirs, const void* w_sigma_f16, const void* w_rgb_f16, int64_t n,
                               const int32_t* n_dev, float* sigmas, float* raws, int c_total, void* h_f16, void* sig_acts_f16,
                               void* x_rgb_f16, void* rgb_acts_f16, void* rgb_out_f16, cudaStream_t st) {
  if ((((uintptr_t)sig_acts_f16 | (uintptr_t)rgb_acts_f16) & 127) != 0) return NCN_E_UNSUPPORTED;      // bulk stores: 16-byte aligned at least
  if ((((uintptr_t)feat_f16 | (uintptr_t)h_f16 | (uintptr_t)x_rgb_f16 | (uintptr_t)rgb_out_f16) & 15) != 0) return NCN_E_UNSUPPORTED;
  using LY = FwdLayout;
  auto k = field_mlp_fwd_tc05_kernel;
  NCN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LY::kBytes));
  NCN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  const int64_t tiles = (n + kTile - 1) / kTile;
  int64_t grid = (int64_t)sm_count() * 2;
  if (grid > tiles) grid = tiles;
  if (grid < 1) grid = 1;
  NCN_CUDA(launch_pdl(k, dim3((unsigned)grid), dim3(kTcThreads), LY::kBytes, st, (const __half*)feat_f16, dirs, (const __half*)w_sigma_f16,
                      (const __half*)w_rgb_f16, n, n_dev, sigmas, raws, c_total, (__half*)h_f16, (__half*)sig_acts_f16, (__half*)x_rgb_f16,
                      (__half*)rgb_acts_f16, (__half*)rgb_out_f16));
  return NCN_OK;
}
