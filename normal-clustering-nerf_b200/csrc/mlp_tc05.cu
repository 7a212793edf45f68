// MLP backward with the weight gradients on the 5th-generation tensor cores (tcgen05) and their
// accumulators resident in tensor memory (TMEM) for the whole kernel.
//
// Why: dL/dW = dZ^T * A_prev contracts over the SAMPLE dimension.  The warp-MMA implementation (mlp.cu)
// has to round-trip every layer's dZ through HBM and re-read the activations in a second, split-K kernel
// per layer (2.4 KB/sample of traffic, 5 extra launches for the sigma + rgb nets).  Here one persistent CTA
// walks 128-sample tiles:
//   * the 8 warps run the dgrad chain in registers exactly like mlp.cu (mma.sync, fp32 accumulate) and drop
//     each layer's dZ tile and input-activation tile into shared memory in the canonical no-swizzle
//     "MN-major" core-matrix layout ([feature/8][sample][8 halfs]: a warp's fragment store is 128 contiguous
//     bytes, conflict free);
//   * ONE thread issues tcgen05.mma (kind::f16, M=64, N=in-width, K=16 per instruction, 8 per tile and layer)
//     whose A and B operands are those two tiles, both transposed for free by the MN-major descriptors, and
//     whose fp32 accumulator D = dW stays in TMEM across all tiles of the CTA (112 columns for the rgb net);
//     tcgen05.commit -> mbarrier releases the tiles while the warps already compute the next tile's dgrad;
//   * at the end four warps read TMEM (tcgen05.ld 32x32b) and add the CTA's dW to the global gradient.
// HBM traffic drops to the unavoidable reads (dL/dout, out, activations, x) and the dL/dx write.
#include "ncn_common.cuh"
#include "mma.cuh"

namespace ncn {

constexpr int kTcThreads = 256;
constexpr int kTile = 128;         // samples per CTA iteration (8 warps x 16 rows)
constexpr int kTcW = 64;
constexpr int kTcPad = 8;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NCN_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NCN_DONE_%=;\n"
      "bra NCN_WAIT_%=;\n"
      "NCN_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// shared-memory matrix descriptor: no swizzle, MN-major.  In a panel laid out [feature/8][sample][8 halfs]:
//   8 K-rows (samples) are 16 B apart, the next 8 samples are LBO = 128 B further (K direction),
//   the next 8 features are SBO = kTile*16 B further (M/N direction).          (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_desc_mn(const void* panel_at_k) {
  const uint64_t addr = (uint64_t)(smem_u32(panel_at_k) >> 4) & 0x3FFF;
  const uint64_t lbo = (128 >> 4), sbo = ((kTile * 16) >> 4);
  return addr | (lbo << 16) | (sbo << 32) | (1ull << 46);   // version = 1 (Blackwell), layout_type = 0 (no swizzle)
}
// instruction descriptor: D=f32, A=B=f16, both MN-major, M=64, N
__host__ __device__ constexpr uint32_t make_idesc(int N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
}

// ------------------------------------------------------------------ fragment helpers (same layouts as mlp.cu)
template <int K, int N>
__device__ __forceinline__ void tc_warp_layer(const uint32_t (*a)[4], const __half* __restrict__ W, float (*c)[4], int g, int t) {
#pragma unroll
  for (int nt = 0; nt < N / 8; ++nt) { c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f; }
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
#pragma unroll
    for (int nt = 0; nt < N / 8; ++nt) {
      const __half* wr = W + (nt * 8 + g) * (K + kTcPad) + kb * 16 + 2 * t;
      mma16816(c[nt], a[kb], *reinterpret_cast<const uint32_t*>(wr), *reinterpret_cast<const uint32_t*>(wr + 8));
    }
  }
}
__device__ __forceinline__ void tc_load_w_t(const __half* __restrict__ w, int rows, int cols, __half* __restrict__ s) {
  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
    const int r = i / cols, c = i % cols;
    s[c * (rows + kTcPad) + r] = w[i];
  }
}
// A-fragment-packed 16 x (16*KB) tile -> panel [feature/8][kTile][8]; `r` = first row of the warp inside the tile
template <int KB>
__device__ __forceinline__ void store_panel(__half* __restrict__ P, int r, const uint32_t (*a)[4], int g, int t) {
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
    __half* p0 = P + ((size_t)(2 * kb) * kTile + r + g) * 8 + 2 * t;
    __half* p1 = P + ((size_t)(2 * kb + 1) * kTile + r + g) * 8 + 2 * t;
    *reinterpret_cast<uint32_t*>(p0) = a[kb][0];
    *reinterpret_cast<uint32_t*>(p0 + 64) = a[kb][1];     // row g+8: 8 rows * 8 halfs further
    *reinterpret_cast<uint32_t*>(p1) = a[kb][2];
    *reinterpret_cast<uint32_t*>(p1 + 64) = a[kb][3];
  }
}
// global row-major (rows, W) fp16 -> panel, 16 B per cp.async; rows >= n are zero filled
template <int W>
__device__ __forceinline__ void stage_panel(const __half* __restrict__ src, int64_t row0, int64_t n, __half* __restrict__ P) {
  for (int i = threadIdx.x; i < kTile * (W / 8); i += blockDim.x) {
    const int r = i / (W / 8), mb = i % (W / 8);
    __half* dst = P + ((size_t)mb * kTile + r) * 8;
    if (row0 + r < n) cp_async16(dst, src + (row0 + r) * W + mb * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
}
// ReLU gate of the accumulators (C layout, 16 x 64) from an activation panel
__device__ __forceinline__ void relu_mask_panel(float (*c)[4], const __half* __restrict__ P, int r, int g, int t) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __half2 a0 = *reinterpret_cast<const __half2*>(P + ((size_t)j * kTile + r + g) * 8 + 2 * t);
    const __half2 a1 = *reinterpret_cast<const __half2*>(P + ((size_t)j * kTile + r + g + 8) * 8 + 2 * t);
    if (!(__low2float(a0) > 0.f)) c[j][0] = 0.f;
    if (!(__high2float(a0) > 0.f)) c[j][1] = 0.f;
    if (!(__low2float(a1) > 0.f)) c[j][2] = 0.f;
    if (!(__high2float(a1) > 0.f)) c[j][3] = 0.f;
  }
}
__device__ __forceinline__ void tc_c_to_a64(const float (*c)[4], uint32_t (*a)[4]) {
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    a[kb][0] = pack_half2(c[2 * kb][0], c[2 * kb][1]);
    a[kb][1] = pack_half2(c[2 * kb][2], c[2 * kb][3]);
    a[kb][2] = pack_half2(c[2 * kb + 1][0], c[2 * kb + 1][1]);
    a[kb][3] = pack_half2(c[2 * kb + 1][2], c[2 * kb + 1][3]);
  }
}

constexpr int tmem_cols_pow2(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

template <int IN, int OUT, int NH>
struct TcLayout {
  // shared memory (halfs unless noted)
  static constexpr int kWlT = 64 * (OUT + kTcPad);
  static constexpr int kWhT = (NH - 1) * 64 * (64 + kTcPad);
  static constexpr int kW0T = IN * (64 + kTcPad);
  static constexpr int kWeights = (kWlT + kWhT + kW0T + 7) / 8 * 8;
  static constexpr int kPdzLast = OUT * kTile, kPdzH = NH * 64 * kTile, kPx = IN * kTile, kPact = NH * 64 * kTile;
  static constexpr size_t kBytes = (size_t)(kWeights + kPdzLast + kPdzH + kPx + kPact) * 2 + 32;
  static constexpr int kTmemColsUsed = IN + 64 * (NH - 1) + OUT;
  static constexpr int kTmemCols = tmem_cols_pow2(kTmemColsUsed);
};

template <int IN, int OUT, int NH>
__global__ void __launch_bounds__(kTcThreads, 1)
mlp_bwd_tc05_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ out,
                    const __half* __restrict__ acts, const __half* __restrict__ dout, int64_t n_cap,
                    const int32_t* __restrict__ n_dev, int out_act, float grad_scale, float* __restrict__ grad_w,
                    __half* __restrict__ dx, int* __restrict__ tile_counter) {
  using LY = TcLayout<IN, OUT, NH>;
  __shared__ int s_next_tile;
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __half* WlT = reinterpret_cast<__half*>(tc_smem);
  __half* WhT = WlT + LY::kWlT;
  __half* W0T = WhT + LY::kWhT;
  __half* P_dz_last = WlT + LY::kWeights;
  __half* P_dz_h = P_dz_last + LY::kPdzLast;
  __half* P_x = P_dz_h + LY::kPdzH;
  __half* P_act = P_x + LY::kPx;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(P_act + LY::kPact);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  tc_load_w_t(w + 64 * IN + (NH - 1) * 64 * 64, OUT, 64, WlT);
  for (int i = 0; i < NH - 1; ++i) tc_load_w_t(w + 64 * IN + i * 64 * 64, 64, 64, WhT + i * 64 * (64 + kTcPad));
  if (dx) tc_load_w_t(w, 64, IN, W0T);
  if (tid == 0) { mbar_init(mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (wid == 0) tmem_alloc<LY::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int64_t n_tiles = (n + kTile - 1) / kTile;
  int it = 0;
  // dynamic tile scheduler: CTAs draw 128-sample tiles from a global counter (even finish times; the static split
  // left 1/7 of the CTAs with an extra tile)
  int64_t tile = blockIdx.x;
  for (; tile < n_tiles; ++it) {
    const int64_t row0 = tile * kTile;
    if (tid == 0) s_next_tile = atomicAdd(tile_counter, 1) + (int)gridDim.x;
    if (it > 0) mbar_wait(mbar, (uint32_t)((it - 1) & 1));     // the previous tile's MMAs have consumed the panels
    // (1) stage the input / hidden activations of this tile
    stage_panel<IN>(x, row0, n, P_x);
    for (int i = 0; i < NH; ++i) stage_panel<64>(acts + (int64_t)i * n_cap * 64, row0, n, P_act + (size_t)i * 64 * kTile);
    // the warp's dL/dout (and out) fragments are fetched while the cp.async panels are still in flight
    const int r = wid * 16;                       // first row of this warp inside the tile
    const int64_t r0 = row0 + r + g, r1 = r0 + 8;
    uint32_t dz[OUT / 16][4], ov[OUT / 16][4];
#pragma unroll
    for (int kb = 0; kb < OUT / 16; ++kb) {
      const int col = kb * 16 + 2 * t;
      dz[kb][0] = r0 < n ? *reinterpret_cast<const uint32_t*>(dout + r0 * OUT + col) : 0u;
      dz[kb][1] = r1 < n ? *reinterpret_cast<const uint32_t*>(dout + r1 * OUT + col) : 0u;
      dz[kb][2] = r0 < n ? *reinterpret_cast<const uint32_t*>(dout + r0 * OUT + col + 8) : 0u;
      dz[kb][3] = r1 < n ? *reinterpret_cast<const uint32_t*>(dout + r1 * OUT + col + 8) : 0u;
      if (out_act == NCN_ACT_SIGMOID || out_act == NCN_ACT_EXP) {
        ov[kb][0] = r0 < n ? *reinterpret_cast<const uint32_t*>(out + r0 * OUT + col) : 0u;
        ov[kb][1] = r1 < n ? *reinterpret_cast<const uint32_t*>(out + r1 * OUT + col) : 0u;
        ov[kb][2] = r0 < n ? *reinterpret_cast<const uint32_t*>(out + r0 * OUT + col + 8) : 0u;
        ov[kb][3] = r1 < n ? *reinterpret_cast<const uint32_t*>(out + r1 * OUT + col + 8) : 0u;
      }
    }
    cp_async_wait_all();
    __syncthreads();
    const int next_tile = s_next_tile;       // read between this barrier and the next one; rewritten only after the latter
    // (2) dgrad chain in registers; every layer's dL/dz goes to its panel
    {
      if (out_act == NCN_ACT_SIGMOID || out_act == NCN_ACT_EXP) {
#pragma unroll
        for (int kb = 0; kb < OUT / 16; ++kb)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t ou = ov[kb][q];
            const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&dz[kb][q]));
            const float2 y = __half22float2(*reinterpret_cast<const __half2*>(&ou));
            if (out_act == NCN_ACT_SIGMOID) dz[kb][q] = pack_half2(d.x * y.x * (1.f - y.x), d.y * y.y * (1.f - y.y));
            else dz[kb][q] = pack_half2(d.x * y.x, d.y * y.y);
          }
      }
      store_panel<OUT / 16>(P_dz_last, r, dz, g, t);
      float c[8][4];
      tc_warp_layer<OUT, 64>(dz, WlT, c, g, t);
      uint32_t dh[4][4];
#pragma unroll
      for (int i = NH - 1; i >= 0; --i) {
        relu_mask_panel(c, P_act + (size_t)i * 64 * kTile, r, g, t);
        tc_c_to_a64(c, dh);
        store_panel<4>(P_dz_h + (size_t)i * 64 * kTile, r, dh, g, t);
        if (i > 0) tc_warp_layer<64, 64>(dh, WhT + (i - 1) * 64 * (64 + kTcPad), c, g, t);
      }
      if (dx) {
        float cx[IN / 8][4];
        tc_warp_layer<64, IN>(dh, W0T, cx, g, t);
#pragma unroll
        for (int j = 0; j < IN / 8; ++j) {
          const int col = j * 8 + 2 * t;
          if (r0 < n) *reinterpret_cast<uint32_t*>(dx + r0 * IN + col) = pack_half2(cx[j][0], cx[j][1]);
          if (r1 < n) *reinterpret_cast<uint32_t*>(dx + r1 * IN + col) = pack_half2(cx[j][2], cx[j][3]);
        }
      }
    }
    // (3) publish the panels to the tensor-core (async) proxy and issue the weight-gradient MMAs
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t acc0 = it > 0 ? 1u : 0u;
#pragma unroll
      for (int ks = 0; ks < kTile / 16; ++ks) {
        const uint32_t acc = (ks > 0) ? 1u : acc0;
        const size_t koff = (size_t)ks * 16 * 8;          // 16 samples * 8 halfs
        // layer 0: dW0[out][in] += dz_h[0]^T x
        tc_mma_f16(tmem + 0, make_desc_mn(P_dz_h + koff), make_desc_mn(P_x + koff), make_idesc(IN), acc);
        // hidden layers i = 1..NH-1: dWi[out][in] += dz_h[i]^T act[i-1]
#pragma unroll
        for (int i = 1; i < NH; ++i)
          tc_mma_f16(tmem + IN + 64 * (i - 1), make_desc_mn(P_dz_h + (size_t)i * 64 * kTile + koff),
                     make_desc_mn(P_act + (size_t)(i - 1) * 64 * kTile + koff), make_idesc(64), acc);
        // last layer, transposed: dWl^T[in][out] += act[NH-1]^T dz_last
        tc_mma_f16(tmem + IN + 64 * (NH - 1), make_desc_mn(P_act + (size_t)(NH - 1) * 64 * kTile + koff),
                   make_desc_mn(P_dz_last + koff), make_idesc(OUT), acc);
      }
      tc_commit(mbar);
    }
    tile = next_tile;
  }
  // (4) epilogue: TMEM -> registers -> global gradient (+=)
  if (it > 0) {
    mbar_wait(mbar, (uint32_t)((it - 1) & 1));
    tc_fence_after();
    if (wid < 4) {
      // M = 64 accumulators live in the lower 16 lanes of each 32-lane TMEM sub-partition: row m = 16*wid + lane
      const int m = 16 * wid + lane;
      const uint32_t lane_base = tmem + ((uint32_t)(32 * wid) << 16);
      uint32_t v[16];
#pragma unroll
      for (int c0 = 0; c0 < IN; c0 += 16) {
        tmem_ld16(lane_base + c0, v);
        if (lane < 16)
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(grad_w + m * IN + c0 + j, __uint_as_float(v[j]) * grad_scale);
      }
#pragma unroll
      for (int i = 1; i < NH; ++i)
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          tmem_ld16(lane_base + IN + 64 * (i - 1) + c0, v);
          if (lane < 16)
#pragma unroll
            for (int j = 0; j < 16; ++j)
              atomicAdd(grad_w + 64 * IN + (i - 1) * 64 * 64 + m * 64 + c0 + j, __uint_as_float(v[j]) * grad_scale);
        }
#pragma unroll
      for (int c0 = 0; c0 < OUT; c0 += 16) {
        tmem_ld16(lane_base + IN + 64 * (NH - 1) + c0, v);
        if (lane < 16)
#pragma unroll
          for (int j = 0; j < 16; ++j)   // D[m = in][n = out] -> W_last[out][in]
            atomicAdd(grad_w + 64 * IN + (NH - 1) * 64 * 64 + (c0 + j) * 64 + m, __uint_as_float(v[j]) * grad_scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<LY::kTmemCols>(tmem);
}

// =====================================================================================================================
// v2: EVERY GEMM of the backward pass on tcgen05 (dgrad as well as wgrad), 128-row tiles, one thread per row.
//
//   dgrad layer:  D[128 x in] = dZ[128 x out] * W[out x in]       A = dZ panel read K-MAJOR  (M = sample), B = W panel MN-major
//   wgrad layer:  dW[out x in] += dZ^T[out x 128] * A_prev[128 x in]   A = the SAME dZ panel read MN-MAJOR (M = feature)
// The panel layout [feature/8][sample][8 halfs] is the canonical no-swizzle layout for both readings, so a dZ tile is
// written once (by the epilogue of the layer above) and consumed by two different MMAs.  The per-tile chain is
//   stage(x, acts; dL/dout * act') -> MMA -> [tcgen05.ld row, ReLU gate, fp16, st.shared row] -> MMA -> ... -> dx rows
// with one thread issuing all MMAs and 128 threads doing ~25 instructions per 16 accumulator columns in the epilogues
// (the warp-MMA v1 kernel spends ~60 instructions per row on fragment shuffling; this one ~10).
constexpr int kV2Threads = 128;

// K-major A operand over a panel [feature/8][kTile][8]: rows (M = samples) 16 B apart, next 8 rows SBO = 128 B,
// next 8 K-elements (next feature group) LBO = kTile*16 B
__device__ __forceinline__ uint64_t make_desc_k(const void* panel_at_k) {
  const uint64_t addr = (uint64_t)(smem_u32(panel_at_k) >> 4) & 0x3FFF;
  const uint64_t lbo = ((kTile * 16) >> 4), sbo = (128 >> 4);
  return addr | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// MN-major B operand over a weight panel [n/8][K][8]: K rows 16 B apart, next 8 K rows LBO = 128 B, next 8 n: SBO = K*16 B
__device__ __forceinline__ uint64_t make_desc_w(const void* panel_at_k, int K) {
  const uint64_t addr = (uint64_t)(smem_u32(panel_at_k) >> 4) & 0x3FFF;
  const uint64_t lbo = (128 >> 4), sbo = (uint64_t)((K * 16) >> 4);
  return addr | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// D=f32, A=f16 K-major, B=f16 MN-major, M=128, N
__host__ __device__ constexpr uint32_t make_idesc_dgrad(int N) {
  return (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// weight matrix W (K=out rows, N=in cols, row-major) -> MN-major panel [n/8][K][8]
__device__ __forceinline__ int perm_col(int c) { return c < 16 ? c + 3 : (c < 19 ? c - 16 : c); }   // fused-forward order -> tcnn order
__device__ __forceinline__ void load_w_panel(const __half* __restrict__ w, int K, int N, __half* __restrict__ P, bool perm = false) {
  for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
    const int k = i / N, n = i % N;                  // n = kernel-internal column
    P[((size_t)(n >> 3) * K + k) * 8 + (n & 7)] = w[k * N + (perm ? perm_col(n) : n)];
  }
}

template <int IN, int OUT, int NH>
struct TcLayout2 {
  static constexpr int kWBl = 64 * OUT, kWBh = (NH - 1) * 64 * 64, kWB0 = IN * 64;
  static constexpr int kWeights = kWBl + kWBh + kWB0;
  static constexpr int kPdzLast = OUT * kTile, kPdzH = NH * 64 * kTile, kPx = IN * kTile, kPact = NH * 64 * kTile;
  static constexpr size_t kBytes = (size_t)(kWeights + kPdzLast + kPdzH + kPx + kPact) * 2 + 64;
  static constexpr int kDcols = 64;                                   // dgrad accumulator tile (reused layer after layer)
  static constexpr int kWgradCols = IN + 64 * (NH - 1) + OUT;
  static constexpr int kTmemCols = tmem_cols_pow2(kDcols + kWgradCols);
};

template <int IN, int OUT, int NH>
__global__ void __launch_bounds__(kV2Threads, 2)
mlp_bwd_tc05_v2_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ out,
                       const __half* __restrict__ acts, const __half* __restrict__ dout, int64_t n_cap,
                       const int32_t* __restrict__ n_dev, int out_act, float grad_scale, float* __restrict__ grad_w,
                       __half* __restrict__ dx, int* __restrict__ tile_counter, ncn_mlp_bwd_src src) {
  using LY = TcLayout2<IN, OUT, NH>;
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __half* WBl = reinterpret_cast<__half*>(tc_smem);          // last layer: K = OUT, N = 64
  __half* WBh = WBl + LY::kWBl;                              // hidden layers: K = 64, N = 64
  __half* WB0 = WBh + LY::kWBh;                              // first layer: K = 64, N = IN
  __half* P_dz_last = WB0 + LY::kWB0;
  __half* P_dz_h = P_dz_last + LY::kPdzLast;
  __half* P_x = P_dz_h + LY::kPdzH;
  __half* P_act = P_x + LY::kPx;
  uint64_t* mbar_w = reinterpret_cast<uint64_t*>(P_act + LY::kPact);
  uint64_t* mbar_d = mbar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar_d + 1);
  __shared__ int s_next_tile;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  load_w_panel(w + 64 * IN + (NH - 1) * 64 * 64, OUT, 64, WBl);
  for (int i = 0; i < NH - 1; ++i) load_w_panel(w + 64 * IN + i * 64 * 64, 64, 64, WBh + i * 64 * 64);
  if (dx) load_w_panel(w, 64, IN, WB0, (src.perm & 1) != 0);
  if (tid == 0) { mbar_init(mbar_w, 1); mbar_init(mbar_d, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (wid == 0) tmem_alloc<LY::kTmemCols>(tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_d = tmem;                              // dgrad tile: columns [0, 64)
  const uint32_t tmem_w = tmem + LY::kDcols;                 // weight-gradient accumulators
  const uint32_t my_lane = tmem_d + ((uint32_t)(32 * wid) << 16);   // this warp's TMEM sub-partition (row = tid)

  const int64_t n_tiles = (n + kTile - 1) / kTile;
  int it = 0;
  uint32_t d_phase = 0;
  int64_t tile = blockIdx.x;
  for (; tile < n_tiles; ++it) {
    const int64_t row0 = tile * kTile;
    const int64_t row = row0 + tid;
    if (tid == 0) s_next_tile = atomicAdd(tile_counter, 1) + (int)gridDim.x;
    if (it > 0) mbar_wait(mbar_w, (uint32_t)((it - 1) & 1));        // the previous tile's wgrad MMAs released the panels
    // (1) stage x / activations; this thread's dL/dout row (times the output activation derivative) -> dz_last panel
    stage_panel<IN>(x, row0, n, P_x);
    for (int i = 0; i < NH; ++i) stage_panel<64>(acts + (int64_t)i * n_cap * 64, row0, n, P_act + (size_t)i * 64 * kTile);
    // dL/dout row of this thread, from one of three sources (src.mode):
    //   0: the (N,OUT) fp16 matrix `dout`
    //   1: colour head - columns [c_off, c_off+n_ch) of the fp32 dL/draws matrix (ncn_field_head_dout fused in)
    //   2: density trunk - dL/dh = dL/dx_rgb[:, 3:19] + e0 * dL/dsigma * exp(clamp(h0,-15,15))  (ncn_field_bwd_h fused in)
    uint32_t drow[OUT / 2];
#pragma unroll
    for (int q = 0; q < OUT / 2; ++q) drow[q] = 0u;
    if (row < n) {
      if (src.mode == 0) {
#pragma unroll
        for (int j = 0; j < OUT / 8; ++j) {
          const uint4 d = *reinterpret_cast<const uint4*>(dout + row * OUT + j * 8);
          drow[4 * j] = d.x; drow[4 * j + 1] = d.y; drow[4 * j + 2] = d.z; drow[4 * j + 3] = d.w;
        }
      } else if (src.mode == 1) {
        const float* rr = src.d_raws + row * src.c_total + src.c_off;
#pragma unroll
        for (int q = 0; q < OUT / 2; ++q) {
          const float a = 2 * q < src.n_ch ? rr[2 * q] * src.scale : 0.f, b = 2 * q + 1 < src.n_ch ? rr[2 * q + 1] * src.scale : 0.f;
          drow[q] = pack_half2(a, b);
        }
      } else {
        const uint4* xr = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(src.dx_rgb) + row * 32);
        const uint4 w0 = xr[0], w1 = xr[1];
        if (src.perm & 2) {                                       // fused-forward order: dL/dh = dx_rgb[:, 0:16]
          drow[0] = w0.x; drow[1] = w0.y; drow[2] = w0.z; drow[3] = w0.w; drow[4] = w1.x; drow[5] = w1.y; drow[6] = w1.z; drow[7] = w1.w;
        } else {
          const uint4 w2 = xr[2];
          const uint32_t wv[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) drow[q] = __funnelshift_r(wv[q + 1], wv[q + 2], 16);   // halfs 3+2q, 4+2q
        }
        const float h0 = __half2float(reinterpret_cast<const __half*>(src.h)[row * 16]);
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&drow[0]));
        drow[0] = pack_half2(f.x + src.d_sigmas[row] * __expf(fminf(fmaxf(h0, -15.f), 15.f)) * src.scale, f.y);
      }
      if (out_act == NCN_ACT_SIGMOID || out_act == NCN_ACT_EXP) {
#pragma unroll
        for (int j = 0; j < OUT / 8; ++j) {
          const uint4 o = *reinterpret_cast<const uint4*>(out + row * OUT + j * 8);
          const uint32_t op[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 dv = __half22float2(*reinterpret_cast<const __half2*>(&drow[4 * j + q]));
            const float2 y = __half22float2(*reinterpret_cast<const __half2*>(&op[q]));
            drow[4 * j + q] = out_act == NCN_ACT_SIGMOID ? pack_half2(dv.x * y.x * (1.f - y.x), dv.y * y.y * (1.f - y.y))
                                                         : pack_half2(dv.x * y.x, dv.y * y.y);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < OUT / 8; ++j) {
      const uint4 d = make_uint4(drow[4 * j], drow[4 * j + 1], drow[4 * j + 2], drow[4 * j + 3]);
      *reinterpret_cast<uint4*>(P_dz_last + ((size_t)j * kTile + tid) * 8) = d;
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    const int next_tile = s_next_tile;
    // (2) dgrad chain: MMA -> epilogue (gate, fp16, panel row) -> MMA ...
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < OUT / 16; ++ks)
        tc_mma_f16(tmem_d, make_desc_k(P_dz_last + (size_t)ks * 2 * kTile * 8), make_desc_w(WBl + (size_t)ks * 16 * 8, OUT),
                   make_idesc_dgrad(64), ks > 0 ? 1u : 0u);
      tc_commit(mbar_d);
    }
#pragma unroll
    for (int i = NH - 1; i >= 0; --i) {
      mbar_wait(mbar_d, d_phase); d_phase ^= 1u;
      tc_fence_after();
      // dL/dh_i row: 64 fp32 accumulators in 4 chunks of 16 columns
      const __half* Pa = P_act + (size_t)i * 64 * kTile;
      __half* Pz = P_dz_h + (size_t)i * 64 * kTile;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(my_lane + c0, v);
        const uint4 a0 = *reinterpret_cast<const uint4*>(Pa + ((size_t)(c0 / 8) * kTile + tid) * 8);
        const uint4 a1 = *reinterpret_cast<const uint4*>(Pa + ((size_t)(c0 / 8 + 1) * kTile + tid) * 8);
        const uint32_t am[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        uint32_t o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&am[q]));
          const float lo = a.x > 0.f ? __uint_as_float(v[2 * q]) : 0.f, hi = a.y > 0.f ? __uint_as_float(v[2 * q + 1]) : 0.f;
          o[q] = pack_half2(lo, hi);
        }
        *reinterpret_cast<uint4*>(Pz + ((size_t)(c0 / 8) * kTile + tid) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(Pz + ((size_t)(c0 / 8 + 1) * kTile + tid) * 8) = make_uint4(o[4], o[5], o[6], o[7]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        if (i > 0) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_f16(tmem_d, make_desc_k(Pz + (size_t)ks * 2 * kTile * 8), make_desc_w(WBh + (size_t)(i - 1) * 64 * 64 + (size_t)ks * 16 * 8, 64),
                       make_idesc_dgrad(64), ks > 0 ? 1u : 0u);
          tc_commit(mbar_d);
        } else {
          if (dx) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_f16(tmem_d, make_desc_k(Pz + (size_t)ks * 2 * kTile * 8), make_desc_w(WB0 + (size_t)ks * 16 * 8, 64),
                         make_idesc_dgrad(IN), ks > 0 ? 1u : 0u);
            tc_commit(mbar_d);
          }
          // (3) all dZ panels are final: weight-gradient MMAs, accumulators stay in TMEM
          const uint32_t acc0 = it > 0 ? 1u : 0u;
#pragma unroll
          for (int ks = 0; ks < kTile / 16; ++ks) {
            const uint32_t acc = (ks > 0) ? 1u : acc0;
            const size_t koff = (size_t)ks * 16 * 8;
            tc_mma_f16(tmem_w + 0, make_desc_mn(P_dz_h + koff), make_desc_mn(P_x + koff), make_idesc(IN), acc);
#pragma unroll
            for (int q = 1; q < NH; ++q)
              tc_mma_f16(tmem_w + IN + 64 * (q - 1), make_desc_mn(P_dz_h + (size_t)q * 64 * kTile + koff),
                         make_desc_mn(P_act + (size_t)(q - 1) * 64 * kTile + koff), make_idesc(64), acc);
            tc_mma_f16(tmem_w + IN + 64 * (NH - 1), make_desc_mn(P_act + (size_t)(NH - 1) * 64 * kTile + koff),
                       make_desc_mn(P_dz_last + koff), make_idesc(OUT), acc);
          }
          tc_commit(mbar_w);
        }
      }
    }
    if (dx) {
      mbar_wait(mbar_d, d_phase); d_phase ^= 1u;
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < IN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(my_lane + c0, v);
        if (row < n) {
          uint32_t o[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) o[q] = pack_half2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
          *reinterpret_cast<uint4*>(dx + row * IN + c0) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(dx + row * IN + c0 + 8) = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      tc_fence_before();     // the next tile's first MMA overwrites the dgrad tile this thread just read
    }
    tile = next_tile;
  }
  // (4) epilogue: weight gradients TMEM -> global (+=)
  if (it > 0) {
    mbar_wait(mbar_w, (uint32_t)((it - 1) & 1));
    tc_fence_after();
    const int m = 16 * wid + lane;        // M = 64 accumulators: rows 16*wid + lane, lanes 0..15 of each sub-partition
    const uint32_t lane_base = tmem_w + ((uint32_t)(32 * wid) << 16);
    uint32_t v[16];
#pragma unroll
    for (int c0 = 0; c0 < IN; c0 += 16) {
      tmem_ld16(lane_base + c0, v);
      if (lane < 16)
#pragma unroll
        for (int j = 0; j < 16; ++j)
          atomicAdd(grad_w + m * IN + ((src.perm & 1) ? perm_col(c0 + j) : c0 + j), __uint_as_float(v[j]) * grad_scale);
    }
#pragma unroll
    for (int i = 1; i < NH; ++i)
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        tmem_ld16(lane_base + IN + 64 * (i - 1) + c0, v);
        if (lane < 16)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            atomicAdd(grad_w + 64 * IN + (i - 1) * 64 * 64 + m * 64 + c0 + j, __uint_as_float(v[j]) * grad_scale);
      }
#pragma unroll
    for (int c0 = 0; c0 < OUT; c0 += 16) {
      tmem_ld16(lane_base + IN + 64 * (NH - 1) + c0, v);
      if (lane < 16)
#pragma unroll
        for (int j = 0; j < 16; ++j)
          atomicAdd(grad_w + 64 * IN + (NH - 1) * 64 * 64 + (c0 + j) * 64 + m, __uint_as_float(v[j]) * grad_scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<LY::kTmemCols>(tmem);
}

}  // namespace ncn

using namespace ncn;

template <int IN, int OUT, int NH>
static int launch_tc05(const void* x, const void* w, const void* out, const void* acts, const void* dout, int64_t n,
                       const int32_t* n_dev, int out_act, float grad_scale, float* grad_w, void* dx, int* tile_counter,
                       cudaStream_t st) {
  using LY = TcLayout<IN, OUT, NH>;
  NCN_CUDA(cudaMemsetAsync(tile_counter, 0, sizeof(int), st));
  auto k = mlp_bwd_tc05_kernel<IN, OUT, NH>;
  NCN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LY::kBytes));
  const int64_t tiles = (n + kTile - 1) / kTile;
  int64_t grid = (int64_t)sm_count() * 2;
  if (grid > tiles) grid = tiles;
  k<<<(int)grid, kTcThreads, LY::kBytes, st>>>((const __half*)x, (const __half*)w, (const __half*)out, (const __half*)acts,
                                               (const __half*)dout, n, n_dev, out_act, grad_scale, grad_w, (__half*)dx, tile_counter);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

template <int IN, int OUT, int NH>
static int launch_tc05_v2(const void* x, const void* w, const void* out, const void* acts, const void* dout, int64_t n,
                          const int32_t* n_dev, int out_act, float grad_scale, float* grad_w, void* dx, int* tile_counter,
                          const ncn_mlp_bwd_src& src, cudaStream_t st) {
  using LY = TcLayout2<IN, OUT, NH>;
  auto k = mlp_bwd_tc05_v2_kernel<IN, OUT, NH>;
  NCN_CUDA(cudaMemsetAsync(tile_counter, 0, sizeof(int), st));
  NCN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LY::kBytes));
  const int64_t tiles = (n + kTile - 1) / kTile;
  // persistent grid = every CTA slot the SMs offer (shared memory bound: 2 for the colour head, 4 for the density trunk);
  // the dynamic tile scheduler keeps them evenly loaded
  static int occ = 0;
  if (occ == 0) {
    int o = 0;
    NCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k, kV2Threads, LY::kBytes));
    occ = o < 1 ? 1 : (o > 512 / LY::kTmemCols ? 512 / LY::kTmemCols : o);      // TMEM: 512 columns per SM
  }
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > tiles) grid = tiles;
  k<<<(int)grid, kV2Threads, LY::kBytes, st>>>((const __half*)x, (const __half*)w, (const __half*)out, (const __half*)acts,
                                               (const __half*)dout, n, n_dev, out_act, grad_scale, grad_w, (__half*)dx, tile_counter, src);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

// returns NCN_E_UNSUPPORTED when the configuration has no tcgen05 instantiation (the caller falls back to mlp.cu)
int ncn_mlp_bwd_tc05_try(int in_pad, int out_pad, int n_hidden, const void* x, const void* w, const void* out, const void* acts,
                         const void* dout, int64_t n, const int32_t* n_dev, int out_act, float grad_scale, float* grad_w,
                         void* dx, int* tile_counter, int impl, const ncn_mlp_bwd_src* src_in, cudaStream_t st) {
  if (!grad_w) return NCN_E_UNSUPPORTED;
  ncn_mlp_bwd_src src;
  if (src_in) src = *src_in; else { src = ncn_mlp_bwd_src(); src.mode = 0; }
  if (src.mode != 0 && (impl != 2 || out_pad != 16)) return NCN_E_UNSUPPORTED;
  if (impl == 2) {
#define NCN_TC2(I, O, H) if (in_pad == I && out_pad == O && n_hidden == H) \
    return launch_tc05_v2<I, O, H>(x, w, out, acts, dout, n, n_dev, out_act, grad_scale, grad_w, dx, tile_counter, src, st);
    NCN_TC2(32, 16, 1) NCN_TC2(32, 16, 2) NCN_TC2(16, 16, 2) NCN_TC2(16, 16, 1) NCN_TC2(16, 48, 2) NCN_TC2(16, 32, 2)
#undef NCN_TC2
  }
#define NCN_TC(I, O, H) if (in_pad == I && out_pad == O && n_hidden == H) \
    return launch_tc05<I, O, H>(x, w, out, acts, dout, n, n_dev, out_act, grad_scale, grad_w, dx, tile_counter, st);
  NCN_TC(32, 16, 1) NCN_TC(32, 16, 2) NCN_TC(16, 16, 2) NCN_TC(16, 16, 1) NCN_TC(16, 48, 2) NCN_TC(16, 32, 2)
#undef NCN_TC
  return NCN_E_UNSUPPORTED;
}
