// MLP backward with EVERY GEMM (dgrad and wgrad) on the 5th-generation tensor cores (tcgen05), accumulators in tensor
// memory (TMEM), operands fed by the bulk-copy engine (TMA).
//
//   dgrad layer:  D[128 x in] = dZ[128 x out] * W[out x in]            A = dZ panel read K-MAJOR  (M = sample), B = W panel
//   wgrad layer:  dW[out x in] += dZ^T[out x 128] * A_prev[128 x in]   A = the SAME dZ panel read MN-MAJOR (M = feature)
//
// One persistent CTA walks 128-sample tiles handed out by an atomic counter.  A panel is [feature/8][sample][8 halfs] -
// the canonical no-swizzle core-matrix layout for both readings - and the forward kernels already store the hidden
// activations in global memory as one such 16 KB panel per tile and layer (ncn_common.cuh act_offset), so a tile's
// activations arrive with ONE cp.async.bulk per layer, completing on an mbarrier, issued a whole tile ahead.
// Per tile (colour head: 3 weight matrices):
//   * thread r assembles its sample's dL/dout row (see ncn_mlp_bwd_src) in registers -> dz_last panel;
//   * one thread issues  D = dz_last * W_last  and  dW_last^T += act^T * dz_last,  tcgen05.commit -> mbarrier;
//   * 128 threads: tcgen05.ld their accumulator row, gate it with the activation row, convert to fp16 and write it IN
//     PLACE over the activation panel (dW_{i+1} has consumed act_i by then: the same commit covers both MMAs);
//   * next layer's dgrad + wgrad MMAs ...  -> dL/dx rows to global memory.
// dW accumulates in TMEM over all tiles of the CTA (112 columns for the colour head) and is added to the global gradient
// once, with red.global.add.v4.f32.  HBM traffic is the unavoidable reads (x, activations, out, dL/dout sources) and the
// dL/dx write; shared memory holds two panel sets so the loads of tile t+1 overlap the MMA -> epilogue chain of tile t.
//
// History (profiles/r1_ncu_mlp_bwd.md): a warp-MMA dgrad + TMEM wgrad hybrid ran 114 us for the colour head at 269 k
// samples; the first all-tcgen05 kernel 110 us (30 % of its stall samples in a scalar-RED epilogue, 24 % in cp.async
// staging loops, a synchronous scheduler atomic on the MMA-issuing thread); in-place panels + prefetch + 2 CTAs/SM 70 us;
// the remaining stalls were LSU back-pressure from 20 cp.async per thread and tile, which the bulk copies remove.
#include <cstdio>
#include "ncn_common.cuh"
#include "mma.cuh"
#include "tc05.cuh"

namespace ncn {

template <int IN, int OUT, int NH>
struct TcLayout {
  static constexpr int kWBl = 64 * OUT, kWBh = (NH - 1) * 64 * 64, kWB0 = IN * 64;
  static constexpr int kWeights = kWBl + kWBh + kWB0;
  static constexpr int kPdzl = OUT * kTile, kPact = NH * 64 * kTile, kPx = IN * kTile;
  static constexpr int kSet = kPdzl + kPact + kPx;                     // halfs per panel set
  static constexpr size_t kBytes = (size_t)(kWeights + 2 * kSet) * 2 + 64;
  static constexpr int kDcols = 64;                                    // dgrad accumulator tile (reused layer after layer)
  static constexpr int kWgradCols = IN + 64 * (NH - 1) + OUT;
  static constexpr int kTmemCols = tmem_cols_pow2(kDcols + kWgradCols);
  static constexpr uint32_t kActBytes = 64 * kTile * 2;                // one layer's activation panel
};

// the not-yet-combined dL/dout row of one sample (see ncn_mlp_bwd_src): fetched one tile ahead
template <int OUT>
struct DoutRaw {
  uint4 a[3];
  uint4 o[OUT / 8];
  float f[3];
};

template <int OUT>
__device__ __forceinline__ void dout_fetch(DoutRaw<OUT>& R, int64_t row, int64_t n, const __half* __restrict__ dout,
                                           const __half* __restrict__ out, int out_act, const ncn_mlp_bwd_src& src) {
  if (row >= n) return;
  if (src.mode == 0) {
    //   0: the (N,OUT) fp16 matrix `dout`
#pragma unroll
    for (int j = 0; j < OUT / 8 && j < 3; ++j) R.a[j] = *reinterpret_cast<const uint4*>(dout + row * OUT + j * 8);
  } else if (src.mode == 1) {
    //   1: colour head - columns [c_off, c_off+n_ch) of the fp32 dL/draws matrix (ncn_field_head_dout fused in)
    const float* rr = src.d_raws + row * src.c_total + src.c_off;
#pragma unroll
    for (int q = 0; q < 3; ++q) R.f[q] = q < src.n_ch ? rr[q] : 0.f;
  } else {
    //   2: density trunk - dL/dh = dL/dx_rgb[:, 3:19] + e0 * dL/dsigma * exp(clamp(h0,-15,15))  (ncn_field_bwd_h fused in)
    const uint4* xr = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(src.dx_rgb) + row * 32);
    R.a[0] = xr[0]; R.a[1] = xr[1];
    if (!(src.perm & 2)) R.a[2] = xr[2];
    R.f[0] = src.d_sigmas[row];
    R.f[1] = __half2float(reinterpret_cast<const __half*>(src.h)[row * 16]);
    if (src.dx_extra != nullptr) {          // dL/dh of a further head on h (sem_net); R.o is free: this net has no output activation
      const uint4* er = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(src.dx_extra) + row * 16);
      R.o[0] = er[0]; R.o[1] = er[1];
    }
  }
  if (out_act == NCN_ACT_SIGMOID || out_act == NCN_ACT_EXP) {
#pragma unroll
    for (int j = 0; j < OUT / 8; ++j) R.o[j] = *reinterpret_cast<const uint4*>(out + row * OUT + j * 8);
  }
}

// L2 prefetch of the rows dout_fetch will read for `row` (no destination registers: fire and forget).  The register-resident
// prefetch this replaces looked free in the source and cost 2600 clocks per tile in the trace: ptxas recycled the loads'
// destination registers for the next address computations, so every load waited for the previous one to RETURN (ncu source
// page: 19 % of the kernel's stall samples on those LDG / LDC instructions, long scoreboard).  Now the rows are pulled into the
// L2 a tile ahead and fetched at the top of their own tile, one L2 round trip.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int OUT>
__device__ __forceinline__ void dout_prefetch(int64_t row, int64_t n, const __half* __restrict__ dout, const __half* __restrict__ out, int out_act,
                                              const ncn_mlp_bwd_src& src) {
  if (row >= n) return;
  if (src.mode == 0) {
    prefetch_l2(dout + row * OUT);
  } else if (src.mode == 1) {
    prefetch_l2(src.d_raws + row * src.c_total + src.c_off);
  } else {
    prefetch_l2(reinterpret_cast<const __half*>(src.dx_rgb) + row * 32);
    prefetch_l2(src.d_sigmas + row);
    prefetch_l2(reinterpret_cast<const __half*>(src.h) + row * 16);
    if (src.dx_extra != nullptr) prefetch_l2(reinterpret_cast<const __half*>(src.dx_extra) + row * 16);
  }
  if (out_act == NCN_ACT_SIGMOID || out_act == NCN_ACT_EXP) prefetch_l2(out + row * OUT);
}

template <int OUT>
__device__ __forceinline__ void dout_finish(const DoutRaw<OUT>& R, bool valid, int out_act, const ncn_mlp_bwd_src& src, uint32_t (&drow)[OUT / 2]) {
#pragma unroll
  for (int q = 0; q < OUT / 2; ++q) drow[q] = 0u;
  if (!valid) return;
  if (src.mode == 0) {
#pragma unroll
    for (int j = 0; j < OUT / 8 && j < 3; ++j) { drow[4 * j] = R.a[j].x; drow[4 * j + 1] = R.a[j].y; drow[4 * j + 2] = R.a[j].z; drow[4 * j + 3] = R.a[j].w; }
  } else if (src.mode == 1) {
    drow[0] = pack_half2(R.f[0] * src.scale, R.f[1] * src.scale);
    drow[1] = pack_half2(R.f[2] * src.scale, 0.f);
  } else {
    if (src.perm & 2) {                                       // fused-forward order: dL/dh = dx_rgb[:, 0:16]
      drow[0] = R.a[0].x; drow[1] = R.a[0].y; drow[2] = R.a[0].z; drow[3] = R.a[0].w;
      drow[4] = R.a[1].x; drow[5] = R.a[1].y; drow[6] = R.a[1].z; drow[7] = R.a[1].w;
    } else {
      const uint32_t wv[12] = {R.a[0].x, R.a[0].y, R.a[0].z, R.a[0].w, R.a[1].x, R.a[1].y, R.a[1].z, R.a[1].w, R.a[2].x, R.a[2].y, R.a[2].z, R.a[2].w};
#pragma unroll
      for (int q = 0; q < 8; ++q) drow[q] = __funnelshift_r(wv[q + 1], wv[q + 2], 16);   // halfs 3+2q, 4+2q
    }
    float2 f = __half22float2(*reinterpret_cast<const __half2*>(&drow[0]));
    if (src.dx_extra != nullptr) {
      const uint32_t ev[8] = {R.o[0].x, R.o[0].y, R.o[0].z, R.o[0].w, R.o[1].x, R.o[1].y, R.o[1].z, R.o[1].w};
#pragma unroll
      for (int q = 1; q < 8; ++q) {
        const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&drow[q]), *reinterpret_cast<const __half2*>(&ev[q]));
        drow[q] = *reinterpret_cast<const uint32_t*>(&sum);
      }
      const float2 e0 = __half22float2(*reinterpret_cast<const __half2*>(&ev[0]));
      f.x += e0.x; f.y += e0.y;
    }
    drow[0] = pack_half2(f.x + R.f[0] * __expf(fminf(fmaxf(R.f[1], -15.f), 15.f)) * src.scale, f.y);
  }
  if (out_act == NCN_ACT_SIGMOID || out_act == NCN_ACT_EXP) {
#pragma unroll
    for (int j = 0; j < OUT / 8; ++j) {
      const uint32_t op[4] = {R.o[j].x, R.o[j].y, R.o[j].z, R.o[j].w};
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

#ifdef NCN_TC05_TRACE      // developer build only (NCN_NVCC_EXTRA=-DNCN_TC05_TRACE): per-tile phase clocks of the first CTAs
__device__ long long g_tc05_trace[16 * 16 * 16];
#define NCN_TRACE(slot) do { if (threadIdx.x == 0 && blockIdx.x < 16 && it < 16) g_tc05_trace[(blockIdx.x * 16 + it) * 16 + (slot)] = clock64(); } while (0)
#define NCN_TRACE_K(slot, v) do { if (threadIdx.x == 0 && blockIdx.x < 16) g_tc05_trace[(blockIdx.x * 16 + 15) * 16 + (slot)] = (v); } while (0)
#else
#define NCN_TRACE(slot) do { } while (0)
#define NCN_TRACE_K(slot, v) do { } while (0)
#endif

template <int IN, int OUT, int NH>
__global__ void __launch_bounds__(kTcThreads, 3)
mlp_bwd_tc05_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ out,
                    const __half* __restrict__ acts, const __half* __restrict__ dout, int64_t n_cap,
                    const int32_t* __restrict__ n_dev, int out_act, float grad_scale, float* __restrict__ grad_w,
                    __half* __restrict__ dx, int* __restrict__ tile_counter, ncn_mlp_bwd_src src) {
  static_assert(OUT == 16, "dL/dout row sources assume a 16-wide (padded) output");
  using LY = TcLayout<IN, OUT, NH>;
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __half* WBl = reinterpret_cast<__half*>(tc_smem);          // last layer: K = OUT, N = 64
  __half* WBh = WBl + LY::kWBl;                              // hidden layers: K = 64, N = 64
  __half* WB0 = WBh + LY::kWBh;                              // first layer: K = 64, N = IN
  __half* sets = WB0 + LY::kWB0;                             // two panel sets: [dz_last | act_0..act_{NH-1} | x]
  uint64_t* mbar_d = reinterpret_cast<uint64_t*>(sets + 2 * LY::kSet);   // MMA chain
  uint64_t* mbar_full = mbar_d + 1;                                       // [2]: activation panels of a set have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar_full + 2);
  int* s_tile = reinterpret_cast<int*>(tmem_slot + 1);

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const bool mma_warp = __shfl_sync(0xffffffffu, wid, 0) == 0;      // provably warp-uniform: the issue path stays on the uniform datapath
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int64_t layer_stride = act_rows(n_cap) * 64;
  const bool perm_in = (src.perm & 1) != 0;
  NCN_TRACE_K(12, clock64());
  NCN_TRACE_K(15, gridDim.x);

  // the bulk copies of one tile's activation panels (one thread)
  auto fetch_acts = [&](int64_t t, int set) {
    __half* A = sets + (size_t)set * LY::kSet + LY::kPdzl;
    mbar_expect_tx(mbar_full + set, NH * LY::kActBytes);
#pragma unroll
    for (int i = 0; i < NH; ++i)
      bulk_g2s(A + (size_t)i * 64 * kTile, acts + (int64_t)i * layer_stride + t * (64 * kTile), LY::kActBytes, mbar_full + set);
  };

  // ---- prologue: barriers, the first tile's loads, weights, TMEM
  int64_t cur = blockIdx.x, nxt = n_tiles;
  int first_next = 0;
  if (tid == 0) {
    mbar_init(mbar_d, 1); mbar_init(mbar_full, 1); mbar_init(mbar_full + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
    if (cur < n_tiles) fetch_acts(cur, 0);
    first_next = 0;
  }
  DoutRaw<OUT> raw;
  // saved activations, inputs and weights come from the forward pass / the optimizer: requested before the wait on the stream
  // predecessor (the kernel that produces the dL/dout sources), so they are in flight while that kernel drains
  if (cur < n_tiles) stage_rows<IN>(x, cur * kTile, n, sets + LY::kPdzl + LY::kPact);
  load_w_panel_async(w + 64 * IN + (NH - 1) * 64 * 64, OUT, 64, WBl);
  for (int i = 0; i < NH - 1; ++i) load_w_panel_async(w + 64 * IN + i * 64 * 64, 64, 64, WBh + i * 64 * 64);
  if (dx) { if (perm_in) load_w_panel_perm(w, 64, IN, WB0); else load_w_panel_async(w, 64, IN, WB0); }
  cp_async_commit();
  pdl_wait(); pdl_trigger();
  if (cur < n_tiles) dout_prefetch<OUT>(cur * kTile + tid, n, dout, out, out_act, src);
  (void)first_next; (void)tile_counter; (void)s_tile;
  // tiles are handed out statically (tile = blockIdx.x + i * gridDim.x): every tile costs the same, and the atomic scheduler put
  // a global round trip (and a memset node per launch) on the thread that issues the MMAs
  __syncwarp();
  if (wid == 0) tmem_alloc<LY::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_d = tmem;                              // dgrad tile: columns [0, 64)
  const uint32_t tmem_w = tmem + LY::kDcols;                 // weight-gradient accumulators
  const uint32_t my_lane = tmem_d + ((uint32_t)(32 * wid) << 16);   // this warp's TMEM sub-partition (row = tid)
  nxt = cur + gridDim.x;

  uint32_t d_phase = 0;
  int it = 0;
  for (; cur < n_tiles; ++it) {
    const int set = it & 1;
    __half* S = sets + (size_t)set * LY::kSet;
    __half* P_dzl = S;
    __half* P_act = S + LY::kPdzl;
    __half* P_x = P_act + LY::kPact;
    const int64_t row = cur * kTile + tid;
    NCN_TRACE(0);
    // (0) this tile's dL/dout row (fetched one tile ago) -> dz_last panel; wait for its staged x rows and activations
    {
      uint32_t drow[OUT / 2];
      dout_fetch<OUT>(raw, row, n, dout, out, out_act, src);      // L2 hits: prefetched one tile ago
      dout_finish<OUT>(raw, row < n, out_act, src, drow);
#pragma unroll
      for (int j = 0; j < OUT / 8; ++j)
        *reinterpret_cast<uint4*>(P_dzl + ((size_t)j * kTile + tid) * 8) = make_uint4(drow[4 * j], drow[4 * j + 1], drow[4 * j + 2], drow[4 * j + 3]);
    }
    NCN_TRACE(1);
    cp_async_wait_all();
    mbar_wait(mbar_full + set, (uint32_t)((it >> 1) & 1));
    if (row >= n) {      // ragged last tile: rows past the live count hold stale activations; they must not reach dW
#pragma unroll
      for (int i = 0; i < NH; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(P_act + (size_t)i * 64 * kTile + ((size_t)c * kTile + tid) * 8) = make_uint4(0, 0, 0, 0);
    }
    NCN_TRACE(2);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    const uint32_t acc0 = it > 0 ? 1u : 0u;
    int after_next = 0;
    if (mma_warp) {
      if (elect_one()) {
        tc_fence_after();
        // dgrad: D = dz_last * W_last ; wgrad: dW_last^T += act_{NH-1}^T * dz_last
        const uint64_t a_k = make_desc_k(P_dzl), b_w = make_desc_w(WBl, OUT);
#pragma unroll
        for (int ks = 0; ks < OUT / 16; ++ks)
          tc_mma_f16(tmem_d, desc_add(a_k, ks * 2 * kTile * 16), desc_add(b_w, ks * 256), make_idesc_dgrad(64), ks > 0 ? 1u : 0u);
        const uint64_t a_mn = make_desc_mn(P_act + (size_t)(NH - 1) * 64 * kTile), b_mn = make_desc_mn(P_dzl);
#pragma unroll
        for (int ks = 0; ks < kTile / 16; ++ks)
          tc_mma_f16(tmem_w + IN + 64 * (NH - 1), desc_add(a_mn, ks * 256), desc_add(b_mn, ks * 256), make_idesc_wgrad(OUT), ks > 0 ? 1u : acc0);
        tc_commit(mbar_d);
      }
      __syncwarp();
    }
    if (tid == 0) {
      after_next = 0;
      if (nxt < n_tiles) fetch_acts(nxt, set ^ 1);             // the other set's last readers (previous tile) were waited for
    }
    NCN_TRACE(3);
    // prefetch the next tile's x rows (cp.async) and dL/dout row (registers)
    if (nxt < n_tiles) {
      stage_rows<IN>(x, nxt * kTile, n, sets + (size_t)(set ^ 1) * LY::kSet + LY::kPdzl + LY::kPact);
      dout_prefetch<OUT>(nxt * kTile + tid, n, dout, out, out_act, src);
    }
    cp_async_commit();
    NCN_TRACE(4);
    // (1) dgrad chain: [tcgen05.ld row, ReLU gate, fp16, st.shared row IN PLACE over the activations] -> MMAs ...
#pragma unroll
    for (int i = NH - 1; i >= 0; --i) {
      mbar_wait(mbar_d, d_phase); d_phase ^= 1u;
      NCN_TRACE(5 + 3 * i);
      tc_fence_after();
      __half* Pz = P_act + (size_t)i * 64 * kTile;
      {
        uint32_t v0[32], v1[32];                               // the whole 64-column accumulator row: two loads in flight, one wait
        tmem_ld32_nowait(my_lane, v0);
        tmem_ld32_nowait(my_lane + 32, v1);
        uint4 am[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) am[c] = *reinterpret_cast<const uint4*>(Pz + ((size_t)c * kTile + tid) * 8);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t* v = (c < 4 ? v0 : v1) + (c & 3) * 8;
          const uint32_t a4[4] = {am[c].x, am[c].y, am[c].z, am[c].w};
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const __half2 gate = __hgt2(*reinterpret_cast<const __half2*>(&a4[q]), __float2half2_rn(0.f));     // 1.0 / 0.0
            const uint32_t pk = pack_half2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
            const __half2 r = __hmul2(*reinterpret_cast<const __half2*>(&pk), gate);
            o[q] = *reinterpret_cast<const uint32_t*>(&r);
          }
          *reinterpret_cast<uint4*>(Pz + ((size_t)c * kTile + tid) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      NCN_TRACE(6 + 3 * i);
      (void)after_next;
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      NCN_TRACE(7 + 3 * i);
      if (mma_warp) {
        if (elect_one()) {
          tc_fence_after();
          const uint64_t a_k = make_desc_k(Pz), a_mn = make_desc_mn(Pz);
          if (i > 0) {
            const uint64_t b_w = make_desc_w(WBh + (size_t)(i - 1) * 64 * 64, 64), b_mn = make_desc_mn(P_act + (size_t)(i - 1) * 64 * kTile);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_f16(tmem_d, desc_add(a_k, ks * 2 * kTile * 16), desc_add(b_w, ks * 256), make_idesc_dgrad(64), ks > 0 ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < kTile / 16; ++ks)
              tc_mma_f16(tmem_w + IN + 64 * (i - 1), desc_add(a_mn, ks * 256), desc_add(b_mn, ks * 256), make_idesc_wgrad(64), ks > 0 ? 1u : acc0);
          } else {
            if (dx) {
              const uint64_t b_w = make_desc_w(WB0, 64);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                tc_mma_f16(tmem_d, desc_add(a_k, ks * 2 * kTile * 16), desc_add(b_w, ks * 256), make_idesc_dgrad(IN), ks > 0 ? 1u : 0u);
            }
            const uint64_t b_mn = make_desc_mn(P_x);
#pragma unroll
            for (int ks = 0; ks < kTile / 16; ++ks)
              tc_mma_f16(tmem_w, desc_add(a_mn, ks * 256), desc_add(b_mn, ks * 256), make_idesc_wgrad(IN), ks > 0 ? 1u : acc0);
          }
          tc_commit(mbar_d);
        }
        __syncwarp();
      }
    }
    // (2) dL/dx rows; the wait also retires the last reads of this panel set
    mbar_wait(mbar_d, d_phase); d_phase ^= 1u;
    NCN_TRACE(11);
    tc_fence_after();
    if (dx) {
      if constexpr (IN == 32) {
        uint32_t v[32];
        tmem_ld32_nowait(my_lane, v);
        tmem_ld_wait();
        if (row < n) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(dx + row * IN + c * 8) =
                make_uint4(pack_half2(__uint_as_float(v[8 * c]), __uint_as_float(v[8 * c + 1])), pack_half2(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3])),
                           pack_half2(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5])), pack_half2(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7])));
        }
      } else {
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
      }
      tc_fence_before();     // the next tile's first MMA overwrites the dgrad tile this thread just read
    }
    NCN_TRACE(12);
    cur = nxt;
    nxt = cur + gridDim.x;
  }
  NCN_TRACE_K(13, clock64());
  // (3) epilogue: weight gradients TMEM -> global (+=).  M = 64 accumulators: rows 16*wid + lane, lanes 0..15
  if (it > 0) {
    const int m = 16 * wid + lane;
    const uint32_t lane_base = tmem_w + ((uint32_t)(32 * wid) << 16);
    uint32_t v[16];
#pragma unroll
    for (int c0 = 0; c0 < IN; c0 += 16) {
      tmem_ld16(lane_base + c0, v);
      if (lane < 16) {
        if (perm_in) {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(grad_w + m * IN + perm_col(c0 + j), __uint_as_float(v[j]) * grad_scale);
        } else {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            red_add_v4(grad_w + m * IN + c0 + j, __uint_as_float(v[j]) * grad_scale, __uint_as_float(v[j + 1]) * grad_scale,
                       __uint_as_float(v[j + 2]) * grad_scale, __uint_as_float(v[j + 3]) * grad_scale);
        }
      }
    }
#pragma unroll
    for (int i = 1; i < NH; ++i)
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        tmem_ld16(lane_base + IN + 64 * (i - 1) + c0, v);
        if (lane < 16)
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            red_add_v4(grad_w + 64 * IN + (i - 1) * 64 * 64 + m * 64 + c0 + j, __uint_as_float(v[j]) * grad_scale,
                       __uint_as_float(v[j + 1]) * grad_scale, __uint_as_float(v[j + 2]) * grad_scale, __uint_as_float(v[j + 3]) * grad_scale);
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
  cp_async_wait_all();
  tc_fence_before();
  __syncthreads();
  NCN_TRACE_K(14, clock64());
  if (wid == 0) tmem_dealloc<LY::kTmemCols>(tmem);
}

}  // namespace ncn

using namespace ncn;

template <int IN, int OUT, int NH>
static int launch_tc05(const void* x, const void* w, const void* out, const void* acts, const void* dout, int64_t n,
                       const int32_t* n_dev, int out_act, float grad_scale, float* grad_w, void* dx, int* tile_counter,
                       const ncn_mlp_bwd_src& src, cudaStream_t st) {
  using LY = TcLayout<IN, OUT, NH>;
  auto k = mlp_bwd_tc05_kernel<IN, OUT, NH>;
  (void)tile_counter;      // static tile assignment: no scheduler state
  NCN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LY::kBytes));
  NCN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  const int64_t tiles = (n + kTile - 1) / kTile;
  // persistent grid = every CTA slot the SMs offer: bounded by shared memory (228 KB per SM at the maximum carveout
  // requested above, 1 KB reserved per CTA - the occupancy API answers for the default carveout) and the 512 TMEM columns
  int occ = (int)((228u * 1024u) / (LY::kBytes + 1024u));
  if (occ > 512 / LY::kTmemCols) occ = 512 / LY::kTmemCols;
  if (occ > 3) occ = 3;                                      // __launch_bounds__(128, 3)
  if (occ < 1) occ = 1;
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > tiles) grid = tiles;
  NCN_CUDA(launch_pdl(k, dim3((unsigned)grid), dim3(kTcThreads), LY::kBytes, st, (const __half*)x, (const __half*)w, (const __half*)out,
                      (const __half*)acts, (const __half*)dout, n, n_dev, out_act, grad_scale, grad_w, (__half*)dx, tile_counter, src));
  return NCN_OK;
}

#ifdef NCN_TC05_TRACE
extern "C" int ncn_debug_tc05_trace(long long* host_dst) {
  return cudaMemcpyFromSymbol(host_dst, ncn::g_tc05_trace, sizeof(long long) * 16 * 16 * 16) == cudaSuccess ? 0 : 1;
}
#endif

// returns NCN_E_UNSUPPORTED when the configuration has no tcgen05 instantiation (the caller falls back to mlp.cu)
int ncn_mlp_bwd_tc05_try(int in_pad, int out_pad, int n_hidden, const void* x, const void* w, const void* out, const void* acts,
                         const void* dout, int64_t n, const int32_t* n_dev, int out_act, float grad_scale, float* grad_w,
                         void* dx, int* tile_counter, int impl, const ncn_mlp_bwd_src* src_in, cudaStream_t st) {
  (void)impl;
  if (!grad_w) return NCN_E_UNSUPPORTED;
  ncn_mlp_bwd_src src;
  if (src_in) src = *src_in; else { src = ncn_mlp_bwd_src(); src.mode = 0; }
  if (src.mode == 1 && src.n_ch > 3) return NCN_E_UNSUPPORTED;
  if (src.mode == 2 && src.dx_extra != nullptr && (out_act == NCN_ACT_SIGMOID || out_act == NCN_ACT_EXP)) return NCN_E_UNSUPPORTED;
  if (((uintptr_t)acts & 127) != 0) return NCN_E_UNSUPPORTED;      // bulk copies want the tiles 128-byte aligned
#define NCN_TC(I, O, H) if (in_pad == I && out_pad == O && n_hidden == H) \
    return launch_tc05<I, O, H>(x, w, out, acts, dout, n, n_dev, out_act, grad_scale, grad_w, dx, tile_counter, src, st);
  NCN_TC(32, 16, 1) NCN_TC(32, 16, 2) NCN_TC(16, 16, 2) NCN_TC(16, 16, 1)
#undef NCN_TC
  return NCN_E_UNSUPPORTED;
}
