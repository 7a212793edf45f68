// Fully fused width-64 MLPs (tcnn "FullyFusedMLP": no biases, ReLU hidden layers, output
// activation None / Sigmoid).  Replaces tcnn.Network as used by models/ngp_mt.py:83-155
// (semantics: SURVEY.md Appendix B; tiny-cuda-nn is not in the reference tree - parity is
// pinned against oracle/mlp.py, fp32-accumulate vs tcnn's fp16-accumulate is deliberate).
//
// This file is the register-chained warp-MMA implementation (mma.sync m16n8k16, fp32
// accumulate): one warp owns a 16-row tile, all weight matrices live in shared memory
// (<= 17 KB), and the accumulator fragments of layer i are re-packed IN REGISTERS into
// the A fragments of layer i+1 - activations never touch shared or global memory inside
// the network.  Forward optionally streams the post-ReLU hidden states out for training.
// Backward = one dgrad kernel (same register chaining through the transposed weights) that
// emits dL/dz per layer, plus a split-K wgrad kernel per layer (ldmatrix.trans operands,
// fp32 register accumulators, one red.global.add.f32 per weight per CTA).
// The tcgen05/TMEM variant for 128-row tiles is in mlp_tc05.cu.
#include "ncn_common.cuh"
#include "mma.cuh"

namespace ncn {

constexpr int kW = 64;            // hidden width
constexpr int kPad = 8;           // smem row padding (halfs): conflict-free fragment loads
constexpr int kMlpThreads = 128;

// C[16 x N] = A[16 x K] * W^T, W stored in smem as rows n (stride K+kPad halfs)
template <int K, int N>
__device__ __forceinline__ void warp_layer(const uint32_t (*a)[4], const __half* __restrict__ W, float (*c)[4], int g, int t) {
#pragma unroll
  for (int nt = 0; nt < N / 8; ++nt) { c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f; }
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
#pragma unroll
    for (int nt = 0; nt < N / 8; ++nt) {
      const __half* wr = W + (nt * 8 + g) * (K + kPad) + kb * 16 + 2 * t;
      mma16816(c[nt], a[kb], *reinterpret_cast<const uint32_t*>(wr), *reinterpret_cast<const uint32_t*>(wr + 8));
    }
  }
}

// accumulators (16 x 64, C layout) -> A fragments of the next layer (4 k-blocks)
__device__ __forceinline__ void c_to_a64(const float (*c)[4], uint32_t (*a)[4]) {
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    a[kb][0] = pack_half2(c[2 * kb][0], c[2 * kb][1]);
    a[kb][1] = pack_half2(c[2 * kb][2], c[2 * kb][3]);
    a[kb][2] = pack_half2(c[2 * kb + 1][0], c[2 * kb + 1][1]);
    a[kb][3] = pack_half2(c[2 * kb + 1][2], c[2 * kb + 1][3]);
  }
}

// load a 16 x K tile of a row-major (rows, K) fp16 matrix as A fragments (rows >= n read as 0)
template <int K>
__device__ __forceinline__ void load_a(const __half* __restrict__ x, int64_t row0, int64_t n, uint32_t (*a)[4], int g, int t) {
  const int64_t r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
    const int col = kb * 16 + 2 * t;
    a[kb][0] = r0 < n ? *reinterpret_cast<const uint32_t*>(x + r0 * K + col) : 0u;
    a[kb][1] = r1 < n ? *reinterpret_cast<const uint32_t*>(x + r1 * K + col) : 0u;
    a[kb][2] = r0 < n ? *reinterpret_cast<const uint32_t*>(x + r0 * K + col + 8) : 0u;
    a[kb][3] = r1 < n ? *reinterpret_cast<const uint32_t*>(x + r1 * K + col + 8) : 0u;
  }
}

// store A-fragment-packed 16 x K tile to a row-major (rows, K) fp16 matrix
template <int K>
__device__ __forceinline__ void store_a(__half* __restrict__ y, int64_t row0, int64_t n, const uint32_t (*a)[4], int g, int t) {
  const int64_t r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
    const int col = kb * 16 + 2 * t;
    if (r0 < n) { *reinterpret_cast<uint32_t*>(y + r0 * K + col) = a[kb][0]; *reinterpret_cast<uint32_t*>(y + r0 * K + col + 8) = a[kb][2]; }
    if (r1 < n) { *reinterpret_cast<uint32_t*>(y + r1 * K + col) = a[kb][1]; *reinterpret_cast<uint32_t*>(y + r1 * K + col + 8) = a[kb][3]; }
  }
}

// store an A-fragment-packed 16 x 64 hidden-state tile into the tiled activation layout (act_offset): the 8 rows x 4
// lanes of one fragment register land in 8 x 16 B = 128 contiguous bytes
__device__ __forceinline__ void store_act(__half* __restrict__ y, int64_t row0, int64_t n, const uint32_t (*a)[4], int g, int t) {
  const int64_t r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    const int col = kb * 16 + 2 * t;
    if (r0 < n) { *reinterpret_cast<uint32_t*>(y + act_offset(r0, col)) = a[kb][0]; *reinterpret_cast<uint32_t*>(y + act_offset(r0, col + 8)) = a[kb][2]; }
    if (r1 < n) { *reinterpret_cast<uint32_t*>(y + act_offset(r1, col)) = a[kb][1]; *reinterpret_cast<uint32_t*>(y + act_offset(r1, col + 8)) = a[kb][3]; }
  }
}

// cooperative copy of a (rows, cols) row-major fp16 matrix into smem with row stride cols+kPad
__device__ __forceinline__ void load_w(const __half* __restrict__ w, int rows, int cols, __half* __restrict__ s) {
  for (int i = threadIdx.x; i < rows * cols / 2; i += blockDim.x) {
    const int r = (2 * i) / cols, c = (2 * i) % cols;
    *reinterpret_cast<uint32_t*>(s + r * (cols + kPad) + c) = *reinterpret_cast<const uint32_t*>(w + 2 * i);
  }
}
// same, transposed: s[c][r] = w[r][c], row stride rows+kPad
__device__ __forceinline__ void load_w_t(const __half* __restrict__ w, int rows, int cols, __half* __restrict__ s) {
  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
    const int r = i / cols, c = i % cols;
    s[c * (rows + kPad) + r] = w[i];
  }
}

__device__ __forceinline__ float act_out(float v, int act) {
  if (act == NCN_ACT_SIGMOID) return 1.0f / (1.0f + __expf(-v));
  if (act == NCN_ACT_RELU) return fmaxf(v, 0.f);
  if (act == NCN_ACT_EXP) return __expf(v);
  return v;
}

// ------------------------------------------------------------------------------ forward
template <int IN, int OUT>
__global__ void __launch_bounds__(kMlpThreads)
mlp_fwd_kernel(const __half* __restrict__ x, const __half* __restrict__ w, int64_t n_cap, const int32_t* __restrict__ n_dev,
               int n_hidden, int out_act, __half* __restrict__ out, __half* __restrict__ acts) {
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  extern __shared__ __align__(16) __half smem[];
  __half* W0 = smem;                                   // [64][IN+8]
  __half* Wh = W0 + kW * (IN + kPad);                  // (n_hidden-1) x [64][72]
  __half* Wl = Wh + (n_hidden - 1) * kW * (kW + kPad); // [OUT][72]
  load_w(w, kW, IN, W0);
  for (int i = 0; i < n_hidden - 1; ++i) load_w(w + kW * IN + i * kW * kW, kW, kW, Wh + i * kW * (kW + kPad));
  load_w(w + kW * IN + (n_hidden - 1) * kW * kW, OUT, kW, Wl);
  __syncthreads();

  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_tiles = (n + 15) >> 4;
  for (int64_t tile = warp; tile < n_tiles; tile += n_warps) {
    const int64_t row0 = tile << 4;
    uint32_t ain[IN / 16][4];
    load_a<IN>(x, row0, n, ain, g, t);
    float c[8][4];
    warp_layer<IN, kW>(ain, W0, c, g, t);
    uint32_t h[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = fmaxf(c[j][0], 0.f); c[j][1] = fmaxf(c[j][1], 0.f); c[j][2] = fmaxf(c[j][2], 0.f); c[j][3] = fmaxf(c[j][3], 0.f); }
    c_to_a64(c, h);
    if (acts) store_act(acts, row0, n, h, g, t);
    for (int i = 1; i < n_hidden; ++i) {
      warp_layer<kW, kW>(h, Wh + (i - 1) * kW * (kW + kPad), c, g, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j][0] = fmaxf(c[j][0], 0.f); c[j][1] = fmaxf(c[j][1], 0.f); c[j][2] = fmaxf(c[j][2], 0.f); c[j][3] = fmaxf(c[j][3], 0.f); }
      c_to_a64(c, h);
      if (acts) store_act(acts + (int64_t)i * act_rows(n_cap) * kW, row0, n, h, g, t);
    }
    float co[OUT / 8][4];
    warp_layer<kW, OUT>(h, Wl, co, g, t);
    const int64_t r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
    for (int j = 0; j < OUT / 8; ++j) {
      const int col = j * 8 + 2 * t;
      if (r0 < n) *reinterpret_cast<uint32_t*>(out + r0 * OUT + col) = pack_half2(act_out(co[j][0], out_act), act_out(co[j][1], out_act));
      if (r1 < n) *reinterpret_cast<uint32_t*>(out + r1 * OUT + col) = pack_half2(act_out(co[j][2], out_act), act_out(co[j][3], out_act));
    }
  }
}

// ------------------------------------------------------------------------------ dgrad
// mask accumulators (C layout, 16 x 64) with relu'(act) read from one layer of the tiled activations (act_offset)
__device__ __forceinline__ void relu_mask(float (*c)[4], const __half* __restrict__ act, int64_t row0, int64_t n, int g, int t) {
  const int64_t r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = j * 8 + 2 * t;
    const __half2 z = __floats2half2_rn(0.f, 0.f);
    const __half2 a0 = r0 < n ? *reinterpret_cast<const __half2*>(act + act_offset(r0, col)) : z;
    const __half2 a1 = r1 < n ? *reinterpret_cast<const __half2*>(act + act_offset(r1, col)) : z;
    if (!(__low2float(a0) > 0.f)) c[j][0] = 0.f;
    if (!(__high2float(a0) > 0.f)) c[j][1] = 0.f;
    if (!(__low2float(a1) > 0.f)) c[j][2] = 0.f;
    if (!(__high2float(a1) > 0.f)) c[j][3] = 0.f;
  }
}

template <int IN, int OUT>
__global__ void __launch_bounds__(kMlpThreads)
mlp_dgrad_kernel(const __half* __restrict__ w, const __half* __restrict__ out, const __half* __restrict__ acts,
                 const __half* __restrict__ dout, int64_t n_cap, const int32_t* __restrict__ n_dev, int n_hidden, int out_act,
                 __half* __restrict__ dz_last, __half* __restrict__ dz_hidden, __half* __restrict__ dx) {
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  extern __shared__ __align__(16) __half smem[];
  __half* WlT = smem;                                        // [64][OUT+8]   (Wl^T)
  __half* WhT = WlT + kW * (OUT + kPad);                     // (n_hidden-1) x [64][72]
  __half* W0T = WhT + (n_hidden - 1) * kW * (kW + kPad);     // [IN][72]      (W0^T), only if dx
  load_w_t(w + kW * IN + (n_hidden - 1) * kW * kW, OUT, kW, WlT);
  for (int i = 0; i < n_hidden - 1; ++i) load_w_t(w + kW * IN + i * kW * kW, kW, kW, WhT + i * kW * (kW + kPad));
  if (dx) load_w_t(w, kW, IN, W0T);
  __syncthreads();

  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_tiles = (n + 15) >> 4;
  for (int64_t tile = warp; tile < n_tiles; tile += n_warps) {
    const int64_t row0 = tile << 4;
    uint32_t dz[OUT / 16][4];
    load_a<OUT>(dout, row0, n, dz, g, t);
    if (out_act == NCN_ACT_SIGMOID || out_act == NCN_ACT_EXP) {
      uint32_t o[OUT / 16][4];
      load_a<OUT>(out, row0, n, o, g, t);
#pragma unroll
      for (int kb = 0; kb < OUT / 16; ++kb)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&dz[kb][q]));
          const float2 y = __half22float2(*reinterpret_cast<const __half2*>(&o[kb][q]));
          if (out_act == NCN_ACT_SIGMOID) dz[kb][q] = pack_half2(d.x * y.x * (1.f - y.x), d.y * y.y * (1.f - y.y));
          else dz[kb][q] = pack_half2(d.x * y.x, d.y * y.y);
        }
    }
    store_a<OUT>(dz_last, row0, n, dz, g, t);
    float c[8][4];
    warp_layer<OUT, kW>(dz, WlT, c, g, t);               // dL/dh (16 x 64)
    uint32_t dh[4][4];
    for (int i = n_hidden - 1; i >= 0; --i) {
      relu_mask(c, acts + (int64_t)i * act_rows(n_cap) * kW, row0, n, g, t);
      c_to_a64(c, dh);
      store_a<kW>(dz_hidden + (int64_t)i * n_cap * kW, row0, n, dh, g, t);
      if (i > 0) warp_layer<kW, kW>(dh, WhT + (i - 1) * kW * (kW + kPad), c, g, t);
    }
    if (dx) {
      float cx[IN / 8][4];
      warp_layer<kW, IN>(dh, W0T, cx, g, t);
      const int64_t r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
      for (int j = 0; j < IN / 8; ++j) {
        const int col = j * 8 + 2 * t;
        if (r0 < n) *reinterpret_cast<uint32_t*>(dx + r0 * IN + col) = pack_half2(cx[j][0], cx[j][1]);
        if (r1 < n) *reinterpret_cast<uint32_t*>(dx + r1 * IN + col) = pack_half2(cx[j][2], cx[j][3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------ wgrad
// grad_w[M x NN] += scale * dz^T (M x n) * a (n x NN);  dz (n, M) row-major fp16; a (n, NN) row-major fp16, or
// (a_tiled, NN == 64) one layer of the tiled activations.
constexpr int kChunk = 64;
template <int M, int NN>
__global__ void __launch_bounds__(kMlpThreads)
mlp_wgrad_kernel(const __half* __restrict__ dz, const __half* __restrict__ a, int a_tiled, int64_t n_cap,
                 const int32_t* __restrict__ n_dev, float scale, float* __restrict__ grad_w) {
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  constexpr int TM = M / 16, TN = NN / 8, TOTAL = TM * TN;
  constexpr int PER = TOTAL >= 4 ? TOTAL / 4 : 1;
  static_assert(TOTAL < 4 || TOTAL % 4 == 0, "tile split");
  __shared__ __align__(16) __half Sd[kChunk][M + kPad];
  __shared__ __align__(16) __half Sa[kChunk][NN + kPad];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  float acc[PER][4];
#pragma unroll
  for (int p = 0; p < PER; ++p) acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0.f;
  const bool active = wid * PER < TOTAL;
  const int64_t n_chunks = (n + kChunk - 1) / kChunk;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t row0 = ch * kChunk;
    // cooperative 16 B loads (rows beyond n -> zeros)
    for (int i = threadIdx.x; i < kChunk * (M / 8); i += kMlpThreads) {
      const int r = i / (M / 8), c8 = i % (M / 8);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row0 + r < n) v = *reinterpret_cast<const uint4*>(dz + (row0 + r) * M + c8 * 8);
      *reinterpret_cast<uint4*>(&Sd[r][c8 * 8]) = v;
    }
    for (int i = threadIdx.x; i < kChunk * (NN / 8); i += kMlpThreads) {
      const int r = i / (NN / 8), c8 = i % (NN / 8);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row0 + r < n) v = *reinterpret_cast<const uint4*>(a_tiled ? a + act_offset(row0 + r, c8 * 8) : a + (row0 + r) * NN + c8 * 8);
      *reinterpret_cast<uint4*>(&Sa[r][c8 * 8]) = v;
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int ks = 0; ks < kChunk / 16; ++ks) {
        const int k0 = ks * 16;
        int prev_mt = -1;
        uint32_t af[4];
#pragma unroll
        for (int p = 0; p < PER; ++p) {
          const int id = wid * PER + p, mt = id / TN, nt = id % TN;
          if (mt != prev_mt) {
            // A[m][k] = Sd[k][m]: matrices (k0.., m0), (k0.., m0+8), (k0+8.., m0), (k0+8.., m0+8)
            const int kr = k0 + (lane & 7) + ((lane & 16) ? 8 : 0);
            const int mc = mt * 16 + ((lane & 8) ? 8 : 0);
            ldmatrix_x4_trans(af, &Sd[kr][mc]);
            prev_mt = mt;
          }
          uint32_t bf[2];
          // B[k][n] = Sa[k][n]: matrices (k0.., n0), (k0+8.., n0)
          const int kr = k0 + (lane & 7) + ((lane & 8) ? 8 : 0);
          ldmatrix_x2_trans(bf, &Sa[kr][nt * 8]);
          mma16816(acc[p], af, bf[0], bf[1]);
        }
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int id = wid * PER + p, mt = id / TN, nt = id % TN;
      const int m0 = mt * 16 + g, c0 = nt * 8 + 2 * t;
      atomicAdd(grad_w + m0 * NN + c0, acc[p][0] * scale);
      atomicAdd(grad_w + m0 * NN + c0 + 1, acc[p][1] * scale);
      atomicAdd(grad_w + (m0 + 8) * NN + c0, acc[p][2] * scale);
      atomicAdd(grad_w + (m0 + 8) * NN + c0 + 1, acc[p][3] * scale);
    }
  }
}

}  // namespace ncn

using namespace ncn;

static inline int pad16(int v) { return (v + 15) / 16 * 16; }

static int check_desc(const ncn_mlp_desc* d, int* in_pad, int* out_pad) {
  if (!d) return NCN_E_NULL;
  if (d->width != 64 || d->n_hidden < 1 || d->n_hidden > 8 || d->n_in < 1 || d->n_out < 1) return NCN_E_CONFIG;
  if (d->activation != NCN_ACT_RELU) return NCN_E_CONFIG;
  *in_pad = pad16(d->n_in); *out_pad = pad16(d->n_out);
  if (*in_pad > 64 || *out_pad > 64) return NCN_E_CONFIG;
  return NCN_OK;
}

extern "C" int64_t ncn_mlp_n_params(const ncn_mlp_desc* d) {
  int ip, op;
  if (check_desc(d, &ip, &op)) return -1;
  return (int64_t)64 * ip + (int64_t)(d->n_hidden - 1) * 64 * 64 + (int64_t)op * 64;
}

extern "C" size_t ncn_mlp_acts_bytes(const ncn_mlp_desc* d, int64_t n) {
  int ip, op;
  if (check_desc(d, &ip, &op) || n < 0) return 0;
  return (size_t)d->n_hidden * (size_t)act_rows(n) * 64 * sizeof(__half);
}

extern "C" size_t ncn_mlp_bwd_workspace_bytes(const ncn_mlp_desc* d, int64_t n) {
  int ip, op;
  if (check_desc(d, &ip, &op) || n < 0) return 0;
  return (size_t)n * (size_t)(op + d->n_hidden * 64) * sizeof(__half) + 256;
}

template <int IN, int OUT>
static int launch_fwd(const ncn_mlp_desc* d, const void* x, const void* w, int64_t n, void* out, void* acts, const int32_t* n_dev,
                      cudaStream_t st) {
  const size_t smem = (size_t)(64 * (IN + kPad) + (d->n_hidden - 1) * 64 * (64 + kPad) + OUT * (64 + kPad)) * sizeof(__half);
  auto k = mlp_fwd_kernel<IN, OUT>;
  if (smem > 48 * 1024) NCN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = persistent_grid(((n + 15) / 16) * 32, kMlpThreads, 8);
  k<<<grid, kMlpThreads, smem, st>>>((const __half*)x, (const __half*)w, n, n_dev, d->n_hidden, d->out_activation, (__half*)out,
                                     (__half*)acts);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

template <int M, int NN>
static int launch_wgrad(const __half* dz, const __half* a, int a_tiled, int64_t n, const int32_t* n_dev, float scale, float* gw, cudaStream_t st) {
  const int64_t chunks = (n + kChunk - 1) / kChunk;
  int64_t grid = (int64_t)sm_count() * 4;
  if (grid > chunks) grid = chunks;
  mlp_wgrad_kernel<M, NN><<<(int)grid, kMlpThreads, 0, st>>>(dz, a, a_tiled, n, n_dev, scale, gw);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

static int wgrad_dispatch(int M, int NN, const __half* dz, const __half* a, int a_tiled, int64_t n, const int32_t* n_dev, float scale,
                          float* gw, cudaStream_t st) {
#define NCN_WG(MM, N2) if (M == MM && NN == N2) return launch_wgrad<MM, N2>(dz, a, a_tiled, n, n_dev, scale, gw, st);
  NCN_WG(64, 16) NCN_WG(64, 32) NCN_WG(64, 48) NCN_WG(64, 64)
  NCN_WG(16, 64) NCN_WG(32, 64) NCN_WG(48, 64)
#undef NCN_WG
  return NCN_E_CONFIG;
}

template <int IN, int OUT>
static int launch_bwd(const ncn_mlp_desc* d, const void* x, const void* w, const void* out, const void* acts,
                      const void* dout, int64_t n, float* grad_w, void* dx, float grad_scale, void* scratch, const int32_t* n_dev,
                      cudaStream_t st) {
  __half* dz_last = (__half*)scratch;
  __half* dz_hidden = dz_last + (size_t)n * OUT;
  const size_t smem = (size_t)(64 * (OUT + kPad) + (d->n_hidden - 1) * 64 * (64 + kPad) + (dx ? IN * (64 + kPad) : 0)) * sizeof(__half);
  auto k = mlp_dgrad_kernel<IN, OUT>;
  if (smem > 48 * 1024) NCN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = persistent_grid(((n + 15) / 16) * 32, kMlpThreads, 8);
  k<<<grid, kMlpThreads, smem, st>>>((const __half*)w, (const __half*)out, (const __half*)acts, (const __half*)dout, n, n_dev,
                                     d->n_hidden, d->out_activation, dz_last, dz_hidden, (__half*)dx);
  NCN_LAUNCH_OK();
  if (grad_w) {
    // layer 0: dz_hidden[0]^T x
    int rc = wgrad_dispatch(64, IN, dz_hidden, (const __half*)x, 0, n, n_dev, grad_scale, grad_w, st);
    if (rc) return rc;
    for (int i = 1; i < d->n_hidden; ++i) {
      rc = wgrad_dispatch(64, 64, dz_hidden + (size_t)i * n * 64, (const __half*)acts + (size_t)(i - 1) * act_rows(n) * 64, 1, n, n_dev, grad_scale,
                          grad_w + 64 * IN + (size_t)(i - 1) * 64 * 64, st);
      if (rc) return rc;
    }
    rc = wgrad_dispatch(OUT, 64, dz_last, (const __half*)acts + (size_t)(d->n_hidden - 1) * act_rows(n) * 64, 1, n, n_dev, grad_scale,
                        grad_w + 64 * IN + (size_t)(d->n_hidden - 1) * 64 * 64, st);
    if (rc) return rc;
  }
  return NCN_OK;
}

#define NCN_MLP_DISPATCH(IP, OP, CALL)                                               \
  if (IP == 16 && OP == 16) { constexpr int kI = 16, kO = 16; return CALL; }         \
  if (IP == 32 && OP == 16) { constexpr int kI = 32, kO = 16; return CALL; }         \
  if (IP == 16 && OP == 32) { constexpr int kI = 16, kO = 32; return CALL; }         \
  if (IP == 16 && OP == 48) { constexpr int kI = 16, kO = 48; return CALL; }         \
  if (IP == 32 && OP == 32) { constexpr int kI = 32, kO = 32; return CALL; }         \
  if (IP == 64 && OP == 16) { constexpr int kI = 64, kO = 16; return CALL; }         \
  if (IP == 64 && OP == 64) { constexpr int kI = 64, kO = 64; return CALL; }         \
  if (IP == 48 && OP == 16) { constexpr int kI = 48, kO = 16; return CALL; }         \
  return NCN_E_CONFIG;

extern "C" int ncn_mlp_fwd(const ncn_mlp_desc* d, const void* x, const void* w, int64_t n, void* out, void* acts,
                           const int32_t* n_dev, ncn_stream_t stream) {
  int ip, op;
  int rc = check_desc(d, &ip, &op); if (rc) return rc;
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(x); NCN_CHECK_PTR(w); NCN_CHECK_PTR(out);
  if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)out | (uintptr_t)acts) & 15) return NCN_E_ALIGN;
  NCN_MLP_DISPATCH(ip, op, (launch_fwd<kI, kO>(d, x, w, n, out, acts, n_dev, as_stream(stream))))
}

// tcgen05 / TMEM implementation (mlp_tc05.cu)
int ncn_mlp_bwd_tc05_try(int in_pad, int out_pad, int n_hidden, const void* x, const void* w, const void* out, const void* acts,
                         const void* dout, int64_t n, const int32_t* n_dev, int out_act, float grad_scale, float* grad_w,
                         void* dx, int* tile_counter, int impl, const ncn_mlp_bwd_src* src, cudaStream_t st);
static int g_mlp_bwd_impl = 1;     // 1 = tcgen05 / TMEM kernel (default; falls back per shape), 0 = warp-MMA dgrad + split-K wgrad kernels
extern "C" int ncn_set_mlp_bwd_impl(int impl) { const int old = g_mlp_bwd_impl; g_mlp_bwd_impl = impl; return old; }

extern "C" int ncn_mlp_bwd(const ncn_mlp_desc* d, const void* x, const void* w, const void* out, const void* acts,
                           const void* dL_dout, int64_t n, float* grad_w, void* dL_dx, float grad_scale, void* scratch,
                           size_t scratch_bytes, const int32_t* n_dev, ncn_stream_t stream) {
  int ip, op;
  int rc = check_desc(d, &ip, &op); if (rc) return rc;
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(x); NCN_CHECK_PTR(w); NCN_CHECK_PTR(out); NCN_CHECK_PTR(acts); NCN_CHECK_PTR(dL_dout); NCN_CHECK_PTR(scratch);
  if (scratch_bytes < ncn_mlp_bwd_workspace_bytes(d, n)) return NCN_E_SIZE;
  if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)out | (uintptr_t)acts | (uintptr_t)dL_dout | (uintptr_t)scratch | (uintptr_t)dL_dx) & 15)
    return NCN_E_ALIGN;
  if (g_mlp_bwd_impl >= 1) {
    // the last 256 bytes of the caller's scratch hold the tile counter of the persistent kernel
    int* tile_counter = (int*)((char*)scratch + ((ncn_mlp_bwd_workspace_bytes(d, n) - 256) & ~(size_t)15));
    rc = ncn_mlp_bwd_tc05_try(ip, op, d->n_hidden, x, w, out, acts, dL_dout, n, n_dev, d->out_activation, grad_scale, grad_w, dL_dx,
                              tile_counter, g_mlp_bwd_impl, nullptr, as_stream(stream));
    if (rc != NCN_E_UNSUPPORTED) return rc;
  }
  NCN_MLP_DISPATCH(ip, op, (launch_bwd<kI, kO>(d, x, w, out, acts, dL_dout, n, grad_w, dL_dx, grad_scale, scratch, n_dev, as_stream(stream))))
}

// ncn_mlp_bwd with the gradient w.r.t. the network output assembled on the fly from the compositing backward's outputs
// (removes the ncn_field_head_dout / ncn_field_bwd_h passes).  tcgen05 implementation only.
extern "C" int ncn_mlp_bwd_src_fused(const ncn_mlp_desc* d, const ncn_mlp_bwd_src* src, const void* x, const void* w, const void* out,
                                     const void* acts, int64_t n, float* grad_w, void* dL_dx, float grad_scale, void* scratch,
                                     size_t scratch_bytes, const int32_t* n_dev, ncn_stream_t stream) {
  int ip, op;
  int rc = check_desc(d, &ip, &op); if (rc) return rc;
  NCN_CHECK_SIZE(n >= 0);
  NCN_CHECK_PTR(src);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(x); NCN_CHECK_PTR(w); NCN_CHECK_PTR(acts); NCN_CHECK_PTR(scratch); NCN_CHECK_PTR(grad_w);
  if (src->mode == 1) { NCN_CHECK_PTR(src->d_raws); NCN_CHECK_PTR(out); }
  else if (src->mode == 2) { NCN_CHECK_PTR(src->dx_rgb); NCN_CHECK_PTR(src->d_sigmas); NCN_CHECK_PTR(src->h); }
  else return NCN_E_CONFIG;
  if (scratch_bytes < ncn_mlp_bwd_workspace_bytes(d, n)) return NCN_E_SIZE;
  if (g_mlp_bwd_impl < 1) return NCN_E_UNSUPPORTED;
  int* tile_counter = (int*)((char*)scratch + ((ncn_mlp_bwd_workspace_bytes(d, n) - 256) & ~(size_t)15));
  return ncn_mlp_bwd_tc05_try(ip, op, d->n_hidden, x, w, out, acts, nullptr, n, n_dev, d->out_activation, grad_scale, grad_w, dL_dx,
                              tile_counter, g_mlp_bwd_impl, src, as_stream(stream));
}
