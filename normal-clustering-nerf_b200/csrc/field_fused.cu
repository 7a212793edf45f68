// Fused NGPMT field forward (models/ngp_mt.py:157-229 for the RGB+density configuration):
//   hash-grid gather -> sigma net (32->64->16) -> sigma = TruncExp(h0) -> [h | d/|d| | 1] -> rgb net (32->64->64->16, sigmoid)
// in ONE kernel.  A warp owns a 16-sample tile; the multiresolution features are gathered straight into the A fragments
// of the first MMA (thread (g,t) interpolates levels {t,t+4,t+8,t+12} of samples g and g+8), every layer's accumulators
// are re-packed in registers into the next layer's A fragments, and the density trunk's output h IS the first k-block of
// the colour head's input - so between the table lookups and the final (sigma, rgb) nothing is read back from memory.
// What training needs later (features, h, hidden activations, colour-head input) is streamed out once, fp16.
//
// Column order of the colour head's input inside this kernel (and in the stored x_rgb, and in the backward kernel when
// its `perm` flag is set): [h (16) | d (3) | ones (13)] instead of the reference's cat([d, h]) + padding - a pure
// re-indexing of the first layer's weight columns (perm[k'] = k'+3 for k'<16, k'-16 for 16<=k'<19, k' otherwise).
#include "ncn_common.cuh"
#include "mma.cuh"

namespace ncn {

constexpr int kFfThreads = 128;
constexpr int kFfPad = 8;

struct FfGridMeta {
  float scale[16];
  uint32_t res[16], size[16], offset[16];
  float lo[3], inv_size_dummy;   // lo / size: input normalisation (x - lo) / size
  float size3[3];
  int xform_on;
};

__device__ __forceinline__ uint32_t ff_grid_index(uint32_t gx, uint32_t gy, uint32_t gz, uint32_t res, uint32_t size) {
  uint32_t stride = 1, index = 0;
  index += gx * stride; stride *= res;
  if (stride <= size) { index += gy * stride; stride *= res;
    if (stride <= size) { index += gz * stride; stride *= res; } }
  if (size < stride) index = gx ^ (gy * 2654435761u) ^ (gz * 805459861u);
  if ((size & (size - 1u)) == 0u) return index & (size - 1u);
  if (index < size) return index;
  return index % size;
}

// trilinear lookup of one (sample, level): returns the two features
__device__ __forceinline__ float2 ff_lookup(const float* __restrict__ xn, const __half2* __restrict__ tl, float scale, uint32_t res,
                                            uint32_t size) {
  uint32_t g[3]; float w[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = fmaf(scale, xn[d], 0.5f);
    const float fl = floorf(pos);
    g[d] = (uint32_t)(int)fl;
    w[d] = pos - fl;
  }
  __half2 v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    v[k] = __ldg(tl + ff_grid_index(g[0] + (k & 1), g[1] + ((k >> 1) & 1), g[2] + ((k >> 2) & 1), res, size));
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float wk = ((k & 1) ? w[0] : 1.f - w[0]) * ((k & 2) ? w[1] : 1.f - w[1]) * ((k & 4) ? w[2] : 1.f - w[2]);
    const float2 f2 = __half22float2(v[k]);
    a0 = fmaf(wk, f2.x, a0); a1 = fmaf(wk, f2.y, a1);
  }
  return make_float2(a0, a1);
}

template <int K, int N>
__device__ __forceinline__ void ff_layer(const uint32_t (*a)[4], const __half* __restrict__ W, float (*c)[4], int g, int t) {
#pragma unroll
  for (int nt = 0; nt < N / 8; ++nt) { c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f; }
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
#pragma unroll
    for (int nt = 0; nt < N / 8; ++nt) {
      const __half* wr = W + (nt * 8 + g) * (K + kFfPad) + kb * 16 + 2 * t;
      mma16816(c[nt], a[kb], *reinterpret_cast<const uint32_t*>(wr), *reinterpret_cast<const uint32_t*>(wr + 8));
    }
  }
}
__device__ __forceinline__ void ff_relu_pack(float (*c)[4], uint32_t (*a)[4]) {
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    a[kb][0] = pack_half2(fmaxf(c[2 * kb][0], 0.f), fmaxf(c[2 * kb][1], 0.f));
    a[kb][1] = pack_half2(fmaxf(c[2 * kb][2], 0.f), fmaxf(c[2 * kb][3], 0.f));
    a[kb][2] = pack_half2(fmaxf(c[2 * kb + 1][0], 0.f), fmaxf(c[2 * kb + 1][1], 0.f));
    a[kb][3] = pack_half2(fmaxf(c[2 * kb + 1][2], 0.f), fmaxf(c[2 * kb + 1][3], 0.f));
  }
}
template <int K>
__device__ __forceinline__ void ff_store_a(__half* __restrict__ y, int64_t row0, int64_t n, const uint32_t (*a)[4], int g, int t) {
  const int64_t r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
    const int col = kb * 16 + 2 * t;
    if (r0 < n) { *reinterpret_cast<uint32_t*>(y + r0 * K + col) = a[kb][0]; *reinterpret_cast<uint32_t*>(y + r0 * K + col + 8) = a[kb][2]; }
    if (r1 < n) { *reinterpret_cast<uint32_t*>(y + r1 * K + col) = a[kb][1]; *reinterpret_cast<uint32_t*>(y + r1 * K + col + 8) = a[kb][3]; }
  }
}
// hidden-state tile -> tiled activation layout (ncn_common.cuh act_offset)
__device__ __forceinline__ void ff_store_act(__half* __restrict__ y, int64_t row0, int64_t n, const uint32_t (*a)[4], int g, int t) {
  const int64_t r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    const int col = kb * 16 + 2 * t;
    if (r0 < n) { *reinterpret_cast<uint32_t*>(y + act_offset(r0, col)) = a[kb][0]; *reinterpret_cast<uint32_t*>(y + act_offset(r0, col + 8)) = a[kb][2]; }
    if (r1 < n) { *reinterpret_cast<uint32_t*>(y + act_offset(r1, col)) = a[kb][1]; *reinterpret_cast<uint32_t*>(y + act_offset(r1, col + 8)) = a[kb][3]; }
  }
}
__device__ __forceinline__ void ff_load_w(const __half* __restrict__ w, int rows, int cols, __half* __restrict__ s, bool perm) {
  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
    const int r = i / cols, c = i % cols;             // c = kernel-internal column
    const int src = perm ? (c < 16 ? c + 3 : (c < 19 ? c - 16 : c)) : c;
    s[r * (cols + kFfPad) + c] = w[r * cols + src];
  }
}

// kGrid = true : hash-grid encode -> density trunk -> colour head (ncn_field_fwd)
// kGrid = false: the two MLPs only; `feat` is the INPUT (N,32) f16 produced by ncn_grid_fwd (ncn_field_mlp_fwd)
template <bool kGrid>
__global__ void __launch_bounds__(kFfThreads)
field_fwd_kernel(const __grid_constant__ FfGridMeta meta, const float* __restrict__ x, const float* __restrict__ dirs,
                 const __half* __restrict__ table, const __half* __restrict__ w_sigma, const __half* __restrict__ w_rgb,
                 int64_t n_cap, const int32_t* __restrict__ n_dev, float* __restrict__ sigmas, float* __restrict__ raws,
                 int c_total, __half* __restrict__ feat, __half* __restrict__ h_out, __half* __restrict__ sig_acts,
                 __half* __restrict__ x_rgb, __half* __restrict__ rgb_acts, __half* __restrict__ rgb_out) {
  __shared__ __align__(16) __half sW[64 * 40 + 16 * 72 + 64 * 40 + 64 * 72 + 16 * 72];
  __shared__ FfGridMeta sm;
  __half* S0 = sW;                      // sigma layer 0  [64][32+8]
  __half* S1 = S0 + 64 * 40;            // sigma layer 1  [16][64+8]
  __half* R0 = S1 + 16 * 72;            // rgb layer 0    [64][32+8]  (columns permuted)
  __half* R1 = R0 + 64 * 40;            // rgb layer 1    [64][64+8]
  __half* R2 = R1 + 64 * 72;            // rgb layer 2    [16][64+8]
  for (int i = threadIdx.x; i < (int)(sizeof(FfGridMeta) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&sm)[i] = reinterpret_cast<const uint32_t*>(&meta)[i];
  ff_load_w(w_sigma, 64, 32, S0, false);
  ff_load_w(w_sigma + 64 * 32, 16, 64, S1, false);
  ff_load_w(w_rgb, 64, 32, R0, true);
  ff_load_w(w_rgb + 64 * 32, 64, 64, R1, false);
  ff_load_w(w_rgb + 64 * 32 + 64 * 64, 16, 64, R2, false);
  __syncthreads();
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_tiles = (n + 15) >> 4;
  for (int64_t tile = warp; tile < n_tiles; tile += n_warps) {
    const int64_t row0 = tile << 4;
    const int64_t r0 = row0 + g, r1 = r0 + 8;
    // ---- hash-grid features straight into the A fragments (levels t, t+4 -> k-block 0; t+8, t+12 -> k-block 1)
    float xn0[3] = {0.f, 0.f, 0.f}, xn1[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 3 && kGrid; ++d) {
      if (r0 < n) { const float v = x[3 * r0 + d]; xn0[d] = sm.xform_on ? __fdiv_rn(__fsub_rn(v, sm.lo[d]), sm.size3[d]) : v; }
      if (r1 < n) { const float v = x[3 * r1 + d]; xn1[d] = sm.xform_on ? __fdiv_rn(__fsub_rn(v, sm.lo[d]), sm.size3[d]) : v; }
    }
    uint32_t af[2][4];
    if (!kGrid) {      // A fragments of the 16 x 32 feature tile straight from the row-major matrix
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const int col = kb * 16 + 2 * t;
        af[kb][0] = r0 < n ? *reinterpret_cast<const uint32_t*>(feat + r0 * 32 + col) : 0u;
        af[kb][1] = r1 < n ? *reinterpret_cast<const uint32_t*>(feat + r1 * 32 + col) : 0u;
        af[kb][2] = r0 < n ? *reinterpret_cast<const uint32_t*>(feat + r0 * 32 + col + 8) : 0u;
        af[kb][3] = r1 < n ? *reinterpret_cast<const uint32_t*>(feat + r1 * 32 + col + 8) : 0u;
      }
    }
#pragma unroll
    for (int q = 0; q < 4 && kGrid; ++q) {
      const int l = t + 4 * q;
      const __half2* tl = reinterpret_cast<const __half2*>(table) + sm.offset[l];
      const float2 f0 = ff_lookup(xn0, tl, sm.scale[l], sm.res[l], sm.size[l]);
      const float2 f1 = ff_lookup(xn1, tl, sm.scale[l], sm.res[l], sm.size[l]);
      // level l -> columns 2l, 2l+1: k-block l>>3, register 0/1 (cols 2t..) for l&7 < 4, register 2/3 (cols 8+2t..) otherwise
      af[q >> 1][(q & 1) * 2 + 0] = pack_half2(f0.x, f0.y);
      af[q >> 1][(q & 1) * 2 + 1] = pack_half2(f1.x, f1.y);
    }
    if (kGrid && feat) ff_store_a<32>(feat, row0, n, af, g, t);
    // ---- density trunk
    float c[8][4];
    ff_layer<32, 64>(af, S0, c, g, t);
    uint32_t hid[4][4];
    ff_relu_pack(c, hid);
    if (sig_acts) ff_store_act(sig_acts, row0, n, hid, g, t);
    float ch[2][4];
    ff_layer<64, 16>(hid, S1, ch, g, t);
    // h (fp16, as the tcnn module returns it) is both an output and k-block 0 of the colour head's input
    uint32_t xin[2][4];
    xin[0][0] = pack_half2(ch[0][0], ch[0][1]); xin[0][1] = pack_half2(ch[0][2], ch[0][3]);
    xin[0][2] = pack_half2(ch[1][0], ch[1][1]); xin[0][3] = pack_half2(ch[1][2], ch[1][3]);
    if (t == 0) {
      if (r0 < n) sigmas[r0] = expf(__low2float(*reinterpret_cast<const __half2*>(&xin[0][0])));
      if (r1 < n) sigmas[r1] = expf(__low2float(*reinterpret_cast<const __half2*>(&xin[0][1])));
    }
    if (h_out) {
      if (r0 < n) { *reinterpret_cast<uint32_t*>(h_out + r0 * 16 + 2 * t) = xin[0][0]; *reinterpret_cast<uint32_t*>(h_out + r0 * 16 + 8 + 2 * t) = xin[0][2]; }
      if (r1 < n) { *reinterpret_cast<uint32_t*>(h_out + r1 * 16 + 2 * t) = xin[0][1]; *reinterpret_cast<uint32_t*>(h_out + r1 * 16 + 8 + 2 * t) = xin[0][3]; }
    }
    // k-block 1 = [d/|d| (3) | ones (13)]: columns 16+2t, 17+2t and 24+2t, 25+2t
    {
      float d0[3] = {0.f, 0.f, 1.f}, d1[3] = {0.f, 0.f, 1.f};
      if (r0 < n) { const float a = dirs[3 * r0], b = dirs[3 * r0 + 1], cc = dirs[3 * r0 + 2]; const float nr = sqrtf(a * a + b * b + cc * cc); d0[0] = a / nr; d0[1] = b / nr; d0[2] = cc / nr; }
      if (r1 < n) { const float a = dirs[3 * r1], b = dirs[3 * r1 + 1], cc = dirs[3 * r1 + 2]; const float nr = sqrtf(a * a + b * b + cc * cc); d1[0] = a / nr; d1[1] = b / nr; d1[2] = cc / nr; }
      const uint32_t ones = pack_half2(1.f, 1.f);
      xin[1][0] = t == 0 ? pack_half2(d0[0], d0[1]) : (t == 1 ? pack_half2(d0[2], 1.f) : ones);
      xin[1][1] = t == 0 ? pack_half2(d1[0], d1[1]) : (t == 1 ? pack_half2(d1[2], 1.f) : ones);
      xin[1][2] = ones; xin[1][3] = ones;
    }
    if (x_rgb) ff_store_a<32>(x_rgb, row0, n, xin, g, t);
    // ---- colour head
    ff_layer<32, 64>(xin, R0, c, g, t);
    ff_relu_pack(c, hid);
    if (rgb_acts) ff_store_act(rgb_acts, row0, n, hid, g, t);
    ff_layer<64, 64>(hid, R1, c, g, t);
    ff_relu_pack(c, hid);
    if (rgb_acts) ff_store_act(rgb_acts + act_rows(n_cap) * 64, row0, n, hid, g, t);
    float co[2][4];
    ff_layer<64, 16>(hid, R2, co, g, t);
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      o[2 * j] = pack_half2(1.0f / (1.0f + __expf(-co[j][0])), 1.0f / (1.0f + __expf(-co[j][1])));
      o[2 * j + 1] = pack_half2(1.0f / (1.0f + __expf(-co[j][2])), 1.0f / (1.0f + __expf(-co[j][3])));
    }
    if (rgb_out) {
      if (r0 < n) { *reinterpret_cast<uint32_t*>(rgb_out + r0 * 16 + 2 * t) = o[0]; *reinterpret_cast<uint32_t*>(rgb_out + r0 * 16 + 8 + 2 * t) = o[2]; }
      if (r1 < n) { *reinterpret_cast<uint32_t*>(rgb_out + r1 * 16 + 2 * t) = o[1]; *reinterpret_cast<uint32_t*>(rgb_out + r1 * 16 + 8 + 2 * t) = o[3]; }
    }
    // raws[:, 0:3] (fp32 of the fp16 network output): columns 0,1 by t == 0, column 2 by t == 1
    if (t == 0) {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&o[0])), b = __half22float2(*reinterpret_cast<const __half2*>(&o[1]));
      if (r0 < n) { raws[r0 * c_total] = a.x; raws[r0 * c_total + 1] = a.y; }
      if (r1 < n) { raws[r1 * c_total] = b.x; raws[r1 * c_total + 1] = b.y; }
    } else if (t == 1) {
      if (r0 < n) raws[r0 * c_total + 2] = __low2float(*reinterpret_cast<const __half2*>(&o[0]));
      if (r1 < n) raws[r1 * c_total + 2] = __low2float(*reinterpret_cast<const __half2*>(&o[1]));
    }
  }
}

// ---------------------------------------------------------------- extra heads on h (ngp_mt.py:217-224)
// sem_net / norm_net (16 -> 64 -> 64 -> n_out <= 16, ReLU hidden, no output activation) evaluated together: h is read
// once, each net's three layers are chained in registers, the hidden states go out in the tiled activation layout (only
// when a backward will need them) and the outputs land directly in their raws columns (fp32 of the fp16 network output,
// as the tcnn module returns half) - replaces 2 x (ncn_mlp_fwd + ncn_field_head_out).
struct FfHead {
  const __half* w;      // tcnn layout: [64][16] | [64][64] | [16][64]; nullptr = head absent
  __half* acts;         // (2, act_rows(n_cap), 64) tiled, or nullptr
  __half* out;          // (N,16) f16 or nullptr
  int c_off, n_ch;
};
constexpr int kHeadW = 64 * (16 + kFfPad) + 64 * (64 + kFfPad) + 16 * (64 + kFfPad);

__global__ void __launch_bounds__(kFfThreads)
field_heads_fwd_kernel(const __half* __restrict__ h, int64_t n_cap, const int32_t* __restrict__ n_dev, float* __restrict__ raws,
                       int c_total, FfHead ha, FfHead hb) {
  __shared__ __align__(16) __half sW[2 * kHeadW];
  const FfHead heads[2] = {ha, hb};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (heads[q].w == nullptr) continue;
    __half* W0 = sW + q * kHeadW;
    ff_load_w(heads[q].w, 64, 16, W0, false);
    ff_load_w(heads[q].w + 64 * 16, 64, 64, W0 + 64 * (16 + kFfPad), false);
    ff_load_w(heads[q].w + 64 * 16 + 64 * 64, 16, 64, W0 + 64 * (16 + kFfPad) + 64 * (64 + kFfPad), false);
  }
  __syncthreads();
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_tiles = (n + 15) >> 4;
  for (int64_t tile = warp; tile < n_tiles; tile += n_warps) {
    const int64_t row0 = tile << 4;
    const int64_t r0 = row0 + g, r1 = r0 + 8;
    uint32_t ah[1][4];
    ah[0][0] = r0 < n ? *reinterpret_cast<const uint32_t*>(h + r0 * 16 + 2 * t) : 0u;
    ah[0][1] = r1 < n ? *reinterpret_cast<const uint32_t*>(h + r1 * 16 + 2 * t) : 0u;
    ah[0][2] = r0 < n ? *reinterpret_cast<const uint32_t*>(h + r0 * 16 + 8 + 2 * t) : 0u;
    ah[0][3] = r1 < n ? *reinterpret_cast<const uint32_t*>(h + r1 * 16 + 8 + 2 * t) : 0u;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const FfHead& hd = heads[q];
      if (hd.w == nullptr) continue;
      const __half* W0 = sW + q * kHeadW;
      const __half* W1 = W0 + 64 * (16 + kFfPad);
      const __half* W2 = W1 + 64 * (64 + kFfPad);
      float c[8][4];
      uint32_t hid[4][4];
      ff_layer<16, 64>(ah, W0, c, g, t);
      ff_relu_pack(c, hid);
      if (hd.acts) ff_store_act(hd.acts, row0, n, hid, g, t);
      ff_layer<64, 64>(hid, W1, c, g, t);
      ff_relu_pack(c, hid);
      if (hd.acts) ff_store_act(hd.acts + act_rows(n_cap) * 64, row0, n, hid, g, t);
      float co[2][4];
      ff_layer<64, 16>(hid, W2, co, g, t);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = j * 8 + 2 * t;
        const uint32_t p0 = pack_half2(co[j][0], co[j][1]), p1 = pack_half2(co[j][2], co[j][3]);
        if (hd.out) {
          if (r0 < n) *reinterpret_cast<uint32_t*>(hd.out + r0 * 16 + col) = p0;
          if (r1 < n) *reinterpret_cast<uint32_t*>(hd.out + r1 * 16 + col) = p1;
        }
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&p0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&p1));
        if (col < hd.n_ch) {
          if (r0 < n) raws[r0 * c_total + hd.c_off + col] = f0.x;
          if (r1 < n) raws[r1 * c_total + hd.c_off + col] = f1.x;
        }
        if (col + 1 < hd.n_ch) {
          if (r0 < n) raws[r0 * c_total + hd.c_off + col + 1] = f0.y;
          if (r1 < n) raws[r1 * c_total + hd.c_off + col + 1] = f1.y;
        }
      }
    }
  }
}

}  // namespace ncn

using namespace ncn;

extern "C" int ncn_field_heads_fwd(const void* h_f16, int64_t n, const int32_t* n_dev, float* raws, int c_total,
                                   const void* w_a_f16, int c_off_a, int n_ch_a, void* acts_a_f16, void* out_a_f16,
                                   const void* w_b_f16, int c_off_b, int n_ch_b, void* acts_b_f16, void* out_b_f16,
                                   ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0 && c_total >= 1);
  if (n == 0 || (w_a_f16 == nullptr && w_b_f16 == nullptr)) return NCN_OK;
  NCN_CHECK_PTR(h_f16); NCN_CHECK_PTR(raws);
  if (w_a_f16) NCN_CHECK_SIZE(n_ch_a >= 1 && n_ch_a <= 16 && c_off_a >= 0 && c_off_a + n_ch_a <= c_total);
  if (w_b_f16) NCN_CHECK_SIZE(n_ch_b >= 1 && n_ch_b <= 16 && c_off_b >= 0 && c_off_b + n_ch_b <= c_total);
  FfHead ha{(const __half*)w_a_f16, (__half*)acts_a_f16, (__half*)out_a_f16, c_off_a, n_ch_a};
  FfHead hb{(const __half*)w_b_f16, (__half*)acts_b_f16, (__half*)out_b_f16, c_off_b, n_ch_b};
  const int grid = persistent_grid(((n + 15) / 16) * 32, kFfThreads, 6);
  field_heads_fwd_kernel<<<grid, kFfThreads, 0, as_stream(stream)>>>((const __half*)h_f16, n, n_dev, raws, c_total, ha, hb);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_field_fwd(const ncn_grid_desc* desc, const float* x, const float* dirs, const void* table_f16,
                             const void* w_sigma_f16, const void* w_rgb_f16, int64_t n, const int32_t* n_dev,
                             const float* xform_host, float* sigmas, float* raws, int c_total, void* feat_f16, void* h_f16,
                             void* sig_acts_f16, void* x_rgb_f16, void* rgb_acts_f16, void* rgb_out_f16, ncn_stream_t stream) {
  NCN_CHECK_PTR(desc);
  if (desc->n_levels != 16 || desc->n_features != 2) return NCN_E_CONFIG;
  NCN_CHECK_SIZE(n >= 0 && c_total >= 3);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(x); NCN_CHECK_PTR(dirs); NCN_CHECK_PTR(table_f16); NCN_CHECK_PTR(w_sigma_f16); NCN_CHECK_PTR(w_rgb_f16);
  NCN_CHECK_PTR(sigmas); NCN_CHECK_PTR(raws);
  FfGridMeta m;
  for (int l = 0; l < 16; ++l) { m.scale[l] = desc->level_scale[l]; m.res[l] = desc->level_res[l]; m.size[l] = desc->level_size[l]; m.offset[l] = desc->level_offset[l]; }
  for (int d = 0; d < 3; ++d) { m.lo[d] = xform_host ? xform_host[d] : 0.f; m.size3[d] = xform_host ? xform_host[3 + d] : 1.f; }
  m.inv_size_dummy = 0.f;
  m.xform_on = xform_host != nullptr;
  const int grid = persistent_grid(((n + 15) / 16) * 32, kFfThreads, 6);
  field_fwd_kernel<true><<<grid, kFfThreads, 0, as_stream(stream)>>>(m, x, dirs, (const __half*)table_f16, (const __half*)w_sigma_f16,
                                                              (const __half*)w_rgb_f16, n, n_dev, sigmas, raws, c_total, (__half*)feat_f16,
                                                              (__half*)h_f16, (__half*)sig_acts_f16, (__half*)x_rgb_f16,
                                                              (__half*)rgb_acts_f16, (__half*)rgb_out_f16);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

// density trunk + colour head in one launch on precomputed hash-grid features (replaces ncn_mlp_fwd(sigma) ->
// ncn_field_prepare_rgb -> ncn_mlp_fwd(rgb) -> ncn_field_head_out; same outputs as ncn_field_fwd, x_rgb in [h | d | 1] order)
// implementation of ncn_field_mlp_fwd: 1 = tcgen05 / TMEM (field_tc05.cu, default), 0 = warp MMA (this file); developer A/B knob
int ncn_field_mlp_fwd_tc05_try(const void* feat_f16, const float* dirs, const void* w_sigma_f16, const void* w_rgb_f16, int64_t n,
                               const int32_t* n_dev, float* sigmas, float* raws, int c_total, void* h_f16, void* sig_acts_f16,
                               void* x_rgb_f16, void* rgb_acts_f16, void* rgb_out_f16, cudaStream_t st);
static int g_field_fwd_impl = 1;
extern "C" int ncn_set_field_fwd_impl(int impl) { const int old = g_field_fwd_impl; g_field_fwd_impl = impl; return old; }

extern "C" int ncn_field_mlp_fwd(const void* feat_f16, const float* dirs, const void* w_sigma_f16, const void* w_rgb_f16, int64_t n,
                                 const int32_t* n_dev, float* sigmas, float* raws, int c_total, void* h_f16, void* sig_acts_f16,
                                 void* x_rgb_f16, void* rgb_acts_f16, void* rgb_out_f16, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0 && c_total >= 3);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(feat_f16); NCN_CHECK_PTR(dirs); NCN_CHECK_PTR(w_sigma_f16); NCN_CHECK_PTR(w_rgb_f16); NCN_CHECK_PTR(sigmas); NCN_CHECK_PTR(raws);
  if (g_field_fwd_impl >= 1) {
    const int rc = ncn_field_mlp_fwd_tc05_try(feat_f16, dirs, w_sigma_f16, w_rgb_f16, n, n_dev, sigmas, raws, c_total, h_f16, sig_acts_f16,
                                              x_rgb_f16, rgb_acts_f16, rgb_out_f16, as_stream(stream));
    if (rc != NCN_E_UNSUPPORTED) return rc;
  }
  FfGridMeta m = {};
  const int grid = persistent_grid(((n + 15) / 16) * 32, kFfThreads, 6);
  field_fwd_kernel<false><<<grid, kFfThreads, 0, as_stream(stream)>>>(m, nullptr, dirs, nullptr, (const __half*)w_sigma_f16,
                                                                     (const __half*)w_rgb_f16, n, n_dev, sigmas, raws, c_total,
                                                                     (__half*)feat_f16, (__half*)h_f16, (__half*)sig_acts_f16,
                                                                     (__half*)x_rgb_f16, (__half*)rgb_acts_f16, (__half*)rgb_out_f16);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
