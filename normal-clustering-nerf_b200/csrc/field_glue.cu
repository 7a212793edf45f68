// Elementwise glue of the NGPMT field (models/ngp_mt.py:157-229) between the encoder / MLP kernels: what the
// reference does with ~20 small torch launches per step (x normalisation is folded into the grid kernel):
//   prepare_rgb : d/||d||, cat[d, h] (19 -> padded to 32 with ones, the tcnn Identity-encoding padding),
//                 sigma = TruncExp(h[:,0])  (custom_functions.py:162-173)
//   head_out    : a head's fp16 output columns -> the fp32 `raws` matrix fed to the compositor (rendering.py:203-212)
//   head_dout   : dL/draws columns -> the head's (padded) fp16 dL/dout, times the loss scale
//   bwd_h       : dL/dh = dL/dx_rgb[:, 3:19] (+ other heads' dL/dx) + e_0 * dL/dsigma * exp(clamp(h0,-15,15))
// All take a device-side live row count (n_dev) so the training step never synchronises.
#include "ncn_common.cuh"

namespace ncn {

__device__ __forceinline__ int64_t live_rows(int64_t n_cap, const int32_t* __restrict__ n_dev) {
  int64_t n = n_cap;
  if (n_dev != nullptr) { const int64_t nd = *n_dev; if (nd < n) n = nd; }
  return n;
}

// thread per sample: two 16-byte loads of h, the direction, four 16-byte stores of the 32-wide row
//   x = [d/|d| (3) | h (16) | 1.0 x 13]:  32-bit words X0=(d0,d1) X1=(d2,h0) X_j=(h_{2j-3},h_{2j-2}) j=2..8, X9=(h15,1), X10..15=(1,1)
__global__ void __launch_bounds__(256)
prepare_rgb_kernel(const float* __restrict__ dirs, const __half* __restrict__ h, int64_t n_cap,
                   const int32_t* __restrict__ n_dev, __half* __restrict__ x_rgb, float* __restrict__ sigmas) {
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += stride) {
    const uint4 ha = *reinterpret_cast<const uint4*>(h + s * 16), hb = *reinterpret_cast<const uint4*>(h + s * 16 + 8);
    const float dx = dirs[3 * s], dy = dirs[3 * s + 1], dz = dirs[3 * s + 2];
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    const uint32_t w[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
    const __half2 d01 = __floats2half2_rn(dx / nrm, dy / nrm);
    const __half d2 = __float2half_rn(dz / nrm);
    uint32_t X[16];
    X[0] = *reinterpret_cast<const uint32_t*>(&d01);
    X[1] = (uint32_t)__half_as_ushort(d2) | (w[0] << 16);
#pragma unroll
    for (int q = 2; q <= 8; ++q) X[q] = __funnelshift_r(w[q - 2], w[q - 1], 16);
    X[9] = (w[7] >> 16) | 0x3C000000u;                       // (h15, 1.0)
#pragma unroll
    for (int q = 10; q < 16; ++q) X[q] = 0x3C003C00u;        // (1.0, 1.0)
    uint4* o = reinterpret_cast<uint4*>(x_rgb + s * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = make_uint4(X[4 * q], X[4 * q + 1], X[4 * q + 2], X[4 * q + 3]);
    if (sigmas) sigmas[s] = expf(__half2float(__ushort_as_half((unsigned short)(w[0] & 0xFFFFu))));
  }
}

// thread per sample: the first n_ch (<= 8) halfs of the 16-byte-aligned head row -> fp32 columns of raws
__global__ void __launch_bounds__(256)
head_out_kernel(const __half* __restrict__ out, int out_pad, int64_t n_cap, const int32_t* __restrict__ n_dev,
                float* __restrict__ raws, int c_total, int c_offset, int n_ch) {
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += stride) {
    float* r = raws + s * c_total + c_offset;
    if (n_ch <= 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(out + s * out_pad);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
        if (2 * q < n_ch) r[2 * q] = f.x;
        if (2 * q + 1 < n_ch) r[2 * q + 1] = f.y;
      }
    } else {
      for (int j = 0; j < n_ch; ++j) r[j] = __half2float(out[s * out_pad + j]);
    }
  }
}

__global__ void __launch_bounds__(256)
head_dout_kernel(const float* __restrict__ d_raws, int c_total, int c_offset, int n_ch, float scale, int64_t n_cap,
                 const int32_t* __restrict__ n_dev, __half* __restrict__ dout, int out_pad) {
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t total = n * out_pad, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t s = i / out_pad;
    const int j = (int)(i - s * out_pad);
    dout[i] = __float2half_rn(j < n_ch ? d_raws[s * c_total + c_offset + j] * scale : 0.f);
  }
}

__global__ void __launch_bounds__(256)
bwd_h_kernel(const __half* __restrict__ dx_rgb, const float* __restrict__ d_sigmas, const __half* __restrict__ h,
             const __half* __restrict__ dx_a, const __half* __restrict__ dx_b, float scale, int64_t n_cap,
             const int32_t* __restrict__ n_dev, __half* __restrict__ dh) {
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t total = n * 16, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t s = i >> 4;
    const int j = (int)(i & 15);
    float g = __half2float(dx_rgb[s * 32 + 3 + j]);
    if (dx_a) g += __half2float(dx_a[i]);
    if (dx_b) g += __half2float(dx_b[i]);
    if (j == 0) {
      const float h0 = __half2float(h[s * 16]);
      g += d_sigmas[s] * expf(fminf(fmaxf(h0, -15.f), 15.f)) * scale;
    }
    dh[i] = __float2half_rn(g);
  }
}

// rays from (image index, pixel index): rays_d = directions[pix] @ c2w[img][:, :3]^T, rays_o = c2w[img][:, 3]
// (NeRFSystem.forward gather + get_rays, train_nerf.py:167-182, datasets/ray_utils.py:46-71) in one launch
__global__ void __launch_bounds__(256)
rays_kernel(const float* __restrict__ poses, const float* __restrict__ directions, const int64_t* __restrict__ img_idx,
            const int64_t* __restrict__ pix_idx, int64_t n, float* __restrict__ rays_o, float* __restrict__ rays_d) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float* c2w = poses + img_idx[i] * 12;
    const float* d = directions + pix_idx[i] * 3;
    const float dx = d[0], dy = d[1], dz = d[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      // matmul accumulation order of a (1x3)@(3x3) product: ((d0*R[r][0] + d1*R[r][1]) + d2*R[r][2])
      rays_d[3 * i + r] = __fmaf_rn(dz, c2w[4 * r + 2], __fmaf_rn(dy, c2w[4 * r + 1], __fmul_rn(dx, c2w[4 * r])));
      rays_o[3 * i + r] = c2w[4 * r + 3];
    }
  }
}

}  // namespace ncn

using namespace ncn;

extern "C" int ncn_rays_from_pixels(const float* poses, const float* directions, const int64_t* img_idx, const int64_t* pix_idx,
                                    int64_t n, float* rays_o, float* rays_d, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(poses); NCN_CHECK_PTR(directions); NCN_CHECK_PTR(img_idx); NCN_CHECK_PTR(pix_idx); NCN_CHECK_PTR(rays_o); NCN_CHECK_PTR(rays_d);
  rays_kernel<<<persistent_grid(n, 256, 8), 256, 0, as_stream(stream)>>>(poses, directions, img_idx, pix_idx, n, rays_o, rays_d);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_field_prepare_rgb(const float* dirs, const void* h_f16, int64_t n, const int32_t* n_dev, void* x_rgb_f16,
                                     float* sigmas, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(dirs); NCN_CHECK_PTR(h_f16); NCN_CHECK_PTR(x_rgb_f16);
  if (((uintptr_t)h_f16 | (uintptr_t)x_rgb_f16) & 15) return NCN_E_ALIGN;
  prepare_rgb_kernel<<<persistent_grid(n, 256, 8), 256, 0, as_stream(stream)>>>(dirs, (const __half*)h_f16, n, n_dev,
                                                                                     (__half*)x_rgb_f16, sigmas);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_field_head_out(const void* out_f16, int out_pad, int64_t n, const int32_t* n_dev, float* raws, int c_total,
                                  int c_offset, int n_ch, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0 && n_ch >= 1 && c_offset >= 0 && c_offset + n_ch <= c_total && n_ch <= out_pad);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(out_f16); NCN_CHECK_PTR(raws);
  if (((uintptr_t)out_f16 & 15) || (out_pad & 7)) return NCN_E_ALIGN;
  head_out_kernel<<<persistent_grid(n, 256, 8), 256, 0, as_stream(stream)>>>((const __half*)out_f16, out_pad, n, n_dev, raws,
                                                                                    c_total, c_offset, n_ch);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_field_head_dout(const float* dL_draws, int c_total, int c_offset, int n_ch, float scale, int64_t n,
                                   const int32_t* n_dev, void* dout_f16, int out_pad, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0 && n_ch >= 1 && c_offset >= 0 && c_offset + n_ch <= c_total && n_ch <= out_pad);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(dL_draws); NCN_CHECK_PTR(dout_f16);
  head_dout_kernel<<<persistent_grid(n * out_pad, 256, 8), 256, 0, as_stream(stream)>>>(dL_draws, c_total, c_offset, n_ch, scale, n,
                                                                                        n_dev, (__half*)dout_f16, out_pad);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_field_bwd_h(const void* dx_rgb_f16, const float* dL_dsigmas, const void* h_f16, const void* dx_a_f16,
                               const void* dx_b_f16, float scale, int64_t n, const int32_t* n_dev, void* dh_f16,
                               ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(dx_rgb_f16); NCN_CHECK_PTR(dL_dsigmas); NCN_CHECK_PTR(h_f16); NCN_CHECK_PTR(dh_f16);
  bwd_h_kernel<<<persistent_grid(n * 16, 256, 8), 256, 0, as_stream(stream)>>>((const __half*)dx_rgb_f16, dL_dsigmas,
                                                                               (const __half*)h_f16, (const __half*)dx_a_f16,
                                                                               (const __half*)dx_b_f16, scale, n, n_dev, (__half*)dh_f16);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
