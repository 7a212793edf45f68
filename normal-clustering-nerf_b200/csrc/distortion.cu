// Mip-NeRF-360 distortion loss and the segmented row sum used by RayMarcher.backward.
// Replaces losses.cu:7-172 of the reference (vren.distortion_loss_fw/_bw) and
// torch_scatter.segment_csr (models/custom_functions.py:107-110).
//
// The loss is off by default in the reference (opt.py:68 loss_distortion_w=0), so this is
// kept as the reference's own evaluation order - one thread walks a ray's samples - which
// makes the forward bit-exact: the reference's prefix sums are sequential in-thread
// thrust scans and its elementwise combination is a chain of individually rounded torch
// ops (wts=ws*ts; 2*(wts_inc*ws_exc - ws_inc*wts_exc) + (1/3*ws)*ws*deltas).
#include "ncn_common.cuh"

namespace ncn {

__global__ void __launch_bounds__(128)
distortion_fw_kernel(const float* __restrict__ ws, const float* __restrict__ deltas, const float* __restrict__ ts,
                     const int64_t* __restrict__ rays_a, int64_t n_rays, float* __restrict__ loss,
                     float* __restrict__ ws_inc, float* __restrict__ wts_inc) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_rays) return;
  const int64_t ray_idx = rays_a[3 * n], start = rays_a[3 * n + 1];
  const int N = (int)rays_a[3 * n + 2];
  float wi = 0.f, wti = 0.f, acc = 0.f;
  const float third = 1.0f / 3;
  for (int k = 0; k < N; ++k) {
    const int64_t s = start + k;
    const float w = ws[s];
    const float wt = __fmul_rn(w, ts[s]);
    const float we = wi, wte = wti;                       // exclusive sums
    wi = k == 0 ? w : __fadd_rn(wi, w);                   // inclusive sums
    wti = k == 0 ? wt : __fadd_rn(wti, wt);
    ws_inc[s] = wi; wts_inc[s] = wti;
    const float cross = __fmul_rn(2.0f, __fsub_rn(__fmul_rn(wti, we), __fmul_rn(wi, wte)));
    const float self = __fmul_rn(__fmul_rn(__fmul_rn(third, w), w), deltas[s]);
    acc = __fadd_rn(acc, __fadd_rn(cross, self));
  }
  loss[ray_idx] = acc;
}

__global__ void __launch_bounds__(128)
distortion_bw_kernel(const float* __restrict__ dL_dloss, const float* __restrict__ ws_inc,
                     const float* __restrict__ wts_inc, const float* __restrict__ ws,
                     const float* __restrict__ deltas, const float* __restrict__ ts,
                     const int64_t* __restrict__ rays_a, int64_t n_rays, float* __restrict__ dL_dws) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_rays) return;
  const int64_t ray_idx = rays_a[3 * n], start = rays_a[3 * n + 1];
  const int N = (int)rays_a[3 * n + 2];
  if (N == 0) return;
  const int64_t end = start + N - 1;
  const float ws_sum = ws_inc[end], wts_sum = wts_inc[end];
  const float g = dL_dloss[ray_idx];
  for (int64_t s = start; s <= end; ++s) {
    const float t = ts[s];
    const float before = s == start ? 0.f : (t * ws_inc[s - 1] - wts_inc[s - 1]);
    const float after = wts_sum - wts_inc[s] - t * (ws_sum - ws_inc[s]);
    float v = g * 2 * (before + after);
    v += g * (2.0f / 3) * ws[s] * deltas[s];
    dL_dws[s] = v;
  }
}

// out[seg][d] = sum_{row in [indptr[seg], indptr[seg+1])} src[row][d] ; one warp per segment
__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ src, const int64_t* __restrict__ indptr, int64_t n_seg, int D,
                   float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t seg = warp; seg < n_seg; seg += n_warps) {
    const int64_t b = indptr[seg], e = indptr[seg + 1];
    for (int d = 0; d < D; ++d) {
      float acc = 0.f;
      for (int64_t r = b + lane; r < e; r += 32) acc += src[r * D + d];
      acc = warp_sum(acc);
      if (lane == 0) out[seg * D + d] = acc;
    }
  }
}

}  // namespace ncn

using namespace ncn;

extern "C" int ncn_distortion_fw(const float* ws, const float* deltas, const float* ts, const int64_t* rays_a,
                                 int64_t n_rays, int64_t n_samples, float* loss, float* ws_inclusive,
                                 float* wts_inclusive, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_rays >= 0 && n_samples >= 0);
  if (n_rays == 0) return NCN_OK;
  NCN_CHECK_PTR(rays_a); NCN_CHECK_PTR(loss);
  if (n_samples > 0) { NCN_CHECK_PTR(ws); NCN_CHECK_PTR(deltas); NCN_CHECK_PTR(ts); NCN_CHECK_PTR(ws_inclusive); NCN_CHECK_PTR(wts_inclusive); }
  distortion_fw_kernel<<<(unsigned)ceil_div(n_rays, 128), 128, 0, as_stream(stream)>>>(ws, deltas, ts, rays_a, n_rays, loss,
                                                                                   ws_inclusive, wts_inclusive);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_distortion_bw(const float* dL_dloss, const float* ws_inclusive, const float* wts_inclusive,
                                 const float* ws, const float* deltas, const float* ts, const int64_t* rays_a,
                                 int64_t n_rays, int64_t n_samples, float* dL_dws, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_rays >= 0 && n_samples >= 0);
  if (n_rays == 0 || n_samples == 0) return NCN_OK;
  NCN_CHECK_PTR(dL_dloss); NCN_CHECK_PTR(ws_inclusive); NCN_CHECK_PTR(wts_inclusive); NCN_CHECK_PTR(ws);
  NCN_CHECK_PTR(deltas); NCN_CHECK_PTR(ts); NCN_CHECK_PTR(rays_a); NCN_CHECK_PTR(dL_dws);
  distortion_bw_kernel<<<(unsigned)ceil_div(n_rays, 128), 128, 0, as_stream(stream)>>>(dL_dloss, ws_inclusive, wts_inclusive,
                                                                                   ws, deltas, ts, rays_a, n_rays, dL_dws);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_segment_csr_sum(const float* src, const int64_t* indptr, int64_t n_segments, int dim, float* out,
                                   ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_segments >= 0 && dim >= 1);
  if (n_segments == 0) return NCN_OK;
  NCN_CHECK_PTR(indptr); NCN_CHECK_PTR(out);
  const int grid = persistent_grid(n_segments * 32, 256, 8);
  segment_sum_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, indptr, n_segments, dim, out);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
