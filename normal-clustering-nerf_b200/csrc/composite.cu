// Front-to-back volume compositing.  Replaces volumerendering.cu:16-585 of the reference
// (vren.composite_train_fw/_multi_fw, composite_train_bw/_multi_bw, composite_test_fw/_multi_fw).
//
// B200 design.  The reference gives one thread a whole ray: strided (uncoalesced) sample
// reads, a global read-modify-write of rend[ray][c] per sample and channel, an in-thread
// sequential thrust scan and a zero-filled (N,C) scratch in the backward pass.  Here one
// WARP owns a ray: lanes load 32 consecutive samples with coalesced 128 B requests,
// evaluate alpha = 1 - __expf(-sigma*delta) in parallel, and the transmittance product is
// replayed in the reference's exact sequential order with register shuffles (T, ws and the
// early-termination sample count are therefore bit-exact); the ray sums (opacity, depth,
// rend) are warp-tree reductions kept in registers (fp32 rounding differs from the
// reference's sequential sum by a few ulp - tolerance stated in tests).  The backward
// pass needs ONE inclusive warp scan per 32 samples: the per-channel running sums of the
// reference are linear in dL/drend and collapse to a single scanned quantity
//   q_k = w_k (dL/ddepth t_k + <dL/drend, raw_k>) + dL/dws_k ws_k.
// Algorithmic traffic: fwd 16+4C B/sample + 24+(16+4C) B/ray, bwd 24+8C B/sample.
#include "ncn_common.cuh"

namespace ncn {

constexpr int kMaxChanWords = 2;   // supports up to 64 render channels

__device__ __forceinline__ float warp_scan_incl_f(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Replays T *= (1-a_j) for j = 0..31 in order (uniform across the warp; lanes past the ray's end carry 1-a = 1, which
// leaves T bit-identical).  Fully unrolled: the 32 broadcasts are independent of the running product and pipeline, so the
// serial part is the 32 dependent multiplies only (the rolled loop with an early break exposed a shuffle latency per
// sample, which made the long rays - hundreds of samples - the tail of the whole kernel).
// in : om = 1-a of this lane's sample, T = transmittance before the chunk
// out: T_before for this lane's sample, T = transmittance after the chunk (unused once a stop was found),
//      returns index of the terminating sample within the chunk or -1.
__device__ __forceinline__ int replay_transmittance(float om, int cnt, float thr, float& T, float& T_before, int lane) {
  (void)cnt;
  int stop = -1;
  T_before = T;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float omj = __shfl_sync(0xffffffffu, om, j);
    if (lane == j) T_before = T;
    T = __fmul_rn(T, omj);
    if (stop < 0 && T <= thr) stop = j;
  }
  return stop;
}

__global__ void __launch_bounds__(256)
composite_train_fw_kernel(const float* __restrict__ sigmas, const float* __restrict__ raws,
                          const float* __restrict__ deltas, const float* __restrict__ ts,
                          const int64_t* __restrict__ rays_a, float thr, int64_t n_rays, int64_t capacity, int C,
                          int64_t* __restrict__ total_samples, float* __restrict__ opacity,
                          float* __restrict__ depth, float* __restrict__ rend, float* __restrict__ ws) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = warp; n < n_rays; n += n_warps) {
    const int64_t ray_idx = rays_a[3 * n], start = rays_a[3 * n + 1];
    int64_t N64 = rays_a[3 * n + 2];
    if (start + N64 > capacity) N64 = capacity > start ? capacity - start : 0;
    const int N = (int)N64;
    float T = 1.0f, acc_o = 0.f, acc_d = 0.f;
    float acc_r[kMaxChanWords] = {0.f, 0.f};
    int samples = N;       // total_samples: #samples composited before termination
    bool dead = false;
    for (int base = 0; base < N; base += 32) {
      const int k = base + lane;
      const int cnt = min(32, N - base);
      const int64_t s = start + k;
      if (dead) { if (k < N) ws[s] = 0.f; continue; }
      float a = 0.f, t = 0.f;
      if (k < N) {
        a = __fsub_rn(1.0f, __expf(__fmul_rn(-sigmas[s], deltas[s])));
        t = ts[s];
      }
      float T_before;
      const int stop = replay_transmittance(__fsub_rn(1.0f, a), cnt, thr, T, T_before, lane);
      const bool active = (k < N) && (stop < 0 || lane <= stop);
      const float w = active ? __fmul_rn(a, T_before) : 0.f;
      if (k < N) ws[s] = w;
      acc_o += warp_sum(w);
      acc_d += warp_sum(w * t);
      for (int c = 0; c < C; ++c) {
        const float v = warp_sum(active ? w * raws[s * C + c] : 0.f);
        if (lane == (c & 31)) acc_r[c >> 5] += v;
      }
      if (stop >= 0) { samples = base + stop; dead = true; }
    }
    if (lane == 0) { total_samples[ray_idx] = samples; opacity[ray_idx] = acc_o; depth[ray_idx] = acc_d; }
    for (int c = lane; c < C; c += 32) rend[ray_idx * C + c] = acc_r[c >> 5];
  }
}

// Fast path for a compile-time channel count: every load of a chunk (sigma, delta, t, CT raws) is issued in ONE round,
// and the next chunk's round is issued before the current chunk is processed, so a ray costs one exposed memory latency
// plus ~300 clocks per 32 samples instead of (2 + CT) dependent latencies per chunk.
template <int CT>
__global__ void __launch_bounds__(256)
composite_train_fw_ct_kernel(const float* __restrict__ sigmas, const float* __restrict__ raws,
                             const float* __restrict__ deltas, const float* __restrict__ ts,
                             const int64_t* __restrict__ rays_a, float thr, int64_t n_rays, int64_t capacity,
                             int64_t* __restrict__ total_samples, float* __restrict__ opacity,
                             float* __restrict__ depth, float* __restrict__ rend, float* __restrict__ ws) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = warp; n < n_rays; n += n_warps) {
    const int64_t ray_idx = rays_a[3 * n], start = rays_a[3 * n + 1];
    int64_t N64 = rays_a[3 * n + 2];
    if (start + N64 > capacity) N64 = capacity > start ? capacity - start : 0;
    const int N = (int)N64;
    float T = 1.0f, acc_o = 0.f, acc_d = 0.f;
    float acc_r[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) acc_r[c] = 0.f;
    int samples = N;
    bool dead = false;
    float sg = 0.f, dl = 0.f, tt = 0.f, rw[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) rw[c] = 0.f;
    auto fetch = [&](int base, float& o_sg, float& o_dl, float& o_tt, float (&o_rw)[CT]) {
      const int k = base + lane;
      o_sg = 0.f; o_dl = 0.f; o_tt = 0.f;
#pragma unroll
      for (int c = 0; c < CT; ++c) o_rw[c] = 0.f;
      if (k < N) {
        const int64_t s = start + k;
        o_sg = sigmas[s]; o_dl = deltas[s]; o_tt = ts[s];
#pragma unroll
        for (int c = 0; c < CT; ++c) o_rw[c] = raws[s * CT + c];
      }
    };
    if (N > 0) fetch(0, sg, dl, tt, rw);
    for (int base = 0; base < N; base += 32) {
      const int k = base + lane;
      const int64_t s = start + k;
      if (dead) { if (k < N) ws[s] = 0.f; continue; }
      float n_sg = 0.f, n_dl = 0.f, n_tt = 0.f, n_rw[CT];
#pragma unroll
      for (int c = 0; c < CT; ++c) n_rw[c] = 0.f;
      if (base + 32 < N) fetch(base + 32, n_sg, n_dl, n_tt, n_rw);
      const float a = k < N ? __fsub_rn(1.0f, __expf(__fmul_rn(-sg, dl))) : 0.f;
      float T_before;
      const int stop = replay_transmittance(__fsub_rn(1.0f, a), 32, thr, T, T_before, lane);
      const bool active = (k < N) && (stop < 0 || lane <= stop);
      const float w = active ? __fmul_rn(a, T_before) : 0.f;
      if (k < N) ws[s] = w;
      acc_o += warp_sum(w);
      acc_d += warp_sum(w * tt);
#pragma unroll
      for (int c = 0; c < CT; ++c) acc_r[c] += warp_sum(active ? w * rw[c] : 0.f);
      if (stop >= 0) { samples = base + stop; dead = true; }
      sg = n_sg; dl = n_dl; tt = n_tt;
#pragma unroll
      for (int c = 0; c < CT; ++c) rw[c] = n_rw[c];
    }
    if (lane == 0) {
      total_samples[ray_idx] = samples; opacity[ray_idx] = acc_o; depth[ray_idx] = acc_d;
#pragma unroll
      for (int c = 0; c < CT; ++c) rend[ray_idx * CT + c] = acc_r[c];
    }
  }
}

__global__ void __launch_bounds__(256)
composite_train_bw_kernel(const float* __restrict__ dL_dopacity, const float* __restrict__ dL_ddepth,
                          const float* __restrict__ dL_drend, const float* __restrict__ dL_dws,
                          const float* __restrict__ sigmas, const float* __restrict__ raws,
                          const float* __restrict__ ws, const float* __restrict__ deltas,
                          const float* __restrict__ ts, const int64_t* __restrict__ rays_a,
                          const float* __restrict__ opacity, const float* __restrict__ depth,
                          const float* __restrict__ rend, float thr, int64_t n_rays, int64_t capacity, int C,
                          float* __restrict__ dL_dsigmas, float* __restrict__ dL_draws) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = warp; n < n_rays; n += n_warps) {
    const int64_t ray_idx = rays_a[3 * n], start = rays_a[3 * n + 1];
    int64_t N64 = rays_a[3 * n + 2];
    if (start + N64 > capacity) N64 = capacity > start ? capacity - start : 0;
    const int N = (int)N64;
    if (N == 0) continue;
    const float gO = dL_dopacity ? dL_dopacity[ray_idx] : 0.f;
    const float gD = dL_ddepth ? dL_ddepth[ray_idx] : 0.f;
    const float* gR = dL_drend ? dL_drend + ray_idx * C : nullptr;
    const float O = opacity[ray_idx], D = depth[ray_idx];
    // Q_total = gD*D + <gR, REND> + sum_k dL_dws_k ws_k
    float part = 0.f;
    if (gR) for (int c = lane; c < C; c += 32) part += gR[c] * rend[ray_idx * C + c];
    if (dL_dws) for (int k = lane; k < N; k += 32) part += dL_dws[start + k] * ws[start + k];
    const float Q_total = warp_sum(part) + gD * D;
    const float gO_term = gO * (1.0f - O);

    float T = 1.0f, carry = 0.f;
    bool dead = false;
    for (int base = 0; base < N; base += 32) {
      const int k = base + lane;
      const int cnt = min(32, N - base);
      const int64_t s = start + k;
      if (dead) {
        if (k < N) { if (dL_dsigmas) dL_dsigmas[s] = 0.f; if (dL_draws) for (int c = 0; c < C; ++c) dL_draws[s * C + c] = 0.f; }
        continue;
      }
      float a = 0.f, t = 0.f, delta = 0.f, g = 0.f, gw = 0.f;
      if (k < N) {
        delta = deltas[s];
        a = __fsub_rn(1.0f, __expf(__fmul_rn(-sigmas[s], delta)));
        t = ts[s];
        if (gR) for (int c = 0; c < C; ++c) g += gR[c] * raws[s * C + c];
        if (dL_dws) gw = dL_dws[s];
      }
      const float om = __fsub_rn(1.0f, a);
      float T_before;
      const int stop = replay_transmittance(om, cnt, thr, T, T_before, lane);
      const bool active = (k < N) && (stop < 0 || lane <= stop);
      const float w = active ? __fmul_rn(a, T_before) : 0.f;
      const float T_after = __fmul_rn(T_before, om);
      const float lin = gD * t + g;                       // d(contribution)/dw of depth + rend
      const float q = active ? (w * lin + gw * w) : 0.f;
      const float incl = warp_scan_incl_f(q, lane) + carry;
      if (k < N) {
        if (dL_dsigmas) dL_dsigmas[s] = active ? delta * (gO_term + T_after * (lin + gw) - (Q_total - incl)) : 0.f;
        if (dL_draws) for (int c = 0; c < C; ++c) dL_draws[s * C + c] = (active && gR) ? gR[c] * w : 0.f;
      }
      carry = __shfl_sync(0xffffffffu, incl, 31);
      if (stop >= 0) dead = true;
    }
  }
}

// Backward fast path for a compile-time channel count: the first chunk's sample loads are issued together with the ray-level
// loads (they only need `start`), and each later chunk is fetched one chunk ahead - a ray costs two exposed memory
// latencies instead of three plus one per chunk.  Same arithmetic, same order as composite_train_bw_kernel.
template <int CT>
__global__ void __launch_bounds__(256)
composite_train_bw_ct_kernel(const float* __restrict__ dL_dopacity, const float* __restrict__ dL_ddepth,
                             const float* __restrict__ dL_drend, const float* __restrict__ dL_dws,
                             const float* __restrict__ sigmas, const float* __restrict__ raws,
                             const float* __restrict__ ws, const float* __restrict__ deltas,
                             const float* __restrict__ ts, const int64_t* __restrict__ rays_a,
                             const float* __restrict__ opacity, const float* __restrict__ depth,
                             const float* __restrict__ rend, float thr, int64_t n_rays, int64_t capacity,
                             float* __restrict__ dL_dsigmas, float* __restrict__ dL_draws) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = warp; n < n_rays; n += n_warps) {
    const int64_t ray_idx = rays_a[3 * n], start = rays_a[3 * n + 1];
    int64_t N64 = rays_a[3 * n + 2];
    if (start + N64 > capacity) N64 = capacity > start ? capacity - start : 0;
    const int N = (int)N64;
    if (N == 0) continue;
    float sg = 0.f, dl = 0.f, tt = 0.f, gwv = 0.f, rw[CT];
    auto fetch = [&](int base, float& o_sg, float& o_dl, float& o_tt, float& o_gw, float (&o_rw)[CT]) {
      const int k = base + lane;
      o_sg = 0.f; o_dl = 0.f; o_tt = 0.f; o_gw = 0.f;
#pragma unroll
      for (int c = 0; c < CT; ++c) o_rw[c] = 0.f;
      if (k < N) {
        const int64_t s = start + k;
        o_sg = sigmas[s]; o_dl = deltas[s]; o_tt = ts[s];
        if (dL_drend) {
#pragma unroll
          for (int c = 0; c < CT; ++c) o_rw[c] = raws[s * CT + c];
        }
        if (dL_dws) o_gw = dL_dws[s];
      }
    };
    fetch(0, sg, dl, tt, gwv, rw);                         // in flight together with the ray-level loads below
    const float gO = dL_dopacity ? dL_dopacity[ray_idx] : 0.f;
    const float gD = dL_ddepth ? dL_ddepth[ray_idx] : 0.f;
    float gR[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) gR[c] = dL_drend ? dL_drend[ray_idx * CT + c] : 0.f;
    const float O = opacity[ray_idx], D = depth[ray_idx];
    float part = 0.f;
    if (dL_drend && lane < CT) part += dL_drend[ray_idx * CT + lane] * rend[ray_idx * CT + lane];
    if (dL_dws) for (int k = lane; k < N; k += 32) part += dL_dws[start + k] * ws[start + k];
    const float Q_total = warp_sum(part) + gD * D;
    const float gO_term = gO * (1.0f - O);
    float T = 1.0f, carry = 0.f;
    bool dead = false;
    for (int base = 0; base < N; base += 32) {
      const int k = base + lane;
      const int64_t s = start + k;
      if (dead) {
        if (k < N) {
          if (dL_dsigmas) dL_dsigmas[s] = 0.f;
          if (dL_draws) {
#pragma unroll
            for (int c = 0; c < CT; ++c) dL_draws[s * CT + c] = 0.f;
          }
        }
        continue;
      }
      float n_sg = 0.f, n_dl = 0.f, n_tt = 0.f, n_gw = 0.f, n_rw[CT];
#pragma unroll
      for (int c = 0; c < CT; ++c) n_rw[c] = 0.f;
      if (base + 32 < N) fetch(base + 32, n_sg, n_dl, n_tt, n_gw, n_rw);
      float a = 0.f, g = 0.f;
      if (k < N) {
        a = __fsub_rn(1.0f, __expf(__fmul_rn(-sg, dl)));
        if (dL_drend) {
#pragma unroll
          for (int c = 0; c < CT; ++c) g += gR[c] * rw[c];
        }
      }
      const float om = __fsub_rn(1.0f, a);
      float T_before;
      const int stop = replay_transmittance(om, 32, thr, T, T_before, lane);
      const bool active = (k < N) && (stop < 0 || lane <= stop);
      const float w = active ? __fmul_rn(a, T_before) : 0.f;
      const float T_after = __fmul_rn(T_before, om);
      const float lin = gD * tt + g;
      const float q = active ? (w * lin + gwv * w) : 0.f;
      const float incl = warp_scan_incl_f(q, lane) + carry;
      if (k < N) {
        if (dL_dsigmas) dL_dsigmas[s] = active ? dl * (gO_term + T_after * (lin + gwv) - (Q_total - incl)) : 0.f;
        if (dL_draws) {
#pragma unroll
          for (int c = 0; c < CT; ++c) dL_draws[s * CT + c] = (active && dL_drend) ? gR[c] * w : 0.f;
        }
      }
      carry = __shfl_sync(0xffffffffu, incl, 31);
      if (stop >= 0) dead = true;
      sg = n_sg; dl = n_dl; tt = n_tt; gwv = n_gw;
#pragma unroll
      for (int c = 0; c < CT; ++c) rw[c] = n_rw[c];
    }
  }
}

// Test-time incremental compositing: one thread per alive ray (S is 1..64 and the layout
// is (A,S[,C]), so consecutive threads read consecutive rows).
__global__ void __launch_bounds__(256)
composite_test_fw_kernel(const float* __restrict__ sigmas, const float* __restrict__ raws,
                         const float* __restrict__ deltas, const float* __restrict__ ts,
                         int64_t* __restrict__ alive, float thr, const int32_t* __restrict__ n_eff,
                         int64_t n_alive, int S, int C, float* __restrict__ opacity,
                         float* __restrict__ depth, float* __restrict__ rend) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_alive) return;
  const int ne = n_eff[n];
  if (ne == 0) { alive[n] = -1; return; }
  const int64_t r = alive[n];
  float o = opacity[r], d = depth[r];
  float T = __fsub_rn(1.0f, o);
  const float* sg = sigmas + n * S; const float* dl = deltas + n * S; const float* tt = ts + n * S;
  const float* rw = raws + n * (int64_t)S * C;
  float* out = rend + r * C;
  for (int s = 0; s < ne; ++s) {
    const float a = __fsub_rn(1.0f, __expf(__fmul_rn(-sg[s], dl[s])));
    const float w = __fmul_rn(a, T);
    for (int c = 0; c < C; ++c) out[c] = __fmaf_rn(w, rw[s * C + c], out[c]);
    d = __fmaf_rn(w, tt[s], d);
    o = __fadd_rn(o, w);
    T = __fmul_rn(T, __fsub_rn(1.0f, a));
    if (T <= thr) { alive[n] = -1; break; }
  }
  opacity[r] = o; depth[r] = d;
}

}  // namespace ncn

using namespace ncn;

extern "C" int ncn_composite_train_fw(const float* sigmas, const float* raws, const float* deltas, const float* ts,
                                      const int64_t* rays_a, float T_threshold, int64_t n_rays, int64_t capacity,
                                      int n_channels, int64_t* total_samples, float* opacity, float* depth,
                                      float* rend, float* ws, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_rays >= 0 && capacity >= 0 && n_channels >= 0 && n_channels <= 32 * kMaxChanWords);
  if (n_rays == 0) return NCN_OK;
  NCN_CHECK_PTR(rays_a); NCN_CHECK_PTR(total_samples); NCN_CHECK_PTR(opacity); NCN_CHECK_PTR(depth);
  if (n_channels > 0) NCN_CHECK_PTR(rend);
  if (capacity > 0) {
    NCN_CHECK_PTR(sigmas); NCN_CHECK_PTR(deltas); NCN_CHECK_PTR(ts); NCN_CHECK_PTR(ws);
    if (n_channels > 0) NCN_CHECK_PTR(raws);
  }
  const int grid = persistent_grid(n_rays * 32, 256, 8);
  if (n_channels == 3) {
    composite_train_fw_ct_kernel<3><<<grid, 256, 0, as_stream(stream)>>>(sigmas, raws, deltas, ts, rays_a, T_threshold, n_rays, capacity,
                                                                         total_samples, opacity, depth, rend, ws);
    NCN_LAUNCH_OK();
    return NCN_OK;
  }
  if (n_channels == 6) {
    composite_train_fw_ct_kernel<6><<<grid, 256, 0, as_stream(stream)>>>(sigmas, raws, deltas, ts, rays_a, T_threshold, n_rays, capacity,
                                                                         total_samples, opacity, depth, rend, ws);
    NCN_LAUNCH_OK();
    return NCN_OK;
  }
  composite_train_fw_kernel<<<grid, 256, 0, as_stream(stream)>>>(sigmas, raws, deltas, ts, rays_a, T_threshold, n_rays,
                                                                 capacity, n_channels, total_samples, opacity, depth,
                                                                 rend, ws);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_composite_train_bw(const float* dL_dopacity, const float* dL_ddepth, const float* dL_drend,
                                      const float* dL_dws, const float* sigmas, const float* raws, const float* ws,
                                      const float* deltas, const float* ts, const int64_t* rays_a,
                                      const float* opacity, const float* depth, const float* rend, float T_threshold,
                                      int64_t n_rays, int64_t capacity, int n_channels, float* dL_dsigmas,
                                      float* dL_draws, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_rays >= 0 && capacity >= 0 && n_channels >= 0 && n_channels <= 32 * kMaxChanWords);
  if (n_rays == 0 || capacity == 0) return NCN_OK;
  NCN_CHECK_PTR(sigmas); NCN_CHECK_PTR(deltas); NCN_CHECK_PTR(ts); NCN_CHECK_PTR(rays_a);
  NCN_CHECK_PTR(opacity); NCN_CHECK_PTR(depth);
  if (!dL_dsigmas && !dL_draws) return NCN_E_NULL;          // either output may be skipped (two-branch backward), not both
  if (n_channels > 0) { NCN_CHECK_PTR(raws); NCN_CHECK_PTR(rend); }
  if (dL_dws) NCN_CHECK_PTR(ws);
  const int grid = persistent_grid(n_rays * 32, 256, 8);
  if (n_channels == 3) {
    composite_train_bw_ct_kernel<3><<<grid, 256, 0, as_stream(stream)>>>(dL_dopacity, dL_ddepth, dL_drend, dL_dws, sigmas, raws, ws, deltas, ts,
                                                                         rays_a, opacity, depth, rend, T_threshold, n_rays, capacity,
                                                                         dL_dsigmas, dL_draws);
    NCN_LAUNCH_OK();
    return NCN_OK;
  }
  composite_train_bw_kernel<<<grid, 256, 0, as_stream(stream)>>>(dL_dopacity, dL_ddepth, dL_drend, dL_dws, sigmas, raws,
                                                                 ws, deltas, ts, rays_a, opacity, depth, rend,
                                                                 T_threshold, n_rays, capacity, n_channels, dL_dsigmas,
                                                                 dL_draws);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_composite_test_fw(const float* sigmas, const float* raws, const float* deltas, const float* ts,
                                     int64_t* alive_indices, float T_threshold, const int32_t* n_eff, int64_t n_alive,
                                     int n_samples, int n_channels, float* opacity, float* depth, float* rend,
                                     ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_alive >= 0 && n_samples >= 1 && n_channels >= 0);
  if (n_alive == 0) return NCN_OK;
  NCN_CHECK_PTR(sigmas); NCN_CHECK_PTR(deltas); NCN_CHECK_PTR(ts); NCN_CHECK_PTR(alive_indices); NCN_CHECK_PTR(n_eff);
  NCN_CHECK_PTR(opacity); NCN_CHECK_PTR(depth);
  if (n_channels > 0) { NCN_CHECK_PTR(raws); NCN_CHECK_PTR(rend); }
  composite_test_fw_kernel<<<(unsigned)ceil_div(n_alive, 256), 256, 0, as_stream(stream)>>>(
      sigmas, raws, deltas, ts, alive_indices, T_threshold, n_eff, n_alive, n_samples, n_channels, opacity, depth, rend);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
