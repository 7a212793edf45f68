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

// ---------------------------------------------------------------------------------------------------------------------
// Sub-warp kernels (default): W lanes (W = 4, 8, 16) own a ray, so a warp composites 32/W rays at once.  At the training
// regime of ~30 samples per ray a whole warp per ray spends most of its issue slots on the 32-step transmittance replay
// of a mostly empty second chunk (728 warp instructions per ray measured); with W = 8 the replay is 8 steps shared by four
// rays, the chunk granularity follows the ray lengths, and the ray sums are kept as per-lane partials that are reduced ONCE
// at the end of the ray instead of five warp reductions per chunk.  The transmittance product is still replayed in the
// reference's sequential order (T, ws, total_samples bit-exact); the ray sums differ from the reference's sequential sums by
// fp32 rounding only (tolerance in tests/).
template <int W>
__device__ __forceinline__ float subwarp_sum(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int W>
__device__ __forceinline__ float subwarp_scan_incl(float v, int sl) {
#pragma unroll
  for (int o = 1; o < W; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, v, o, W);
    if (sl >= o) v += t;
  }
  return v;
}
// W-step replay inside every sub-warp of the warp at once (see replay_transmittance)
template <int W>
__device__ __forceinline__ int subwarp_replay(float om, float thr, float& T, float& T_before, int sl) {
  int stop = -1;
  T_before = T;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const float omj = __shfl_sync(0xffffffffu, om, j, W);
    if (sl == j) T_before = T;
    T = __fmul_rn(T, omj);
    if (stop < 0 && T <= thr) stop = j;
  }
  return stop;
}

// Photometric epilogue of the fused form (ncn_composite_train_fw_photometric): the lane that holds a ray's final sums evaluates
// rgb = rend[:3] + bg (1 - opacity), the squared error against the target colour and the opacity entropy term, and writes
// dL/drend, dL/dopacity (losses.py:347-361, models/rendering.py:231-241) - the separate loss launch and its dependent
// round trip through memory leave the step's critical path.  Same arithmetic as photometric_kernel (loss.cu).
struct PhotoArgs {
  const float* target;      // (R,3)
  float bg[3];
  float opacity_w, gscale, inv3n, invn;
  float* rgb_out;           // (R,3) or null
  float* sums;              // [0] += sum sq err, [1] += sum entropy
  float* d_rend;            // (R,CT) or null
  float* d_opacity;         // (R) or null
  int64_t n_gt;             // rays [0, n_gt) have a target colour; the rest (random_tr_poses, losses.py:271-297) only the opacity term
};

template <int CT, int W, bool PHOTO>
__global__ void __launch_bounds__(256)
composite_train_fw_sw_kernel(const float* __restrict__ sigmas, const float* __restrict__ raws,
                             const float* __restrict__ deltas, const float* __restrict__ ts,
                             const int64_t* __restrict__ rays_a, float thr, int64_t n_rays, int64_t capacity,
                             int64_t* __restrict__ total_samples, float* __restrict__ opacity,
                             float* __restrict__ depth, float* __restrict__ rend, float* __restrict__ ws, PhotoArgs ph) {
  constexpr int RPW = 32 / W;
  pdl_wait(); pdl_trigger();
  float ph_se = 0.f, ph_ent = 0.f;
  const int lane = threadIdx.x & 31, sub = lane / W, sl = lane % W;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n0 = warp * RPW; n0 < n_rays; n0 += n_warps * RPW) {
    const int64_t n = n0 + sub;
    const bool have = n < n_rays;
    int64_t ray_idx = 0, start = 0, N64 = 0;
    if (have) { ray_idx = rays_a[3 * n]; start = rays_a[3 * n + 1]; N64 = rays_a[3 * n + 2]; }
    if (start + N64 > capacity) N64 = capacity > start ? capacity - start : 0;
    const int N = (int)N64;
    const int N_max = __reduce_max_sync(0xffffffffu, N);
    float T = 1.0f, acc_o = 0.f, acc_d = 0.f, acc_r[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) acc_r[c] = 0.f;
    int samples = N;
    bool dead = false;
    float sg = 0.f, dl = 0.f, tt = 0.f, rw[CT];
    auto fetch = [&](int base, float& o_sg, float& o_dl, float& o_tt, float (&o_rw)[CT]) {
      const int k = base + sl;
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
    fetch(0, sg, dl, tt, rw);
    for (int base = 0; base < N_max; base += W) {
      const int k = base + sl;
      const int64_t s = start + k;
      float n_sg, n_dl, n_tt, n_rw[CT];
      fetch(base + W, n_sg, n_dl, n_tt, n_rw);
      const float a = k < N ? __fsub_rn(1.0f, __expf(__fmul_rn(-sg, dl))) : 0.f;
      float T_before;
      const int stop = subwarp_replay<W>(__fsub_rn(1.0f, a), thr, T, T_before, sl);
      const bool active = (k < N) && !dead && (stop < 0 || sl <= stop);
      const float w = active ? __fmul_rn(a, T_before) : 0.f;
      if (k < N) ws[s] = w;
      acc_o += w;
      acc_d += w * tt;
#pragma unroll
      for (int c = 0; c < CT; ++c) acc_r[c] += w * rw[c];
      if (!dead && stop >= 0) { samples = base + stop; dead = true; }
      sg = n_sg; dl = n_dl; tt = n_tt;
#pragma unroll
      for (int c = 0; c < CT; ++c) rw[c] = n_rw[c];
    }
    acc_o = subwarp_sum<W>(acc_o); acc_d = subwarp_sum<W>(acc_d);
#pragma unroll
    for (int c = 0; c < CT; ++c) acc_r[c] = subwarp_sum<W>(acc_r[c]);
    if (have && sl == 0) {
      total_samples[ray_idx] = samples; opacity[ray_idx] = acc_o; depth[ray_idx] = acc_d;
#pragma unroll
      for (int c = 0; c < CT; ++c) rend[ray_idx * CT + c] = acc_r[c];
      if (PHOTO) {
        float go = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = acc_r[c] + ph.bg[c] * (1.f - acc_o);
          if (ph.rgb_out) ph.rgb_out[3 * ray_idx + c] = v;
          const float e = ray_idx < ph.n_gt ? v - ph.target[3 * ray_idx + c] : 0.f;
          ph_se += e * e;
          const float g = 2.f * e * ph.inv3n * ph.gscale;
          if (ph.d_rend) ph.d_rend[ray_idx * CT + c] = g;
          go -= ph.bg[c] * g;
        }
        if (ph.d_rend) {
#pragma unroll
          for (int c = 3; c < CT; ++c) ph.d_rend[ray_idx * CT + c] = 0.f;
        }
        if (ph.opacity_w > 0.f) {
          const float oe = acc_o + 1e-10f;
          const float lg = logf(oe);
          ph_ent += -oe * lg;
          go += ph.opacity_w * (-(lg + 1.f)) * ph.invn * ph.gscale;
        }
        if (ph.d_opacity) ph.d_opacity[ray_idx] = go;
      }
    }
  }
  if (PHOTO) {
    ph_se = warp_sum(ph_se); ph_ent = warp_sum(ph_ent);
    __shared__ float s_a[8], s_b[8];
    const int wid = threadIdx.x >> 5;
    if (lane == 0) { s_a[wid] = ph_se; s_b[wid] = ph_ent; }
    __syncthreads();
    if (wid == 0) {
      ph_se = lane < 8 ? s_a[lane] : 0.f; ph_ent = lane < 8 ? s_b[lane] : 0.f;
      ph_se = warp_sum(ph_se); ph_ent = warp_sum(ph_ent);
      if (lane == 0 && (ph_se != 0.f || ph_ent != 0.f)) { atomicAdd(ph.sums, ph_se); atomicAdd(ph.sums + 1, ph_ent); }
    }
  }
}

template <int CT, int W>
__global__ void __launch_bounds__(256)
composite_train_bw_sw_kernel(const float* __restrict__ dL_dopacity, const float* __restrict__ dL_ddepth,
                             const float* __restrict__ dL_drend, const float* __restrict__ dL_dws,
                             const float* __restrict__ sigmas, const float* __restrict__ raws,
                             const float* __restrict__ ws, const float* __restrict__ deltas,
                             const float* __restrict__ ts, const int64_t* __restrict__ rays_a,
                             const float* __restrict__ opacity, const float* __restrict__ depth,
                             const float* __restrict__ rend, float thr, int64_t n_rays, int64_t capacity,
                             float* __restrict__ dL_dsigmas, float* __restrict__ dL_draws) {
  constexpr int RPW = 32 / W;
  pdl_wait(); pdl_trigger();
  const int lane = threadIdx.x & 31, sub = lane / W, sl = lane % W;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n0 = warp * RPW; n0 < n_rays; n0 += n_warps * RPW) {
    const int64_t n = n0 + sub;
    const bool have = n < n_rays;
    int64_t ray_idx = 0, start = 0, N64 = 0;
    if (have) { ray_idx = rays_a[3 * n]; start = rays_a[3 * n + 1]; N64 = rays_a[3 * n + 2]; }
    if (start + N64 > capacity) N64 = capacity > start ? capacity - start : 0;
    const int N = (int)N64;
    const int N_max = __reduce_max_sync(0xffffffffu, N);
    if (N_max == 0) continue;
    float sg = 0.f, dl = 0.f, tt = 0.f, gwv = 0.f, rw[CT];
    auto fetch = [&](int base, float& o_sg, float& o_dl, float& o_tt, float& o_gw, float (&o_rw)[CT]) {
      const int k = base + sl;
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
    float gO = 0.f, gD = 0.f, O = 0.f, D = 0.f, gR[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) gR[c] = 0.f;
    float part = 0.f;
    if (have && N > 0) {
      gO = dL_dopacity ? dL_dopacity[ray_idx] : 0.f;
      gD = dL_ddepth ? dL_ddepth[ray_idx] : 0.f;
      O = opacity[ray_idx]; D = depth[ray_idx];
      if (dL_drend) {
#pragma unroll
        for (int c = 0; c < CT; ++c) gR[c] = dL_drend[ray_idx * CT + c];
        if (sl == 0) {
#pragma unroll
          for (int c = 0; c < CT; ++c) part += gR[c] * rend[ray_idx * CT + c];
        }
      }
      if (dL_dws) for (int k = sl; k < N; k += W) part += dL_dws[start + k] * ws[start + k];
    }
    const float Q_total = subwarp_sum<W>(part) + gD * D;
    const float gO_term = gO * (1.0f - O);
    float T = 1.0f, carry = 0.f;
    bool dead = false;
    for (int base = 0; base < N_max; base += W) {
      const int k = base + sl;
      const int64_t s = start + k;
      float n_sg, n_dl, n_tt, n_gw, n_rw[CT];
      fetch(base + W, n_sg, n_dl, n_tt, n_gw, n_rw);
      float a = 0.f, g = 0.f;
      if (k < N) {
        a = __fsub_rn(1.0f, __expf(__fmul_rn(-sg, dl)));
#pragma unroll
        for (int c = 0; c < CT; ++c) g += gR[c] * rw[c];
      }
      const float om = __fsub_rn(1.0f, a);
      float T_before;
      const int stop = subwarp_replay<W>(om, thr, T, T_before, sl);
      const bool active = (k < N) && !dead && (stop < 0 || sl <= stop);
      const float w = active ? __fmul_rn(a, T_before) : 0.f;
      const float T_after = __fmul_rn(T_before, om);
      const float lin = gD * tt + g;
      const float q = active ? (w * lin + gwv * w) : 0.f;
      const float incl = subwarp_scan_incl<W>(q, sl) + carry;
      if (k < N) {
        if (dL_dsigmas) dL_dsigmas[s] = active ? dl * (gO_term + T_after * (lin + gwv) - (Q_total - incl)) : 0.f;
        if (dL_draws) {
#pragma unroll
          for (int c = 0; c < CT; ++c) dL_draws[s * CT + c] = active ? gR[c] * w : 0.f;
        }
      }
      carry = __shfl_sync(0xffffffffu, incl, W - 1, W);
      if (!dead && stop >= 0) dead = true;
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

// lanes per ray of the training compositing kernels: 4 / 8 / 16 = sub-warp kernels (C = 3, 6, 9), 32 = one warp per ray
static int g_composite_width = 16;
extern "C" int ncn_set_composite_width(int w) {
  const int old = g_composite_width;
  if (w == 4 || w == 8 || w == 16 || w == 32) g_composite_width = w;
  return old;
}

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
  if (g_composite_width < 32 && (n_channels == 3 || n_channels == 6 || n_channels == 9)) {
    const int gsw = persistent_grid(n_rays * g_composite_width, 256, 8);
    const PhotoArgs ph = PhotoArgs();
#define NCN_CFW(CT, W) composite_train_fw_sw_kernel<CT, W, false><<<gsw, 256, 0, as_stream(stream)>>>(sigmas, raws, deltas, ts, rays_a, T_threshold, \
                                                                                         n_rays, capacity, total_samples, opacity, depth, rend, ws, ph)
#define NCN_CFW_W(CT) do { if (g_composite_width == 4) NCN_CFW(CT, 4); else if (g_composite_width == 8) NCN_CFW(CT, 8); else NCN_CFW(CT, 16); } while (0)
    if (n_channels == 3) NCN_CFW_W(3); else if (n_channels == 6) NCN_CFW_W(6); else NCN_CFW_W(9);
#undef NCN_CFW_W
#undef NCN_CFW
    NCN_LAUNCH_OK();
    return NCN_OK;
  }
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

// compositing forward + the photometric / opacity loss terms and their gradients in ONE launch (C = 3, 6 or 9)
// n_gt_rays: rays [0, n_gt_rays) carry a target colour (the squared error is a mean over them), every ray the opacity term
extern "C" int ncn_composite_train_fw_photometric_gt(const float* sigmas, const float* raws, const float* deltas, const float* ts,
                                                     const int64_t* rays_a, float T_threshold, int64_t n_rays, int64_t capacity,
                                                     int n_channels, int64_t* total_samples, float* opacity, float* depth, float* rend,
                                                     float* ws, const float* target_rgb, int64_t n_gt_rays, const float* bg_rgb_host,
                                                     float opacity_w, float grad_scale, float* rgb_out, float* sums, float* dL_drend,
                                                     float* dL_dopacity, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_rays >= 0 && capacity >= 0 && n_gt_rays >= 0 && n_gt_rays <= n_rays);
  if (n_channels != 3 && n_channels != 6 && n_channels != 9) return NCN_E_UNSUPPORTED;
  if (n_rays == 0) return NCN_OK;
  NCN_CHECK_PTR(rays_a); NCN_CHECK_PTR(total_samples); NCN_CHECK_PTR(opacity); NCN_CHECK_PTR(depth); NCN_CHECK_PTR(rend);
  NCN_CHECK_PTR(target_rgb); NCN_CHECK_PTR(bg_rgb_host); NCN_CHECK_PTR(sums);
  if (capacity > 0) { NCN_CHECK_PTR(sigmas); NCN_CHECK_PTR(deltas); NCN_CHECK_PTR(ts); NCN_CHECK_PTR(ws); NCN_CHECK_PTR(raws); }
  PhotoArgs ph;
  ph.target = target_rgb; ph.bg[0] = bg_rgb_host[0]; ph.bg[1] = bg_rgb_host[1]; ph.bg[2] = bg_rgb_host[2];
  ph.opacity_w = opacity_w; ph.gscale = grad_scale; ph.inv3n = n_gt_rays > 0 ? 1.0f / (3.0f * (float)n_gt_rays) : 0.f; ph.invn = 1.0f / (float)n_rays;
  ph.rgb_out = rgb_out; ph.sums = sums; ph.d_rend = dL_drend; ph.d_opacity = dL_dopacity; ph.n_gt = n_gt_rays;
  const int w = g_composite_width < 32 ? g_composite_width : 16;
  const int gsw = persistent_grid(n_rays * w, 256, 8);
#define NCN_CFP(CT, W) NCN_CUDA(launch_pdl(composite_train_fw_sw_kernel<CT, W, true>, dim3(gsw), dim3(256), 0, as_stream(stream), sigmas, raws, deltas, ts, \
                                          rays_a, T_threshold, n_rays, capacity, total_samples, opacity, depth, rend, ws, ph))
#define NCN_CFP_W(CT) do { if (w == 4) NCN_CFP(CT, 4); else if (w == 8) NCN_CFP(CT, 8); else NCN_CFP(CT, 16); } while (0)
  if (n_channels == 3) NCN_CFP_W(3); else if (n_channels == 6) NCN_CFP_W(6); else NCN_CFP_W(9);
#undef NCN_CFP_W
#undef NCN_CFP
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_composite_train_fw_photometric(const float* sigmas, const float* raws, const float* deltas, const float* ts,
                                                  const int64_t* rays_a, float T_threshold, int64_t n_rays, int64_t capacity,
                                                  int n_channels, int64_t* total_samples, float* opacity, float* depth, float* rend,
                                                  float* ws, const float* target_rgb, const float* bg_rgb_host, float opacity_w,
                                                  float grad_scale, float* rgb_out, float* sums, float* dL_drend, float* dL_dopacity,
                                                  ncn_stream_t stream) {
  return ncn_composite_train_fw_photometric_gt(sigmas, raws, deltas, ts, rays_a, T_threshold, n_rays, capacity, n_channels, total_samples,
                                               opacity, depth, rend, ws, target_rgb, n_rays, bg_rgb_host, opacity_w, grad_scale, rgb_out,
                                               sums, dL_drend, dL_dopacity, stream);
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
  if (g_composite_width < 32 && (n_channels == 3 || n_channels == 6 || n_channels == 9)) {
    const int gsw = persistent_grid(n_rays * g_composite_width, 256, 8);
#define NCN_CBW(CT, W) NCN_CUDA(launch_pdl(composite_train_bw_sw_kernel<CT, W>, dim3(gsw), dim3(256), 0, as_stream(stream), dL_dopacity, dL_ddepth, dL_drend, \
                                          dL_dws, sigmas, raws, ws, deltas, ts, rays_a, opacity, depth, rend, T_threshold, n_rays, capacity,        \
                                          dL_dsigmas, dL_draws))
#define NCN_CBW_W(CT) do { if (g_composite_width == 4) NCN_CBW(CT, 4); else if (g_composite_width == 8) NCN_CBW(CT, 8); else NCN_CBW(CT, 16); } while (0)
    if (n_channels == 3) NCN_CBW_W(3); else if (n_channels == 6) NCN_CBW_W(6); else NCN_CBW_W(9);
#undef NCN_CBW_W
#undef NCN_CBW
    NCN_LAUNCH_OK();
    return NCN_OK;
  }
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
