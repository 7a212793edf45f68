// Optimizer tail of the training step: the dense per-parameter passes that, at T=2^19,
// move more bytes per step than all per-sample traffic.
// Replaces apex FusedAdam (adam_w_mode, train_nerf.py:262-285: eps 1e-15, weight decay 0 for
// the hash table / 1e-6 for the MLPs), GradScaler.unscale_, clip_grad_norm_(0.05)
// (train_nerf.py:954-955, opt.py:159) and the per-call fp32->fp16 parameter cast of the tcnn
// binding - fused into ONE streaming pass: read p, g, m, v (16 B) - write p, m, v, g=0, p16 (18 B).
#include <cstdlib>
#include "ncn_common.cuh"

namespace ncn {

// CTAs per SM of the streaming pass.  2 x 256 threads x 8 float4 loads in flight still saturate HBM, and leave half of the
// register file / thread slots to the latency-bound march kernels that run beside the deferred optimizer in the step graph
// (at 4 the optimizer owned every register of the SM and the march simply queued behind it).
static int adam_ctas_per_sm() {      // developer knob NCN_ADAM_CTAS (A/B measurements); default 2
  static int v = 0;
  if (v == 0) { const char* e = getenv("NCN_ADAM_CTAS"); v = e ? atoi(e) : 2; if (v < 1 || v > 8) v = 2; }
  return v;
}
#define kAdamCtasPerSm adam_ctas_per_sm()

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2;
  // several parameter groups in one launch (ncn_adam_step_groups): group q covers [start[q], start[q+1]) with its own
  // weight decay; max_norm > 0: the clip coefficient is derived in the kernel from the squared gradient norm
  int n_groups;
  long long start[NCN_ADAM_MAX_GROUPS];
  float wd[NCN_ADAM_MAX_GROUPS];
  float max_norm;
};

// apex FusedAdam (adam_w_mode) update.  The two bias corrections are applied as reciprocals computed once per thread
// and the final quotient uses the fast divider (<= 2 ulp from the IEEE quotient apex computes; the pass stays
// bandwidth bound instead of spending ~40 instructions per element on three IEEE divisions)
__device__ __forceinline__ float adam_wd_at(const AdamArgs& a, long long idx) {
  float wd = a.weight_decay;
#pragma unroll
  for (int q = 0; q < NCN_ADAM_MAX_GROUPS; ++q) if (q < a.n_groups && idx >= a.start[q]) wd = a.wd[q];
  return wd;
}

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, AdamArgs a, float gmul, float inv_bc1,
                                         float inv_bc2) {
  const float gr = g * gmul;
  m = a.beta1 * m + (1.f - a.beta1) * gr;
  v = a.beta2 * v + (1.f - a.beta2) * gr * gr;
  const float denom = sqrtf(v * inv_bc2) + a.eps;
  const float upd = __fdividef(m * inv_bc1, denom) + a.weight_decay * p;
  p = p - a.lr * upd;
  g = 0.f;
}

__global__ void __launch_bounds__(256, 4)
adam_kernel(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
            __half* __restrict__ p16, int64_t n, AdamArgs a, const float* __restrict__ grad_div,
            const int32_t* __restrict__ skip, const float* __restrict__ clip_coef, const float* __restrict__ lr_bc) {
  if (lr_bc != nullptr) { a.lr = lr_bc[0]; a.bc1 = lr_bc[1]; a.bc2 = lr_bc[2]; }   // device-side schedule (CUDA-graph replay)
  const bool do_skip = skip != nullptr && *skip != 0;
  float gmul = 1.f;
  if (grad_div) gmul = 1.f / *grad_div;
  if (clip_coef) {
    if (a.max_norm > 0.f) {        // clip_coef points at the squared gradient norm: min(1, max_norm / (|g| + 1e-6)) (clip_grad_norm_)
      const float c = a.max_norm / (sqrtf(*clip_coef) + 1e-6f);
      gmul *= c < 1.f ? c : 1.f;
    } else {
      gmul *= *clip_coef;
    }
  }
  const float inv_bc1 = 1.f / a.bc1, inv_bc2 = 1.f / a.bc2;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // two float4 per array per iteration: 8 independent 16 B loads in flight per thread
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
    const int64_t i1 = i0 + stride;
    const bool two = i1 < n4;
    float4 g0 = __ldcs(reinterpret_cast<const float4*>(grad) + i0);
    float4 g1 = two ? __ldcs(reinterpret_cast<const float4*>(grad) + i1) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (do_skip) {
      reinterpret_cast<float4*>(grad)[i0] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (two) reinterpret_cast<float4*>(grad)[i1] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    float4 p0 = __ldcs(reinterpret_cast<const float4*>(param) + i0), m0 = __ldcs(reinterpret_cast<const float4*>(m) + i0),
           v0 = __ldcs(reinterpret_cast<const float4*>(v) + i0);
    float4 p1 = p0, m1 = m0, v1 = v0;
    const float wd1 = a.n_groups > 0 ? adam_wd_at(a, i1 << 2) : a.weight_decay;
    if (a.n_groups > 0) a.weight_decay = adam_wd_at(a, i0 << 2);
    if (two) { p1 = __ldcs(reinterpret_cast<const float4*>(param) + i1); m1 = __ldcs(reinterpret_cast<const float4*>(m) + i1); v1 = __ldcs(reinterpret_cast<const float4*>(v) + i1); }
    adam_one(p0.x, g0.x, m0.x, v0.x, a, gmul, inv_bc1, inv_bc2); adam_one(p0.y, g0.y, m0.y, v0.y, a, gmul, inv_bc1, inv_bc2);
    adam_one(p0.z, g0.z, m0.z, v0.z, a, gmul, inv_bc1, inv_bc2); adam_one(p0.w, g0.w, m0.w, v0.w, a, gmul, inv_bc1, inv_bc2);
    reinterpret_cast<float4*>(param)[i0] = p0; reinterpret_cast<float4*>(m)[i0] = m0; reinterpret_cast<float4*>(v)[i0] = v0;
    reinterpret_cast<float4*>(grad)[i0] = g0;
    if (p16) {
      const __half2 lo = __floats2half2_rn(p0.x, p0.y), hi = __floats2half2_rn(p0.z, p0.w);
      uint2 pk; pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
      reinterpret_cast<uint2*>(p16)[i0] = pk;
    }
    if (two) {
      a.weight_decay = wd1;
      adam_one(p1.x, g1.x, m1.x, v1.x, a, gmul, inv_bc1, inv_bc2); adam_one(p1.y, g1.y, m1.y, v1.y, a, gmul, inv_bc1, inv_bc2);
      adam_one(p1.z, g1.z, m1.z, v1.z, a, gmul, inv_bc1, inv_bc2); adam_one(p1.w, g1.w, m1.w, v1.w, a, gmul, inv_bc1, inv_bc2);
      reinterpret_cast<float4*>(param)[i1] = p1; reinterpret_cast<float4*>(m)[i1] = m1; reinterpret_cast<float4*>(v)[i1] = v1;
      reinterpret_cast<float4*>(grad)[i1] = g1;
      if (p16) {
        const __half2 lo = __floats2half2_rn(p1.x, p1.y), hi = __floats2half2_rn(p1.z, p1.w);
        uint2 pk; pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(p16)[i1] = pk;
      }
    }
  }
  const int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    float g = grad[t];
    if (!do_skip) {
      float p = param[t], mm = m[t], vv = v[t];
      if (a.n_groups > 0) a.weight_decay = adam_wd_at(a, t);
      adam_one(p, g, mm, vv, a, gmul, inv_bc1, inv_bc2);
      param[t] = p; m[t] = mm; v[t] = vv;
      if (p16) p16[t] = __float2half_rn(p);
    }
    grad[t] = 0.f;
  }
}

// Squared L2 norm of the (unscaled) gradient, DETERMINISTIC: block partials are parked in a device buffer and the last
// block to finish (atomic ticket) adds them up in a fixed order, so the same gradient gives the same bits on every rank
// and every run - the clip coefficient, and with it the replicated parameters of a data-parallel job, stay bit-identical
// (an atomicAdd of the block partials made the ranks drift apart by ulps).  Not re-entrant across streams of one device.
constexpr int kSumsqMaxBlocks = 2048;
__device__ float g_sumsq_partials[kSumsqMaxBlocks];
__device__ unsigned int g_sumsq_ticket = 0;

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ grad, int64_t n, const float* __restrict__ grad_div, float* __restrict__ out,
             int32_t* __restrict__ flag) {
  const float gmul = grad_div ? 1.f / *grad_div : 1.f;
  float acc = 0.f;
  bool bad = false;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g = reinterpret_cast<const float4*>(grad)[i];
    const float a = g.x * gmul, b = g.y * gmul, c = g.z * gmul, d = g.w * gmul;
    acc += a * a + b * b + c * c + d * d;
    bad |= !(isfinite(a) && isfinite(b) && isfinite(c) && isfinite(d));
  }
  const int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) { const float a = grad[t] * gmul; acc += a * a; bad |= !isfinite(a); }
  acc = warp_sum(acc);
  __shared__ float s[8];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s[wid] = acc;
  if (bad && flag) atomicOr(flag, 1);
  __syncthreads();
  if (wid == 0) {
    acc = lane < 8 ? s[lane] : 0.f;
    acc = warp_sum(acc);
    if (lane == 0) {
      g_sumsq_partials[blockIdx.x] = acc;
      __threadfence();
      s_last = atomicAdd(&g_sumsq_ticket, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float tot = 0.f;                                     // fixed order: thread t adds partials t, t+256, ...; then a fixed tree
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) tot += __ldcg(&g_sumsq_partials[b]);
  tot = warp_sum(tot);
  if (lane == 0) s[wid] = tot;
  __syncthreads();
  if (wid == 0) {
    tot = lane < 8 ? s[lane] : 0.f;
    tot = warp_sum(tot);
    if (lane == 0) { *out += tot; g_sumsq_ticket = 0; }
  }
}

// clip_coef = min(1, max_norm / (sqrt(sumsq) + 1e-6))  (torch.nn.utils.clip_grad_norm_)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ coef) {
  const float nrm = sqrtf(*sumsq);
  const float c = max_norm / (nrm + 1e-6f);
  *coef = c < 1.f ? c : 1.f;
}

}  // namespace ncn

using namespace ncn;

extern "C" int ncn_adam_step(float* param, float* grad, float* m, float* v, void* param_f16, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step,
                             const float* grad_div_dev, const int32_t* skip_dev, const float* clip_coef_dev,
                             const float* lr_bc_dev, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0 && step >= 1);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(param); NCN_CHECK_PTR(grad); NCN_CHECK_PTR(m); NCN_CHECK_PTR(v);
  if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) return NCN_E_ALIGN;
  if ((uintptr_t)param_f16 & 7) return NCN_E_ALIGN;
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bc1 = 1.0f - powf(beta1, (float)step); a.bc2 = 1.0f - powf(beta2, (float)step);
  a.n_groups = 0; a.max_norm = 0.f;
  const int grid = persistent_grid((n + 7) / 8, 256, kAdamCtasPerSm);
  adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(param, grad, m, v, (__half*)param_f16, n, a, grad_div_dev, skip_dev,
                                                   clip_coef_dev, lr_bc_dev);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_adam_step_groups(float* param, float* grad, float* m, float* v, void* param_f16, int64_t n,
                                    const ncn_adam_groups* groups, float beta1, float beta2, float eps,
                                    const float* grad_div_dev, const int32_t* skip_dev, const float* sumsq_dev,
                                    const float* lr_bc_dev, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0);
  NCN_CHECK_PTR(groups); NCN_CHECK_PTR(lr_bc_dev);
  if (groups->n_groups < 1 || groups->n_groups > NCN_ADAM_MAX_GROUPS || groups->start[0] != 0) return NCN_E_CONFIG;
  if (groups->max_norm > 0.f) NCN_CHECK_PTR(sumsq_dev);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(param); NCN_CHECK_PTR(grad); NCN_CHECK_PTR(m); NCN_CHECK_PTR(v);
  if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) return NCN_E_ALIGN;
  if ((uintptr_t)param_f16 & 7) return NCN_E_ALIGN;
  AdamArgs a;
  a.lr = 0.f; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = groups->weight_decay[0]; a.bc1 = 1.f; a.bc2 = 1.f;
  a.n_groups = groups->n_groups; a.max_norm = groups->max_norm;
  for (int q = 0; q < NCN_ADAM_MAX_GROUPS; ++q) {
    const bool on = q < groups->n_groups;
    a.start[q] = on ? groups->start[q] : n; a.wd[q] = on ? groups->weight_decay[q] : 0.f;
    if (on && (groups->start[q] & 3)) return NCN_E_ALIGN;            // a float4 never straddles two groups
    if (on && q > 0 && groups->start[q] < groups->start[q - 1]) return NCN_E_CONFIG;
  }
  const int grid = persistent_grid((n + 7) / 8, 256, kAdamCtasPerSm);
  adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(param, grad, m, v, (__half*)param_f16, n, a, grad_div_dev, skip_dev,
                                                   groups->max_norm > 0.f ? sumsq_dev : nullptr, lr_bc_dev);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_grad_sumsq(const float* grad, int64_t n, const float* grad_div_dev, float* out, int32_t* flag,
                              ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(grad); NCN_CHECK_PTR(out);
  if ((uintptr_t)grad & 15) return NCN_E_ALIGN;
  int grid = persistent_grid((n + 3) / 4, 256, 8);
  if (grid > kSumsqMaxBlocks) grid = kSumsqMaxBlocks;
  sumsq_kernel<<<grid, 256, 0, as_stream(stream)>>>(grad, n, grad_div_dev, out, flag);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_clip_coef(const float* sumsq_dev, float max_norm, float* coef_dev, ncn_stream_t stream) {
  NCN_CHECK_PTR(sumsq_dev); NCN_CHECK_PTR(coef_dev);
  clip_coef_kernel<<<1, 1, 0, as_stream(stream)>>>(sumsq_dev, max_norm, coef_dev);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
