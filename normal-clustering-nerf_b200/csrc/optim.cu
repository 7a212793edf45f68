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
  // four independent 16-byte loads in flight per thread and iteration (the gradient is L2 resident right after the backward:
  // with one load per iteration the pass ran at 2.2 TB/s, a latency figure); the summation order stays fixed for a given grid
  float acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 g0 = reinterpret_cast<const float4*>(grad)[i], g1 = reinterpret_cast<const float4*>(grad)[i + stride],
                 g2 = reinterpret_cast<const float4*>(grad)[i + 2 * stride], g3 = reinterpret_cast<const float4*>(grad)[i + 3 * stride];
    { const float a = g0.x * gmul, b = g0.y * gmul, c = g0.z * gmul, d = g0.w * gmul; acc += a * a + b * b + c * c + d * d; bad |= !(isfinite(a) && isfinite(b) && isfinite(c) && isfinite(d)); }
    { const float a = g1.x * gmul, b = g1.y * gmul, c = g1.z * gmul, d = g1.w * gmul; acc1 += a * a + b * b + c * c + d * d; bad |= !(isfinite(a) && isfinite(b) && isfinite(c) && isfinite(d)); }
    { const float a = g2.x * gmul, b = g2.y * gmul, c = g2.z * gmul, d = g2.w * gmul; acc2 += a * a + b * b + c * c + d * d; bad |= !(isfinite(a) && isfinite(b) && isfinite(c) && isfinite(d)); }
    { const float a = g3.x * gmul, b = g3.y * gmul, c = g3.z * gmul, d = g3.w * gmul; acc3 += a * a + b * b + c * c + d * d; bad |= !(isfinite(a) && isfinite(b) && isfinite(c) && isfinite(d)); }
  }
  for (; i < n4; i += stride) {
    const float4 g = reinterpret_cast<const float4*>(grad)[i];
    const float a = g.x * gmul, b = g.y * gmul, c = g.z * gmul, d = g.w * gmul;
    acc += a * a + b * b + c * c + d * d;
    bad |= !(isfinite(a) && isfinite(b) && isfinite(c) && isfinite(d));
  }
  acc = (acc + acc1) + (acc2 + acc3);
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
  static int resident = 0;
  int grid = resident_grid(sumsq_kernel, 256, 0, &resident, ceil_div((n + 3) / 4, (int64_t)256));
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

// =====================================================================================================================
// Data-parallel optimizer over NVLink PEER MEMORY (the one exchange step of the path).
// Replaces [ncclAllReduce(flat gradient) -> ||g||^2 -> Adam on every rank] (train_nerf.py:949-955) by a sharded form whose
// result is the same replicated parameter set:
//   K1  rank r sums shard r of the gradient straight out of every peer's gradient buffer (P2P loads, fixed rank order ->
//       deterministic), parks the sum in its own buffer, reduces ||g_shard||^2 and posts it to every peer;
//   K2  total norm (same fixed-order sum on every rank -> identical clip coefficient), Adam on the shard only (1/W of the
//       p/m/v traffic), the fp16 working copy of the shard is STORED INTO EVERY PEER's fp16 parameter buffer (the only form
//       of the parameters the forward reads), the whole local gradient buffer is zeroed for the next backward.
// Per rank and step: (W-1)/W * 4 B/param in over NVLink, (W-1)/W * 2 B/param out - vs 2 * (W-1)/W * 4 B each way for a ring
// all-reduce - no NCCL call, so the whole step stays ONE CUDA graph.  Cross-GPU ordering uses epoch flags in each rank's
// sync block (release/acquire at system scope); every wait has a wall-clock bound (20 s) and raises `error` instead of hanging.
namespace ncn {

constexpr int kPeerMax = 8;
struct PeerSync {                       // lives in each rank's own device memory, written by peers
  unsigned int flag[4][kPeerMax];       // [phase][source rank] = last epoch that rank signalled (phase 3: early range complete)
  float norm[2][kPeerMax];              // [epoch & 1][source rank] partial squared norms (NaN = non-finite gradient there)
  unsigned int epoch;                   // local: completed steps
  unsigned int ticket[2];               // local: block tickets of K1 / K2
  unsigned int error;                   // local: a wait timed out
  float partials[kSumsqMaxBlocks];      // local: K1 block partials
  float partials0[kSumsqMaxBlocks];     // local: block partials of the early reduction (ncn_peer_early), folded in and cleared by K1
  // local, developer timeline (ns, %globaltimer) of the LAST step, written by one thread per event (ncn_peer_debug_times):
  // [0] K1 start  [1] K1 all peers' backward done (wait 0 over)  [2] K1 last block finished its reduction
  // [3] K2 start  [4] K2 norms in / peers done reading (wait 1 over)  [5] K2 last block finished Adam + publish  [6] K2 wait 2 over
  unsigned long long times[8];
  unsigned long long times0[4];         // early reduction: [0] start  [1] every peer's early range complete  [2] a block finished
};
struct PeerPtrs {
  const float* grad[kPeerMax];
  __half* p16[kPeerMax];
  PeerSync* sync[kPeerMax];
  unsigned int* err_host;               // this rank's error word in mapped pinned host memory (polled by the host without a sync)
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// flag fan-out to W peers: ONE system-scope fence (orders every earlier write of this thread - and, through the preceding block /
// grid synchronisation, of the kernel - before the flags) followed by W RELAXED stores that are all in flight together.  W
// st.release stores in a row cost W serialised NVLink round trips (measured: 18 us between the last K1 block finishing and K2
// starting on 8 GPUs), because every release waits for the previous remote store to be acknowledged.
__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// one thread: wait until every rank's flag of `phase` reached `epoch` (bounded: 20 s - ranks enter a step within microseconds of
// each other in steady state; the bound only has to survive start-up skew and must never turn a dead peer into a hung GPU)
// A timed-out wait is FATAL for the exchange: the error word is latched (device + host-mapped copy), the step that hit it and
// every later step become skipped steps on this rank (no Adam, nothing published, gradient zeroed) and - when the time-out
// happens in phase 0 - NaN is posted as this rank's partial norm so that the live peers skip the step as well.  The host polls
// the mapped word on every step and raises (ncn_b200.fused / trainer), instead of training on with half-reduced gradients.
__device__ unsigned long long g_peer_timeout_ns = 20000000000ull;
__device__ __forceinline__ void peer_wait(PeerSync* me, int phase, int world, unsigned int epoch, unsigned int* err_host) {
  const unsigned long long t0 = peer_now_ns();
  const unsigned long long limit = g_peer_timeout_ns;
  for (int q = 0; q < world; ++q) {
    while ((int)(ld_acquire_sys(&me->flag[phase][q]) - epoch) < 0) {
      __nanosleep(64);
      if (peer_now_ns() - t0 > limit) {
        atomicExch(&me->error, 1u + (unsigned)phase);
        if (err_host != nullptr) { *reinterpret_cast<volatile unsigned int*>(err_host) = 1u + (unsigned)phase; __threadfence_system(); }
        return;
      }
    }
  }
}
__device__ __forceinline__ bool peer_failed(PeerSync* me) { return *reinterpret_cast<volatile unsigned int*>(&me->error) != 0u; }

template <int U, int B>
__global__ void __launch_bounds__(256, 2)
peer_reduce_kernel(PeerPtrs pp, int rank, int world, int64_t lo4, int64_t hi4, float* my_grad /* == pp.grad[rank] */,
                   const float* __restrict__ grad_div, int early) {
  // early = 1 (ncn_peer_early): the same reduction over the EARLY range of this rank's shard, behind its own flag phase (3):
  // "the early range of my gradient is complete" - the rest of the backward is still running on every rank.  The block
  // partials of the norm are parked in partials0; the final launch (early = 0) folds them in, in a fixed order.
  PeerSync* me = pp.sync[rank];
  const int phase = early ? 3 : 0;
  unsigned long long* tm = early ? me->times0 : me->times;
  __shared__ unsigned int s_epoch;
  if (threadIdx.x == 0) {
    const unsigned int e = *reinterpret_cast<volatile unsigned int*>(&me->epoch) + 1u;
    if (blockIdx.x == 0) {                     // "my backward [its early range] is complete": stream order put this kernel after it
      __threadfence_system();
      for (int q = 0; q < world; ++q) st_relaxed_sys(&pp.sync[q]->flag[phase][rank], e);
    }
    if (blockIdx.x == 0) tm[0] = peer_now_ns();
    peer_wait(me, phase, world, e, pp.err_host);
    if (blockIdx.x == 0) tm[1] = peer_now_ns();
    s_epoch = e;
  }
  __syncthreads();
  const unsigned int epoch = s_epoch;
  const float gmul = grad_div ? 1.f / *grad_div : 1.f;
  float acc = 0.f;
  bool bad = peer_failed(me);                    // a peer never arrived: the sums below may be incomplete -> poison the norm
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // U float4 per peer and iteration: W * U independent 16-byte loads in flight per thread (sums in fixed rank order).
  // Measured on 8 GPUs (profiles/r2_timeline_8gpu_*, same box): 4 loads in flight per thread (U = 1, four peers at a time) pull the
  // 40 MB in 73-81 us, 8 in 104-106 us, 16 in 103-139 us - more requests in flight congest the fabric / the serving GPUs'
  // memory systems instead of hiding latency, so the default is the shallowest form.
  for (int64_t i0 = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi4; i0 += U * stride) {
    float4 sum[U];
#pragma unroll
    for (int u = 0; u < U; ++u) sum[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q0 = 0; q0 < kPeerMax; q0 += B) {          // B peers at a time: B * U loads in flight
      float4 v[B][U];
#pragma unroll
      for (int qq = 0; qq < B; ++qq)
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (q0 + qq < world && i0 + u * stride < hi4) v[qq][u] = __ldcs(reinterpret_cast<const float4*>(pp.grad[q0 + qq]) + i0 + u * stride);
#pragma unroll
      for (int qq = 0; qq < B; ++qq)
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (q0 + qq < world && i0 + u * stride < hi4) { sum[u].x += v[qq][u].x; sum[u].y += v[qq][u].y; sum[u].z += v[qq][u].z; sum[u].w += v[qq][u].w; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * stride < hi4) {
        reinterpret_cast<float4*>(my_grad)[i0 + u * stride] = sum[u];
        const float x = sum[u].x * gmul, y = sum[u].y * gmul, z = sum[u].z * gmul, w = sum[u].w * gmul;
        acc += x * x + y * y + z * z + w * w; bad |= !(isfinite(x) && isfinite(y) && isfinite(z) && isfinite(w));
      }
    }
  }
  if (bad) acc = __int_as_float(0x7fc00000);
  acc = warp_sum(acc);
  __shared__ float s[8];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    acc = lane < 8 ? s[lane] : 0.f;
    acc = warp_sum(acc);
    if (lane == 0) {
      if (early) { me->partials0[blockIdx.x] = acc; tm[2] = peer_now_ns(); s_last = false; }
      else {
        me->partials[blockIdx.x] = acc;
        __threadfence();
        s_last = atomicAdd(&me->ticket[0], 1u) == gridDim.x - 1;
      }
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float tot = 0.f;                               // fixed order (see sumsq_kernel); the early launch used the same grid
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
    tot += __ldcg(&me->partials[b]) + __ldcg(&me->partials0[b]);
    me->partials0[b] = 0.f;
  }
  tot = warp_sum(tot);
  if (lane == 0) s[wid] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s[w];
    me->ticket[0] = 0;
    me->times[2] = peer_now_ns();
    // every block of this rank has finished reading the peers' gradients: post the partial norm, then the flag
    for (int q = 0; q < world; ++q) *reinterpret_cast<volatile float*>(&pp.sync[q]->norm[epoch & 1][rank]) = tot;
    __threadfence_system();
    for (int q = 0; q < world; ++q) st_relaxed_sys(&pp.sync[q]->flag[1][rank], epoch);
  }
}

__global__ void __launch_bounds__(256, 2)
peer_adam_kernel(PeerPtrs pp, int rank, int world, int64_t lo4, int64_t hi4, int64_t lo4e, int64_t hi4e, int64_t n4, float* __restrict__ param,
                 float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v, AdamArgs a,
                 const float* __restrict__ grad_div, const int32_t* __restrict__ skip, const float* __restrict__ lr_bc,
                 float* __restrict__ sumsq_out, int zero_all) {
  PeerSync* me = pp.sync[rank];
  __shared__ unsigned int s_epoch;
  __shared__ float s_tot;
  if (threadIdx.x == 0) {
    const unsigned int e = *reinterpret_cast<volatile unsigned int*>(&me->epoch) + 1u;
    if (blockIdx.x == 0) me->times[3] = peer_now_ns();
    peer_wait(me, 1, world, e, pp.err_host);   // every rank's partial norm is here AND every rank is done reading my gradient
    if (blockIdx.x == 0) me->times[4] = peer_now_ns();
    float tot = 0.f;
    for (int q = 0; q < world; ++q) tot += *reinterpret_cast<volatile float*>(&me->norm[e & 1][q]);
    s_tot = tot; s_epoch = e;
  }
  __syncthreads();
  const unsigned int epoch = s_epoch;
  const float total = s_tot;
  if (blockIdx.x == 0 && threadIdx.x == 0 && sumsq_out) *sumsq_out = total;
  if (lr_bc != nullptr) { a.lr = lr_bc[0]; a.bc1 = lr_bc[1]; a.bc2 = lr_bc[2]; }
  // (the error word is read after the block-wide barrier above, i.e. after thread 0's own wait: uniform within the block)
  const bool do_skip = (skip != nullptr && *skip != 0) || !isfinite(total) || peer_failed(me);
  float gmul = grad_div ? 1.f / *grad_div : 1.f;
  if (a.max_norm > 0.f) { const float c = a.max_norm / (sqrtf(total) + 1e-6f); gmul *= c < 1.f ? c : 1.f; }
  const float inv_bc1 = 1.f / a.bc1, inv_bc2 = 1.f / a.bc2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (!do_skip) {
    // this rank's shard = [lo4, hi4) of the late range followed by [lo4e, hi4e) of the early range (empty without a cut)
    const int64_t len_l = hi4 - lo4, len_all = len_l + (hi4e - lo4e);
    for (int64_t j = tid; j < len_all; j += stride) {
      const int64_t i = j < len_l ? lo4 + j : lo4e + (j - len_l);
      float4 g = reinterpret_cast<const float4*>(grad)[i];
      float4 p = __ldcs(reinterpret_cast<const float4*>(param) + i), mm = __ldcs(reinterpret_cast<const float4*>(m) + i),
             vv = __ldcs(reinterpret_cast<const float4*>(v) + i);
      if (a.n_groups > 0) a.weight_decay = adam_wd_at(a, i << 2);
      adam_one(p.x, g.x, mm.x, vv.x, a, gmul, inv_bc1, inv_bc2); adam_one(p.y, g.y, mm.y, vv.y, a, gmul, inv_bc1, inv_bc2);
      adam_one(p.z, g.z, mm.z, vv.z, a, gmul, inv_bc1, inv_bc2); adam_one(p.w, g.w, mm.w, vv.w, a, gmul, inv_bc1, inv_bc2);
      reinterpret_cast<float4*>(param)[i] = p; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
      reinterpret_cast<float4*>(grad)[i] = g;      // adam_one zeroed it
      const __half2 lo = __floats2half2_rn(p.x, p.y), hi = __floats2half2_rn(p.z, p.w);
      uint2 pk; pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
#pragma unroll
      for (int q = 0; q < kPeerMax; ++q)         // publish the shard's fp16 working copy into every rank's parameter buffer
        if (q < world) reinterpret_cast<uint2*>(pp.p16[q])[i] = pk;
    }
  }
  // the next backward accumulates into a zeroed buffer; peers are done reading it (barrier above)
  // (the shard itself was zeroed element by element by the thread that consumed it)
  // (zero_all == 0: the caller zeroes the buffer itself after this kernel, off the critical path - ncn_peer_set_external_zero)
  if (zero_all)
    for (int64_t i = tid; i < n4; i += stride)
      if (do_skip || !((i >= lo4 && i < hi4) || (i >= lo4e && i < hi4e))) reinterpret_cast<float4*>(grad)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __threadfence_system();
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&me->ticket[1], 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) {
    me->ticket[1] = 0;
    me->times[5] = peer_now_ns();
    __threadfence_system();
    for (int q = 0; q < world; ++q) st_relaxed_sys(&pp.sync[q]->flag[2][rank], epoch);
    peer_wait(me, 2, world, epoch, pp.err_host);   // every shard of MY fp16 parameter buffer has been written by its owner
    me->times[6] = peer_now_ns();
    *reinterpret_cast<volatile unsigned int*>(&me->epoch) = epoch;
    __threadfence();
  }
}

}  // namespace ncn

struct ncn_peer {
  int rank, world, device;
  int64_t n;
  int64_t cut;                          // [cut, n) = the early range (ncn_peer_set_cut); cut == n: none
  float* grad;
  void* p16;
  ncn::PeerSync* sync;
  ncn::PeerPtrs ptrs;
  void* opened[3][ncn::kPeerMax];
  bool connected;
  bool external_zero;                   // the caller zeroes the gradient buffer after ncn_peer_step (ncn_peer_set_external_zero)
  unsigned int* err_host;               // cudaHostAllocMapped: written by the kernels on a time-out, read by ncn_peer_poll
};

// developer A/B knob: 16-byte loads in flight per peer and thread in the reduce kernel (1, 2 [default] or 4)
static int g_peer_loads = 1, g_peer_batch = 4, g_peer_ctas = 2, g_peer_early_loads = 0;
extern "C" int ncn_peer_set_early_loads(int u) { const int old = g_peer_early_loads; if (u == 0 || u == 1 || u == 2 || u == 4) g_peer_early_loads = u; return old; }
extern "C" int ncn_peer_set_loads(int u) { const int old = g_peer_loads; if (u == 1 || u == 2 || u == 4) g_peer_loads = u; return old; }
// further A/B knobs of the reduce kernel: peers loaded per batch (2, 4 [default], 8; with loads = 1) and CTAs per SM (1 or 2 [default])
extern "C" int ncn_peer_set_shape(int peers_per_batch, int ctas_per_sm) {
  if (peers_per_batch == 2 || peers_per_batch == 4 || peers_per_batch == 8) g_peer_batch = peers_per_batch;
  if (ctas_per_sm == 1 || ctas_per_sm == 2) g_peer_ctas = ctas_per_sm;
  return NCN_OK;
}

extern "C" int ncn_peer_create(ncn_peer** out, int rank, int world, int64_t n_params) {
  NCN_CHECK_PTR(out);
  NCN_CHECK_SIZE(world >= 1 && world <= ncn::kPeerMax && rank >= 0 && rank < world && n_params > 0 && (n_params & 3) == 0);
  ncn_peer* p = new ncn_peer();
  p->rank = rank; p->world = world; p->n = n_params; p->cut = n_params; p->connected = false; p->external_zero = false;
  NCN_CUDA(cudaGetDevice(&p->device));
  NCN_CUDA(cudaMalloc(&p->grad, (size_t)n_params * 4));
  NCN_CUDA(cudaMalloc(&p->p16, (size_t)n_params * 2));
  NCN_CUDA(cudaMalloc(&p->sync, sizeof(ncn::PeerSync)));
  NCN_CUDA(cudaMemset(p->grad, 0, (size_t)n_params * 4));
  NCN_CUDA(cudaMemset(p->p16, 0, (size_t)n_params * 2));
  NCN_CUDA(cudaMemset(p->sync, 0, sizeof(ncn::PeerSync)));
  NCN_CUDA(cudaHostAlloc((void**)&p->err_host, sizeof(unsigned int), cudaHostAllocMapped));
  *p->err_host = 0u;
  { void* dptr = nullptr; NCN_CUDA(cudaHostGetDevicePointer(&dptr, p->err_host, 0)); p->ptrs.err_host = (unsigned int*)dptr; }
  NCN_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < 3; ++k) for (int q = 0; q < ncn::kPeerMax; ++q) p->opened[k][q] = nullptr;
  if (world == 1) {
    p->ptrs.grad[0] = p->grad; p->ptrs.p16[0] = (__half*)p->p16; p->ptrs.sync[0] = p->sync;
    p->connected = true;
  }
  *out = p;
  return NCN_OK;
}

extern "C" float* ncn_peer_grad(ncn_peer* p) { return p ? p->grad : nullptr; }
extern "C" void* ncn_peer_p16(ncn_peer* p) { return p ? p->p16 : nullptr; }

extern "C" int ncn_peer_handles(ncn_peer* p, void* handles_out) {
  NCN_CHECK_PTR(p); NCN_CHECK_PTR(handles_out);
  cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)handles_out;
  NCN_CUDA(cudaIpcGetMemHandle(&h[0], p->grad));
  NCN_CUDA(cudaIpcGetMemHandle(&h[1], p->p16));
  NCN_CUDA(cudaIpcGetMemHandle(&h[2], p->sync));
  return NCN_OK;
}

extern "C" int ncn_peer_connect(ncn_peer* p, const void* all_handles) {
  NCN_CHECK_PTR(p); NCN_CHECK_PTR(all_handles);
  const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)all_handles;
  for (int q = 0; q < p->world; ++q) {
    if (q == p->rank) {
      p->ptrs.grad[q] = p->grad; p->ptrs.p16[q] = (__half*)p->p16; p->ptrs.sync[q] = p->sync;
      continue;
    }
    for (int k = 0; k < 3; ++k)
      NCN_CUDA(cudaIpcOpenMemHandle(&p->opened[k][q], h[3 * q + k], cudaIpcMemLazyEnablePeerAccess));
    p->ptrs.grad[q] = (const float*)p->opened[0][q];
    p->ptrs.p16[q] = (__half*)p->opened[1][q];
    p->ptrs.sync[q] = (ncn::PeerSync*)p->opened[2][q];
  }
  p->connected = true;
  return NCN_OK;
}

extern "C" void ncn_peer_shard(int64_t n_params, int rank, int world, int64_t* lo, int64_t* hi) {
  const int64_t n4 = n_params >> 2;
  if (lo) *lo = (n4 * rank / world) << 2;
  if (hi) *hi = (n4 * (rank + 1) / world) << 2;
}

// rank q's shard as two segments: [seg[0], seg[1]) of the late range [0, cut) and [seg[2], seg[3]) of the early range [cut, n)
static void peer_segments(int64_t n, int64_t cut, int q, int world, int64_t seg[4]) {
  const int64_t c4 = cut >> 2, n4 = n >> 2;
  seg[0] = (c4 * q / world) << 2; seg[1] = (c4 * (q + 1) / world) << 2;
  seg[2] = (c4 + (n4 - c4) * q / world) << 2; seg[3] = (c4 + (n4 - c4) * (q + 1) / world) << 2;
}
// the same arithmetic without a peer object (host only)
extern "C" void ncn_peer_segments_of(int64_t n_params, int64_t cut, int rank, int world, int64_t* seg4_out) {
  if (seg4_out) peer_segments(n_params, cut, rank, world, seg4_out);
}
extern "C" int ncn_peer_segments(ncn_peer* p, int rank_q, int64_t* seg4_out) {
  NCN_CHECK_PTR(p); NCN_CHECK_PTR(seg4_out);
  NCN_CHECK_SIZE(rank_q >= 0 && rank_q < p->world);
  peer_segments(p->n, p->cut, rank_q, p->world, seg4_out);
  return NCN_OK;
}
extern "C" int ncn_peer_set_cut(ncn_peer* p, int64_t cut) {
  NCN_CHECK_PTR(p);
  NCN_CHECK_SIZE(cut >= 0 && cut <= p->n && (cut & 3) == 0);
  p->cut = cut;
  return NCN_OK;
}

static int peer_launch_reduce(ncn_peer* p, int64_t lo4, int64_t hi4, const float* grad_div_dev, int early, cudaStream_t st) {
  // the early launch co-runs with the rest of the table backward: ONE CTA per SM (52 registers x 256 threads fit beside four of that
  // kernel's CTAs; the pull is NVLink bound, its time did not depend on the CTA count - profiles/r2_timeline_8gpu_*)
  int grid1 = ncn::sm_count() * (early ? 1 : g_peer_ctas);
  if (grid1 > ncn::kSumsqMaxBlocks) grid1 = ncn::kSumsqMaxBlocks;
#define NCN_K1(U, B) ncn::peer_reduce_kernel<U, B><<<grid1, 256, 0, st>>>(p->ptrs, p->rank, p->world, lo4, hi4, p->grad, grad_div_dev, early)
  // the early launch has half the threads of the final one and shares its SMs: keep about 8 loads in flight per thread on small worlds
  int loads = g_peer_loads;
  if (early && g_peer_early_loads > 0) loads = g_peer_early_loads;
  else if (early) loads = p->world <= 2 ? 4 : (p->world <= 4 ? 2 : 1);
  if (loads == 4) NCN_K1(4, 4);
  else if (loads == 2) NCN_K1(2, 4);
  else if (g_peer_batch == 2) NCN_K1(1, 2);
  else if (g_peer_batch == 8) NCN_K1(1, 8);
  else NCN_K1(1, 4);
#undef NCN_K1
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_peer_early(ncn_peer* p, const float* grad_div_dev, ncn_stream_t stream) {
  NCN_CHECK_PTR(p);
  if (!p->connected || p->cut >= p->n) return NCN_E_CONFIG;
  int64_t seg[4];
  peer_segments(p->n, p->cut, p->rank, p->world, seg);
  return peer_launch_reduce(p, seg[2] >> 2, seg[3] >> 2, grad_div_dev, 1, ncn::as_stream(stream));
}

extern "C" int ncn_peer_step(ncn_peer* p, float* param, float* m, float* v, const ncn_adam_groups* groups, float beta1,
                             float beta2, float eps, const float* grad_div_dev, const int32_t* skip_dev,
                             const float* lr_bc_dev, float* sumsq_out_dev, ncn_stream_t stream) {
  NCN_CHECK_PTR(p); NCN_CHECK_PTR(param); NCN_CHECK_PTR(m); NCN_CHECK_PTR(v); NCN_CHECK_PTR(groups); NCN_CHECK_PTR(lr_bc_dev);
  if (!p->connected) return NCN_E_CONFIG;
  if (groups->n_groups < 1 || groups->n_groups > NCN_ADAM_MAX_GROUPS || groups->start[0] != 0) return NCN_E_CONFIG;
  if (((uintptr_t)param | (uintptr_t)m | (uintptr_t)v) & 15) return NCN_E_ALIGN;
  ncn::AdamArgs a;
  a.lr = 0.f; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = groups->weight_decay[0]; a.bc1 = 1.f; a.bc2 = 1.f;
  a.n_groups = groups->n_groups; a.max_norm = groups->max_norm;
  for (int q = 0; q < NCN_ADAM_MAX_GROUPS; ++q) {
    const bool on = q < groups->n_groups;
    a.start[q] = on ? groups->start[q] : p->n; a.wd[q] = on ? groups->weight_decay[q] : 0.f;
    if (on && (groups->start[q] & 3)) return NCN_E_ALIGN;
  }
  int64_t seg[4];
  peer_segments(p->n, p->cut, p->rank, p->world, seg);
  const int64_t n4 = p->n >> 2;
  int grid = ncn::sm_count() * 2;
  if (grid > ncn::kSumsqMaxBlocks) grid = ncn::kSumsqMaxBlocks;
  cudaStream_t st = ncn::as_stream(stream);
  { const int rc = peer_launch_reduce(p, seg[0] >> 2, seg[1] >> 2, grad_div_dev, 0, st); if (rc) return rc; }
  ncn::peer_adam_kernel<<<grid, 256, 0, st>>>(p->ptrs, p->rank, p->world, seg[0] >> 2, seg[1] >> 2, seg[2] >> 2, seg[3] >> 2, n4, param, p->grad,
                                              m, v, a, grad_div_dev, skip_dev, lr_bc_dev, sumsq_out_dev, p->external_zero ? 0 : 1);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_peer_error(ncn_peer* p, unsigned int* error_host) {
  NCN_CHECK_PTR(p); NCN_CHECK_PTR(error_host);
  NCN_CUDA(cudaMemcpy(error_host, &p->sync->error, sizeof(unsigned int), cudaMemcpyDeviceToHost));
  return NCN_OK;
}

// on = 1: ncn_peer_step leaves the gradient buffer as it is (its own shard is consumed to zero, the rest still holds this rank's
// gradient) and the CALLER zeroes the whole buffer after the step, on whatever stream lets it overlap the next forward pass -
// the 45.8 MB memset then leaves the optimizer's critical path.  The buffer must be zero again before the next backward starts.
extern "C" int ncn_peer_set_external_zero(ncn_peer* p, int on) {
  NCN_CHECK_PTR(p);
  p->external_zero = on != 0;
  return NCN_OK;
}

// developer timeline of the last completed step (synchronises): 8 x %globaltimer ns, see PeerSync::times
extern "C" int ncn_peer_debug_times(ncn_peer* p, unsigned long long* times8_host) {
  NCN_CHECK_PTR(p); NCN_CHECK_PTR(times8_host);
  NCN_CUDA(cudaMemcpy(times8_host, p->sync->times, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost));
  return NCN_OK;
}

extern "C" int ncn_peer_debug_times_early(ncn_peer* p, unsigned long long* times4_host) {
  NCN_CHECK_PTR(p); NCN_CHECK_PTR(times4_host);
  NCN_CUDA(cudaMemcpy(times4_host, p->sync->times0, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost));
  return NCN_OK;
}

// the error word WITHOUT a device synchronisation (host-mapped copy): 0 = ok, 1 + phase of the wait that timed out
extern "C" unsigned int ncn_peer_poll(ncn_peer* p) {
  return p && p->err_host ? *reinterpret_cast<volatile unsigned int*>(p->err_host) : 0u;
}

// bound of every cross-GPU wait in seconds (default 20; applies to all ncn_peer objects of the process)
extern "C" int ncn_peer_set_timeout(double seconds) {
  NCN_CHECK_SIZE(seconds > 0.0 && seconds < 3600.0);
  const unsigned long long ns = (unsigned long long)(seconds * 1e9);
  NCN_CUDA(cudaMemcpyToSymbol(ncn::g_peer_timeout_ns, &ns, sizeof(ns)));
  return NCN_OK;
}

extern "C" int ncn_peer_destroy(ncn_peer* p) {
  if (!p) return NCN_OK;
  cudaDeviceSynchronize();
  if (p->err_host) cudaFreeHost(p->err_host);
  for (int k = 0; k < 3; ++k) for (int q = 0; q < ncn::kPeerMax; ++q) if (p->opened[k][q]) cudaIpcCloseMemHandle(p->opened[k][q]);
  cudaFree(p->grad); cudaFree(p->p16); cudaFree(p->sync);
  delete p;
  return NCN_OK;
}
