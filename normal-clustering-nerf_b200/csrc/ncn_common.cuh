// Shared helpers for libncn.so (sm_100a).  Compiled with -fmad=false: every fused
// multiply-add in this library is written out (__fmaf_rn / fmaf) so the fp32
// rounding sequence of the bit-exact kernels is fixed in the source, not chosen
// by the compiler.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/ncn.h"

#define NCN_CHECK_PTR(p) do { if ((p) == nullptr) return NCN_E_NULL; } while (0)
#define NCN_CHECK_SIZE(c) do { if (!(c)) return NCN_E_SIZE; } while (0)
#define NCN_LAUNCH_OK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)
#define NCN_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return (int)e__; } while (0)

namespace ncn {

constexpr int kWarp = 32;
constexpr float kSqrt3 = 1.73205080757f;  // raymarching.cu:4

static inline cudaStream_t as_stream(ncn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// SM count of the current device, cached per device.
int sm_count();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Saved hidden activations of the MLPs ("acts", include/ncn.h ncn_mlp_fwd): per layer ceil(N/128) tiles of 128 rows,
// each tile stored as the [feature/8][row][8 halfs] panel the tcgen05 backward multiplies from - one contiguous
// 16 KB block per tile and layer, fetched by a single bulk copy (TMA) without any layout change on the way.
constexpr int kActTile = 128;
__host__ __device__ __forceinline__ int64_t act_rows(int64_t n) { return (n + (kActTile - 1)) & ~(int64_t)(kActTile - 1); }
__host__ __device__ __forceinline__ int64_t act_offset(int64_t row, int f) {        // halfs, within one layer
  return (row >> 7) * (64 * kActTile) + ((int64_t)((f >> 3) * kActTile + (int)(row & (kActTile - 1))) << 3) + (f & 7);
}

// grid for a grid-stride kernel: enough CTAs for `n` items but at most `waves`
// resident waves of the 148-SM part (blocks_per_sm resident CTAs each).
static inline int persistent_grid(int64_t n, int threads, int blocks_per_sm) {
  int64_t need = ceil_div(n > 0 ? n : 1, threads);
  int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  return (int)(need < cap ? need : cap);
}

// CTAs of `kernel` that are co-resident on the device (register / shared-memory limited), for grid-stride kernels whose CTAs
// all carry the same share of the work: a grid larger than this runs a second, partly filled wave (1184 CTAs at 5 per SM were
// 1.6 waves: ncu showed 15 % of the elapsed cycles with idle SMs).  `cached` is a per-call-site static.
template <typename K>
static inline int resident_grid(K kernel, int threads, size_t smem, int* cached, int64_t need_blocks) {
  if (*cached == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    *cached = per_sm * sm_count();
  }
  if (need_blocks < 1) need_blocks = 1;
  return (int)(need_blocks < (int64_t)*cached ? need_blocks : (int64_t)*cached);
}

// Programmatic dependent launch (sm_90+).  A kernel launched through launch_pdl() may be scheduled while its stream predecessor
// is still running; it must execute pdl_wait() before it touches anything an earlier kernel wrote (or overwrites anything an
// earlier kernel reads) - the wait returns when the predecessor grid has completed and flushed.  pdl_trigger() lets the NEXT
// kernel of the stream be scheduled; every kernel here triggers only AFTER its own wait, so at most one dependent is ever
// resident beside a running kernel and completion stays transitive along the stream.  Both are no-ops in a kernel that was
// launched without the attribute (ncn_set_pdl(0), or a predecessor that is not a kernel).
int pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive warp prefix sum
__device__ __forceinline__ int warp_scan_incl_i(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// streaming (read-once) loads / (write-once) stores
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

}  // namespace ncn
