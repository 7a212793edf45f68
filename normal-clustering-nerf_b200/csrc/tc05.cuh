// tcgen05 / TMEM / TMA building blocks shared by the tensor-core MLP kernels (mlp_tc05.cu: backward, field_tc05.cu: forward):
// PTX wrappers, shared-memory matrix descriptors for the no-swizzle [feature/8][row][8 halfs] panels, panel staging helpers.
#pragma once
#include "ncn_common.cuh"
#include "mma.cuh"

namespace ncn {

constexpr int kTile = kActTile;    // samples per CTA iteration = one activation tile
constexpr int kTcThreads = 128;    // one thread per sample row (TMEM lane)

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NCN_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NCN_DONE_%=;\n"
      "bra NCN_WAIT_%=;\n"
      "NCN_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// contiguous global -> shared bulk copy (TMA, no tensor map), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 accumulator columns of this thread's TMEM lane (no wait: pair with tmem_ld_wait)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// one lane of a converged warp (CUTLASS elect_one_sync): keeps the tcgen05 issue path on the uniform datapath
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred px;\n"
      "elect.sync _|px, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, px;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ------------------------------------------------------------------ descriptors (cute/arch/mma_sm100_desc.hpp)
// shared-memory matrix descriptor, no swizzle: bits [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, bit 46 version = 1.
// MN-major operand over a panel [feature/8][kTile][8]: 8 K-rows (samples) are 16 B apart, the next 8 samples LBO = 128 B
// further (K direction), the next 8 features SBO = kTile*16 B further (M/N direction)
__device__ __forceinline__ uint64_t make_desc_mn(const void* panel_at_k) {
  const uint64_t addr = (uint64_t)(smem_u32(panel_at_k) >> 4) & 0x3FFF;
  const uint64_t lbo = (128 >> 4), sbo = ((kTile * 16) >> 4);
  return addr | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// K-major A operand over the same panel: rows (M = samples) 16 B apart, next 8 rows SBO = 128 B, next 8 K-elements (next
// feature group) LBO = kTile*16 B
__device__ __forceinline__ uint64_t make_desc_k(const void* panel_at_k) {
  const uint64_t addr = (uint64_t)(smem_u32(panel_at_k) >> 4) & 0x3FFF;
  const uint64_t lbo = ((kTile * 16) >> 4), sbo = (128 >> 4);
  return addr | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// MN-major B operand over a weight panel [n/8][K][8]: K rows 16 B apart, next 8 K rows LBO = 128 B, next 8 n: SBO = K*16 B
__device__ __forceinline__ uint64_t make_desc_w(const void* panel_at_k, int K) {
  const uint64_t addr = (uint64_t)(smem_u32(panel_at_k) >> 4) & 0x3FFF;
  const uint64_t lbo = (128 >> 4), sbo = (uint64_t)((K * 16) >> 4);
  return addr | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// instruction descriptors: D = f32 (bit 4), A/B = f16, a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t make_idesc_wgrad(int N) {      // A, B MN-major, M = 64
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_dgrad(int N) {      // A K-major, B MN-major, M = 128
  return (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// advance a descriptor's start address by `bytes` (address field = addr >> 4 in the low 14 bits; no carry out below 256 KB)
__device__ __forceinline__ uint64_t desc_add(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
constexpr int tmem_cols_pow2(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// ------------------------------------------------------------------ staging helpers
__device__ __forceinline__ int perm_col(int c) { return c < 16 ? c + 3 : (c < 19 ? c - 16 : c); }   // fused-forward order -> tcnn order
// weight matrix (K rows, N cols, row-major fp16) -> MN-major panel [n/8][K][8]: a permutation of 16-byte chunks
__device__ __forceinline__ void load_w_panel_async(const __half* __restrict__ w, int K, int N, __half* __restrict__ P) {
  const int nb_count = N >> 3;
  for (int i = threadIdx.x; i < K * nb_count; i += blockDim.x) {
    const int k = i / nb_count, nb = i - k * nb_count;
    cp_async16(P + ((size_t)nb * K + k) * 8, w + k * N + nb * 8);
  }
}
// same with the input columns of the first layer permuted (ncn_mlp_bwd_src.perm bit 0)
__device__ __forceinline__ void load_w_panel_perm(const __half* __restrict__ w, int K, int N, __half* __restrict__ P) {
  for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
    const int k = i / N, n = i % N;
    P[((size_t)(n >> 3) * K + k) * 8 + (n & 7)] = w[k * N + perm_col(n)];
  }
}
// rows [row0, row0+128) of a row-major (rows, W) fp16 matrix -> panel [W/8][128][8].  Lane pairs fetch one 32-byte
// sector (row r, chunks 2j and 2j+1); rows >= n are zero filled by the copy engine
template <int W>
__device__ __forceinline__ void stage_rows(const __half* __restrict__ src, int64_t row0, int64_t n, __half* __restrict__ P) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = (tid >> 1) + 64 * h;
    const bool ok = row0 + r < n;
    const __half* g = src + (ok ? (row0 + r) : row0) * W + (tid & 1) * 8;
    __half* d = P + ((size_t)(tid & 1) * kTile + r) * 8;
#pragma unroll
    for (int j = 0; j < W / 16; ++j) cp_async16_zfill(d + (size_t)j * 2 * kTile * 8, g + j * 16, ok);
  }
}

}  // namespace ncn
