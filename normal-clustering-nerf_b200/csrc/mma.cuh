// Warp-level tensor-core helpers (mma.sync m16n8k16 fp16 -> fp32, ldmatrix).
// Fragment layouts (g = lane>>2, t = lane&3):
//   A (16x16, row): a0=(g, 2t..2t+1) a1=(g+8, 2t..) a2=(g, 2t+8..) a3=(g+8, 2t+8..)
//   B (16x8,  col): b0=(k=2t..2t+1, n=g) b1=(k=2t+8.., n=g)
//   C (16x8):       c0,c1=(g, 2t..2t+1)  c2,c3=(g+8, 2t..2t+1)
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace ncn {

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row_ptr) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row_ptr);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* smem_row_ptr) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row_ptr);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
               : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

}  // namespace ncn
