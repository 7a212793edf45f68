// 10-bit-per-axis Morton (Z-order) codes: the cell index of the 128^3 (<=1024^3) occupancy
// grid.  Same bit layout as the reference (raymarching.cu:35-60): x -> bits 0,3,6..,
// y -> bits 1,4,7.., z -> bits 2,5,8...  Usable from host code too (oracle-free unit tests).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define NCN_HD __host__ __device__ __forceinline__
#else
#define NCN_HD inline
#endif

namespace ncn {

// spread the low 10 bits of v so that bit i lands at bit 3i
NCN_HD uint32_t spread3(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
NCN_HD uint32_t morton3d(uint32_t x, uint32_t y, uint32_t z) {
  return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
// gather bits 0,3,6,... of x into the low 11 bits
NCN_HD uint32_t morton3d_invert(uint32_t x) {
  x &= 0x49249249u;
  x = (x | (x >> 2)) & 0xC30C30C3u;
  x = (x | (x >> 4)) & 0x0F00F00Fu;
  x = (x | (x >> 8)) & 0xFF0000FFu;
  x = (x | (x >> 16)) & 0x0000FFFFu;
  return x;
}

}  // namespace ncn
