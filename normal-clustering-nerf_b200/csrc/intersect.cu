// Ray / volume intersection.  Replaces intersection.cu:5-197 of the reference
// (vren.ray_aabb_intersect, vren.ray_sphere_intersect) and fuses the near clamp of
// models/rendering.py:28 into the one-box fast path used by render().
//
// Arithmetic is written with explicit round-to-nearest intrinsics in the reference's
// operation order ((c -/+ h) - o) * (1/d), so hits_t is bit-exact with the reference.
// The reference fills hit slots through an atomicAdd race and then torch::sort's the
// slots by t1 (so unfilled -1 slots come FIRST); here one thread owns a ray, walks the
// volumes in index order (one of the orders the race can produce) and emits the same
// sorted layout directly: no fill kernels, no sort, no gathers (8 launches -> 1).
//
// B200 notes: 44 B/ray of traffic, pure streaming; rays are read as 3 consecutive
// floats per thread (the (R,3) layout is fixed by the boundary), which the L1 coalesces
// into full 128 B lines per warp.
#include "ncn_common.cuh"

namespace ncn {

struct Hit { float t1, t2; };

// slab test; returns t2 <= 0 style "no hit" exactly like the reference: (-1,-1) when t1 > t2
__device__ __forceinline__ Hit aabb_hit(float ox, float oy, float oz, float ix, float iy, float iz,
                                        float cx, float cy, float cz, float hx, float hy, float hz) {
  const float tminx = __fmul_rn(__fsub_rn(__fsub_rn(cx, hx), ox), ix);
  const float tminy = __fmul_rn(__fsub_rn(__fsub_rn(cy, hy), oy), iy);
  const float tminz = __fmul_rn(__fsub_rn(__fsub_rn(cz, hz), oz), iz);
  const float tmaxx = __fmul_rn(__fsub_rn(__fadd_rn(cx, hx), ox), ix);
  const float tmaxy = __fmul_rn(__fsub_rn(__fadd_rn(cy, hy), oy), iy);
  const float tmaxz = __fmul_rn(__fsub_rn(__fadd_rn(cz, hz), oz), iz);
  const float t1 = fmaxf(fmaxf(fminf(tminx, tmaxx), fminf(tminy, tmaxy)), fminf(tminz, tmaxz));
  const float t2 = fminf(fminf(fmaxf(tminx, tmaxx), fmaxf(tminy, tmaxy)), fmaxf(tminz, tmaxz));
  Hit h;
  if (t1 > t2) { h.t1 = -1.f; h.t2 = -1.f; } else { h.t1 = t1; h.t2 = t2; }
  return h;
}

// intersection.cu:103-121
__device__ __forceinline__ Hit sphere_hit(float ox, float oy, float oz, float dx, float dy, float dz,
                                          float cx, float cy, float cz, float radius) {
  const float cox = __fsub_rn(ox, cx), coy = __fsub_rn(oy, cy), coz = __fsub_rn(oz, cz);
  // dot(a,b) = a.x*b.x + a.y*b.y + a.z*b.z is contracted by nvcc (reference build, SASS checked)
  // to fma(z, z', fma(x, x', y*y'))
  const float a = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
  const float half_b = __fmaf_rn(dz, coz, __fmaf_rn(dx, cox, __fmul_rn(dy, coy)));
  const float cc = __fmaf_rn(coz, coz, __fmaf_rn(cox, cox, __fmul_rn(coy, coy)));
  const float c = __fmaf_rn(-radius, radius, cc);          // dot(co,co) - radius*radius
  const float disc = __fmaf_rn(half_b, half_b, -__fmul_rn(a, c));  // half_b*half_b - a*c
  Hit h;
  if (disc < 0.f) { h.t1 = -1.f; h.t2 = -1.f; return h; }
  const float s = __fsqrt_rn(disc);
  h.t1 = __fdiv_rn(__fsub_rn(-half_b, s), a);
  h.t2 = __fdiv_rn(__fadd_rn(-half_b, s), a);
  return h;
}

// Fast path: one box, one slot, optional near clamp (near < 0 disables it).
__global__ void __launch_bounds__(256)
aabb_one_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                const float* __restrict__ center, const float* __restrict__ half_size,
                float near_distance, int64_t n_rays, int32_t* __restrict__ hit_cnt,
                float* __restrict__ hits_t, int64_t* __restrict__ hits_idx) {
  const float cx = center[0], cy = center[1], cz = center[2];
  const float hx = half_size[0], hy = half_size[1], hz = half_size[2];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += stride) {
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const Hit h = aabb_hit(ox, oy, oz, __frcp_rn(dx), __frcp_rn(dy), __frcp_rn(dz), cx, cy, cz, hx, hy, hz);
    float t1 = -1.f, t2 = -1.f;
    int64_t idx = -1;
    int cnt = 0;
    if (h.t2 > 0.f) {
      t1 = fmaxf(h.t1, 0.f); t2 = h.t2; idx = 0; cnt = 1;
      if (near_distance >= 0.f && t1 >= 0.f && t1 < near_distance) t1 = near_distance;  // rendering.py:28
    }
    reinterpret_cast<float2*>(hits_t)[r] = make_float2(t1, t2);
    if (hit_cnt) hit_cnt[r] = cnt;
    if (hits_idx) hits_idx[r] = idx;
  }
}

// Generic path: V volumes, max_hits slots.  kSphere selects the quadratic test.
template <bool kSphere>
__global__ void __launch_bounds__(256)
intersect_many_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                      const float* __restrict__ centers, const float* __restrict__ extents,
                      int64_t n_rays, int64_t n_vol, int max_hits, int32_t* __restrict__ hit_cnt,
                      float* __restrict__ hits_t, int64_t* __restrict__ hits_idx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += stride) {
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float ix = __frcp_rn(dx), iy = __frcp_rn(dy), iz = __frcp_rn(dz);
    float2* ht = reinterpret_cast<float2*>(hits_t) + r * max_hits;
    int64_t* hi = hits_idx + r * max_hits;
    int cnt = 0;
    for (int64_t v = 0; v < n_vol; ++v) {
      Hit h;
      if (kSphere) h = sphere_hit(ox, oy, oz, dx, dy, dz, centers[3 * v], centers[3 * v + 1], centers[3 * v + 2], extents[v]);
      else h = aabb_hit(ox, oy, oz, ix, iy, iz, centers[3 * v], centers[3 * v + 1], centers[3 * v + 2],
                        extents[3 * v], extents[3 * v + 1], extents[3 * v + 2]);
      if (h.t2 > 0.f) {
        if (cnt < max_hits) {
          // stable insertion by t1 into the first `cnt` slots
          const float t1 = fmaxf(h.t1, 0.f);
          int j = cnt;
          while (j > 0 && ht[j - 1].x > t1) { ht[j] = ht[j - 1]; hi[j] = hi[j - 1]; --j; }
          ht[j] = make_float2(t1, h.t2); hi[j] = v;
        }
        ++cnt;
      }
    }
    hit_cnt[r] = cnt;
    // ascending sort over ALL slots puts the unfilled (-1) ones first: shift hits to the back
    const int k = cnt < max_hits ? cnt : max_hits;
    const int pad = max_hits - k;
    if (pad > 0) {
      for (int j = k - 1; j >= 0; --j) { ht[j + pad] = ht[j]; hi[j + pad] = hi[j]; }
      for (int j = 0; j < pad; ++j) { ht[j] = make_float2(-1.f, -1.f); hi[j] = -1; }
    }
  }
}

}  // namespace ncn

using namespace ncn;

static int intersect_many(bool sphere, const float* rays_o, const float* rays_d, const float* centers,
                          const float* extents, int64_t n_rays, int64_t n_vol, int max_hits,
                          int32_t* hit_cnt, float* hits_t, int64_t* hits_idx, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_rays >= 0 && n_vol >= 0 && max_hits >= 1);
  if (n_rays == 0) return NCN_OK;
  NCN_CHECK_PTR(rays_o); NCN_CHECK_PTR(rays_d); NCN_CHECK_PTR(hit_cnt); NCN_CHECK_PTR(hits_t); NCN_CHECK_PTR(hits_idx);
  if (n_vol > 0) { NCN_CHECK_PTR(centers); NCN_CHECK_PTR(extents); }
  if ((uintptr_t)hits_t & 7) return NCN_E_ALIGN;
  const int grid = persistent_grid(n_rays, 256, 8);
  if (!sphere && n_vol == 1 && max_hits == 1) {
    aabb_one_kernel<<<grid, 256, 0, as_stream(stream)>>>(rays_o, rays_d, centers, extents, -1.f, n_rays,
                                                         hit_cnt, hits_t, hits_idx);
  } else if (sphere) {
    intersect_many_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(rays_o, rays_d, centers, extents, n_rays,
                                                                     n_vol, max_hits, hit_cnt, hits_t, hits_idx);
  } else {
    intersect_many_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(rays_o, rays_d, centers, extents, n_rays,
                                                                      n_vol, max_hits, hit_cnt, hits_t, hits_idx);
  }
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_ray_aabb_intersect(const float* rays_o, const float* rays_d, const float* centers,
                                      const float* half_sizes, int64_t n_rays, int64_t n_boxes, int max_hits,
                                      int32_t* hit_cnt, float* hits_t, int64_t* hits_idx, ncn_stream_t stream) {
  return intersect_many(false, rays_o, rays_d, centers, half_sizes, n_rays, n_boxes, max_hits, hit_cnt, hits_t,
                        hits_idx, stream);
}

extern "C" int ncn_ray_sphere_intersect(const float* rays_o, const float* rays_d, const float* centers,
                                        const float* radii, int64_t n_rays, int64_t n_spheres, int max_hits,
                                        int32_t* hit_cnt, float* hits_t, int64_t* hits_idx, ncn_stream_t stream) {
  return intersect_many(true, rays_o, rays_d, centers, radii, n_rays, n_spheres, max_hits, hit_cnt, hits_t,
                        hits_idx, stream);
}

extern "C" int ncn_ray_aabb_near(const float* rays_o, const float* rays_d, const float* center,
                                 const float* half_size, float near_distance, int64_t n_rays, float* hits_t,
                                 ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_rays >= 0);
  if (n_rays == 0) return NCN_OK;
  NCN_CHECK_PTR(rays_o); NCN_CHECK_PTR(rays_d); NCN_CHECK_PTR(center); NCN_CHECK_PTR(half_size); NCN_CHECK_PTR(hits_t);
  if ((uintptr_t)hits_t & 7) return NCN_E_ALIGN;
  const int grid = persistent_grid(n_rays, 256, 8);
  aabb_one_kernel<<<grid, 256, 0, as_stream(stream)>>>(rays_o, rays_d, center, half_size, near_distance, n_rays,
                                                       nullptr, hits_t, nullptr);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
