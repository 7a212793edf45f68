/* TEST INFRASTRUCTURE ONLY - plain C restatement of the reference's ray/AABB test and train-time
 * occupancy-grid march, used as the CPU oracle and as the CPU baseline of bench.py.
 *
 *   ncn_oracle_aabb         intersection.cu:5-56   (one box, max_hits=1) + near clamp rendering.py:28
 *   ncn_oracle_march_train  raymarching.cu:166-280 (count pass + write pass merged; rows in ray order)
 *   ncn_oracle_packbits     raymarching.cu:122-141
 *   ncn_oracle_morton3d     raymarching.cu:35-60
 *
 * fp32 rounding: the reference is compiled by nvcc with FMA contraction; the contracted operations (verified
 * in the SASS of the reference build) are written as fmaf() here, everything else is a plain
 * single-precision operation.  Build with: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC
 * (oracle/build_oracle.py) so the compiler neither fuses nor re-associates.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define SQRT3 1.73205080757f

static uint32_t expand_bits(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
static uint32_t morton3d(uint32_t x, uint32_t y, uint32_t z) {
  return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
void ncn_oracle_morton3d(const int32_t* coords, int64_t n, int32_t* out) {
  for (int64_t i = 0; i < n; ++i)
    out[i] = (int32_t)morton3d((uint32_t)coords[3 * i], (uint32_t)coords[3 * i + 1], (uint32_t)coords[3 * i + 2]);
}
void ncn_oracle_packbits(const float* grid, int64_t n_bytes, float thr, uint8_t* out) {
  for (int64_t n = 0; n < n_bytes; ++n) {
    uint8_t bits = 0;
    for (int i = 0; i < 8; ++i) bits |= (grid[8 * n + i] > thr) ? (uint8_t)(1u << i) : 0;
    out[n] = bits;
  }
}

static float clampf(float v, float lo, float hi) { return fmaxf(lo, fminf(v, hi)); }

static float calc_dt(float t, float esf, int max_samples, int grid_size, float scale) {
  volatile float hi_num = scale * 3.4641015529632568359f; /* SQRT3*2 folded by the compiler, then *scale */
  return clampf(t * esf, SQRT3 / (float)max_samples, hi_num / (float)grid_size);
}
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
static int mip_from_pos(float x, float y, float z, int cascades) {
  const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
  int e; frexpf(mx, &e);
  return imin(cascades - 1, imax(0, e + 1));
}
static int mip_from_dt(float dt, int grid_size, int cascades) {
  int e; frexpf(dt * (float)grid_size, &e);
  return imin(cascades - 1, imax(0, e));
}

/* hits_t (R,2); near < 0 disables the clamp */
void ncn_oracle_aabb(const float* rays_o, const float* rays_d, const float* center, const float* half, float near_d,
                     int64_t n_rays, float* hits_t) {
  for (int64_t r = 0; r < n_rays; ++r) {
    float tmin[3], tmax[3];
    for (int c = 0; c < 3; ++c) {
      const float inv = 1.0f / rays_d[3 * r + c];
      const float a = ((center[c] - half[c]) - rays_o[3 * r + c]) * inv;
      const float b = ((center[c] + half[c]) - rays_o[3 * r + c]) * inv;
      tmin[c] = fminf(a, b); tmax[c] = fmaxf(a, b);
    }
    float t1 = fmaxf(fmaxf(tmin[0], tmin[1]), tmin[2]);
    float t2 = fminf(fminf(tmax[0], tmax[1]), tmax[2]);
    if (t1 > t2) { t1 = -1.f; t2 = -1.f; }
    float o1 = -1.f, o2 = -1.f;
    if (t2 > 0) { o1 = fmaxf(t1, 0.f); o2 = t2; if (near_d >= 0 && o1 >= 0 && o1 < near_d) o1 = near_d; }
    hits_t[2 * r] = o1; hits_t[2 * r + 1] = o2;
  }
}

/* Returns the total number of samples.  If xyzs == NULL only counts (rays_a is still filled).
 * rays_a (R,3) = [r, start, n]; sample arrays must hold `capacity` rows (extra samples are dropped). */
int64_t ncn_oracle_march_train(const float* rays_o, const float* rays_d, const float* hits_t, const uint8_t* bitfield,
                               int cascades, float scale, float esf, const float* noise, int grid_size, int max_samples,
                               int64_t n_rays, int64_t capacity, int64_t* rays_a, float* xyzs, float* dirs, float* deltas,
                               float* ts) {
  const uint32_t g3 = (uint32_t)grid_size * grid_size * grid_size;
  const float gs_inv = 1.0f / (float)grid_size, gs_f = (float)grid_size, gs_m1 = (float)grid_size - 1.0f;
  int64_t total = 0;
  for (int64_t r = 0; r < n_rays; ++r) {
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    float t = hits_t[2 * r];
    const float t2 = hits_t[2 * r + 1];
    if (t >= 0 && noise) t = fmaf(calc_dt(t, esf, max_samples, grid_size, scale), noise[r], t);
    int n = 0;
    const int64_t start = total;
    while (0 <= t && t < t2 && n < max_samples) {
      const float x = fmaf(dx, t, ox), y = fmaf(dy, t, oy), z = fmaf(dz, t, oz);
      const float dt = calc_dt(t, esf, max_samples, grid_size, scale);
      const int mip = imax(mip_from_pos(x, y, z, cascades), mip_from_dt(dt, grid_size, cascades));
      const float mip_bound = fminf(scalbnf(1.0f, mip - 1), scale);
      const float mb_inv = 1.0f / mip_bound;
      const int nx = (int)clampf((fmaf(x, mb_inv, 1.0f) * 0.5f) * gs_f, 0.0f, gs_m1);
      const int ny = (int)clampf((fmaf(y, mb_inv, 1.0f) * 0.5f) * gs_f, 0.0f, gs_m1);
      const int nz = (int)clampf((fmaf(z, mb_inv, 1.0f) * 0.5f) * gs_f, 0.0f, gs_m1);
      const uint32_t idx = (uint32_t)mip * g3 + morton3d((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
      const int occ = bitfield[idx / 8] & (1 << (idx % 8));
      if (occ) {
        const int64_t s = start + n;
        if (xyzs && s < capacity) {
          xyzs[3 * s] = x; xyzs[3 * s + 1] = y; xyzs[3 * s + 2] = z;
          dirs[3 * s] = dx; dirs[3 * s + 1] = dy; dirs[3 * s + 2] = dz;
          ts[s] = t; deltas[s] = dt;
        }
        t += dt; n++;
      } else {
        const float tx = fmaf(mip_bound, fmaf((((float)nx + 0.5f) + copysignf(0.5f, dx)) * gs_inv, 2.0f, -1.0f), -x) * ix;
        const float ty = fmaf(mip_bound, fmaf((((float)ny + 0.5f) + copysignf(0.5f, dy)) * gs_inv, 2.0f, -1.0f), -y) * iy;
        const float tz = fmaf(mip_bound, fmaf((((float)nz + 0.5f) + copysignf(0.5f, dz)) * gs_inv, 2.0f, -1.0f), -z) * iz;
        const float t_target = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
        do { t += calc_dt(t, esf, max_samples, grid_size, scale); } while (t < t_target);
      }
    }
    rays_a[3 * r] = r; rays_a[3 * r + 1] = start; rays_a[3 * r + 2] = n;
    total += n;
  }
  return total;
}
