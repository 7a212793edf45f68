// Training-batch sampling on the device: the index half of BaseDataset.__getitem__
// (datasets/base.py:94-173) and its target gather (:175-183) - replaces numpy sampling in DataLoader workers + the H2D copy
// of the batch by two small kernels, so a step needs NO host input at all once the images live in HBM.
//   strategies:
//     0 all_images_triang_patch   n_patches = R / p^2 patches, image per patch, corner per patch, p x p pixels each
//     1 same_image_triang_patch   one image for the whole batch
//     2 all_images_triang         n_tri = R / 3 triangles (x1, x2 = up, x3 = left), image per triangle
//     3 same_image_triang         one image
//   random_tr_poses: ncn_sample_random_pose_half appends the same pixels seen from generated poses (second half of the batch)
//   max_expand > 0 (triangle strategies, base.py:130-141): x1 moves `expand` rows down when that stays inside the image, x2
//   `expand` rows up when that stays inside, x3 `expand` pixels left when that stays in its row (python floor division)
// Reference quirk kept: the patch corner is drawn as an INDEX into valid_idx['patch_corners'] (0 <= c < (H-p+1)(W-p+1)) and
// that index - not the pixel id it points at - is what the patch offsets are added to (base.py:164-166).
// Random numbers: counter-based splitmix64 of (seed, item, draw); the seed lives in device memory and is advanced by the
// kernel itself, so CUDA-graph replays draw fresh batches.  (numpy's MT19937 stream is not reproduced: the parity is
// distributional - uniform images / corners / triangles - and exact for the index arithmetic.)
#include "ncn_common.cuh"

namespace ncn {

__device__ __forceinline__ uint64_t smp_hash(uint64_t seed, uint64_t i, uint32_t k) {      // splitmix64 finaliser
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (i * 4ull + k + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// unbiased enough for n << 2^32: multiply-shift of the high 32 bits
__device__ __forceinline__ int64_t smp_below(uint64_t h, int64_t n) { return (int64_t)(((h >> 32) * (uint64_t)n) >> 32); }

__global__ void __launch_bounds__(256)
sample_batch_kernel(int strategy, int64_t* __restrict__ seed_dev, int n_rays, int n_poses, int H, int W, int patch, int max_expand,
                    int64_t* __restrict__ img_idx, int64_t* __restrict__ pix_idx) {
  const uint64_t seed = (uint64_t)*seed_dev;
  const bool patches = strategy <= 1, same = (strategy & 1) != 0;
  const int group = patches ? patch * patch : 3;
  const int n_used = (n_rays / group) * group;
  const int64_t same_img = smp_below(smp_hash(seed, 0xFFFFFFFFull, 0), n_poses);
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += gridDim.x * blockDim.x) {
    if (r >= n_used) { img_idx[r] = 0; pix_idx[r] = 0; continue; }       // rays beyond the last whole patch / triangle (none when R % group == 0)
    const int q = r / group, j = r - q * group;
    const int64_t img = same ? same_img : smp_below(smp_hash(seed, (uint64_t)q, 0), n_poses);
    int64_t pix;
    if (patches) {
      const int64_t n_corners = (int64_t)(H - patch + 1) * (W - patch + 1);
      const int64_t c = smp_below(smp_hash(seed, (uint64_t)q, 1), n_corners);
      pix = c + (int64_t)(j / patch) * W + (j % patch);
    } else {
      const int64_t n_valid = (int64_t)(H - 2) * (W - 2);
      const int64_t t = smp_below(smp_hash(seed, (uint64_t)q, 1), n_valid);
      const int64_t y = 1 + t / (W - 2), x = 1 + t % (W - 2);
      int64_t x1 = y * W + x, x2 = x1 - W, x3 = x1 - 1;
      if (max_expand > 0) {
        const int64_t e = max_expand, NP = (int64_t)H * W;
        if (x1 + e * W < NP) x1 += e * W;
        if (x2 - e * W >= 0) x2 -= e * W;
        const int64_t x3n = x3 - e;
        const int64_t row_n = x3n >= 0 ? x3n / W : -((-x3n + W - 1) / W);      // floor division, as numpy's //
        if (row_n == x3 / W) x3 = x3n;
      }
      pix = j == 0 ? x1 : (j == 1 ? x2 : x3);
    }
    img_idx[r] = img; pix_idx[r] = pix;
  }
}

__global__ void sample_advance_kernel(int64_t* seed_dev) { *seed_dev += 1; }

// random_tr_poses (datasets/base.py:106-126, 148-159; train_nerf.py:169-172): the second half of the batch repeats the pixels of the
// first half, seen from GENERATED poses - one pose per patch / triangle (all_images_*) or one for the whole batch (same_image_*).
// Draw index 2 of the counter hash: independent of the image / corner draws (0, 1) of sample_batch_kernel for any seed.
__global__ void __launch_bounds__(256)
sample_random_half_kernel(int strategy, const int64_t* __restrict__ seed_dev, int n_gt, int n_random_poses, int64_t pose_offset, int group,
                          int64_t* __restrict__ img_idx, int64_t* __restrict__ pix_idx) {
  const uint64_t seed = (uint64_t)*seed_dev;
  const bool same = (strategy & 1) != 0;
  const int64_t same_pose = smp_below(smp_hash(seed, 0xFFFFFFFFull, 2), n_random_poses);
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_gt; r += gridDim.x * blockDim.x) {
    const int q = r / group;
    const int64_t pose = same ? same_pose : smp_below(smp_hash(seed, (uint64_t)q, 2), n_random_poses);
    img_idx[n_gt + r] = pose_offset + pose;
    pix_idx[n_gt + r] = pix_idx[r];
  }
}

// out[r, :] = table[img[r], pix[r], :]  (rows of `width` 4-byte words: rgb f32 x3, a depth f32, an int32 label ...)
__global__ void __launch_bounds__(256)
gather_pixels_kernel(const uint32_t* __restrict__ table, const int64_t* __restrict__ img_idx, const int64_t* __restrict__ pix_idx,
                     int64_t n, int64_t hw, int width, uint32_t* __restrict__ out) {
  const int64_t total = n * width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / width;
    const int c = (int)(i - r * width);
    out[i] = __ldg(table + (img_idx[r] * hw + pix_idx[r]) * width + c);
  }
}

}  // namespace ncn

using namespace ncn;

extern "C" int ncn_sample_ray_batch_ex(int strategy, int64_t* seed_dev, int n_rays, int n_poses, int height, int width, int patch_size,
                                       int max_expand, int64_t* img_idx, int64_t* pix_idx, ncn_stream_t stream);
extern "C" int ncn_sample_ray_batch(int strategy, int64_t* seed_dev, int n_rays, int n_poses, int height, int width, int patch_size,
                                    int64_t* img_idx, int64_t* pix_idx, ncn_stream_t stream) {
  return ncn_sample_ray_batch_ex(strategy, seed_dev, n_rays, n_poses, height, width, patch_size, 0, img_idx, pix_idx, stream);
}

extern "C" int ncn_sample_ray_batch_ex(int strategy, int64_t* seed_dev, int n_rays, int n_poses, int height, int width, int patch_size,
                                       int max_expand, int64_t* img_idx, int64_t* pix_idx, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_rays >= 0 && n_poses >= 1 && strategy >= 0 && strategy <= 3 && max_expand >= 0);
  if (strategy <= 1) NCN_CHECK_SIZE(patch_size >= 2 && height >= patch_size && width >= patch_size);
  else NCN_CHECK_SIZE(height >= 3 && width >= 3);
  if (n_rays == 0) return NCN_OK;
  NCN_CHECK_PTR(seed_dev); NCN_CHECK_PTR(img_idx); NCN_CHECK_PTR(pix_idx);
  sample_batch_kernel<<<(unsigned)ceil_div(n_rays, 256), 256, 0, as_stream(stream)>>>(strategy, seed_dev, n_rays, n_poses, height, width,
                                                                                     patch_size, max_expand, img_idx, pix_idx);
  NCN_LAUNCH_OK();
  sample_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(seed_dev);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_sample_random_pose_half(int strategy, const int64_t* seed_dev, int n_gt_rays, int n_random_poses, int64_t pose_offset,
                                           int patch_size, int64_t* img_idx, int64_t* pix_idx, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n_gt_rays >= 0 && n_random_poses >= 1 && pose_offset >= 0 && strategy >= 0 && strategy <= 3);
  if (strategy <= 1) NCN_CHECK_SIZE(patch_size >= 2);
  if (n_gt_rays == 0) return NCN_OK;
  NCN_CHECK_PTR(seed_dev); NCN_CHECK_PTR(img_idx); NCN_CHECK_PTR(pix_idx);
  const int group = strategy <= 1 ? patch_size * patch_size : 3;
  sample_random_half_kernel<<<(unsigned)ceil_div(n_gt_rays, 256), 256, 0, as_stream(stream)>>>(strategy, seed_dev, n_gt_rays, n_random_poses,
                                                                                               pose_offset, group, img_idx, pix_idx);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

extern "C" int ncn_gather_pixels(const void* table, const int64_t* img_idx, const int64_t* pix_idx, int64_t n, int64_t pixels_per_image,
                                 int words_per_pixel, void* out, ncn_stream_t stream) {
  NCN_CHECK_SIZE(n >= 0 && pixels_per_image >= 1 && words_per_pixel >= 1);
  if (n == 0) return NCN_OK;
  NCN_CHECK_PTR(table); NCN_CHECK_PTR(img_idx); NCN_CHECK_PTR(pix_idx); NCN_CHECK_PTR(out);
  gather_pixels_kernel<<<persistent_grid(n * words_per_pixel, 256, 8), 256, 0, as_stream(stream)>>>(
      (const uint32_t*)table, img_idx, pix_idx, n, pixels_per_image, words_per_pixel, (uint32_t*)out);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
