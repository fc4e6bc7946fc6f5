// image.cuh -- image / gradient / depth pyramids of the dense pipeline on the device: SURVEY 8 f3.
// Replaces DenseKeyframe.compute_image_pyramid / compute_jacobian_pyramid and the disparity / depth
// sub-sampling of pyslam/pipelines/keyframes.py:30-46,59-72,92-114:
//   pyr_down_u8_kernel   cv2.pyrDown of an 8-bit image: 5x5 binomial kernel [1 4 6 4 1]^2 / 256 with
//                        BORDER_REFLECT_101, output ((w + 1) / 2, (h + 1) / 2), rounded (sum + 128) >> 8
//   u8_to_unit_kernel    im.astype(float) / 255
//   sobel_half_kernel    0.5 * cv2.Sobel(im, -1, 1, 0) and 0.5 * cv2.Sobel(im, -1, 0, 1): 3x3, BORDER_REFLECT_101
//   subsample2_kernel    a[0::2, 0::2] * scale  (disparity: scale 0.5 per level; depth: 1)
#pragma once
#include "common.cuh"

namespace bs {

BS_D int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

__global__ void __launch_bounds__(256) pyr_down_u8_kernel(const unsigned char* __restrict__ src, int w, int h,
                                                          unsigned char* __restrict__ dst, int wo, int ho) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (x >= wo || y >= ho) return;
  const int k[5] = {1, 4, 6, 4, 1};
  int sum = 0;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
    const int yy = reflect101(2 * y + dy, h);
    int row = 0;
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) row += k[dx + 2] * (int)src[(size_t)yy * w + reflect101(2 * x + dx, w)];
    sum += k[dy + 2] * row;
  }
  dst[(size_t)y * wo + x] = (unsigned char)((sum + 128) >> 8);
}

__global__ void __launch_bounds__(256) u8_to_unit_kernel(const unsigned char* __restrict__ src, size_t n, double* __restrict__ dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)src[i] / 255.0;
}

__global__ void __launch_bounds__(256) sobel_half_kernel(const double* __restrict__ im, int w, int h, double* __restrict__ gx,
                                                         double* __restrict__ gy) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (x >= w || y >= h) return;
  const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w), ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
  auto at = [&](int yy, int xx) { return im[(size_t)yy * w + xx]; };
  const double sx = (at(ym, xp) - at(ym, xm)) + 2.0 * (at(y, xp) - at(y, xm)) + (at(yp, xp) - at(yp, xm));
  const double sy = (at(yp, xm) - at(ym, xm)) + 2.0 * (at(yp, x) - at(ym, x)) + (at(yp, xp) - at(ym, xp));
  gx[(size_t)y * w + x] = 0.5 * sx;
  gy[(size_t)y * w + x] = 0.5 * sy;
}

__global__ void __launch_bounds__(256) subsample2_kernel(const double* __restrict__ src, int w, int h, double* __restrict__ dst, int wo,
                                                         int ho, double scale) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (x >= wo || y >= ho) return;
  dst[(size_t)y * wo + x] = src[(size_t)(2 * y) * w + 2 * x] * scale;
}

}  // namespace bs
