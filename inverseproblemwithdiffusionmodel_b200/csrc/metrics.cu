// Image-quality metrics on the device: the NRMSE / MSE / MAE sums and skimage-default SSIM that the reference
// computes with numpy + skimage on the host after every reconstruction (helpers/metrics.py:21-74).
// HBM-bound reductions: one read of each image, fp64 accumulation.
#include "common.cuh"

namespace ipdm {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums[b] = { sum (a-b)^2, sum a^2, sum b^2, sum |a-b| } over one image; grid (chunks, images)
__global__ void __launch_bounds__(256) k_image_sums(const float* __restrict__ a, const float* __restrict__ ref, double* __restrict__ sums,
                                                    size_t n, int ref_images) {
  const int img = blockIdx.y;
  const float* pa = a + (size_t)img * n;
  const float* pb = ref + (size_t)(ref_images == 1 ? 0 : img) * n;
  double s[4] = {0, 0, 0, 0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double x = pa[i], y = pb[i], d = x - y;
    s[0] += d * d;
    s[1] += x * x;
    s[2] += y * y;
    s[3] += fabs(d);
  }
  __shared__ double red[8][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double v = warp_sum(s[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    atomicAdd(&sums[(size_t)img * 4 + threadIdx.x], v);
  }
}

// skimage.metrics.structural_similarity defaults for one 2-D float image pair: 7x7 uniform window, sample
// covariance (n/(n-1)), K1 = 0.01, K2 = 0.03, mean of S over the pixels whose window lies inside the image.
// out[b] += sum of S over this CTA's 32x8 output tile; the host divides by (H-6)*(W-6).  grid (tiles_x, tiles_y, images)
constexpr int SS_W = 7, SS_TX = 32, SS_TY = 8;
__global__ void __launch_bounds__(SS_TX * SS_TY) k_ssim(const float* __restrict__ a, const float* __restrict__ ref, double* __restrict__ out,
                                                        int H, int W, int ref_images, double C1, double C2) {
  __shared__ float ta[SS_TY + SS_W - 1][SS_TX + SS_W - 1], tb[SS_TY + SS_W - 1][SS_TX + SS_W - 1];
  const int img = blockIdx.z, x0 = blockIdx.x * SS_TX, y0 = blockIdx.y * SS_TY;
  const float* pa = a + (size_t)img * H * W;
  const float* pb = ref + (size_t)(ref_images == 1 ? 0 : img) * H * W;
  for (int i = threadIdx.x; i < (SS_TY + SS_W - 1) * (SS_TX + SS_W - 1); i += SS_TX * SS_TY) {
    const int ty = i / (SS_TX + SS_W - 1), tx = i % (SS_TX + SS_W - 1);
    const int y = y0 + ty, x = x0 + tx;
    const bool in = y < H && x < W;
    ta[ty][tx] = in ? pa[(size_t)y * W + x] : 0.f;
    tb[ty][tx] = in ? pb[(size_t)y * W + x] : 0.f;
  }
  __syncthreads();
  const int lx = threadIdx.x % SS_TX, ly = threadIdx.x / SS_TX;
  double S = 0.0;
  if (x0 + lx < W - (SS_W - 1) && y0 + ly < H - (SS_W - 1)) {
    double sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
#pragma unroll
    for (int dy = 0; dy < SS_W; ++dy)
#pragma unroll
      for (int dx = 0; dx < SS_W; ++dx) {
        const double u = ta[ly + dy][lx + dx], v = tb[ly + dy][lx + dx];
        sa += u; sb += v; saa += u * u; sbb += v * v; sab += u * v;
      }
    const double n = SS_W * SS_W, cn = n / (n - 1.0);
    const double ua = sa / n, ub = sb / n;
    const double va = cn * (saa / n - ua * ua), vb = cn * (sbb / n - ub * ub), vab = cn * (sab / n - ua * ub);
    S = ((2 * ua * ub + C1) * (2 * vab + C2)) / ((ua * ua + ub * ub + C1) * (va + vb + C2));
  }
  __shared__ double red[SS_TX * SS_TY / 32];
  const double v = warp_sum(S);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < SS_TX * SS_TY / 32; ++w) t += red[w];
    atomicAdd(&out[img], t);
  }
}

}  // namespace ipdm

using namespace ipdm;

extern "C" int ipdm_image_sums(const float* img, const float* ref, double* sums, int images, int ref_images, size_t n, void* stream) {
  IPDM_REQUIRE(img && ref && sums, IPDM_E_BADARG, "image_sums: null pointer");
  IPDM_REQUIRE(images >= 1 && (ref_images == 1 || ref_images == images) && n >= 1, IPDM_E_BADARG, "image_sums: bad shape");
  cudaStream_t s = as_stream(stream);
  IPDM_CUDA(cudaMemsetAsync(sums, 0, (size_t)images * 4 * sizeof(double), s));
  size_t chunks = (n + 256 * 16 - 1) / (256 * 16);
  if (chunks > 64) chunks = 64;
  k_image_sums<<<dim3((unsigned)chunks, images), 256, 0, s>>>(img, ref, sums, n, ref_images);
  return launched("k_image_sums");
}

extern "C" int ipdm_ssim(const float* img, const float* ref, double* out, int images, int ref_images, int H, int W,
                         double data_range, void* stream) {
  IPDM_REQUIRE(img && ref && out, IPDM_E_BADARG, "ssim: null pointer");
  IPDM_REQUIRE(images >= 1 && (ref_images == 1 || ref_images == images), IPDM_E_BADARG, "ssim: bad image count");
  IPDM_REQUIRE(H >= SS_W && W >= SS_W, IPDM_E_UNSUPPORTED, "ssim: image %dx%d is smaller than the 7x7 window", H, W);
  IPDM_REQUIRE(data_range > 0.0, IPDM_E_BADARG, "ssim: data_range must be given and positive");
  cudaStream_t s = as_stream(stream);
  IPDM_CUDA(cudaMemsetAsync(out, 0, (size_t)images * sizeof(double), s));
  const double C1 = (0.01 * data_range) * (0.01 * data_range), C2 = (0.03 * data_range) * (0.03 * data_range);
  dim3 grid((W - SS_W + 1 + SS_TX - 1) / SS_TX, (H - SS_W + 1 + SS_TY - 1) / SS_TY, images);
  k_ssim<<<grid, SS_TX * SS_TY, 0, s>>>(img, ref, out, H, W, ref_images, C1, C2);
  return launched("k_ssim");
}
