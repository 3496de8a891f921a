// CUDA-core kernels of the NCSNv2 score network (everything that is not the tensor-core implicit
// GEMM): first / last 3x3 convolutions with one channel on one side, InstanceNorm++ statistics and
// apply (+ELU, f16 operand store), ELU / casts, 5x5 max-pool, bilinear (align_corners) accumulate,
// 2x2 mean-pool, weight repack, and a direct convolution with the igemm's exact epilogue contract.
// Activations are NHWC; fp32 for the residual streams, f16 for the tensor-core operands.
#include "common.cuh"

namespace ipdm {

static int grid1d(size_t n, int block, int cap_mult = 32) {
  size_t g = (n + block - 1) / block;
  const size_t cap = (size_t)148 * cap_mult;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---------------------------------------------------------------------------- begin_conv (Cin = 1)
// thread = (pixel, 4 output channels); the 9 taps of the pixel are shared by the C/4 threads of it.
__global__ void k_conv_first(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                             float* __restrict__ out, int N, int H, int W, int Cout, int affine) {
  extern __shared__ float sw[];  // [9][Cout] + [Cout]
  for (int i = threadIdx.x; i < 9 * Cout; i += blockDim.x) {
    const int co = i % Cout, tap = i / Cout;
    sw[i] = w[co * 9 + tap];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[9 * Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int groups = Cout / 4;
  const size_t total = (size_t)N * H * W * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const size_t pix = i / groups;
    const int xw = (int)(pix % W), yh = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
    float4 acc = *reinterpret_cast<const float4*>(&sw[9 * Cout + 4 * g]);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = yh + ky - 1, xx = xw + kx - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        float v = x[((size_t)n * H + yy) * W + xx];
        if (affine) v = 2.f * v - 1.f;
        const float4 ww = *reinterpret_cast<const float4*>(&sw[(ky * 3 + kx) * Cout + 4 * g]);
        acc.x += v * ww.x; acc.y += v * ww.y; acc.z += v * ww.z; acc.w += v * ww.w;
      }
    *reinterpret_cast<float4*>(&out[pix * Cout + 4 * g]) = acc;
  }
}

// ---------------------------------------------------------------------------- end_conv (Cout = 1)
// 16 lanes per pixel, each lane 8 channels (one 16-byte load) per tap.
__global__ void k_conv_last(const __half* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                            const float* __restrict__ sigmas, const int64_t* __restrict__ labels,
                            float* __restrict__ out, int N, int H, int W, int Cin) {
  extern __shared__ float sw[];  // [9][Cin]
  for (int i = threadIdx.x; i < 9 * Cin; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int lane16 = threadIdx.x & 15;
  const size_t npix = (size_t)N * H * W;
  const size_t gstride = (size_t)gridDim.x * (blockDim.x / 16);
  // all 16 lanes of a group (and both groups of a warp) iterate together: pad the loop so shuffles stay converged
  const size_t iters = (npix + gstride - 1) / gstride;
  size_t pix = blockIdx.x * (size_t)(blockDim.x / 16) + threadIdx.x / 16;
  for (size_t it = 0; it < iters; ++it, pix += gstride) {
    const bool live = pix < npix;
    float acc = 0.f;
    int n = 0;
    if (live) {
      const int xw = (int)(pix % W), yh = (int)((pix / W) % H);
      n = (int)(pix / ((size_t)W * H));
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = yh + ky - 1, xx = xw + kx - 1;
          if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
          const __half* p = in + (((size_t)n * H + yy) * W + xx) * Cin;
          const float* wt = sw + (ky * 3 + kx) * Cin;
          for (int c0 = lane16 * 8; c0 < Cin; c0 += 128) {
            const uint4 raw = *reinterpret_cast<const uint4*>(p + c0);
            const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __half22float2(h2[j]);
              acc += f.x * wt[c0 + 2 * j] + f.y * wt[c0 + 2 * j + 1];
            }
          }
        }
    }
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (live && lane16 == 0) out[pix] = (acc + (bias ? bias[0] : 0.f)) / sigmas[labels[n]];
  }
}

// ---------------------------------------------------------------------------- InstanceNorm++
// stats[n][c] = (sum(x - p), sum((x - p)^2)) over HW, p = x[n,0,c]  (pivot kills the cancellation
// in E[x^2] - E[x]^2).  grid (chunks, N), block 256 = (C/4 lanes) x (256/(C/4) pixel rows).
__global__ void k_instnorm_stats(const float* __restrict__ x, float* __restrict__ stats, int HW, int C, int pivoted) {
  const int n = blockIdx.y;
  const int lanes = C / 4;
  const int rows = blockDim.x / lanes;
  const int lane = threadIdx.x % lanes, row = threadIdx.x / lanes;
  const float* base = x + (size_t)n * HW * C;
  float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
  if (row < rows) {
    const float4 p = pivoted ? *reinterpret_cast<const float4*>(base + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int pix = blockIdx.x * rows + row; pix < HW; pix += gridDim.x * rows) {
      float4 v = *reinterpret_cast<const float4*>(base + (size_t)pix * C + 4 * lane);
      v.x -= p.x; v.y -= p.y; v.z -= p.z; v.w -= p.w;
      s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
      s2.x += v.x * v.x; s2.y += v.y * v.y; s2.z += v.z * v.z; s2.w += v.w * v.w;
    }
  }
  extern __shared__ float red[];  // [rows][C][2]
  if (row < rows) {
    float* r = red + ((size_t)row * C + 4 * lane) * 2;
    r[0] = s1.x; r[1] = s2.x; r[2] = s1.y; r[3] = s2.y; r[4] = s1.z; r[5] = s2.z; r[6] = s1.w; r[7] = s2.w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float t = 0.f;
    for (int rr = 0; rr < rows; ++rr) t += red[(size_t)rr * C * 2 + i];
    atomicAdd(&stats[(size_t)n * C * 2 + i], t);
  }
}

// out = f16(ELU(gamma*((x-m)*rstd + alpha*m_hat) + beta)); per-(n,c) A,B precomputed in smem.
__global__ void k_instnorm_apply(const float* __restrict__ x, const float* __restrict__ stats, int pivoted,
                                 const float* __restrict__ alpha, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, __half* __restrict__ out, int HW, int C) {
  extern __shared__ float sm[];  // A[C], B[C], mean[C], red[64]
  float* A = sm;
  float* Bv = sm + C;
  float* mean = sm + 2 * C;
  float* red = sm + 3 * C;
  const int n = blockIdx.y;
  const float* base = x + (size_t)n * HW * C;
  const float inv = 1.0f / (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float p = pivoted ? base[c] : 0.f;
    const float d = stats[((size_t)n * C + c) * 2] * inv;
    mean[c] = p + d;
    const float var = fmaxf(stats[((size_t)n * C + c) * 2 + 1] * inv - d * d, 0.f);
    A[c] = rsqrtf(var + 1e-5f);  // rstd for now
  }
  __syncthreads();
  // mean and unbiased variance of the channel means (two-pass, block reduce)
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += mean[c];
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float mu = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) mu += red[i];
  mu /= (float)C;
  __syncthreads();
  float q = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) q += (mean[c] - mu) * (mean[c] - mu);
  for (int off = 16; off >= 1; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  float vv = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) vv += red[i];
  const float rs = rsqrtf(vv / (float)(C - 1) + 1e-5f);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float rstd = A[c];
    const float mhat = (mean[c] - mu) * rs;
    const float g = gamma[c];
    A[c] = g * rstd;
    Bv[c] = g * (alpha[c] * mhat - mean[c] * rstd) + (beta ? beta[c] : 0.f);
  }
  __syncthreads();
  const int lanes = C / 4;
  const size_t total = (size_t)HW * lanes;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % lanes) * 4;
    const float4 v = *reinterpret_cast<const float4*>(base + (i / lanes) * C + c4);
    const float a0 = elu1(v.x * A[c4] + Bv[c4]), a1 = elu1(v.y * A[c4 + 1] + Bv[c4 + 1]);
    const float a2 = elu1(v.z * A[c4 + 2] + Bv[c4 + 2]), a3 = elu1(v.w * A[c4 + 3] + Bv[c4 + 3]);
    __half2 lo = __floats2half2_rn(a0, a1), hi = __floats2half2_rn(a2, a3);
    uint2 pk;
    pk.x = *reinterpret_cast<unsigned*>(&lo);
    pk.y = *reinterpret_cast<unsigned*>(&hi);
    *reinterpret_cast<uint2*>(out + (size_t)n * HW * C + (i / lanes) * C + c4) = pk;
  }
}

// ---------------------------------------------------------------------------- elementwise
__global__ void k_act_to_f16(const float* __restrict__ x, __half* __restrict__ out, size_t n4, int elu) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    if (elu) { v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w); }
    __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<unsigned*>(&lo);
    pk.y = *reinterpret_cast<unsigned*>(&hi);
    reinterpret_cast<uint2*>(out)[i] = pk;
  }
}

// thread = (pixel, 8 channels)
__global__ void k_maxpool5(const __half* __restrict__ in, __half* __restrict__ out, int N, int H, int W, int C) {
  const int groups = C / 8;
  const size_t total = (size_t)N * H * W * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const size_t pix = i / groups;
    const int xw = (int)(pix % W), yh = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
    __half2 m[4];
    const __half2 ninf = __float2half2_rn(-INFINITY);
    m[0] = m[1] = m[2] = m[3] = ninf;
    for (int dy = -2; dy <= 2; ++dy) {
      const int yy = yh + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -2; dx <= 2; ++dx) {
        const int xx = xw + dx;
        if (xx < 0 || xx >= W) continue;
        const uint4 raw = *reinterpret_cast<const uint4*>(in + (((size_t)n * H + yy) * W + xx) * C + 8 * g);
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], h2[j]);
      }
    }
    uint4 o;
    o.x = *reinterpret_cast<unsigned*>(&m[0]); o.y = *reinterpret_cast<unsigned*>(&m[1]);
    o.z = *reinterpret_cast<unsigned*>(&m[2]); o.w = *reinterpret_cast<unsigned*>(&m[3]);
    *reinterpret_cast<uint4*>(out + pix * C + 8 * g) = o;
  }
}

// dst (+)= bilinear(src), align_corners=True, PyTorch's index arithmetic (upsample_bilinear2d).
__global__ void k_bilinear_add(const float* __restrict__ src, float* __restrict__ dst, __half* __restrict__ out16, int N,
                               int h, int w, int H, int W, int C, int accumulate) {
  const int lanes = C / 4;
  const size_t total = (size_t)N * H * W * lanes;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % lanes) * 4;
    const size_t pix = i / lanes;
    const int X = (int)(pix % W), Y = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
    const float fy = sy * Y, fx = sx * X;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* b = src + (size_t)n * h * w * C + c4;
    const float4 v00 = *reinterpret_cast<const float4*>(b + ((size_t)y0 * w + x0) * C);
    const float4 v01 = *reinterpret_cast<const float4*>(b + ((size_t)y0 * w + x1) * C);
    const float4 v10 = *reinterpret_cast<const float4*>(b + ((size_t)y1 * w + x0) * C);
    const float4 v11 = *reinterpret_cast<const float4*>(b + ((size_t)y1 * w + x1) * C);
    float4 r;
    r.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
    r.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    r.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
    r.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
    float4* d = reinterpret_cast<float4*>(dst + pix * C + c4);
    if (accumulate) {
      const float4 o = *d;
      r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
    }
    *d = r;
    if (out16) {
      __half2 lo = __floats2half2_rn(elu1(r.x), elu1(r.y)), hi = __floats2half2_rn(elu1(r.z), elu1(r.w));
      uint2 pk;
      pk.x = *reinterpret_cast<unsigned*>(&lo);
      pk.y = *reinterpret_cast<unsigned*>(&hi);
      *reinterpret_cast<uint2*>(out16 + pix * C + c4) = pk;
    }
  }
}

__global__ void k_meanpool2(const float* __restrict__ in, const float* __restrict__ add, float* __restrict__ out, int N,
                            int H, int W, int C) {
  const int lanes = C / 4, Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * lanes;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % lanes) * 4;
    const size_t pix = i / lanes;
    const int X = (int)(pix % Wo), Y = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
    const float* b = in + (((size_t)n * H + 2 * Y) * W + 2 * X) * C + c4;
    const float4 a00 = *reinterpret_cast<const float4*>(b);
    const float4 a01 = *reinterpret_cast<const float4*>(b + C);
    const float4 a10 = *reinterpret_cast<const float4*>(b + (size_t)W * C);
    const float4 a11 = *reinterpret_cast<const float4*>(b + (size_t)W * C + C);
    float4 r;
    r.x = (((a00.x + a10.x) + a01.x) + a11.x) * 0.25f;
    r.y = (((a00.y + a10.y) + a01.y) + a11.y) * 0.25f;
    r.z = (((a00.z + a10.z) + a01.z) + a11.z) * 0.25f;
    r.w = (((a00.w + a10.w) + a01.w) + a11.w) * 0.25f;
    if (add) {
      const float4 o = *reinterpret_cast<const float4*>(add + pix * C + c4);
      r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
    }
    *reinterpret_cast<float4*>(out + pix * C + c4) = r;
  }
}

__global__ void k_pack_weights(const float* __restrict__ w, __half* __restrict__ out, int Cout, int Cin, int taps) {
  const size_t total = (size_t)Cout * taps * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin), tap = (int)((i / Cin) % taps), co = (int)(i / ((size_t)Cin * taps));
    out[i] = __float2half_rn(w[((size_t)co * Cin + ci) * taps + tap]);
  }
}

// ---------------------------------------------------------------------------- direct convolution
// One thread per (output pixel, output channel); same epilogue contract as the igemm.
__device__ __forceinline__ float conv_at(const ipdm_conv_desc& d, int n, int y, int x, int co) {
  const __half* in = reinterpret_cast<const __half*>(d.in_f16);
  const __half* wt = reinterpret_cast<const __half*>(d.w_f16) + (size_t)co * d.taps * d.Cin;
  float acc = 0.f;
  for (int tap = 0; tap < d.taps; ++tap) {
    const int ky = d.taps == 9 ? tap / 3 - 1 : 0, kx = d.taps == 9 ? tap % 3 - 1 : 0;
    const int yy = y + ky * d.dilation, xx = x + kx * d.dilation;
    if (yy < 0 || yy >= d.H || xx < 0 || xx >= d.W) continue;
    const __half* p = in + (((size_t)n * d.H + yy) * d.W + xx) * d.Cin;
    const __half* q = wt + (size_t)tap * d.Cin;
    for (int ci = 0; ci < d.Cin; ++ci) acc += __half2float(p[ci]) * __half2float(q[ci]);
  }
  return acc;
}

__global__ void k_conv_direct(ipdm_conv_desc d) {
  const bool pool = (d.flags & IPDM_CONV_POOL2) != 0;
  const int Ho = pool ? d.H / 2 : d.H, Wo = pool ? d.W / 2 : d.W;
  const size_t total = (size_t)d.N * Ho * Wo * d.Cout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % d.Cout);
    const size_t pix = i / d.Cout;
    const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
    float acc;
    if (pool) {
      acc = (((conv_at(d, n, 2 * y, 2 * x, co) + conv_at(d, n, 2 * y + 1, 2 * x, co)) + conv_at(d, n, 2 * y, 2 * x + 1, co)) +
             conv_at(d, n, 2 * y + 1, 2 * x + 1, co)) * 0.25f;
    } else {
      acc = conv_at(d, n, y, x, co);
    }
    float v = acc + (d.bias ? d.bias[co] : 0.f);
    const float pre = v;
    if (d.residual) {
      float r = d.residual[i];
      if (d.flags & IPDM_CONV_RES_ELU) r = elu1(r);
      v += r;
    }
    if (d.out_f32) d.out_f32[i] = v;
    if (d.out_f16) {
      float s = (d.flags & IPDM_CONV_F16_PRE_RES) ? pre : v;
      if (d.flags & IPDM_CONV_F16_ELU) s = elu1(s);
      reinterpret_cast<__half*>(d.out_f16)[i] = __float2half_rn(s);
    }
  }
}

int stats_after_conv(const ipdm_conv_desc& d, cudaStream_t s);

}  // namespace ipdm

using namespace ipdm;

extern "C" int ipdm_instnorm_stats(const float* x, float* stats, int N, int HW, int C, int pivoted, void* stream) {
  IPDM_REQUIRE(x && stats, IPDM_E_BADARG, "instnorm_stats: null pointer");
  IPDM_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024 && N >= 1 && HW >= 1, IPDM_E_BADARG, "instnorm_stats: C=%d must be a multiple of 4 (<=1024)", C);
  cudaStream_t s = as_stream(stream);
  IPDM_CUDA(cudaMemsetAsync(stats, 0, (size_t)N * C * 2 * sizeof(float), s));
  const int lanes = C / 4;
  const int block = lanes >= 256 ? lanes : 256;
  const int rows = block / lanes;
  int chunks = (HW + rows * 8 - 1) / (rows * 8);
  const int cap = (148 * 8 + N - 1) / N;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  const size_t smem = (size_t)rows * C * 2 * sizeof(float);
  k_instnorm_stats<<<dim3(chunks, N), block, smem, s>>>(x, stats, HW, C, pivoted);
  return launched("k_instnorm_stats");
}

int ipdm::stats_after_conv(const ipdm_conv_desc& d, cudaStream_t s) {
  if (!d.stats) return 0;
  IPDM_REQUIRE(d.out_f32, IPDM_E_BADARG, "conv: stats needs out_f32");
  const bool pool = (d.flags & IPDM_CONV_POOL2) != 0;
  const int HW = (pool ? d.H / 2 : d.H) * (pool ? d.W / 2 : d.W);
  return ipdm_instnorm_stats(d.out_f32, d.stats, d.N, HW, d.Cout, 0, s);
}

extern "C" int ipdm_instnorm_apply_elu(const float* x, const float* stats, int stats_pivoted, const float* alpha,
                                       const float* gamma, const float* beta, void* out_f16, int N, int HW, int C,
                                       void* stream) {
  IPDM_REQUIRE(x && stats && alpha && gamma && out_f16, IPDM_E_BADARG, "instnorm_apply_elu: null pointer");
  IPDM_REQUIRE(C % 4 == 0 && C >= 4 && C <= 2048, IPDM_E_BADARG, "instnorm_apply_elu: C=%d must be a multiple of 4", C);
  const size_t smem = (size_t)(3 * C + 64) * sizeof(float);
  int chunks = grid1d((size_t)HW * (C / 4), 256, 8);
  const int cap = (148 * 8 + N - 1) / N;
  if (chunks > cap) chunks = cap;
  k_instnorm_apply<<<dim3(chunks, N), 256, smem, as_stream(stream)>>>(x, stats, stats_pivoted, alpha, gamma, beta,
                                                                      reinterpret_cast<__half*>(out_f16), HW, C);
  return launched("k_instnorm_apply");
}

extern "C" int ipdm_conv_first(const float* x, const float* w, const float* bias, float* out, float* stats, int N, int H,
                               int W, int Cout, int affine, void* stream) {
  IPDM_REQUIRE(x && w && out, IPDM_E_BADARG, "conv_first: null pointer");
  IPDM_REQUIRE(Cout % 4 == 0 && Cout <= 2048, IPDM_E_BADARG, "conv_first: Cout=%d must be a multiple of 4", Cout);
  const size_t total = (size_t)N * H * W * (Cout / 4);
  k_conv_first<<<grid1d(total, 256), 256, (size_t)10 * Cout * sizeof(float), as_stream(stream)>>>(x, w, bias, out, N, H, W, Cout, affine);
  if (int e = launched("k_conv_first")) return e;
  if (stats) return ipdm_instnorm_stats(out, stats, N, H * W, Cout, 0, stream);
  return 0;
}

extern "C" int ipdm_conv_last(const void* in_f16, const float* w, const float* bias, const float* sigmas,
                              const int64_t* labels, float* out, int N, int H, int W, int Cin, void* stream) {
  IPDM_REQUIRE(in_f16 && w && sigmas && labels && out, IPDM_E_BADARG, "conv_last: null pointer");
  IPDM_REQUIRE(Cin % 8 == 0 && Cin <= 1024, IPDM_E_BADARG, "conv_last: Cin=%d must be a multiple of 8", Cin);
  const size_t npix = (size_t)N * H * W;
  k_conv_last<<<grid1d(npix, 16), 256, (size_t)9 * Cin * sizeof(float), as_stream(stream)>>>(
      reinterpret_cast<const __half*>(in_f16), w, bias, sigmas, labels, out, N, H, W, Cin);
  return launched("k_conv_last");
}

extern "C" int ipdm_act_to_f16(const float* x, void* out_f16, size_t n, int elu, void* stream) {
  IPDM_REQUIRE(x && out_f16, IPDM_E_BADARG, "act_to_f16: null pointer");
  IPDM_REQUIRE(n % 4 == 0, IPDM_E_BADARG, "act_to_f16: n must be a multiple of 4");
  k_act_to_f16<<<grid1d(n / 4, 256), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<__half*>(out_f16), n / 4, elu);
  return launched("k_act_to_f16");
}

extern "C" int ipdm_maxpool5_f16(const void* in_f16, void* out_f16, int N, int H, int W, int C, void* stream) {
  IPDM_REQUIRE(in_f16 && out_f16, IPDM_E_BADARG, "maxpool5: null pointer");
  IPDM_REQUIRE(C % 8 == 0, IPDM_E_BADARG, "maxpool5: C=%d must be a multiple of 8", C);
  const size_t total = (size_t)N * H * W * (C / 8);
  k_maxpool5<<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(in_f16),
                                                                  reinterpret_cast<__half*>(out_f16), N, H, W, C);
  return launched("k_maxpool5");
}

extern "C" int ipdm_bilinear_add(const float* src, float* dst, void* out_elu_f16, int N, int h, int w, int H, int W, int C,
                                 int accumulate, void* stream) {
  IPDM_REQUIRE(src && dst, IPDM_E_BADARG, "bilinear_add: null pointer");
  IPDM_REQUIRE(C % 4 == 0, IPDM_E_BADARG, "bilinear_add: C=%d must be a multiple of 4", C);
  const size_t total = (size_t)N * H * W * (C / 4);
  k_bilinear_add<<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(src, dst, reinterpret_cast<__half*>(out_elu_f16), N, h, w, H, W, C, accumulate);
  return launched("k_bilinear_add");
}

extern "C" int ipdm_meanpool2(const float* in, const float* add, float* out, int N, int H, int W, int C, void* stream) {
  IPDM_REQUIRE(in && out, IPDM_E_BADARG, "meanpool2: null pointer");
  IPDM_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, IPDM_E_BADARG, "meanpool2: C %% 4, H %% 2, W %% 2 must be 0");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 4);
  k_meanpool2<<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(in, add, out, N, H, W, C);
  return launched("k_meanpool2");
}

extern "C" int ipdm_pack_weights_f16(const float* w_oihw, void* w_f16, int Cout, int Cin, int taps, void* stream) {
  IPDM_REQUIRE(w_oihw && w_f16 && (taps == 9 || taps == 1), IPDM_E_BADARG, "pack_weights: bad argument");
  const size_t total = (size_t)Cout * taps * Cin;
  k_pack_weights<<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(w_oihw, reinterpret_cast<__half*>(w_f16), Cout, Cin, taps);
  return launched("k_pack_weights");
}

extern "C" int ipdm_conv_direct(const ipdm_conv_desc* dh, void* stream) {
  IPDM_REQUIRE(dh && dh->in_f16 && dh->w_f16, IPDM_E_BADARG, "conv_direct: null pointer");
  IPDM_REQUIRE(dh->taps == 9 || dh->taps == 1, IPDM_E_BADARG, "conv_direct: taps must be 9 or 1");
  IPDM_REQUIRE(dh->out_f32 || dh->out_f16, IPDM_E_BADARG, "conv_direct: no output");
  const ipdm_conv_desc d = *dh;
  const bool pool = (d.flags & IPDM_CONV_POOL2) != 0;
  IPDM_REQUIRE(!pool || (d.H % 2 == 0 && d.W % 2 == 0), IPDM_E_BADARG, "conv_direct: pooling needs even H, W");
  const size_t total = (size_t)d.N * (pool ? d.H / 2 : d.H) * (pool ? d.W / 2 : d.W) * d.Cout;
  k_conv_direct<<<grid1d(total, 128, 64), 128, 0, as_stream(stream)>>>(d);
  if (int e = launched("k_conv_direct")) return e;
  return stats_after_conv(d, as_stream(stream));
}
