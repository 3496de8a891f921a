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
// thread = (4 consecutive pixels of a row, 4 output channels): 18 input values and 9 float4 weights feed 16
// outputs.  grid (chunks, N): a block stays inside one image so the InstanceNorm++ sums can be reduced in the
// block and added with 2 atomics per channel.
template <bool OUT16>   // OUT16: the output (the start of the residual stream) is stored as f16
__global__ void __launch_bounds__(256, 2) k_conv_first(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                             void* __restrict__ out, double* __restrict__ stats, int H, int W, int Cout, int affine) {
  extern __shared__ float sw[];  // [9][Cout] + [Cout] + per-thread stats scratch [blockDim][8]
  float* sst = sw + 10 * Cout;
  for (int i = threadIdx.x; i < 9 * Cout; i += blockDim.x) {
    const int co = i % Cout, tap = i / Cout;
    sw[i] = w[co * 9 + tap];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[9 * Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int groups = Cout / 4;
  const int g = threadIdx.x % groups;
  const int W4 = (W + 3) / 4;
  const int quads = H * W4;                                     // 4-pixel groups in the image (32-bit: checked on the host)
  const int qper = blockDim.x / groups;
  const float* xin = x + (size_t)n * H * W;
  float* o = reinterpret_cast<float*>(out) + (OUT16 ? 0 : (size_t)n * H * W * Cout);
  __half* o16 = reinterpret_cast<__half*>(out) + (size_t)n * H * W * Cout;
  float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
  const float4 b4 = *reinterpret_cast<const float4*>(&sw[9 * Cout + 4 * g]);
  float4 wt[9];                                                  // this thread's 9 x 4 weights stay in registers
#pragma unroll
  for (int t = 0; t < 9; ++t) wt[t] = *reinterpret_cast<const float4*>(&sw[t * Cout + 4 * g]);
  // the 3 x 6 input patch of quad q, RAW values (the affine map 2v - 1 is applied when the patch is consumed, so that
  // nothing waits on these loads here); outside the image the pad value maps to 0 under the same transform.  Loads of
  // the NEXT quad are issued before the FMAs and stores of the current one.  Interior quads (all but the image border)
  // take three unconditional loads per row: one aligned float4 and its two neighbours.
  const float pad = affine ? 0.5f : 0.f;
  const bool vec_ok = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  auto load_patch = [&](int q, float (&in)[3][6]) {
    if (q >= quads) return;
    const int yh = q / W4, x0 = (q - yh * W4) * 4;
    if (vec_ok && yh >= 1 && yh + 1 < H && x0 >= 4 && x0 + 8 <= W) {
      const float* r0 = xin + (size_t)(yh - 1) * W + x0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float* rp = r0 + (size_t)r * W;
        const float4 mid = *reinterpret_cast<const float4*>(rp);
        in[r][0] = rp[-1]; in[r][1] = mid.x; in[r][2] = mid.y; in[r][3] = mid.z; in[r][4] = mid.w; in[r][5] = rp[4];
      }
      return;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yy = yh + r - 1;
      const bool rok = yy >= 0 && yy < H;
      const float* rp = xin + (size_t)(rok ? yy : 0) * W;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const int xx = x0 + c - 1;
        in[r][c] = (rok && xx >= 0 && xx < W) ? rp[xx] : pad;
      }
    }
  };
  if (threadIdx.x / groups < qper) {
    const int qstep = gridDim.x * qper;
    int q = blockIdx.x * qper + threadIdx.x / groups;
    float nxt[3][6];
    load_patch(q, nxt);
    for (; q < quads; q += qstep) {
      const int yh = q / W4, x0 = (q - yh * W4) * 4;
      float in[3][6];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) in[r][c] = affine ? 2.f * nxt[r][c] - 1.f : nxt[r][c];
      load_patch(q + qstep, nxt);
      float4 acc[4] = {b4, b4, b4, b4};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 ww = wt[ky * 3 + kx];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float v = in[ky][kx + p];
            acc[p].x += v * ww.x; acc[p].y += v * ww.y; acc[p].z += v * ww.z; acc[p].w += v * ww.w;
          }
        }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (x0 + p < W) {
          if (OUT16) {
            uint2 pk;
            pk.x = pack_half2_sat(acc[p].x, acc[p].y);
            pk.y = pack_half2_sat(acc[p].z, acc[p].w);
            *reinterpret_cast<uint2*>(&o16[((size_t)yh * W + x0 + p) * Cout + 4 * g]) = pk;
          } else {
            *reinterpret_cast<float4*>(&o[((size_t)yh * W + x0 + p) * Cout + 4 * g]) = acc[p];
          }
          s1.x += acc[p].x; s1.y += acc[p].y; s1.z += acc[p].z; s1.w += acc[p].w;
          s2.x += acc[p].x * acc[p].x; s2.y += acc[p].y * acc[p].y; s2.z += acc[p].z * acc[p].z; s2.w += acc[p].w * acc[p].w;
        }
      }
    }
  }
  if (stats == nullptr) return;
  // fixed-order block reduction (thread partials -> smem -> one thread per channel), then one fp64 atomic per
  // value: the only order-dependent step is a double-precision add, so results are run-to-run stable in fp32
  float* mine = sst + threadIdx.x * 8;
  mine[0] = s1.x; mine[1] = s2.x; mine[2] = s1.y; mine[3] = s2.y; mine[4] = s1.z; mine[5] = s2.z; mine[6] = s1.w; mine[7] = s2.w;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x) {
    const int ch = i >> 1, which = i & 1;
    const int gg = ch >> 2, k = ch & 3;
    float t = 0.f;
    for (int r = 0; r < qper; ++r) t += sst[(r * groups + gg) * 8 + k * 2 + which];
    atomicAdd(&stats[(size_t)n * Cout * 2 + i], (double)t);
  }
}

// ---------------------------------------------------------------------------- end_conv (Cout = 1)
// Two steps so that every input pixel (Cin f16 values) is read exactly once:
//   dots[p][t] = sum_c in[p][c] * w[t][c]            (16 lanes per pixel, 8 channels per lane per 128)
//   out[y][x]  = (bias + sum_t dots[(y,x) + off_t][t]) / sigma[label]     (9 reads of an L2-resident table)
constexpr int LAST_SLOT = 76;   // floats per 8-channel weight slot (72 used)
__global__ void k_conv_last_dots(const __half* __restrict__ in, const float* __restrict__ w, float* __restrict__ dots,
                                 size_t npix, int Cin) {
  // weights re-laid out per 8-channel lane slot: sw[slot][tap][8], slot stride 76 floats (conflict-free 16-byte reads)
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < 9 * Cin; i += blockDim.x) {
    const int t = i / Cin, c = i % Cin;
    sw[(c >> 3) * LAST_SLOT + t * 8 + (c & 7)] = w[i];
  }
  __syncthreads();
  const int lane16 = threadIdx.x & 15;
  const size_t gstride = (size_t)gridDim.x * (blockDim.x / 16);
  const size_t iters = (npix + gstride - 1) / gstride;
  size_t pix = blockIdx.x * (size_t)(blockDim.x / 16) + threadIdx.x / 16;
  for (size_t it = 0; it < iters; ++it, pix += gstride) {
    const bool live = pix < npix;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    if (live) {
      const __half* p = in + pix * Cin;
      for (int c0 = lane16 * 8; c0 < Cin; c0 += 128) {
        const uint4 raw = *reinterpret_cast<const uint4*>(p + c0);
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t2 = __half22float2(h2[j]);
          f[2 * j] = t2.x;
          f[2 * j + 1] = t2.y;
        }
        const float* slot = sw + (c0 >> 3) * LAST_SLOT;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 w0 = *reinterpret_cast<const float4*>(slot + t * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(slot + t * 8 + 4);
          acc[t] += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y + f[6] * w1.z + f[7] * w1.w;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], off);
    if (live && lane16 < 9) {
      float v = acc[0];
#pragma unroll
      for (int t = 1; t < 9; ++t) v = lane16 == t ? acc[t] : v;
      dots[pix * 9 + lane16] = v;
    }
  }
}

// Cin == 128 (the end_conv of every NCSNv2 at ngf 128): the lane's 9 x 8 weights live in registers, so a pixel costs
// one 16-byte load, 72 FMAs and a 15-shuffle transposing reduction (lane t of the 16 ends with the sum of tap t)
// instead of 18 shared-memory reads and 36 shuffles.  Two pixels per iteration keep two loads in flight.
__global__ void __launch_bounds__(256, 2) k_conv_last_dots128(const __half* __restrict__ in, const float* __restrict__ w,
                                                              float* __restrict__ dots, size_t npix) {
  const int lane16 = threadIdx.x & 15;
  float wr[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = *reinterpret_cast<const float4*>(w + t * 128 + lane16 * 8);
    const float4 b = *reinterpret_cast<const float4*>(w + t * 128 + lane16 * 8 + 4);
    wr[t][0] = a.x; wr[t][1] = a.y; wr[t][2] = a.z; wr[t][3] = a.w;
    wr[t][4] = b.x; wr[t][5] = b.y; wr[t][6] = b.z; wr[t][7] = b.w;
  }
  const bool up8 = lane16 & 8, up4 = lane16 & 4, up2 = lane16 & 2, up1 = lane16 & 1;
  const size_t gstride = (size_t)gridDim.x * (blockDim.x / 16);
  const size_t iters = (npix + 2 * gstride - 1) / (2 * gstride);
  size_t pix = blockIdx.x * (size_t)(blockDim.x / 16) + threadIdx.x / 16;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  uint4 nxt[2];                                  // the next iteration's two pixels are in flight while these two are reduced
#pragma unroll
  for (int u = 0; u < 2; ++u)
    nxt[u] = pix + u * gstride < npix ? *reinterpret_cast<const uint4*>(in + (pix + u * gstride) * 128 + lane16 * 8) : zero4;
  for (size_t it = 0; it < iters; ++it, pix += 2 * gstride) {
    const size_t px[2] = {pix, pix + gstride};
    uint4 raw[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      raw[u] = nxt[u];
      const size_t pn = px[u] + 2 * gstride;
      nxt[u] = pn < npix ? *reinterpret_cast<const uint4*>(in + pn * 128 + lane16 * 8) : zero4;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw[u]);
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t2 = __half22float2(h2[j]);
        f[2 * j] = t2.x;
        f[2 * j + 1] = t2.y;
      }
      float v[16];
#pragma unroll
      for (int t = 0; t < 9; ++t)
        v[t] = f[0] * wr[t][0] + f[1] * wr[t][1] + f[2] * wr[t][2] + f[3] * wr[t][3] + f[4] * wr[t][4] + f[5] * wr[t][5] +
               f[6] * wr[t][6] + f[7] * wr[t][7];
#pragma unroll
      for (int t = 9; t < 16; ++t) v[t] = 0.f;
      // after the step with offset o, v[t] (t < o) holds the partial sum of index t + (lane16 & ~(o - 1) & 15)
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] = (up8 ? v[t + 8] : v[t]) + __shfl_xor_sync(0xffffffffu, up8 ? v[t] : v[t + 8], 8);
#pragma unroll
      for (int t = 0; t < 4; ++t) v[t] = (up4 ? v[t + 4] : v[t]) + __shfl_xor_sync(0xffffffffu, up4 ? v[t] : v[t + 4], 4);
#pragma unroll
      for (int t = 0; t < 2; ++t) v[t] = (up2 ? v[t + 2] : v[t]) + __shfl_xor_sync(0xffffffffu, up2 ? v[t] : v[t + 2], 2);
      v[0] = (up1 ? v[1] : v[0]) + __shfl_xor_sync(0xffffffffu, up1 ? v[0] : v[1], 1);
      if (px[u] < npix && lane16 < 9) dots[px[u] * 9 + lane16] = v[0];
    }
  }
}

__global__ void k_conv_last_sum(const float* __restrict__ dots, const float* __restrict__ bias, const float* __restrict__ sigmas,
                                const int64_t* __restrict__ labels, float* __restrict__ out, int N, int H, int W) {
  const size_t npix = (size_t)N * H * W;
  for (size_t pix = blockIdx.x * (size_t)blockDim.x + threadIdx.x; pix < npix; pix += (size_t)gridDim.x * blockDim.x) {
    const int xw = (int)(pix % W), yh = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
    float acc = bias ? bias[0] : 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = yh + ky - 1, xx = xw + kx - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        acc += dots[(((size_t)n * H + yy) * W + xx) * 9 + ky * 3 + kx];
      }
    out[pix] = acc / sigmas[labels[n]];
  }
}

// ---------------------------------------------------------------------------- InstanceNorm++
// stats[n][c] = (sum(x - p), sum((x - p)^2)) over HW, p = x[n,0,c]  (pivot kills the cancellation
// in E[x^2] - E[x]^2).  grid (chunks, N), block 256 = (C/4 lanes) x (256/(C/4) pixel rows).
__global__ void k_instnorm_stats(const float* __restrict__ x, double* __restrict__ stats, int HW, int C, int pivoted) {
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
    atomicAdd(&stats[(size_t)n * C * 2 + i], (double)t);
  }
}

// out = f16(ELU(gamma*((x-m)*rstd + alpha*m_hat) + beta)); per-(n,c) A,B precomputed in smem.
template <bool IN16>   // IN16: x is the 16-bit residual stream
__global__ void k_instnorm_apply(const void* __restrict__ xv, const double* __restrict__ stats, int pivoted,
                                 const float* __restrict__ alpha, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, __half* __restrict__ out, int HW, int C, int s2d_w) {
  // s2d_w != 0: the image is s2d_w pixels wide and the output is written in space-to-depth layout
  // [N][H/2][W/2][(y&1)*2 + (x&1)][C] -- the operand of ConvMeanPool in its 4x4 stride-2 form
  extern __shared__ float sm[];  // A[C], B[C], mean[C], red[64]
  float* A = sm;
  float* Bv = sm + C;
  float* mean = sm + 2 * C;
  float* red = sm + 3 * C;
  const int n = blockIdx.y;
  const float* base = reinterpret_cast<const float*>(xv) + (IN16 ? 0 : (size_t)n * HW * C);
  const __half* base16 = reinterpret_cast<const __half*>(xv) + (size_t)n * HW * C;
  const float inv = 1.0f / (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float p = pivoted ? (IN16 ? __half2float(base16[c]) : base[c]) : 0.f;
    const float d = (float)(stats[((size_t)n * C + c) * 2] * (double)inv);
    mean[c] = p + d;
    const float var = fmaxf((float)(stats[((size_t)n * C + c) * 2 + 1] * (double)inv - (double)d * (double)d), 0.f);
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
  auto out_pix = [&](int pp) -> size_t {      // element offset of pixel pp's channel 0 inside the image's output
    if (s2d_w == 0) return (size_t)pp * C;
    const int y = pp / s2d_w, x = pp - y * s2d_w;
    return ((size_t)((y >> 1) * (s2d_w >> 1) + (x >> 1)) * 4 + (size_t)((y & 1) * 2 + (x & 1))) * C;
  };
  if (IN16 && (C & 7) == 0) {
    // 16-bit stream: thread = 8 channels (16-byte loads and stores), 4 pixels in flight per thread
    const int lanes = C / 8;
    const int c8 = (int)(threadIdx.x % lanes) * 8;
    const int rows = blockDim.x / lanes;
    const int row = threadIdx.x / lanes;
    if (row >= rows) return;
    float a[8], bb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = A[c8 + k]; bb[k] = Bv[c8 + k]; }
    __half* obase = out + (size_t)n * HW * C;
    const int stride = gridDim.x * rows;
    for (int pix = blockIdx.x * rows + row; pix < HW; pix += 4 * stride) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pp = pix + u * stride;
        if (pp < HW) v[u] = *reinterpret_cast<const uint4*>(base16 + (size_t)pp * C + c8);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pp = pix + u * stride;
        if (pp < HW) {
          const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
            o[k] = pack_half2_sat(elu_f16bound(f.x * a[2 * k] + bb[2 * k]), elu_f16bound(f.y * a[2 * k + 1] + bb[2 * k + 1]));
          }
          *reinterpret_cast<uint4*>(obase + out_pix(pp) + c8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    return;
  }
  // main pass: thread = 4 channels; 4 pixels in flight per thread (independent 16-byte loads)
  const int lanes = C / 4;
  const int c4 = (int)(threadIdx.x % lanes) * 4;
  const int rows = blockDim.x / lanes;                       // pixels per block pass
  const int row = threadIdx.x / lanes;
  if (row >= rows) return;
  const float a0 = A[c4], a1 = A[c4 + 1], a2 = A[c4 + 2], a3 = A[c4 + 3];
  const float b0 = Bv[c4], b1 = Bv[c4 + 1], b2 = Bv[c4 + 2], b3 = Bv[c4 + 3];
  __half* obase = out + (size_t)n * HW * C;
  const int stride = gridDim.x * rows;
  for (int pix = blockIdx.x * rows + row; pix < HW; pix += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pp = pix + u * stride;
      if (pp < HW) {
        if (IN16) {
          const uint2 hv = *reinterpret_cast<const uint2*>(base16 + (size_t)pp * C + c4);
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&hv.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&hv.y));
          v[u] = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
          v[u] = *reinterpret_cast<const float4*>(base + (size_t)pp * C + c4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pp = pix + u * stride;
      if (pp < HW) {
        uint2 pk;
        pk.x = pack_half2_sat(elu_f16bound(v[u].x * a0 + b0), elu_f16bound(v[u].y * a1 + b1));
        pk.y = pack_half2_sat(elu_f16bound(v[u].z * a2 + b2), elu_f16bound(v[u].w * a3 + b3));
        *reinterpret_cast<uint2*>(obase + out_pix(pp) + c4) = pk;
      }
    }
  }
}

// ---------------------------------------------------------------------------- elementwise
__global__ void k_act_to_f16(const float* __restrict__ x, __half* __restrict__ out, size_t n4, int elu) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    if (elu) { v.x = elu_f16bound(v.x); v.y = elu_f16bound(v.y); v.z = elu_f16bound(v.z); v.w = elu_f16bound(v.w); }
    uint2 pk;
    pk.x = pack_half2_sat(v.x, v.y);
    pk.y = pack_half2_sat(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = pk;
  }
}

__global__ void k_act_to_f16_scaled(const float* __restrict__ x, __half* __restrict__ out, size_t n4, int elu, float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    if (elu) { v.x = elu_f16bound(v.x); v.y = elu_f16bound(v.y); v.z = elu_f16bound(v.z); v.w = elu_f16bound(v.w); }
    uint2 pk;
    pk.x = pack_half2_sat(v.x * scale, v.y * scale);
    pk.y = pack_half2_sat(v.z * scale, v.w * scale);
    reinterpret_cast<uint2*>(out)[i] = pk;
  }
}

#ifndef IPDM_MAXPOOL_NC
#define IPDM_MAXPOOL_NC 2
#endif
// 5x5/s1 max-pool, separable: thread = (NC consecutive columns, 8 channels).  Per input row it loads the 8 columns
// x-2 .. x+5 once (16 bytes each, 2 loads per output instead of 5) and forms the four horizontal 5-maxima from shared
// partial maxima; the last five such rows live in a register ring (the row loop is unrolled by 5, so the ring index is
// a compile-time constant and nothing is ever moved) for the vertical max.
// grid (ceil(W / (4*XQ)), ceil(H / YS), N * C/8/CG); block = XQ * CG threads.  The strip length YS trades redundant
// priming rows (4 per strip) for parallelism.
// NC output columns from the NC + 4 input columns x-2 .. x+NC+1: o[j] = max(in[j .. j+4])
template <int NC>
__device__ __forceinline__ void hmaxN(const uint4* in, __half2 (*o)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __half2 v[NC + 4];
#pragma unroll
    for (int c = 0; c < NC + 4; ++c) v[c] = reinterpret_cast<const __half2*>(&in[c])[q];
    if (NC == 4) {
      const __half2 a = __hmax2(v[3], v[4]), b = __hmax2(v[1], v[2]), c = __hmax2(v[5], v[6]);
      o[0][q] = __hmax2(__hmax2(v[0], b), a);
      o[1][q] = __hmax2(__hmax2(b, a), v[5]);
      o[2][q] = __hmax2(__hmax2(v[2], a), c);
      o[3][q] = __hmax2(__hmax2(a, c), v[7]);
    } else {
      const __half2 a = __hmax2(__hmax2(v[1], v[2]), __hmax2(v[3], v[4]));
      o[0][q] = __hmax2(v[0], a);
      o[1][q] = __hmax2(a, v[5]);
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(256, NC == 4 ? 2 : 3) k_maxpool5(const __half* __restrict__ in, __half* __restrict__ out, int N, int H, int W,
                                                                   int C, int CG, int XQ, int MP_YS /* rows per strip */) {
  const int cg_per = C / 8 / CG;                       // channel-group blocks per image
  const int n = blockIdx.z / cg_per;
  const int g = (blockIdx.z % cg_per) * CG + (threadIdx.x % CG);
  const int x = (blockIdx.x * XQ + threadIdx.x / CG) * NC;    // first of this thread's NC columns
  const int y0 = blockIdx.y * MP_YS;
  if (x >= W) return;
  const uint4 ninf4 = make_uint4(0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u);   // -inf halves
  const __half2 ninf = __float2half2_rn(-INFINITY);
  __half2 win[5][NC][4];                                // [ring slot][column][half2 lane]
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int j = 0; j < NC; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) win[k][j][q] = ninf;
  const __half* base = in + (size_t)n * H * W * C + 8 * g;
  auto hrow = [&](int y, __half2 (*o)[4]) {
    uint4 raw[NC + 4];
#pragma unroll
    for (int c = 0; c < NC + 4; ++c) {
      const int xx = x - 2 + c;
      raw[c] = ninf4;
      if (y >= 0 && y < H && xx >= 0 && xx < W) raw[c] = *reinterpret_cast<const uint4*>(base + ((size_t)y * W + xx) * C);
    }
    hmaxN<NC>(raw, o);
  };
  // prime ring slots 1..4 with rows y0-2 .. y0+1 (slot 0 is overwritten first)
#pragma unroll
  for (int k = 0; k < 4; ++k) hrow(y0 - 2 + k, win[k + 1]);
  const int yend = min(y0 + MP_YS, H);
  for (int yb = y0; yb < yend; yb += 5) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      const int y = yb + s;
      if (y >= yend) break;
      hrow(y + 2, win[s]);                              // row y+2 replaces row y-3: slots hold rows y-2 .. y+2
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        if (x + j >= W) continue;
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          oh[q] = __hmax2(__hmax2(__hmax2(win[0][j][q], win[1][j][q]), __hmax2(win[2][j][q], win[3][j][q])), win[4][j][q]);
        *reinterpret_cast<uint4*>(out + (((size_t)n * H + y) * W + x + j) * C + 8 * g) = o;
      }
    }
  }
}

// dst (+)= bilinear(src), align_corners=True, PyTorch's index arithmetic (upsample_bilinear2d).
// One CTA per output row (n, Y): the vertical taps / weights are per-CTA constants and all index arithmetic is 32-bit
// (the first version spent most of its instructions on 64-bit div/mod per element); a thread handles 16-byte channel
// quads of consecutive pixels, two per iteration so that ten independent loads are in flight.
template <bool T16> struct Quad;      // four consecutive channels of the residual stream: f32 or f16 in memory, f32 in registers
template <> struct Quad<false> {
  static __device__ __forceinline__ float4 ld(const void* p, size_t i) { return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i); }
  static __device__ __forceinline__ void st(void* p, size_t i, float4 v) { *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = v; }
};
template <> struct Quad<true> {
  static __device__ __forceinline__ float4 ld(const void* p, size_t i) {
    const uint2 hv = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p) + i);
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&hv.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&hv.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  }
  static __device__ __forceinline__ void st(void* p, size_t i, float4 v) {
    uint2 pk;
    pk.x = pack_half2_sat(v.x, v.y);
    pk.y = pack_half2_sat(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p) + i) = pk;
  }
};

template <bool T16>
__global__ void __launch_bounds__(256) k_bilinear_add(const void* __restrict__ src, void* __restrict__ dst, __half* __restrict__ out16,
                                                      int h, int w, int H, int W, int C, int accumulate) {
  using Q = Quad<T16>;
  const int lanes = C >> 2;
  const int n = blockIdx.x / H, Y = blockIdx.x % H;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const float fy = sy * Y;
  const int y0 = (int)fy, y1 = y0 + (y0 < h - 1 ? 1 : 0);
  const float ly = fy - y0, hy = 1.f - ly;
  const size_t r0 = ((size_t)n * h + y0) * w * C, r1 = ((size_t)n * h + y1) * w * C;   // element offsets
  const size_t drow = ((size_t)n * H + Y) * W * C;
  __half* hrow = out16 ? out16 + ((size_t)n * H + Y) * W * C : nullptr;
  const int total = W * lanes;
  for (int i0 = threadIdx.x; i0 < total; i0 += 2 * blockDim.x) {
    float4 v00[2], v01[2], v10[2], v11[2], old[2];
    float lx[2];
    int off[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = i0 + u * blockDim.x;
      live[u] = i < total;
      if (!live[u]) continue;
      const int X = i / lanes, c4 = (i - X * lanes) * 4;
      const float fx = sx * X;
      const int x0 = (int)fx, x1 = x0 + (x0 < w - 1 ? 1 : 0);
      lx[u] = fx - x0;
      v00[u] = Q::ld(src, r0 + x0 * C + c4);
      v01[u] = Q::ld(src, r0 + x1 * C + c4);
      v10[u] = Q::ld(src, r1 + x0 * C + c4);
      v11[u] = Q::ld(src, r1 + x1 * C + c4);
      off[u] = X * C + c4;
      if (accumulate) old[u] = Q::ld(dst, drow + off[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!live[u]) continue;
      const float hx = 1.f - lx[u];
      float4 r;
      r.x = hy * (hx * v00[u].x + lx[u] * v01[u].x) + ly * (hx * v10[u].x + lx[u] * v11[u].x);
      r.y = hy * (hx * v00[u].y + lx[u] * v01[u].y) + ly * (hx * v10[u].y + lx[u] * v11[u].y);
      r.z = hy * (hx * v00[u].z + lx[u] * v01[u].z) + ly * (hx * v10[u].z + lx[u] * v11[u].z);
      r.w = hy * (hx * v00[u].w + lx[u] * v01[u].w) + ly * (hx * v10[u].w + lx[u] * v11[u].w);
      if (accumulate) { r.x += old[u].x; r.y += old[u].y; r.z += old[u].z; r.w += old[u].w; }
      Q::st(dst, drow + off[u], r);
      if (hrow) {
        uint2 pk;
        pk.x = pack_half2_sat(elu_f16bound(r.x), elu_f16bound(r.y));
        pk.y = pack_half2_sat(elu_f16bound(r.z), elu_f16bound(r.w));
        *reinterpret_cast<uint2*>(hrow + off[u]) = pk;
      }
    }
  }
}

// The same on the 16-bit stream with EIGHT channels per thread (16-byte accesses; C % 8 == 0): the 4-channel version above moved
// 1.5 GB in 0.46 ms at 256^2 x 128 channels (3.3 TB/s, 8-byte requests).
struct H8 {
  float v[8];
  static __device__ __forceinline__ H8 ld(const __half* p) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    H8 r;
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&q.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&q.z)), d = __half22float2(*reinterpret_cast<const __half2*>(&q.w));
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
    return r;
  }
};
__global__ void __launch_bounds__(256) k_bilinear_add_h8(const __half* __restrict__ src, __half* __restrict__ dst, __half* __restrict__ out16,
                                                         int h, int w, int H, int W, int C, int accumulate) {
  const int lanes = C >> 3;
  const int n = blockIdx.x / H, Y = blockIdx.x % H;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const float fy = sy * Y;
  const int y0 = (int)fy, y1 = y0 + (y0 < h - 1 ? 1 : 0);
  const float ly = fy - y0, hy = 1.f - ly;
  const __half* s0 = src + ((size_t)n * h + y0) * w * C;
  const __half* s1 = src + ((size_t)n * h + y1) * w * C;
  __half* drow = dst + ((size_t)n * H + Y) * W * C;
  __half* hrow = out16 ? out16 + ((size_t)n * H + Y) * W * C : nullptr;
  const int total = W * lanes;
  for (int i0 = threadIdx.x; i0 < total; i0 += 2 * blockDim.x) {
    H8 v00[2], v01[2], v10[2], v11[2], old[2];
    float lx[2];
    int off[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = i0 + u * blockDim.x;
      live[u] = i < total;
      if (!live[u]) continue;
      const int X = i / lanes, c8 = (i - X * lanes) * 8;
      const float fx = sx * X;
      const int x0 = (int)fx, x1 = x0 + (x0 < w - 1 ? 1 : 0);
      lx[u] = fx - x0;
      v00[u] = H8::ld(s0 + x0 * C + c8);
      v01[u] = H8::ld(s0 + x1 * C + c8);
      v10[u] = H8::ld(s1 + x0 * C + c8);
      v11[u] = H8::ld(s1 + x1 * C + c8);
      off[u] = X * C + c8;
      if (accumulate) old[u] = H8::ld(drow + off[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!live[u]) continue;
      const float hx = 1.f - lx[u];
      float r[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        r[k] = hy * (hx * v00[u].v[k] + lx[u] * v01[u].v[k]) + ly * (hx * v10[u].v[k] + lx[u] * v11[u].v[k]);
        if (accumulate) r[k] += old[u].v[k];
      }
      uint4 pk;
      pk.x = pack_half2_sat(r[0], r[1]); pk.y = pack_half2_sat(r[2], r[3]); pk.z = pack_half2_sat(r[4], r[5]); pk.w = pack_half2_sat(r[6], r[7]);
      *reinterpret_cast<uint4*>(drow + off[u]) = pk;
      if (hrow) {
        pk.x = pack_half2_sat(elu_f16bound(r[0]), elu_f16bound(r[1])); pk.y = pack_half2_sat(elu_f16bound(r[2]), elu_f16bound(r[3]));
        pk.z = pack_half2_sat(elu_f16bound(r[4]), elu_f16bound(r[5])); pk.w = pack_half2_sat(elu_f16bound(r[6]), elu_f16bound(r[7]));
        *reinterpret_cast<uint4*>(hrow + off[u]) = pk;
      }
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
    float v = acc * (d.acc_scale != 0.f ? d.acc_scale : 1.f) + (d.bias ? d.bias[co] : 0.f);
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
      s *= d.out_f16_scale != 0.f ? d.out_f16_scale : 1.f;
      reinterpret_cast<__half*>(d.out_f16)[i] = __float2half_rn(fminf(fmaxf(s, -65504.f), 65504.f));
    }
  }
}

int stats_after_conv(const ipdm_conv_desc& d, cudaStream_t s);

}  // namespace ipdm

using namespace ipdm;

extern "C" int ipdm_instnorm_stats(const float* x, double* stats, int N, int HW, int C, int pivoted, void* stream) {
  IPDM_REQUIRE(x && stats, IPDM_E_BADARG, "instnorm_stats: null pointer");
  IPDM_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024 && N >= 1 && HW >= 1, IPDM_E_BADARG, "instnorm_stats: C=%d must be a multiple of 4 (<=1024)", C);
  cudaStream_t s = as_stream(stream);
  IPDM_CUDA(cudaMemsetAsync(stats, 0, (size_t)N * C * 2 * sizeof(double), s));
  const int lanes = C / 4;
  const int block = lanes >= 256 ? lanes : 256;
  const int rows = block / lanes;
  int chunks = (HW + rows * 8 - 1) / (rows * 8);
  int cap = (148 * 4) / N;          // one resident wave (>= 4 blocks of 256 threads per SM)
  if (cap < 1) cap = 1;
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

extern "C" int ipdm_instnorm_apply_elu(const float* x, const double* stats, int stats_pivoted, const float* alpha,
                                       const float* gamma, const float* beta, void* out_f16, int N, int HW, int C,
                                       void* stream) {
  IPDM_REQUIRE(x && stats && alpha && gamma && out_f16, IPDM_E_BADARG, "instnorm_apply_elu: null pointer");
  IPDM_REQUIRE(C % 4 == 0 && C >= 4 && C <= 2048, IPDM_E_BADARG, "instnorm_apply_elu: C=%d must be a multiple of 4", C);
  const size_t smem = (size_t)(3 * C + 64) * sizeof(float);
  int chunks = grid1d((size_t)HW * (C / 4), 256 * 4, 8);
  int cap = (148 * 4) / N;          // one resident wave
  if (cap < 1) cap = 1;
  if (chunks > cap) chunks = cap;
  k_instnorm_apply<false><<<dim3(chunks, N), 256, smem, as_stream(stream)>>>(x, stats, stats_pivoted, alpha, gamma, beta,
                                                                             reinterpret_cast<__half*>(out_f16), HW, C, 0);
  return launched("k_instnorm_apply");
}

extern "C" int ipdm_instnorm_apply_elu_f16in(const void* x_f16, const double* stats, int stats_pivoted, const float* alpha,
                                             const float* gamma, const float* beta, void* out_f16, int N, int HW, int C,
                                             void* stream) {
  IPDM_REQUIRE(x_f16 && stats && alpha && gamma && out_f16, IPDM_E_BADARG, "instnorm_apply_elu_f16in: null pointer");
  IPDM_REQUIRE(C % 4 == 0 && C >= 4 && C <= 2048, IPDM_E_BADARG, "instnorm_apply_elu_f16in: C=%d must be a multiple of 4", C);
  const size_t smem = (size_t)(3 * C + 64) * sizeof(float);
  int chunks = grid1d((size_t)HW * (C / 4), 256 * 4, 8);
  int cap = (148 * 4) / N;          // one resident wave
  if (cap < 1) cap = 1;
  if (chunks > cap) chunks = cap;
  k_instnorm_apply<true><<<dim3(chunks, N), 256, smem, as_stream(stream)>>>(x_f16, stats, stats_pivoted, alpha, gamma, beta,
                                                                            reinterpret_cast<__half*>(out_f16), HW, C, 0);
  return launched("k_instnorm_apply");
}

extern "C" int ipdm_instnorm_apply_elu_s2d(const void* x, int x_is_f16, const double* stats, int stats_pivoted, const float* alpha,
                                           const float* gamma, const float* beta, void* out_f16, int N, int H, int W, int C,
                                           void* stream) {
  IPDM_REQUIRE(x && stats && alpha && gamma && out_f16, IPDM_E_BADARG, "instnorm_apply_elu_s2d: null pointer");
  IPDM_REQUIRE(C % 4 == 0 && C >= 4 && C <= 2048, IPDM_E_BADARG, "instnorm_apply_elu_s2d: C=%d must be a multiple of 4", C);
  IPDM_REQUIRE(H % 2 == 0 && W % 2 == 0 && H >= 2 && W >= 2, IPDM_E_BADARG, "instnorm_apply_elu_s2d: H=%d, W=%d must be even", H, W);
  const int HW = H * W;
  const size_t smem = (size_t)(3 * C + 64) * sizeof(float);
  int chunks = grid1d((size_t)HW * (C / 4), 256 * 4, 8);
  int cap = (148 * 4) / N;          // one resident wave
  if (cap < 1) cap = 1;
  if (chunks > cap) chunks = cap;
  if (x_is_f16)
    k_instnorm_apply<true><<<dim3(chunks, N), 256, smem, as_stream(stream)>>>(x, stats, stats_pivoted, alpha, gamma, beta,
                                                                              reinterpret_cast<__half*>(out_f16), HW, C, W);
  else
    k_instnorm_apply<false><<<dim3(chunks, N), 256, smem, as_stream(stream)>>>(x, stats, stats_pivoted, alpha, gamma, beta,
                                                                               reinterpret_cast<__half*>(out_f16), HW, C, W);
  return launched("k_instnorm_apply");
}

extern "C" int ipdm_conv_first(const float* x, const float* w, const float* bias, float* out, double* stats, int N, int H,
                               int W, int Cout, int affine, void* stream) {
  IPDM_REQUIRE(x && w && out, IPDM_E_BADARG, "conv_first: null pointer");
  IPDM_REQUIRE(Cout % 4 == 0 && Cout <= 2048, IPDM_E_BADARG, "conv_first: Cout=%d must be a multiple of 4", Cout);
  IPDM_REQUIRE(Cout / 4 <= 256, IPDM_E_BADARG, "conv_first: Cout too large");
  cudaStream_t s = as_stream(stream);
  if (stats) IPDM_CUDA(cudaMemsetAsync(stats, 0, (size_t)N * Cout * 2 * sizeof(double), s));
  const int groups = Cout / 4;
  const int qper = 256 / groups;
  const size_t quads = (size_t)H * ((W + 3) / 4);
  int chunks = (int)((quads + qper - 1) / qper);
  int cap = (148 * 2) / N;          // one resident wave (2 blocks per SM at ~95 registers)
  if (cap < 1) cap = 1;
  if (chunks > cap) chunks = cap;
  k_conv_first<false><<<dim3(chunks, N), 256, (size_t)(10 * Cout + 256 * 8) * sizeof(float), s>>>(x, w, bias, out, stats, H, W, Cout, affine);
  return launched("k_conv_first");
}

extern "C" int ipdm_conv_first_f16out(const float* x, const float* w, const float* bias, void* out_f16, double* stats, int N, int H,
                                      int W, int Cout, int affine, void* stream) {
  IPDM_REQUIRE(x && w && out_f16, IPDM_E_BADARG, "conv_first_f16out: null pointer");
  IPDM_REQUIRE(Cout % 4 == 0 && Cout / 4 <= 256, IPDM_E_BADARG, "conv_first_f16out: Cout=%d must be a multiple of 4, at most 1024", Cout);
  cudaStream_t s = as_stream(stream);
  if (stats) IPDM_CUDA(cudaMemsetAsync(stats, 0, (size_t)N * Cout * 2 * sizeof(double), s));
  const int groups = Cout / 4;
  const int qper = 256 / groups;
  const size_t quads = (size_t)H * ((W + 3) / 4);
  int chunks = (int)((quads + qper - 1) / qper);
  int cap = (148 * 2) / N;
  if (cap < 1) cap = 1;
  if (chunks > cap) chunks = cap;
  k_conv_first<true><<<dim3(chunks, N), 256, (size_t)(10 * Cout + 256 * 8) * sizeof(float), s>>>(x, w, bias, out_f16, stats, H, W, Cout, affine);
  return launched("k_conv_first");
}

extern "C" int ipdm_conv_last(const void* in_f16, const float* w, const float* bias, const float* sigmas,
                              const int64_t* labels, float* out, float* workspace, int N, int H, int W, int Cin, void* stream) {
  IPDM_REQUIRE(in_f16 && w && sigmas && labels && out && workspace, IPDM_E_BADARG, "conv_last: null pointer");
  IPDM_REQUIRE(Cin % 8 == 0 && Cin <= 1024, IPDM_E_BADARG, "conv_last: Cin=%d must be a multiple of 8", Cin);
  const size_t npix = (size_t)N * H * W;
  if (Cin == 128) {
    k_conv_last_dots128<<<148 * 2, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(in_f16), w, workspace, npix);
  } else {
    k_conv_last_dots<<<grid1d(npix, 16, 4), 256, (size_t)(Cin / 8) * LAST_SLOT * sizeof(float), as_stream(stream)>>>(
        reinterpret_cast<const __half*>(in_f16), w, workspace, npix, Cin);
  }
  if (int e = launched("k_conv_last_dots")) return e;
  k_conv_last_sum<<<grid1d(npix, 256), 256, 0, as_stream(stream)>>>(workspace, bias, sigmas, labels, out, N, H, W);
  return launched("k_conv_last_sum");
}

extern "C" int ipdm_act_to_f16(const float* x, void* out_f16, size_t n, int elu, void* stream) {
  IPDM_REQUIRE(x && out_f16, IPDM_E_BADARG, "act_to_f16: null pointer");
  IPDM_REQUIRE(n % 4 == 0, IPDM_E_BADARG, "act_to_f16: n must be a multiple of 4");
  k_act_to_f16<<<grid1d(n / 4, 256), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<__half*>(out_f16), n / 4, elu);
  return launched("k_act_to_f16");
}

extern "C" int ipdm_act_to_f16_scaled(const float* x, void* out_f16, size_t n, int elu, float scale, void* stream) {
  IPDM_REQUIRE(x && out_f16, IPDM_E_BADARG, "act_to_f16_scaled: null pointer");
  IPDM_REQUIRE(n % 4 == 0, IPDM_E_BADARG, "act_to_f16_scaled: n must be a multiple of 4");
  IPDM_REQUIRE(scale > 0.f, IPDM_E_BADARG, "act_to_f16_scaled: scale must be positive");
  k_act_to_f16_scaled<<<grid1d(n / 4, 256), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<__half*>(out_f16), n / 4, elu, scale);
  return launched("k_act_to_f16_scaled");
}

extern "C" int ipdm_maxpool5_f16(const void* in_f16, void* out_f16, int N, int H, int W, int C, void* stream) {
  IPDM_REQUIRE(in_f16 && out_f16, IPDM_E_BADARG, "maxpool5: null pointer");
  IPDM_REQUIRE(C % 8 == 0, IPDM_E_BADARG, "maxpool5: C=%d must be a multiple of 8", C);
  const int groups = C / 8;
  int CG = 16;
  while (groups % CG != 0) CG >>= 1;                      // channel groups per block (power of two dividing C/8)
  constexpr int NC = IPDM_MAXPOOL_NC;                     // columns per thread
  int XQ = 256 / CG;                                      // column groups per block
  while (XQ > 1 && NC * (XQ / 2) >= W) XQ >>= 1;          // narrow images: no idle threads
  const int XT = NC * XQ;                                 // columns per block
  int ys = 30;                                            // multiples of 5 (the ring unroll)
  while (ys > 10 && (size_t)((W + XT - 1) / XT) * ((H + ys - 1) / ys) * N * (groups / CG) < 148 * 4) ys -= 10;
  dim3 grid((W + XT - 1) / XT, (H + ys - 1) / ys, N * (groups / CG));
  k_maxpool5<NC><<<grid, XQ * CG, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(in_f16),
                                                      reinterpret_cast<__half*>(out_f16), N, H, W, C, CG, XQ, ys);
  return launched("k_maxpool5");
}

extern "C" int ipdm_bilinear_add(const float* src, float* dst, void* out_elu_f16, int N, int h, int w, int H, int W, int C,
                                 int accumulate, void* stream) {
  IPDM_REQUIRE(src && dst, IPDM_E_BADARG, "bilinear_add: null pointer");
  IPDM_REQUIRE(C % 4 == 0, IPDM_E_BADARG, "bilinear_add: C=%d must be a multiple of 4", C);
  IPDM_REQUIRE((size_t)N * H < ((size_t)1 << 31) && (size_t)W * C < ((size_t)1 << 30), IPDM_E_UNSUPPORTED, "bilinear_add: image too large");
  k_bilinear_add<false><<<N * H, 256, 0, as_stream(stream)>>>(src, dst, reinterpret_cast<__half*>(out_elu_f16), h, w, H, W, C, accumulate);
  return launched("k_bilinear_add");
}

extern "C" int ipdm_bilinear_add_f16(const void* src_f16, void* dst_f16, void* out_elu_f16, int N, int h, int w, int H, int W, int C,
                                     int accumulate, void* stream) {
  IPDM_REQUIRE(src_f16 && dst_f16, IPDM_E_BADARG, "bilinear_add_f16: null pointer");
  IPDM_REQUIRE(C % 4 == 0, IPDM_E_BADARG, "bilinear_add_f16: C=%d must be a multiple of 4", C);
  IPDM_REQUIRE((size_t)N * H < ((size_t)1 << 31) && (size_t)W * C < ((size_t)1 << 30), IPDM_E_UNSUPPORTED, "bilinear_add_f16: image too large");
  if (C % 8 == 0)
    k_bilinear_add_h8<<<N * H, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(src_f16), reinterpret_cast<__half*>(dst_f16),
                                                            reinterpret_cast<__half*>(out_elu_f16), h, w, H, W, C, accumulate);
  else
    k_bilinear_add<true><<<N * H, 256, 0, as_stream(stream)>>>(src_f16, dst_f16, reinterpret_cast<__half*>(out_elu_f16), h, w, H, W, C, accumulate);
  return launched("k_bilinear_add");
}

// ---------------------------------------------------------------------------- f16 range audit
// The tensor-core path stores activations in f16 with saturation (|x| > 65504 -> +-65504, never inf / NaN): a clipped
// value is silent.  This audit makes it visible: max |x| and the number of values at the end of the range (or not finite)
// of one f16 tensor; the score networks run it over every f16 buffer of a forward on request (`range_audit()`).
__global__ void k_f16_range_audit(const __half* __restrict__ x, size_t n8, size_t n, float* __restrict__ max_abs,
                                  unsigned long long* __restrict__ n_sat) {
  float m = 0.f;
  unsigned cnt = 0;
  auto see = [&](__half2 h) {
    const float2 f = __half22float2(__habs2(h));
    m = fmaxf(m, fmaxf(f.x, f.y));      // fmaxf drops NaN: count those separately
    cnt += !(f.x < 65504.f) + !(f.y < 65504.f);
  };
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(x)[i];
    see(*reinterpret_cast<const __half2*>(&v.x));
    see(*reinterpret_cast<const __half2*>(&v.y));
    see(*reinterpret_cast<const __half2*>(&v.z));
    see(*reinterpret_cast<const __half2*>(&v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t i = n8 * 8; i < n; ++i) {
      const float f = fabsf(__half2float(x[i]));
      m = fmaxf(m, f);
      cnt += !(f < 65504.f);
    }
  for (int off = 16; off >= 1; off >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(reinterpret_cast<int*>(max_abs), __float_as_int(m));      // non-negative floats order like their bit patterns
    if (cnt) atomicAdd(n_sat, (unsigned long long)cnt);
  }
}

extern "C" int ipdm_f16_range_audit(const void* x_f16, size_t n, float* max_abs, unsigned long long* n_saturated, void* stream) {
  IPDM_REQUIRE(x_f16 && max_abs && n_saturated, IPDM_E_BADARG, "f16_range_audit: null pointer");
  IPDM_REQUIRE((reinterpret_cast<uintptr_t>(x_f16) & 15) == 0, IPDM_E_BADARG, "f16_range_audit: tensor must be 16-byte aligned");
  if (n == 0) return 0;
  k_f16_range_audit<<<grid1d(n / 8 + 1, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(x_f16), n / 8, n, max_abs, n_saturated);
  return launched("k_f16_range_audit");
}

extern "C" int ipdm_meanpool2(const float* in, const float* add, float* out, int N, int H, int W, int C, void* stream) {
  IPDM_REQUIRE(in && out, IPDM_E_BADARG, "meanpool2: null pointer");
  IPDM_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, IPDM_E_BADARG, "meanpool2: C %% 4, H %% 2, W %% 2 must be 0");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 4);
  k_meanpool2<<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(in, add, out, N, H, W, C);
  return launched("k_meanpool2");
}

// 2x2 mean-pool of an f16 NHWC tensor, fp32 arithmetic, one rounding: the operand of a pooled 1x1 shortcut convolution
// (mean-pool and a 1x1 convolution commute: pooling first is 4x fewer FLOPs and bytes for the convolution).
__global__ void k_meanpool2_f16(const __half* __restrict__ in, __half* __restrict__ out, int N, int H, int W, int C) {
  const int lanes = C / 8, Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * lanes;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % lanes) * 8;
    const size_t pix = i / lanes;
    const int X = (int)(pix % Wo), Y = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
    const __half* b = in + (((size_t)n * H + 2 * Y) * W + 2 * X) * C + c8;
    const H8 a00 = H8::ld(b), a01 = H8::ld(b + C), a10 = H8::ld(b + (size_t)W * C), a11 = H8::ld(b + (size_t)W * C + C);
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = (((a00.v[k] + a10.v[k]) + a01.v[k]) + a11.v[k]) * 0.25f;
    uint4 pk;
    pk.x = pack_half2_sat(r[0], r[1]); pk.y = pack_half2_sat(r[2], r[3]); pk.z = pack_half2_sat(r[4], r[5]); pk.w = pack_half2_sat(r[6], r[7]);
    *reinterpret_cast<uint4*>(out + pix * C + c8) = pk;
  }
}

extern "C" int ipdm_meanpool2_f16(const void* in_f16, void* out_f16, int N, int H, int W, int C, void* stream) {
  IPDM_REQUIRE(in_f16 && out_f16, IPDM_E_BADARG, "meanpool2_f16: null pointer");
  IPDM_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0 && N >= 1, IPDM_E_BADARG, "meanpool2_f16: C=%d must be a multiple of 8, H and W even", C);
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 8);
  k_meanpool2_f16<<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(in_f16), reinterpret_cast<__half*>(out_f16), N, H, W, C);
  return launched("k_meanpool2_f16");
}

extern "C" int ipdm_pack_weights_f16(const float* w_oihw, void* w_f16, int Cout, int Cin, int taps, void* stream) {
  IPDM_REQUIRE(w_oihw && w_f16 && taps >= 1 && Cout >= 1 && Cin >= 1, IPDM_E_BADARG, "pack_weights: bad argument");
  const size_t total = (size_t)Cout * taps * Cin;
  k_pack_weights<<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(w_oihw, reinterpret_cast<__half*>(w_f16), Cout, Cin, taps);
  return launched("k_pack_weights");
}

extern "C" int ipdm_conv_direct(const ipdm_conv_desc* dh, void* stream) {
  IPDM_REQUIRE(dh && dh->in_f16 && dh->w_f16, IPDM_E_BADARG, "conv_direct: null pointer");
  IPDM_REQUIRE(dh->taps == 9 || dh->taps == 1, IPDM_E_BADARG, "conv_direct: taps must be 9 or 1");
  IPDM_REQUIRE(dh->slices <= 1 && dh->slice_shift == 0, IPDM_E_UNSUPPORTED, "conv_direct: volumes (slices > 1) need the tensor-core path (Cin %% 64 == 0, Cout %% 128 == 0)");
  IPDM_REQUIRE(dh->out_f32 || dh->out_f16, IPDM_E_BADARG, "conv_direct: no output");
  const ipdm_conv_desc d = *dh;
  const bool pool = (d.flags & IPDM_CONV_POOL2) != 0;
  IPDM_REQUIRE(!pool || (d.H % 2 == 0 && d.W % 2 == 0), IPDM_E_BADARG, "conv_direct: pooling needs even H, W");
  const size_t total = (size_t)d.N * (pool ? d.H / 2 : d.H) * (pool ? d.W / 2 : d.W) * d.Cout;
  k_conv_direct<<<grid1d(total, 128, 64), 128, 0, as_stream(stream)>>>(d);
  if (int e = launched("k_conv_direct")) return e;
  return stats_after_conv(d, as_stream(stream));
}
