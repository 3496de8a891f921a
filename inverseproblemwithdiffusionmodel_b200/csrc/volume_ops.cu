// CUDA-core kernels of the 3-D (patch x time) score network NCSN3DShallow (ncsn/models/ncsn3d.py:123-224,
// layers3d.py): everything around its 3x3x3 convolutions, which run on the 2-D tensor-core kernels one kx-plane per
// launch (ipdm_conv_desc.slices / slice_shift).  Volume layout: [P][X][T][Y][C] -- P patches, X slices, each slice an
// NHWC "image" of H = T rows and W = Y columns.  All HBM-bound.
#include "common.cuh"

namespace ipdm {

static int vgrid(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = 148 * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// out[p][x] = max over x-2 .. x+2 (inside the volume) of in[p][.]: the slice axis of MaxPool3d(5, 1, 2) (layers3d.py:71)
__global__ void k_maxpool5_slices(const uint4* __restrict__ in, uint4* __restrict__ out, int P, int X, size_t plane8) {
  const size_t total = (size_t)P * X * plane8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i % plane8;
    const int x = (int)((i / plane8) % X);
    const size_t vol = i / (plane8 * X);
    const uint4* base = in + vol * X * plane8 + e;
    uint4 m = base[(size_t)x * plane8];
    __half2* mh = reinterpret_cast<__half2*>(&m);
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const int xx = x + dx;
      if (dx == 0 || xx < 0 || xx >= X) continue;
      const uint4 v = base[(size_t)xx * plane8];
      const __half2* vh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) mh[j] = __hmax2(mh[j], vh[j]);
    }
    out[i] = m;
  }
}

// begin_conv: Conv3d(1 -> Cout, 3, padding 1) on (2x-1 if affine) with zero padding; w [Cout][27] taps ordered (kx, kt, ky)
__global__ void k_conv3d_first(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                               float* __restrict__ out, int P, int X, int T, int Y, int Cout, int affine) {
  extern __shared__ float sw[];   // [27][Cout] + bias [Cout]
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) sw[(i % 27) * Cout + i / 27] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[27 * Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int cq = Cout / 4;
  const size_t total = (size_t)P * X * T * Y * cq;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cq) * 4;
    size_t v = i / cq;
    const int y = (int)(v % Y); v /= Y;
    const int t = (int)(v % T); v /= T;
    const int xs = (int)(v % X);
    const size_t p = v / X;
    float4 acc = *reinterpret_cast<const float4*>(sw + 27 * Cout + c4);
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int kt = 0; kt < 3; ++kt)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int xx = xs + kx - 1, tt = t + kt - 1, yy = y + ky - 1;
          if (xx < 0 || xx >= X || tt < 0 || tt >= T || yy < 0 || yy >= Y) continue;
          float h = x[((p * X + xx) * T + tt) * Y + yy];
          if (affine) h = 2.f * h - 1.f;
          const float4 wv = *reinterpret_cast<const float4*>(sw + ((kx * 3 + kt) * 3 + ky) * Cout + c4);
          acc.x += h * wv.x; acc.y += h * wv.y; acc.z += h * wv.z; acc.w += h * wv.w;
        }
    *reinterpret_cast<float4*>(out + (i / cq) * Cout + c4) = acc;
  }
}

// end_conv: Conv3d(C -> 1, 3, padding 1) of the f16 operand, + bias, / sigma[label[p]].  One warp per output voxel,
// lane = 4 channels per 128-channel group; w [27][C] taps ordered (kx, kt, ky).
__global__ void __launch_bounds__(256) k_conv3d_last(const __half* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                                                     const float* __restrict__ sigmas, const int64_t* __restrict__ labels,
                                                     float* __restrict__ out, int P, int X, int T, int Y, int C) {
  extern __shared__ float sw[];   // [27][C]
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t total = (size_t)P * X * T * Y;
  const size_t wstride = (size_t)gridDim.x * (blockDim.x / 32);
  for (size_t v = blockIdx.x * (size_t)(blockDim.x / 32) + threadIdx.x / 32; v < total; v += wstride) {
    size_t r = v;
    const int y = (int)(r % Y); r /= Y;
    const int t = (int)(r % T); r /= T;
    const int xs = (int)(r % X);
    const size_t p = r / X;
    float acc = 0.f;
    for (int tap = 0; tap < 27; ++tap) {
      const int xx = xs + tap / 9 - 1, tt = t + (tap / 3) % 3 - 1, yy = y + tap % 3 - 1;
      if (xx < 0 || xx >= X || tt < 0 || tt >= T || yy < 0 || yy >= Y) continue;
      const __half* src = a + (((p * X + xx) * T + tt) * Y + yy) * C;
      for (int c0 = lane * 4; c0 < C; c0 += 128) {
        const uint2 raw = *reinterpret_cast<const uint2*>(src + c0);
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        const float4 wv = *reinterpret_cast<const float4*>(sw + tap * C + c0);
        acc += f0.x * wv.x + f0.y * wv.y + f1.x * wv.z + f1.y * wv.w;
      }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[v] = (acc + (bias ? bias[0] : 0.f)) / sigmas[labels[p]];
  }
}

// out[n][t2][y][k*C + c] = in[n][stride*t2 + offset0 + k][y][c] (0 outside [0,T)): the T-taps of the (1,1,K) temporal
// convolutions laid side by side so they run as ONE 1x1 implicit GEMM with Cin' = K*C.  16-byte vectors.
__global__ void k_gather_t(const uint4* __restrict__ in, uint4* __restrict__ out, size_t NS, int T, int T2, int Y, int C8, int stride,
                           int offset0, int K) {
  const size_t total = NS * T2 * Y * K * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    size_t r = i / C8;
    const int k = (int)(r % K); r /= K;
    const int y = (int)(r % Y); r /= Y;
    const int t2 = (int)(r % T2);
    const size_t n = r / T2;
    const int t = stride * t2 + offset0 + k;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (t >= 0 && t < T) v = in[((n * T + t) * Y + y) * C8 + c];
    out[i] = v;
  }
}

// out[n][2m+ph][y][c] = in[n][m][y][ph*C + c]: the two output phases of the stride-2 transposed temporal convolution,
// computed side by side by one GEMM, back onto the time axis; also writes f16(ELU(.)) for the next convolution.
__global__ void k_interleave_t(const float4* __restrict__ in, float4* __restrict__ out32, uint2* __restrict__ out16, size_t NS, int T,
                               int Y, int C4) {
  const size_t total = NS * T * 2 * Y * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    size_t r = i / C4;
    const int y = (int)(r % Y); r /= Y;
    const int t = (int)(r % (2 * T));
    const size_t n = r / (2 * T);
    const float4 v = in[(((n * T + (t >> 1)) * Y + y) * 2 + (t & 1)) * C4 + c];
    out32[i] = v;
    if (out16) {
      uint2 pk;
      pk.x = pack_half2_sat(elu_f16bound(v.x), elu_f16bound(v.y));
      pk.y = pack_half2_sat(elu_f16bound(v.z), elu_f16bound(v.w));
      out16[i] = pk;
    }
  }
}

// out = a + (elu_b ? ELU(b) : b)
__global__ void k_add_act(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out, size_t n4, int elu_b) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 x = a[i];
    float4 y = b[i];
    if (elu_b) { y.x = elu1(y.x); y.y = elu1(y.y); y.z = elu1(y.z); y.w = elu1(y.w); }
    out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
}

// Patch fold of the 2D+time sampler (helpers/utils.py:330-359 `reshape_temporal_dim`, with the optional random roll
// of ALD_optimizers.py:466-470,495-499 folded in): vol[p][kx][t][ky] <-> state[pl][b][t][(h1*k + kx - sh) mod H]
// [(w1*k + ky - sw) mod W], p = ((pl*B + b)*H/k + h1)*W/k + w1.  unfold = the inverse scatter (same index map).
__global__ void k_patch_fold(float* __restrict__ state, float* __restrict__ vol, int B, int T, int H, int W, int k, int sh, int sw,
                             const int* __restrict__ shifts, const int* __restrict__ cursor, int unfold) {
  if (shifts != nullptr) {                                       // per-step shifts from a device table (captured graphs)
    const int c = *cursor;
    sh = shifts[2 * c];
    sw = shifts[2 * c + 1];
  }
  const int H1 = H / k, W1 = W / k;
  const size_t total = (size_t)2 * B * T * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ky = (int)(i % k);
    size_t r = i / k;
    const int t = (int)(r % T); r /= T;
    const int kx = (int)(r % k); r /= k;
    const int w1 = (int)(r % W1); r /= W1;
    const int h1 = (int)(r % H1);
    const size_t plb = r / H1;                                   // pl*B + b
    int h = h1 * k + kx - sh, w = w1 * k + ky - sw;
    h = ((h % H) + H) % H;
    w = ((w % W) + W) % W;
    const size_t si = ((plb * T + t) * H + h) * W + w;
    if (unfold) state[si] = vol[i];
    else vol[i] = state[si];
  }
}

}  // namespace ipdm

using namespace ipdm;

extern "C" int ipdm_patch_fold(float* state, float* vol, int B, int T, int H, int W, int k, int shift_h, int shift_w, int unfold,
                               void* stream) {
  IPDM_REQUIRE(state && vol && B >= 1 && T >= 1 && k >= 1, IPDM_E_BADARG, "patch_fold: bad argument");
  IPDM_REQUIRE(H % k == 0 && W % k == 0, IPDM_E_BADARG, "patch_fold: H=%d, W=%d must be multiples of the patch size %d", H, W, k);
  const size_t n = (size_t)2 * B * T * H * W;
  k_patch_fold<<<vgrid(n, 256), 256, 0, as_stream(stream)>>>(state, vol, B, T, H, W, k, shift_h, shift_w, nullptr, nullptr, unfold);
  return launched("k_patch_fold");
}

extern "C" int ipdm_patch_fold_sched(float* state, float* vol, int B, int T, int H, int W, int k, const int* shifts, const int* cursor,
                                     int unfold, void* stream) {
  IPDM_REQUIRE(state && vol && shifts && cursor && B >= 1 && T >= 1 && k >= 1, IPDM_E_BADARG, "patch_fold_sched: bad argument");
  IPDM_REQUIRE(H % k == 0 && W % k == 0, IPDM_E_BADARG, "patch_fold_sched: H=%d, W=%d must be multiples of the patch size %d", H, W, k);
  const size_t n = (size_t)2 * B * T * H * W;
  k_patch_fold<<<vgrid(n, 256), 256, 0, as_stream(stream)>>>(state, vol, B, T, H, W, k, 0, 0, shifts, cursor, unfold);
  return launched("k_patch_fold");
}

extern "C" int ipdm_maxpool5_slices_f16(const void* in_f16, void* out_f16, int P, int X, size_t plane_elems, void* stream) {
  IPDM_REQUIRE(in_f16 && out_f16 && P >= 1 && X >= 1, IPDM_E_BADARG, "maxpool5_slices: bad argument");
  IPDM_REQUIRE(plane_elems % 8 == 0, IPDM_E_UNSUPPORTED, "maxpool5_slices: slice size must be a multiple of 8 elements");
  const size_t n = (size_t)P * X * (plane_elems / 8);
  k_maxpool5_slices<<<vgrid(n, 256), 256, 0, as_stream(stream)>>>((const uint4*)in_f16, (uint4*)out_f16, P, X, plane_elems / 8);
  return launched("k_maxpool5_slices");
}

extern "C" int ipdm_conv3d_first(const float* x, const float* w, const float* bias, float* out, int P, int X, int T, int Y, int Cout,
                                 int affine, void* stream) {
  IPDM_REQUIRE(x && w && out, IPDM_E_BADARG, "conv3d_first: null pointer");
  IPDM_REQUIRE(Cout % 4 == 0 && Cout <= 1024, IPDM_E_UNSUPPORTED, "conv3d_first: Cout=%d", Cout);
  const size_t n = (size_t)P * X * T * Y * (Cout / 4);
  k_conv3d_first<<<vgrid(n, 256), 256, (size_t)28 * Cout * sizeof(float), as_stream(stream)>>>(x, w, bias, out, P, X, T, Y, Cout, affine);
  return launched("k_conv3d_first");
}

extern "C" int ipdm_conv3d_last(const void* in_f16, const float* w, const float* bias, const float* sigmas, const int64_t* labels,
                                float* out, int P, int X, int T, int Y, int C, void* stream) {
  IPDM_REQUIRE(in_f16 && w && sigmas && labels && out, IPDM_E_BADARG, "conv3d_last: null pointer");
  IPDM_REQUIRE(C % 128 == 0 && C <= 512, IPDM_E_UNSUPPORTED, "conv3d_last: C=%d must be a multiple of 128 (<= 512)", C);
  const size_t smem = (size_t)27 * C * sizeof(float);
  if (smem > 48 * 1024) IPDM_CUDA(cudaFuncSetAttribute(k_conv3d_last, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t n = (size_t)P * X * T * Y;
  k_conv3d_last<<<vgrid(n, 8), 256, smem, as_stream(stream)>>>((const __half*)in_f16, w, bias, sigmas, labels, out, P, X, T, Y, C);
  return launched("k_conv3d_last");
}

extern "C" int ipdm_gather_t_f16(const void* in_f16, void* out_f16, size_t NS, int T, int T2, int Y, int C, int stride, int offset0, int K,
                                 void* stream) {
  IPDM_REQUIRE(in_f16 && out_f16 && NS >= 1 && T >= 1 && T2 >= 1 && K >= 1 && stride >= 1, IPDM_E_BADARG, "gather_t: bad argument");
  IPDM_REQUIRE(C % 8 == 0, IPDM_E_UNSUPPORTED, "gather_t: C must be a multiple of 8");
  const size_t n = NS * T2 * Y * K * (C / 8);
  k_gather_t<<<vgrid(n, 256), 256, 0, as_stream(stream)>>>((const uint4*)in_f16, (uint4*)out_f16, NS, T, T2, Y, C / 8, stride, offset0, K);
  return launched("k_gather_t");
}

extern "C" int ipdm_interleave_t(const float* in, float* out_f32, void* out_elu_f16, size_t NS, int T, int Y, int C, void* stream) {
  IPDM_REQUIRE(in && out_f32 && NS >= 1, IPDM_E_BADARG, "interleave_t: bad argument");
  IPDM_REQUIRE(C % 4 == 0, IPDM_E_UNSUPPORTED, "interleave_t: C must be a multiple of 4");
  const size_t n = NS * T * 2 * Y * (C / 4);
  k_interleave_t<<<vgrid(n, 256), 256, 0, as_stream(stream)>>>((const float4*)in, (float4*)out_f32, (uint2*)out_elu_f16, NS, T, Y, C / 4);
  return launched("k_interleave_t");
}

extern "C" int ipdm_add_act(const float* a, const float* b, float* out, size_t n, int elu_b, void* stream) {
  IPDM_REQUIRE(a && b && out, IPDM_E_BADARG, "add_act: null pointer");
  IPDM_REQUIRE(n % 4 == 0, IPDM_E_UNSUPPORTED, "add_act: n must be a multiple of 4");
  k_add_act<<<vgrid(n / 4, 256), 256, 0, as_stream(stream)>>>((const float4*)a, (const float4*)b, (float4*)out, n / 4, elu_b);
  return launched("k_add_act");
}
