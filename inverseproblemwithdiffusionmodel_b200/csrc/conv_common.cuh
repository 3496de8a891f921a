// Shared pieces of the implicit-GEMM convolution kernels: parameter block, the TMEM -> global epilogue,
// tensor-map cache entry points.
#pragma once
#include "tc_common.cuh"

namespace ipdm {

constexpr int BLOCK_M = 128;   // output channels per accumulator (TMEM lanes)
constexpr int BLOCK_N = 256;   // pixels per accumulator (TMEM columns)
constexpr int BLOCK_K = 64;    // f16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;

struct IgemmParams {
  const float* bias;
  const float* residual;      // f32 residual stream, or (trunk16) the same pointer reinterpreted as f16
  float* out_f32;             // f32 result, or (trunk16) f16 result ("raw" 16-bit residual stream)
  __half* out_f16;
  double* stats;
  int N, H, W, Cin, Cout, taps, dilation, flags;
  int tiles_w, tiles_h;
  int slices, slice_shift;   // volumes of `slices` consecutive images; input slice = output slice + slice_shift
  float acc_scale, out16_scale;   // operand exponent shift (1 = off): accumulator * acc_scale, f16 operand output * out16_scale
};

int get_weight_map(const void* w, int Cout, int K, CUtensorMap* out);
int get_act_map(const void* x, int N, int H, int W, int C, int box_w, int box_h, int slices, CUtensorMap* out);
int launch_conv_halo(const ipdm_conv_desc& d, cudaStream_t s);
bool conv_halo_supports(const ipdm_conv_desc& d);
extern int g_conv_res_prefetch;  // halo kernel: L2 bulk prefetch of the residual tile by the producer warp (0 = off, the default: measured 5-10 % SLOWER when on)
extern int g_conv_pdl;           // halo kernel launched with programmatic stream serialization (PDL)
extern int g_conv_variant;       // 0 = auto, 1 = force the per-tap tile kernel (diagnostics)

// ---------------------------------------------------------------------------------------------------------------
// Epilogue WITHOUT shared memory: the 32 (channel) x 32 (pixel) block a warp pulls from TMEM is re-distributed with two
// xor-shuffle stages so that a thread ends up with 4 consecutive channels of 8 pixels -- 16-byte global accesses,
// every 8 lanes covering one 128-byte line of the NHWC tensors -- instead of being transposed through a slab.  The
// main loop is bound by the shared-memory port (UMMA operand reads + TMA fills); an epilogue that stays off it costs
// the MMA nothing, needs no team barrier (every warp is independent) and frees 66 KB per CTA.  (Measured equal to the
// slab-transposing epilogue it replaced, within noise, in all three output modes: profiles/experiments/.)
// `wait_acc()` is called once, after the first residual loads are in flight and before the first TMEM read; residual
// tiles are software-pipelined one chunk ahead.
//   before: lane = (c0..c4), register j = (p0..p4)            (channel = TMEM lane, pixel = TMEM column)
//   stage xor 1 swaps c0 <-> p0, stage xor 2 swaps c1 <-> p1
//   after:  lane = (p0, p1, c2, c3, c4), register = (c0, c1, p2, p3, p4): float4 #i = channels 4*(lane>>2)..+3 of pixel
//           (lane & 3) + 4*i of the chunk
// MODE bits: 1 = residual, 2 = result output, 4 = f16 operand output, 8 = 2x2 mean-pool (pooled before the shuffles: 8
// values), 16 = the residual stream is kept in 16 bits: `residual` and the result output (`out_f32`) are f16 tensors.
// The 16-bit stream takes the residual-variant convolution from 12 to 8 bytes per output element (f16 in, f16 residual,
// f16 result, f16 operand copy), i.e. from 192 to 288 FLOP/B at 128 channels: over the ridge of the measured peaks.
template <bool T16> struct ResElem;
template <> struct ResElem<false> {
  using type = float4;
  static __device__ __forceinline__ float4 to_f32(float4 v) { return v; }
};
template <> struct ResElem<true> {
  using type = uint2;   // four f16
  static __device__ __forceinline__ float4 to_f32(uint2 h) {
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  }
};

template <int NV>
__device__ __forceinline__ void lane_register_swap(float* v, int lane) {
  // NV values; exchanges register bit 0 with lane bit 0, then register bit 1 with lane bit 1
#pragma unroll
  for (int bit = 0; bit < 2; ++bit) {
    const bool up = (lane >> bit) & 1;
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      if ((r >> bit) & 1) continue;
      const int r1 = r | (1 << bit);
      float send = up ? v[r] : v[r1];
      send = __shfl_xor_sync(0xffffffffu, send, 1 << bit);
      if (up) v[r] = send; else v[r1] = send;
    }
  }
}

// The chunk loop of the epilogue.  FAST = 0: every feature read from the descriptor at run time, per-pixel bounds checks.
// FAST = 1 / 2: the item's tile lies inside the image and the descriptor has no bias, no InstanceNorm++ sums and no ELU
// on the residual -- every RCU / CRP convolution of the RefineNet, 82 of 113 launches -- with the f16 output being
// ELU(result) (1, RCU) or the pre-residual value (2, CRP).  There the unused work is compiled out rather than predicated
// and the eight pixels are one straight-line block: the epilogue warps (two per scheduler) are ISSUE-bound in the
// residual mode, so an instruction that is not there is time.  SASS, executed instructions per 4-channel pixel:
// ~110 in the first version (predicated ELU rescue paths, per-pixel flag tests, 64-bit index arithmetic) -> ~27.
template <int MODE, int TW, int FAST, class WaitAcc>
__device__ __forceinline__ void conv_epilogue_chunks(const IgemmParams& p, uint32_t tmem_acc, int quad, int lane, int n, int h0, int w0,
                                                     int m0, int chunk0, int NCHUNK, WaitAcc wait_acc) {
  constexpr bool kRes = (MODE & 1) != 0, kOut32 = (MODE & 2) != 0, kOut16 = (MODE & 4) != 0, pool = (MODE & 8) != 0;
  constexpr bool T16 = (MODE & 16) != 0;
  constexpr bool LEAN = FAST != 0;
  constexpr int NPX = pool ? 2 : 8;             // float4 groups (pixels) per thread per chunk
  constexpr int OW = pool ? TW / 2 : TW;        // output pixels per chunk row
  constexpr int OROWS = pool ? (32 / TW) / 2 : 32 / TW;   // output rows per chunk
  constexpr int IPR = OW / 4;                   // of a thread's pixels q = psub + 4*i, IPR consecutive i share a row
  static_assert(OW % 4 == 0 && NPX % IPR == 0, "chunk geometry");
  const int psub = lane & 3;                    // pixel sub-index of this thread
  const int c4 = quad * 32 + (lane >> 2) * 4;   // first of this thread's 4 channels within the 128-channel tile
  const int Ho = pool ? p.H / 2 : p.H, Wo = pool ? p.W / 2 : p.W;
  const int oy0 = pool ? h0 / 2 : h0, ox0 = pool ? w0 / 2 : w0;
  const bool f16_elu = FAST == 1 ? true : FAST == 2 ? false : (p.flags & IPDM_CONV_F16_ELU) != 0;
  const bool f16_pre = FAST == 1 ? false : FAST == 2 ? true : (p.flags & IPDM_CONV_F16_PRE_RES) != 0;
  const bool res_elu = !LEAN && (p.flags & IPDM_CONV_RES_ELU) != 0;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!LEAN && p.bias) bias4 = *reinterpret_cast<const float4*>(p.bias + m0 + c4);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16);
  // element offset of this thread's first pixel of chunk 0 (row oy0, column ox0 + psub); pixel i of a chunk adds
  // i/IPR rows and 4*(i%IPR) pixels: a 32-bit element offset eo(i), so an address is one IMAD.WIDE on the chunk's base
  const uint32_t row_stride = (uint32_t)Wo * p.Cout, col_stride = 4u * p.Cout;
  const size_t off0 = (((size_t)n * Ho + oy0) * Wo + ox0 + psub) * p.Cout + m0 + c4;
  const int cols_left = Wo - ox0 - psub;        // pixel i is inside the image iff 4*(i%IPR) < cols_left and its row < Ho
  auto eo = [&](int i) -> uint32_t { return (uint32_t)(i / IPR) * row_stride + (uint32_t)(i % IPR) * col_stride; };
  auto inside = [&](int i, int rows_left) -> bool { return LEAN || (i / IPR < rows_left && 4 * (i % IPR) < cols_left); };
  // residual tiles: four f32 (float4) or four f16 (uint2) per pixel, software-pipelined one chunk ahead
  using RT = typename ResElem<T16>::type;
  RT rcur[NPX], rnext[NPX];
  auto issue_res = [&](int chunk, RT* r) {
    const int rows_left = Ho - oy0 - chunk * OROWS;
    const char* rp = reinterpret_cast<const char*>(p.residual) + (off0 + (size_t)(chunk * OROWS) * row_stride) * sizeof(RT) / 4;
#pragma unroll
    for (int i = 0; i < NPX; ++i) {
      r[i] = RT{};
      if (inside(i, rows_left)) r[i] = *reinterpret_cast<const RT*>(rp + (size_t)eo(i) * (sizeof(RT) / 4));
    }
  };
  if (kRes) issue_res(chunk0, rcur);
  wait_acc();
#pragma unroll 1
  for (int cc = 0; cc < NCHUNK; ++cc) {
    const int chunk = chunk0 + cc;
    float v[32];
    tmem_ld32(taddr + chunk * 32, v);
    if (kRes && cc + 1 < NCHUNK) issue_res(chunk + 1, rnext);
    float w[4 * NPX];
    if (!pool) {
#pragma unroll
      for (int j = 0; j < 32; ++j) w[j] = v[j];
    } else {
      // pooled pixel q' = r*PW + cx  <-  rows 2r, 2r+1 and columns 2cx, 2cx+1 of the chunk (TW = 8: 2 x 4 pooled pixels)
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) {
        const int r = qq / (TW / 2), cx = qq % (TW / 2);
        const int a = (2 * r) * TW + 2 * cx, b = (2 * r + 1) * TW + 2 * cx;
        w[qq] = (((v[a] + v[b]) + v[a + 1]) + v[b + 1]) * 0.25f;
      }
    }
    lane_register_swap<4 * NPX>(w, lane);
    const size_t offc = off0 + (size_t)(chunk * OROWS) * row_stride;
    char* o32 = T16 ? reinterpret_cast<char*>(p.out_f32) + offc * 2 : reinterpret_cast<char*>(p.out_f32 + offc);
    char* o16 = reinterpret_cast<char*>(p.out_f16 + offc);
    const int rows_left = Ho - oy0 - chunk * OROWS;
#pragma unroll
    for (int i = 0; i < NPX; ++i) {
      if (inside(i, rows_left)) {
        float4 a = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        if (!LEAN) {
          a.x = a.x * p.acc_scale + bias4.x; a.y = a.y * p.acc_scale + bias4.y; a.z = a.z * p.acc_scale + bias4.z; a.w = a.w * p.acc_scale + bias4.w;
        }
        const float4 pre = a;
        if (kRes) {
          float4 r = ResElem<T16>::to_f32(rcur[i]);
          if (res_elu) { r.x = elu_fast(r.x); r.y = elu_fast(r.y); r.z = elu_fast(r.z); r.w = elu_fast(r.w); }
          a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
        }
        if (kOut32) {
          if (T16) {
            uint2 pk;
            pk.x = pack_half2_sat(a.x, a.y);
            pk.y = pack_half2_sat(a.z, a.w);
            *reinterpret_cast<uint2*>(o32 + (size_t)eo(i) * 2) = pk;
          } else {
            *reinterpret_cast<float4*>(o32 + (size_t)eo(i) * 4) = a;
          }
        }
        if (kOut16) {
          float4 h = (kRes && f16_pre) ? pre : a;
          if (f16_elu) { h.x = elu_fast(h.x); h.y = elu_fast(h.y); h.z = elu_fast(h.z); h.w = elu_fast(h.w); }
          if (!LEAN) { h.x *= p.out16_scale; h.y *= p.out16_scale; h.z *= p.out16_scale; h.w *= p.out16_scale; }
          uint2 pk;
          pk.x = pack_half2_sat(h.x, h.y);
          pk.y = pack_half2_sat(h.z, h.w);
          *reinterpret_cast<uint2*>(o16 + (size_t)eo(i) * 2) = pk;
        }
        if (!LEAN) {
          s1[0] += a.x; s1[1] += a.y; s1[2] += a.z; s1[3] += a.w;
          s2[0] += a.x * a.x; s2[1] += a.y * a.y; s2[2] += a.z * a.z; s2[3] += a.w * a.w;
        }
      }
    }
    if (kRes) {
#pragma unroll
      for (int i = 0; i < NPX; ++i) rcur[i] = rnext[i];
    }
  }
  if (!LEAN && p.stats) {
    // the four lanes that share a channel group (lane bits 0, 1) combine, then 8 fp64 atomics from one of them
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 1);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], 1);
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 2);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], 2);
    }
    if (psub == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        atomicAdd(&p.stats[((size_t)(n / p.slices) * p.Cout + m0 + c4 + k) * 2], (double)s1[k]);
        atomicAdd(&p.stats[((size_t)(n / p.slices) * p.Cout + m0 + c4 + k) * 2 + 1], (double)s2[k]);
      }
    }
  }
}

template <int MODE, int TW, class WaitAcc>
__device__ __forceinline__ void conv_epilogue_shfl(const IgemmParams& p, uint32_t tmem_acc, int quad, int lane, int n, int h0, int w0,
                                                   int m0, int chunk0, int NCHUNK, WaitAcc wait_acc) {
  constexpr bool kRes = (MODE & 1) != 0, kOut16 = (MODE & 4) != 0, pool = (MODE & 8) != 0;
  constexpr int OROWS = pool ? (32 / TW) / 2 : 32 / TW, OW = pool ? TW / 2 : TW;
  const int Ho = pool ? p.H / 2 : p.H, Wo = pool ? p.W / 2 : p.W;
  const int oy0 = pool ? h0 / 2 : h0, ox0 = pool ? w0 / 2 : w0;
  // warp-uniform: the chunks [chunk0, chunk0 + NCHUNK) of this item lie inside the image and nothing optional is asked for
  const bool lean = p.bias == nullptr && p.stats == nullptr && (p.flags & IPDM_CONV_RES_ELU) == 0 && p.acc_scale == 1.f && p.out16_scale == 1.f &&
                    oy0 + (chunk0 + NCHUNK) * OROWS <= Ho && ox0 + OW <= Wo;
  const bool elu = (p.flags & IPDM_CONV_F16_ELU) != 0, pre = kRes && (p.flags & IPDM_CONV_F16_PRE_RES) != 0;
  if (lean && (!kOut16 || (elu && !pre)))
    conv_epilogue_chunks<MODE, TW, 1>(p, tmem_acc, quad, lane, n, h0, w0, m0, chunk0, NCHUNK, wait_acc);
  else if (lean && !elu && (pre || !kRes))
    conv_epilogue_chunks<MODE, TW, 2>(p, tmem_acc, quad, lane, n, h0, w0, m0, chunk0, NCHUNK, wait_acc);
  else
    conv_epilogue_chunks<MODE, TW, 0>(p, tmem_acc, quad, lane, n, h0, w0, m0, chunk0, NCHUNK, wait_acc);
}

}  // namespace ipdm
