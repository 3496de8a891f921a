// Shared pieces of the implicit-GEMM convolution kernels: parameter block, the TMEM -> global epilogue,
// tensor-map cache entry points.
#pragma once
#include "tc_common.cuh"

namespace ipdm {

constexpr int BLOCK_M = 128;   // output channels per accumulator (TMEM lanes)
constexpr int BLOCK_N = 256;   // pixels per accumulator (TMEM columns)
constexpr int BLOCK_K = 64;    // f16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_PITCH = BLOCK_M + 4;          // floats per slab row: +4 keeps float4 alignment and staggers banks
constexpr int SLAB_BYTES = 2 * 32 * EPI_PITCH * 4;   // double-buffered [32 pixels][128 ch] fp32 staging

struct IgemmParams {
  const float* bias;
  const float* residual;
  float* out_f32;
  __half* out_f16;
  double* stats;
  int N, H, W, Cin, Cout, taps, dilation, flags;
  int tiles_w, tiles_h;
  int slices, slice_shift;   // volumes of `slices` consecutive images; input slice = output slice + slice_shift
};

int get_weight_map(const void* w, int Cout, int K, CUtensorMap* out);
int get_act_map(const void* x, int N, int H, int W, int C, int box_w, int box_h, int slices, CUtensorMap* out);
int launch_conv_halo(const ipdm_conv_desc& d, cudaStream_t s);
bool conv_halo_supports(const ipdm_conv_desc& d);
extern int g_conv_variant;       // 0 = auto, 1 = force the per-tap tile kernel (diagnostics)

// Epilogue of `NCHUNK` (runtime) 32-column chunks of one accumulator (128 channels x up to 256 pixels) by a TEAM of 4 warps
// (128 threads, named barrier `BAR`, a runtime value so that both teams share ONE copy of this code -- the
// epilogue is instruction-cache bound otherwise); chunks [chunk0, chunk0 + NCHUNK).
// TMEM lane = output channel, column = pixel j = py*TW + px of a (256/TW) x TW pixel tile at (h0, w0).
// Each warp pulls 32 columns for its 32 channels, transposes them through the team's shared-memory slab, and
// the team then streams the [32 pixels][128 ch] slab with 16-byte accesses: thread = 4 consecutive channels of
// one pixel, so a warp touches one whole 512-byte pixel row of the NHWC tensor per instruction.  `wait_acc()` is
// called once, after the first residual loads are in flight and before the first TMEM read.
// SLABS = 2: double-buffered slab, one barrier per chunk; SLABS = 1: single slab, two barriers per chunk.
// MODE bits: 1 = residual, 2 = fp32 output, 4 = f16 output, 8 = 2x2 mean-pool.
template <int MODE, int TW, int SLABS, class WaitAcc>
__device__ __forceinline__ void conv_epilogue(const IgemmParams& p, float* slab, uint32_t tmem_acc, int quad, int lane,
                                              int n, int h0, int w0, int m0, int chunk0, int NCHUNK, int BAR, WaitAcc wait_acc) {
  constexpr bool kRes = (MODE & 1) != 0, kOut32 = (MODE & 2) != 0, kOut16 = (MODE & 4) != 0, pool = (MODE & 8) != 0;
  constexpr int ROWS_PER_CHUNK = 32 / TW;       // tile rows covered by 32 columns
  constexpr int PW = TW / 2;                    // pooled pixels per pooled row
  constexpr int NPX = pool ? 2 : 8;             // output pixels per thread per chunk
  const int te = quad * 32 + lane;              // 0..127 within the team
  const int c4 = (te & 31) * 4;                 // first of this thread's 4 channels (within the 128)
  const int prow = te >> 5;                     // pixel sub-row 0..3
  const int Ho = pool ? p.H / 2 : p.H, Wo = pool ? p.W / 2 : p.W;
  const int oy0 = pool ? h0 / 2 : h0, ox0 = pool ? w0 / 2 : w0;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) bias4 = *reinterpret_cast<const float4*>(p.bias + m0 + c4);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16);
  // this thread's pixels of a chunk are q = prow + 4*i: row i / (TW/4) of the chunk, column prow + 4*(i % (TW/4))
  // (pooled: TW/2 columns, 8 pixels per chunk) -> offsets are a per-chunk base plus compile-time multiples
  constexpr int CPR = (pool ? PW : TW) / 4 > 0 ? (pool ? PW : TW) / 4 : 1;      // thread-pixels per output row
  constexpr int OROWS = pool ? ROWS_PER_CHUNK / 2 : ROWS_PER_CHUNK;             // output rows per chunk
  const size_t row_stride = (size_t)Wo * p.Cout;
  const size_t col_stride = (size_t)4 * p.Cout;
  const bool col_in_tile = pool ? (prow < PW) : true;                             // PW = 4 (TW = 8): all four sub-rows valid
  auto chunk_base_of = [&](int chunk) { return (((size_t)n * Ho + oy0 + chunk * OROWS) * Wo + ox0 + prow) * p.Cout + m0 + c4; };
  auto pixel_ok = [&](int chunk, int i) {
    return col_in_tile && (oy0 + chunk * OROWS + i / CPR) < Ho && (ox0 + prow + 4 * (i % CPR)) < Wo;
  };
  // Residual tiles are software-pipelined one chunk ahead: the loads of chunk c+1 are issued right after chunk c's
  // accumulator values have been staged (their registers are dead by then), so a load has a whole chunk period to
  // land; the first chunk's loads go out before the wait for the accumulator.
  float4 rcur[NPX], rnext[NPX];
  auto issue_res = [&](int chunk, float4* r) {
    const size_t base = chunk_base_of(chunk);
#pragma unroll
    for (int i = 0; i < NPX; ++i) {
      r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pixel_ok(chunk, i)) r[i] = *reinterpret_cast<const float4*>(p.residual + base + (i / CPR) * row_stride + (i % CPR) * col_stride);
    }
  };
  if (kRes) issue_res(chunk0, rcur);
  wait_acc();
#pragma unroll 1
  for (int cc = 0; cc < NCHUNK; ++cc) {
    const int chunk = chunk0 + cc;
    float v[32];
    tmem_ld32(taddr + chunk * 32, v);
    float* buf = slab + (SLABS == 2 ? (cc & 1) * (32 * EPI_PITCH) : 0);
    if (SLABS == 1 && cc > 0) asm volatile("bar.sync %0, 128;" ::"r"(BAR) : "memory");   // previous chunk fully consumed
    if (!pool) {
#pragma unroll
      for (int j = 0; j < 32; ++j) buf[j * EPI_PITCH + te] = v[j];
    } else {
      // pooled pixel q' = r*PW + cx  <-  rows 2r, 2r+1 and columns 2cx, 2cx+1 of the chunk
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) {
        const int r = qq / PW, cx = qq % PW;
        const int a = (2 * r) * TW + 2 * cx, b = (2 * r + 1) * TW + 2 * cx;
        buf[qq * EPI_PITCH + te] = (((v[a] + v[b]) + v[a + 1]) + v[b + 1]) * 0.25f;
      }
    }
    asm volatile("bar.sync %0, 128;" ::"r"(BAR) : "memory");
    if (kRes && cc + 1 < NCHUNK) issue_res(chunk + 1, rnext);
    const size_t chunk_base = chunk_base_of(chunk);
#pragma unroll
    for (int i = 0; i < NPX; ++i) {
      if (pixel_ok(chunk, i)) {
        const size_t off = chunk_base + (i / CPR) * row_stride + (i % CPR) * col_stride;
        const int q = prow + 4 * i;
        float4 a = *reinterpret_cast<const float4*>(buf + q * EPI_PITCH + c4);
        a.x += bias4.x; a.y += bias4.y; a.z += bias4.z; a.w += bias4.w;
        const float4 pre = a;
        if (kRes) {
          float4 r = rcur[i];
          if (p.flags & IPDM_CONV_RES_ELU) { r.x = elu_fast(r.x); r.y = elu_fast(r.y); r.z = elu_fast(r.z); r.w = elu_fast(r.w); }
          a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
        }
        if (kOut32) *reinterpret_cast<float4*>(p.out_f32 + off) = a;
        if (kOut16) {
          float4 h = (p.flags & IPDM_CONV_F16_PRE_RES) ? pre : a;
          if (p.flags & IPDM_CONV_F16_ELU) { h.x = elu_fast(h.x); h.y = elu_fast(h.y); h.z = elu_fast(h.z); h.w = elu_fast(h.w); }
          uint2 pk;
          pk.x = pack_half2_sat(h.x, h.y);
          pk.y = pack_half2_sat(h.z, h.w);
          *reinterpret_cast<uint2*>(p.out_f16 + off) = pk;
        }
        s1[0] += a.x; s1[1] += a.y; s1[2] += a.z; s1[3] += a.w;
        s2[0] += a.x * a.x; s2[1] += a.y * a.y; s2[2] += a.z * a.z; s2[3] += a.w * a.w;
      }
    }
    if (kRes) {
#pragma unroll
      for (int i = 0; i < NPX; ++i) rcur[i] = rnext[i];
    }
  }
  // the slab is reused (by the statistics below and by the next accumulator): everyone must be done reading it
  asm volatile("bar.sync %0, 128;" ::"r"(BAR) : "memory");
  if (p.stats) {
    // combine the four pixel sub-rows that share a channel group, then 2 atomics per channel
    float* red = slab;                                    // [4][128][2]
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      red[(prow * 128 + c4 + k) * 2] = s1[k];
      red[(prow * 128 + c4 + k) * 2 + 1] = s2[k];
    }
    asm volatile("bar.sync %0, 128;" ::"r"(BAR) : "memory");
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      t1 += red[(r * 128 + te) * 2];
      t2 += red[(r * 128 + te) * 2 + 1];
    }
    atomicAdd(&p.stats[((size_t)(n / p.slices) * p.Cout + m0 + te) * 2], (double)t1);
    atomicAdd(&p.stats[((size_t)(n / p.slices) * p.Cout + m0 + te) * 2 + 1], (double)t2);
    asm volatile("bar.sync %0, 128;" ::"r"(BAR) : "memory");
  }
}

}  // namespace ipdm
