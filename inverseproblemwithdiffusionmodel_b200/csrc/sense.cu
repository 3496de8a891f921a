// SENSE operator, centred FFTs and the fused Langevin + data-consistency step for sm_100a.
//
// Data flow: i2k(x) = sigma * P * FFT2(P * x) / sqrt(HW) with P = (-1)^(h+w) and
// sigma = (-1)^(H/2+W/2), so the four fftshift copies of the reference become sign flips folded
// into loads and stores.  A 2-D transform is a row pass (along W, contiguous) and a column pass
// (along H) that meet in a TRANSPOSED scratch T[c][b][k][h]: the row kernel transforms 16 rows of
// one image for every coil and writes 128-byte h-chunks per k-space column, the column kernel reads
// whole columns contiguously and writes 16-column output tiles -- every global access is a full
// 128-byte segment.  Columns the sampling mask removes are never written, read or transformed; the
// coil multiply lives in the first pass of the forward row kernel and the conj-coil reduction in the
// last pass of the adjoint row kernel, so coil images never exist in HBM.
// For the per-step data-consistency term A^H A z the H-axis transforms cancel and one row-only
// kernel does update + prox (k_ald_sense).
#include "common.cuh"
#include "fft_core.cuh"
#include "sense_plan.h"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <utility>

namespace ipdm {

__device__ __forceinline__ int spad(int i) { return i + (i >> 3); }
template <int L> struct Tile {
  static constexpr int TPF = FftPlan<L>::TPF;
  static constexpr int E = FftRegs<L>::E;
  static constexpr int ROWS = (TPF * 16 > 512) ? 8 : 16;   // transforms per CTA
  static constexpr int NT = TPF * ROWS;
  static constexpr int PITCH = (L + L / 8) | 1;
  static constexpr size_t SMEM = (size_t)(ROWS * PITCH + L) * sizeof(cf32);
};

template <int L>
__device__ __forceinline__ void fill_twiddles(cf32* tw, int tid, int nt) {
  for (int m = tid; m < L; m += nt) {
    float s, c;
    sincospif(-2.0f * (float)m / (float)L, &s, &c);
    tw[m] = cf32{c, s};
  }
}

// register permutations between "q order" (u[q] <-> position t + q*TPF) and the leg order of pass P
template <int L, int P>
__device__ __forceinline__ void q_to_regs(const cf32* u, cf32* v) {
  constexpr int R = PassRadix<L, P>::value, NB = (L / R) / FftPlan<L>::TPF;
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int r = 0; r < R; ++r) v[i * R + r] = u[i + r * NB];
}
template <int L, int P>
__device__ __forceinline__ void regs_to_q(const cf32* v, cf32* u) {
  constexpr int R = PassRadix<L, P>::value, NB = (L / R) / FftPlan<L>::TPF;
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int r = 0; r < R; ++r) u[i + r * NB] = v[i * R + r];
}

// Per-thread twiddle registers for all passes of a length-L transform (filled once per kernel).
template <int L> struct TwRegs {
  static constexpr int N1 = PassTw<L, 1>::N, N2 = FftPlan<L>::NP > 2 ? PassTw<L, 2>::N : 0, N3 = FftPlan<L>::NP > 3 ? PassTw<L, 3>::N : 0;
  cf32 p1[N1 > 0 ? N1 : 1], p2[N2 > 0 ? N2 : 1], p3[N3 > 0 ? N3 : 1];
  __device__ __forceinline__ void init(int t, const cf32* tw) {
    if constexpr (FftPlan<L>::NP > 1) pass_twiddles<L, 1>(t, p1, tw);
    if constexpr (FftPlan<L>::NP > 2) pass_twiddles<L, 2>(t, p2, tw);
    if constexpr (FftPlan<L>::NP > 3) pass_twiddles<L, 3>(t, p3, tw);
  }
};

// Barrier among the TPF threads that share one transform.  TPF <= 32: they sit in one warp (tid = row*TPF + t),
// so a warp-level sync is enough and rows / warps run decoupled; TPF = 64: two warps, named barrier 1 + row.
template <int L>
__device__ __forceinline__ void row_sync(int row_id) {
  if constexpr (FftPlan<L>::TPF <= 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(row_id + 1), "n"(FftPlan<L>::TPF) : "memory");
  }
}

// Length-L transform of the values held in q order by the TPF threads of one row; `row` is that row's private
// smem line.  In: u[q] = input at position t+q*TPF.  Out: u[q] = output at position t+q*TPF.
// Only the row's own threads synchronise (row_sync), so the caller needs a CTA barrier only when different
// rows exchange data through the tile.
template <int L, int DIR>
__device__ __forceinline__ void fft_regs(cf32* u, cf32* row, const TwRegs<L>& T, int t, int row_id) {
  using PL = FftPlan<L>;
  constexpr int E = FftRegs<L>::E;
  cf32 v[E];
  auto ld = [&](int i) { return row[spad(i)]; };
  auto st = [&](int i, cf32 val) { row[spad(i)] = val; };
  q_to_regs<L, 0>(u, v);
  pass_compute_regtw<L, 0, DIR>(v, T.p1);
  if constexpr (PL::NP == 1) {
    regs_to_q<L, 0>(v, u);
  } else {
    row_sync<L>(row_id);  // earlier readers of this row's line are done
    pass_store<L, 0>(t, v, st);
    row_sync<L>(row_id);
    pass_load<L, 1>(t, v, ld);
    pass_compute_regtw<L, 1, DIR>(v, T.p1);
    if constexpr (PL::NP == 2) {
      regs_to_q<L, 1>(v, u);
    } else {
      row_sync<L>(row_id);
      pass_store<L, 1>(t, v, st);
      row_sync<L>(row_id);
      pass_load<L, 2>(t, v, ld);
      pass_compute_regtw<L, 2, DIR>(v, T.p2);
      if constexpr (PL::NP == 3) {
        regs_to_q<L, 2>(v, u);
      } else {
        row_sync<L>(row_id);
        pass_store<L, 2>(t, v, st);
        row_sync<L>(row_id);
        pass_load<L, 3>(t, v, ld);
        pass_compute_regtw<L, 3, DIR>(v, T.p3);
        regs_to_q<L, 3>(v, u);
      }
    }
  }
}

struct SenseArgs {
  const cf32* in;
  cf32* out;
  cf32* ws;
  const float* mre;
  const float* mim;
  const uint8_t* mask;
  int mask_frames, ncoils, batch, H, W, ssos, sparse;
  float scale;  // 1/sqrt(HW) * sigma
  int b0, nb;   // pruned kernels: this launch covers images [b0, b0 + nb) of the batch (strides still use `batch`)
};

__device__ __forceinline__ float sgn(int i) { return (i & 1) ? -1.f : 1.f; }
__device__ __forceinline__ bool col_on(const SenseArgs& a, int b, int k) {
  return a.mask == nullptr || a.mask[(size_t)(b % a.mask_frames) * a.W + k] != 0;
}

// ---- forward, pass 1: rows.  grid (ceil(H/ROWS), batch) ---------------------------------------
template <int L>
__global__ void __launch_bounds__(Tile<L>::NT) k_fwd_rows(SenseArgs a) {
  using TL = Tile<L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tile = reinterpret_cast<cf32*>(smem_raw);
  cf32* tw = tile + TL::ROWS * TL::PITCH;
  const int tid = threadIdx.x, r = tid / TL::TPF, t = tid % TL::TPF;
  const int b = blockIdx.y, h0 = blockIdx.x * TL::ROWS, h = h0 + r;
  const bool valid = h < a.H;
  fill_twiddles<L>(tw, tid, TL::NT);
  __syncthreads();
  TwRegs<L> T;
  T.init(t, tw);
  cf32 xq[TL::E];
#pragma unroll
  for (int q = 0; q < TL::E; ++q) {
    const int w = t + q * TL::TPF;
    xq[q] = valid ? cscale(a.in[((size_t)b * a.H + h) * L + w], sgn(h + w)) : cf32{0.f, 0.f};
  }
  __syncthreads();
  // coil maps of the next coil are fetched while the current coil is transformed
  cf32 mnext[TL::E];
  auto fetch_maps = [&](int c) {
#pragma unroll
    for (int q = 0; q < TL::E; ++q) {
      mnext[q] = cf32{1.f, 0.f};
      if (a.mre != nullptr && valid && c < a.ncoils) {
        const size_t mi = ((size_t)c * a.H + h) * L + t + q * TL::TPF;
        mnext[q] = cf32{a.mre[mi], a.mim ? a.mim[mi] : 0.f};
      }
    }
  };
  // which of this thread's k-space columns survive the mask (sparse masks: write them straight from registers)
  bool keep[TL::E];
#pragma unroll
  for (int q = 0; q < TL::E; ++q) keep[q] = valid && col_on(a, b, t + q * TL::TPF);
  fetch_maps(0);
  for (int c = 0; c < a.ncoils; ++c) {
    cf32 u[TL::E];
#pragma unroll
    for (int q = 0; q < TL::E; ++q) u[q] = a.mre != nullptr ? cmul(xq[q], mnext[q]) : xq[q];
    fetch_maps(c + 1);
    fft_regs<L, -1>(u, tile + r * TL::PITCH, T, t, r);
    const size_t img = (size_t)c * a.batch + b;
    if (a.sparse) {
      // few sampled columns: 8-byte scattered stores of just those, no transposition through the tile
#pragma unroll
      for (int q = 0; q < TL::E; ++q)
        if (keep[q]) a.ws[(img * L + t + q * TL::TPF) * a.H + h] = u[q];
      continue;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < TL::E; ++q) tile[r * TL::PITCH + spad(t + q * TL::TPF)] = u[q];
    __syncthreads();
    for (int idx = tid; idx < TL::ROWS * L; idx += TL::NT) {
      const int rr = idx % TL::ROWS, k = idx / TL::ROWS;
      if (h0 + rr < a.H && col_on(a, b, k)) a.ws[(img * L + k) * a.H + h0 + rr] = tile[rr * TL::PITCH + spad(k)];
    }
    __syncthreads();  // the gather read every row's line; the next coil's passes overwrite them
  }
}

// ---- forward, pass 2: columns.  grid (ceil(W/ROWS), ncoils*batch) -------------------------------
template <int L>
__global__ void __launch_bounds__(Tile<L>::NT) k_fwd_cols(SenseArgs a) {
  using TL = Tile<L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tile = reinterpret_cast<cf32*>(smem_raw);
  cf32* tw = tile + TL::ROWS * TL::PITCH;
  const int tid = threadIdx.x, cc = tid / TL::TPF, t = tid % TL::TPF;
  const size_t img = blockIdx.y;
  const int b = (int)(img % a.batch), k0 = blockIdx.x * TL::ROWS, k = k0 + cc;
  const bool active = k < a.W && col_on(a, b, k);
  const int any = __syncthreads_or(active ? 1 : 0);
  if (!any) {
    for (int idx = tid; idx < TL::ROWS * L; idx += TL::NT) {
      const int kk = idx % TL::ROWS, h = idx / TL::ROWS;
      if (k0 + kk < a.W) a.out[(img * L + h) * a.W + k0 + kk] = cf32{0.f, 0.f};
    }
    return;
  }
  fill_twiddles<L>(tw, tid, TL::NT);
  __syncthreads();
  TwRegs<L> T;
  T.init(t, tw);
  cf32 u[TL::E];
#pragma unroll
  for (int q = 0; q < TL::E; ++q)
    u[q] = active ? a.ws[(img * a.W + k) * L + t + q * TL::TPF] : cf32{0.f, 0.f};
  __syncthreads();
  fft_regs<L, -1>(u, tile + cc * TL::PITCH, T, t, cc);
  __syncthreads();
#pragma unroll
  for (int q = 0; q < TL::E; ++q) tile[cc * TL::PITCH + spad(t + q * TL::TPF)] = u[q];
  __syncthreads();
  for (int idx = tid; idx < TL::ROWS * L; idx += TL::NT) {
    const int kk = idx % TL::ROWS, h = idx / TL::ROWS;
    if (k0 + kk < a.W)
      a.out[(img * L + h) * a.W + k0 + kk] = cscale(tile[kk * TL::PITCH + spad(h)], a.scale * sgn(h + k0 + kk));
  }
}

// ---- adjoint, pass 1: columns (inverse along H).  grid (ceil(W/ROWS), ncoils*batch) -------------
template <int L>
__global__ void __launch_bounds__(Tile<L>::NT) k_adj_cols(SenseArgs a) {
  using TL = Tile<L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tile = reinterpret_cast<cf32*>(smem_raw);
  cf32* tw = tile + TL::ROWS * TL::PITCH;
  const int tid = threadIdx.x, cc = tid / TL::TPF, t = tid % TL::TPF;
  const size_t img = blockIdx.y;
  const int b = (int)(img % a.batch), k0 = blockIdx.x * TL::ROWS, k = k0 + cc;
  const bool active = k < a.W && col_on(a, b, k);
  const int any = __syncthreads_or(active ? 1 : 0);
  if (!any) return;
  fill_twiddles<L>(tw, tid, TL::NT);
  __syncthreads();
  TwRegs<L> T;
  T.init(t, tw);
  for (int idx = tid; idx < TL::ROWS * L; idx += TL::NT) {
    const int kk = idx % TL::ROWS, h = idx / TL::ROWS;
    cf32 v{0.f, 0.f};
    if (k0 + kk < a.W && col_on(a, b, k0 + kk)) v = cscale(a.in[(img * L + h) * a.W + k0 + kk], sgn(h + k0 + kk));
    tile[kk * TL::PITCH + spad(h)] = v;
  }
  __syncthreads();
  cf32 u[TL::E];
#pragma unroll
  for (int q = 0; q < TL::E; ++q) u[q] = tile[cc * TL::PITCH + spad(t + q * TL::TPF)];
  fft_regs<L, +1>(u, tile + cc * TL::PITCH, T, t, cc);
  if (active) {
#pragma unroll
    for (int q = 0; q < TL::E; ++q) a.ws[(img * a.W + k) * L + t + q * TL::TPF] = u[q];
  }
}

// ---- adjoint, pass 2: rows (inverse along W) + conj-coil sum.  grid (ceil(H/ROWS), batch) ------
template <int L>
__global__ void __launch_bounds__(Tile<L>::NT) k_adj_rows(SenseArgs a) {
  using TL = Tile<L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tile = reinterpret_cast<cf32*>(smem_raw);
  cf32* tw = tile + TL::ROWS * TL::PITCH;
  const int tid = threadIdx.x, r = tid / TL::TPF, t = tid % TL::TPF;
  const int b = blockIdx.y, h0 = blockIdx.x * TL::ROWS, h = h0 + r;
  const bool valid = h < a.H;
  fill_twiddles<L>(tw, tid, TL::NT);
  __syncthreads();
  TwRegs<L> T;
  T.init(t, tw);
  cf32 acc[TL::E];
#pragma unroll
  for (int q = 0; q < TL::E; ++q) acc[q] = cf32{0.f, 0.f};
  for (int c = 0; c < a.ncoils; ++c) {
    const size_t img = (size_t)c * a.batch + b;
    __syncthreads();
    for (int idx = tid; idx < TL::ROWS * L; idx += TL::NT) {
      const int rr = idx % TL::ROWS, k = idx / TL::ROWS;
      cf32 v{0.f, 0.f};
      if (h0 + rr < a.H && col_on(a, b, k)) v = a.ws[(img * L + k) * a.H + h0 + rr];
      tile[rr * TL::PITCH + spad(k)] = v;
    }
    __syncthreads();
    cf32 u[TL::E];
#pragma unroll
    for (int q = 0; q < TL::E; ++q) u[q] = tile[r * TL::PITCH + spad(t + q * TL::TPF)];
    fft_regs<L, +1>(u, tile + r * TL::PITCH, T, t, r);
#pragma unroll
    for (int q = 0; q < TL::E; ++q) {
      const int w = t + q * TL::TPF;
      const cf32 v = cscale(u[q], a.scale * sgn(h + w));
      if (a.ssos) {
        acc[q].x += v.x * v.x + v.y * v.y;
      } else if (a.mre != nullptr && valid) {
        const size_t mi = ((size_t)c * a.H + h) * L + w;
        acc[q] = cadd(acc[q], cmulc(v, cf32{a.mre[mi], a.mim ? a.mim[mi] : 0.f}));
      } else {
        acc[q] = cadd(acc[q], v);
      }
    }
  }
  if (!valid) return;
#pragma unroll
  for (int q = 0; q < TL::E; ++q) {
    const size_t o = ((size_t)b * a.H + h) * L + t + q * TL::TPF;
    if (a.ssos) reinterpret_cast<float*>(a.out)[o] = sqrtf(acc[q].x);
    else a.out[o] = acc[q];
  }
}

// ---- fused Langevin update + SENSE L2-penalty step (row-only).  grid (ceil(H/ROWS), batch) -----
struct AldArgs {
  float* x;
  const float* grad;
  const float* noise;
  const float* bvec;
  const float* mre;
  const float* mim;
  const uint8_t* mask;
  int mask_frames, ncoils, batch, H, W;
  ipdm_ald_scalars sc;
  const ipdm_ald_scalars* sched;
  const int* cursor;
  RngArgs rng;
};

template <int L>
__global__ void __launch_bounds__(Tile<L>::NT) k_ald_sense(AldArgs a) {
  using TL = Tile<L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tile = reinterpret_cast<cf32*>(smem_raw);
  cf32* tw = tile + TL::ROWS * TL::PITCH;
  const int tid = threadIdx.x, r = tid / TL::TPF, t = tid % TL::TPF;
  const int b = blockIdx.y, h = blockIdx.x * TL::ROWS + r;
  const bool valid = h < a.H;
  ipdm_ald_scalars sc = a.sc;
  uint32_t rstep = a.rng.step;
  if (a.sched != nullptr) {
    const int cur = *a.cursor;
    sc = a.sched[cur];
    rstep += (uint32_t)cur;
  }
  fill_twiddles<L>(tw, tid, TL::NT);
  __syncthreads();
  TwRegs<L> T;
  T.init(t, tw);
  const size_t plane = (size_t)a.batch * a.H * L;
  const size_t rowoff = ((size_t)b * a.H + (valid ? h : 0)) * L;
  const uint8_t* mrow = a.mask ? a.mask + (size_t)(b % a.mask_frames) * L : nullptr;
  cf32 z[TL::E], acc[TL::E];
#pragma unroll
  for (int q = 0; q < TL::E; ++q) {
    const size_t o = rowoff + t + q * TL::TPF;
    z[q].x = a.x[o] + sc.step * a.grad[o];
    z[q].y = a.x[plane + o] + sc.step * a.grad[plane + o];
    if (a.noise != nullptr) {
      z[q].x += sc.noise_scale * a.noise[o];
      z[q].y += sc.noise_scale * a.noise[plane + o];
    }
    acc[q] = cf32{0.f, 0.f};
  }
  if (a.noise == nullptr && sc.noise_scale != 0.f) {
    const uint64_t seed = rng_seed(a.rng);
    const uint32_t chain = rng_chain(a.rng, b);
#pragma unroll
    for (int q = 0; q < TL::E / 2; ++q) {   // pixels w and w + W/2 share one Philox call, as in every kernel family
      float n[4];
      philox_chain_normal4(seed, chain, (uint32_t)((valid ? h : 0) * L + t + q * TL::TPF), rstep, n);
      z[q].x += sc.noise_scale * n[0];
      z[q].y += sc.noise_scale * n[1];
      z[q + TL::E / 2].x += sc.noise_scale * n[2];
      z[q + TL::E / 2].y += sc.noise_scale * n[3];
    }
  }
  __syncthreads();
  for (int c = 0; c < a.ncoils; ++c) {
    cf32 u[TL::E], m[TL::E];
#pragma unroll
    for (int q = 0; q < TL::E; ++q) {
      const int w = t + q * TL::TPF;
      const size_t mi = ((size_t)c * a.H + (valid ? h : 0)) * L + w;
      m[q] = cf32{a.mre[mi], a.mim ? a.mim[mi] : 0.f};
      u[q] = cscale(cmul(z[q], m[q]), sgn(w));
    }
    fft_regs<L, -1>(u, tile + r * TL::PITCH, T, t, r);
    if (mrow != nullptr) {
#pragma unroll
      for (int q = 0; q < TL::E; ++q)
        if (mrow[t + q * TL::TPF] == 0) u[q] = cf32{0.f, 0.f};
    }
    fft_regs<L, +1>(u, tile + r * TL::PITCH, T, t, r);
#pragma unroll
    for (int q = 0; q < TL::E; ++q) acc[q] = cadd(acc[q], cscale(cmulc(u[q], m[q]), sgn(t + q * TL::TPF)));
  }
  if (!valid) return;
  const float invW = 1.0f / (float)L;
#pragma unroll
  for (int q = 0; q < TL::E; ++q) {
    const size_t o = rowoff + t + q * TL::TPF;
    a.x[o] = z[q].x - sc.kappa * (acc[q].x * invW - a.bvec[o]);
    a.x[plane + o] = z[q].y - sc.kappa * (acc[q].y * invW - a.bvec[plane + o]);
  }
}

}  // namespace ipdm
#include "sense_fast.cuh"
#include "sense_pruned.cuh"
namespace ipdm {

// ---- launch helpers ---------------------------------------------------------------------------
template <typename K>
static int set_smem(K kernel, size_t bytes) {
  // (no carve-out hint: asking for the maximum shared-memory carve-out for the column kernels shrank the L1 under the
  // adjoint's 8-byte cp.async.ca copies: 1.10 instead of 0.97 ms at 32 coils x 512^2 x 64)
  if (bytes > 48 * 1024) IPDM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

// Lengths served by the two-pass engine (sense_fast.cuh); shorter transforms use the generic Stockham kernels.
static bool fast_len(int n) { return n == 64 || n == 128 || n == 256 || n == 512; }
static bool use_fast(int H, int W) {
  static const bool legacy = getenv("IPDM_SENSE_LEGACY") != nullptr;   // A/B switch for profiling only
  return !legacy && fast_len(H) && fast_len(W);
}

// per-thread twiddles in registers (true) or read from the per-CTA table (false), per kernel family
#ifndef TWREG_FWD
#define TWREG_FWD false
#endif
#ifndef TWREG_ADJ
#define TWREG_ADJ false
#endif
#ifndef TWREG_ALD
#define TWREG_ALD false
#endif

#define IPDM_FOR_FAST_LEN(LEN, MACRO) \
  switch (LEN) {                      \
    case 64: MACRO(64); break;        \
    case 128: MACRO(128); break;      \
    case 256: MACRO(256); break;      \
    default: MACRO(512); break;       \
  }

static int launch_rows_fast(bool fwd, const SenseArgs& a, cudaStream_t s) {
  const bool cplx = a.mim != nullptr, dense = a.mask == nullptr;
#define ROWS2_CASE(LL)                                                                              \
  {                                                                                                 \
    using G = Geo<LL>;                                                                              \
    dim3 grid(a.batch, a.H / G::TPC);                                                               \
    if (fwd) {                                                                                      \
      if (cplx && dense) k2_fwd_rows<LL, true, TWREG_FWD, true><<<grid, G::NT, G::SMEM_ROWS, s>>>(a);        \
      else if (cplx) k2_fwd_rows<LL, true, TWREG_FWD, false><<<grid, G::NT, G::SMEM_ROWS, s>>>(a);         \
      else if (dense) k2_fwd_rows<LL, false, TWREG_FWD, true><<<grid, G::NT, G::SMEM_ROWS, s>>>(a);        \
      else k2_fwd_rows<LL, false, TWREG_FWD, false><<<grid, G::NT, G::SMEM_ROWS, s>>>(a);                  \
    } else {                                                                                        \
      if (cplx && dense) k2_adj_rows<LL, true, TWREG_ADJ, true><<<grid, G::NT, G::SMEM_ROWS, s>>>(a);        \
      else if (cplx) k2_adj_rows<LL, true, TWREG_ADJ, false><<<grid, G::NT, G::SMEM_ROWS, s>>>(a);         \
      else if (dense) k2_adj_rows<LL, false, TWREG_ADJ, true><<<grid, G::NT, G::SMEM_ROWS, s>>>(a);        \
      else k2_adj_rows<LL, false, TWREG_ADJ, false><<<grid, G::NT, G::SMEM_ROWS, s>>>(a);                  \
    }                                                                                               \
  }
  IPDM_FOR_FAST_LEN(a.W, ROWS2_CASE)
#undef ROWS2_CASE
  return launched(fwd ? "k2_fwd_rows" : "k2_adj_rows");
}

// Column CTAs walk the active-sector list in chunks of 4; with many images a few CTAs per image already fill the
// GPU and the rest would only compile the mask and exit.
static int cols_grid_x(const SenseArgs& a) {
  const int chunks = a.W / 16, images = a.ncoils * a.batch;
  int want = (148 * 8 + images - 1) / images;
  if (want < 4) want = 4;
  return want < chunks ? want : chunks;
}

static int launch_cols_fast(bool fwd, const SenseArgs& a, cudaStream_t s) {
  const bool dense = a.mask == nullptr;
#define COLS2_CASE(LL)                                                                              \
  {                                                                                                 \
    using G = Geo<LL>;                                                                              \
    dim3 grid(a.ncoils * a.batch, cols_grid_x(a));                                                  \
    if (fwd) {                                                                                      \
      if (int e = set_smem(k2_fwd_cols<LL, true>, G::SMEM_COLS)) return e;                          \
      if (int e = set_smem(k2_fwd_cols<LL, false>, G::SMEM_COLS)) return e;                         \
      if (dense) k2_fwd_cols<LL, true><<<grid, G::NT_COLS, G::SMEM_COLS, s>>>(a);                   \
      else k2_fwd_cols<LL, false><<<grid, G::NT_COLS, G::SMEM_COLS, s>>>(a);                        \
    } else {                                                                                        \
      if (int e = set_smem(k2_adj_cols<LL, true>, G::SMEM_COLS)) return e;                          \
      if (int e = set_smem(k2_adj_cols<LL, false>, G::SMEM_COLS)) return e;                         \
      if (dense) k2_adj_cols<LL, true><<<grid, G::NT_COLS, G::SMEM_COLS, s>>>(a);                   \
      else k2_adj_cols<LL, false><<<grid, G::NT_COLS, G::SMEM_COLS, s>>>(a);                        \
    }                                                                                               \
  }
  IPDM_FOR_FAST_LEN(a.H, COLS2_CASE)
#undef COLS2_CASE
  return launched(fwd ? "k2_fwd_cols" : "k2_adj_cols");
}

#define IPDM_FOR_LEN(LEN, MACRO)                                       \
  switch (LEN) {                                                       \
    case 8: MACRO(8); break;                                           \
    case 16: MACRO(16); break;                                         \
    case 32: MACRO(32); break;                                         \
    case 64: MACRO(64); break;                                         \
    case 128: MACRO(128); break;                                       \
    case 256: MACRO(256); break;                                       \
    case 512: MACRO(512); break;                                       \
    default:                                                           \
      set_error("transform length %d unsupported (power of two in [8,512])", LEN); \
      return IPDM_E_UNSUPPORTED;                                       \
  }

static int launch_rows(bool fwd, const SenseArgs& a, cudaStream_t s) {
#define ROWS_CASE(LL)                                                                              \
  {                                                                                                \
    using TL = Tile<LL>;                                                                           \
    dim3 grid((a.H + TL::ROWS - 1) / TL::ROWS, a.batch);                                           \
    if (fwd) {                                                                                     \
      if (int e = set_smem(k_fwd_rows<LL>, TL::SMEM)) return e;                                    \
      k_fwd_rows<LL><<<grid, TL::NT, TL::SMEM, s>>>(a);                                            \
    } else {                                                                                       \
      if (int e = set_smem(k_adj_rows<LL>, TL::SMEM)) return e;                                    \
      k_adj_rows<LL><<<grid, TL::NT, TL::SMEM, s>>>(a);                                            \
    }                                                                                              \
  }
  IPDM_FOR_LEN(a.W, ROWS_CASE)
#undef ROWS_CASE
  return launched(fwd ? "k_fwd_rows" : "k_adj_rows");
}

static int launch_cols(bool fwd, const SenseArgs& a, cudaStream_t s) {
#define COLS_CASE(LL)                                                                              \
  {                                                                                                \
    using TL = Tile<LL>;                                                                           \
    dim3 grid((a.W + TL::ROWS - 1) / TL::ROWS, a.ncoils * a.batch);                                \
    if (fwd) {                                                                                     \
      if (int e = set_smem(k_fwd_cols<LL>, TL::SMEM)) return e;                                    \
      k_fwd_cols<LL><<<grid, TL::NT, TL::SMEM, s>>>(a);                                            \
    } else {                                                                                       \
      if (int e = set_smem(k_adj_cols<LL>, TL::SMEM)) return e;                                    \
      k_adj_cols<LL><<<grid, TL::NT, TL::SMEM, s>>>(a);                                            \
    }                                                                                              \
  }
  IPDM_FOR_LEN(a.H, COLS_CASE)
#undef COLS_CASE
  return launched(fwd ? "k_fwd_cols" : "k_adj_cols");
}


// ---- mask plans -----------------------------------------------------------------------------------------------
// A forward / adjoint call on a big batch is split into image sub-ranges whose two kernels overlap: the row kernel of
// sub-range k+1 (bound by instruction issue and the L1 / shared-memory pipe) runs on the caller's stream while the column
// kernel of sub-range k (bound by the latency of scattered DRAM accesses) runs on the plan's side stream.  Fork and join
// are events, so the call stays asynchronous, stream-ordered and capturable; the mutex only serialises the few host calls
// that enqueue one fork-join (two threads sharing a plan must not interleave records of the same events).
constexpr int PLAN_SPLIT_MAX = 4;
struct SensePlan {
  uint32_t magic;
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join, ev_part[PLAN_SPLIT_MAX];
  std::mutex* mu;
  int device, frames, H, W, ns_max, ns_pad, ng_max, nout, cmax, nchunks_max;
  bool pruned_rows, pruned_2d;
  unsigned char* buf;       // one device allocation holding every table
  const uint8_t* mask_dev;  // [frames][W]
  PlanView view;
};
static constexpr uint32_t PLAN_MAGIC = 0x53504c4eu;   // "SPLN"

static const SensePlan* as_plan(const void* p) {
  const SensePlan* pl = static_cast<const SensePlan*>(p);
  return (pl != nullptr && pl->magic == PLAN_MAGIC) ? pl : nullptr;
}

template <int LH> static std::vector<float> tws_for() { return build_tws_host(LH, P2<LH>::R0, P2<LH>::R1); }

// (row length, outputs per thread, entries per residue class) -> kernel instance
#define IPDM_PRUNED_SWITCH(W_, NOUT_, CMAX_, MACRO)                                          \
  switch ((W_) * 100 + (NOUT_) * 10 + (CMAX_)) {                                             \
    case 51212: MACRO(512, 1, 2); break;  case 51214: MACRO(512, 1, 4); break;               \
    case 25612: MACRO(256, 1, 2); break;  case 25614: MACRO(256, 1, 4); break;               \
    case 25622: MACRO(256, 2, 2); break;  case 25624: MACRO(256, 2, 4); break;               \
    case 12812: MACRO(128, 1, 2); break;  case 12814: MACRO(128, 1, 4); break;               \
    case 12822: MACRO(128, 2, 2); break;  case 12824: MACRO(128, 2, 4); break;               \
    default:                                                                                 \
      set_error("pruned SENSE: no kernel for W=%d, %d outputs per thread, classes of %d", W_, NOUT_, CMAX_); \
      return IPDM_E_UNSUPPORTED;                                                             \
  }

static int launch_pruned_rows_any(bool fwd, const SenseArgs& a, const SensePlan* pl, cudaStream_t s) {
#define PROWS_CASE(LL, NO, CM)                                                               \
  {                                                                                          \
    using G = PGeo<LL>;                                                                      \
    dim3 grid(a.nb, a.H / G::TPC);                                                           \
    if (fwd) kp_fwd_rows<LL, NO><<<grid, G::NT, 0, s>>>(a, pl->view);                        \
    else kp_adj_rows<LL, NO, CM><<<grid, G::NT, 0, s>>>(a, pl->view);                        \
  }
  IPDM_PRUNED_SWITCH(a.W, pl->nout, pl->cmax, PROWS_CASE)
#undef PROWS_CASE
  return launched(fwd ? "kp_fwd_rows" : "kp_adj_rows");
}

// Column kernels: one work item per CTA and ONE staging tile while all items fit on the device at once (small problems: half
// the shared memory, so the SM keeps its L1 -- measured 52 vs 59 us for the masked adjoint at 4 coils x 256^2 x 64); otherwise
// persistent CTAs with two tiles, as many as fit (occupancy query, cached per kernel / device / size), or fewer so that
// every CTA walks the same number of items.
template <typename K>
static int cols_slots(K kernel, int threads, size_t smem, int* slots) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, size_t>, int> slots_of;
  int dev = 0;
  IPDM_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_tuple((const void*)kernel, dev, smem);
  auto it = slots_of.find(key);
  if (it == slots_of.end()) {
    int per_sm = 0, sms = 0;
    IPDM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    IPDM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    it = slots_of.emplace(key, std::max(1, per_sm * sms)).first;
  }
  *slots = it->second;
  return 0;
}
template <typename K>
static int cols_grid(K kernel, int threads, size_t smem1, size_t smem2, int n_items, int* grid, size_t* smem, int which) {
  if (int e = set_smem(kernel, smem2)) return e;
  int slots = 0;
  if (int e = cols_slots(kernel, threads, smem1, &slots)) return e;
  static const int force_one = getenv("IPDM_COLS_ONE") ? atoi(getenv("IPDM_COLS_ONE")) : 0;   // A/B: 1 = fwd, 2 = adj, 3 = both
  if (n_items <= slots || (force_one & which)) {
    *grid = n_items;
    *smem = smem1;
    return 0;
  }
  if (int e = cols_slots(kernel, threads, smem2, &slots)) return e;
  const int rounds = (n_items + slots - 1) / slots;
  *grid = (n_items + rounds - 1) / rounds;
  *smem = smem2;
  return 0;
}

static int launch_pruned_cols(bool fwd, const SenseArgs& a, const SensePlan* pl, cudaStream_t s) {
  const int n_items = a.ncoils * a.nb * pl->nchunks_max;
  int grid = 0;
  size_t smem = 0;
#define PCOLS_CASE(LL)                                                                                                   \
  {                                                                                                                      \
    if (fwd) {                                                                                                           \
      if (int e = cols_grid(kp_fwd_cols<LL>, CGeo<LL>::NT, CGeo<LL>::SMEM1, CGeo<LL>::SMEM, n_items, &grid, &smem, 1)) return e; \
      kp_fwd_cols<LL><<<grid, CGeo<LL>::NT, smem, s>>>(a, pl->view);                                                     \
    } else {                                                                                                             \
      if (int e = cols_grid(kp_adj_cols<LL>, CGeo<LL>::NT, CGeo<LL>::SMEM1, CGeo<LL>::SMEM, n_items, &grid, &smem, 2)) return e; \
      kp_adj_cols<LL><<<grid, CGeo<LL>::NT, smem, s>>>(a, pl->view);                                                     \
    }                                                                                                                    \
  }
  IPDM_FOR_FAST_LEN(a.H, PCOLS_CASE)
#undef PCOLS_CASE
  return launched(fwd ? "kp_fwd_cols" : "kp_adj_cols");
}

static int launch_pruned_ald(const AldArgs& a, const SensePlan* pl, cudaStream_t s) {
#define PALD_CASE(LL, NO, CM)                                                                \
  {                                                                                          \
    using G = PGeo<LL>;                                                                      \
    dim3 grid(a.batch, a.H / G::TPC);                                                        \
    kp_ald_sense<LL, NO, CM><<<grid, G::NT, 0, s>>>(a, pl->view);                            \
  }
  IPDM_PRUNED_SWITCH(a.W, pl->nout, pl->cmax, PALD_CASE)
#undef PALD_CASE
  return launched("kp_ald_sense");
}

static RngArgs rng_args(const ipdm_rng* r) {
  RngArgs o{};
  if (r != nullptr) {
    o.seed = r->seed;
    o.seed_dev = reinterpret_cast<const unsigned long long*>(r->seed_dev);
    o.step = r->rng_step;
    o.chain_base = r->chain_base;
    o.chain_ids = r->chain_ids;
  }
  return o;
}

static bool pow2_ok(int n) { return n >= 8 && n <= 512 && (n & (n - 1)) == 0; }

// ---- small elementwise kernels ------------------------------------------------------------------
__global__ void k_kspace_combine(cf32* S, const cf32* Y, const uint8_t* mask, int mask_frames, float a, int mode,
                                 int H, int W, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % W);
    const int b = (int)(i / ((size_t)H * W));
    const float m = mask[(size_t)(b % mask_frames) * W + k] ? 1.f : 0.f;
    cf32 s = S[i];
    if (mode == 0) {
      const float f = 1.0f / (1.0f + m * a);
      s = cscale(s, f);
    } else if (mode == 2) {
      s = cscale(s, m);
    } else {
      const cf32 y = Y[i];
      const float keep = (1.f - a) * m + (1.f - m);
      s = cf32{a * y.x + keep * s.x, a * y.y + keep * s.y};
    }
    S[i] = s;
  }
}

__global__ void k_caxpy(cf32* out, const cf32* a, const cf32* b, float s, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const cf32 x = a[i], y = b[i];
    out[i] = cf32{x.x + s * y.x, x.y + s * y.y};
  }
}

__global__ void k_planar_to_c64(const float* p, cf32* c, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    c[i] = cf32{p[i], p[n + i]};
}
__global__ void k_c64_to_planar(const cf32* c, float* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const cf32 v = c[i];
    p[i] = v.x;
    p[n + i] = v.y;
  }
}

struct LangevinArgs {
  float* x;
  const float* grad;
  const float* noise;
  float* x_mean;
  size_t n;
  ipdm_ald_scalars sc;
  const ipdm_ald_scalars* sched;
  const int* cursor;
  const float* step_per_sample;
  size_t per_sample;
  RngArgs rng;
  size_t chain_elems;   // floats per chain (0: the whole buffer is chain rng_chain(0))
};

__global__ void k_langevin(LangevinArgs a) {
  ipdm_ald_scalars sc = a.sc;
  uint32_t rstep = a.rng.step;
  if (a.sched != nullptr) {
    const int cur = *a.cursor;
    sc = a.sched[cur];
    rstep += (uint32_t)cur;
  }
  const bool draw = a.noise == nullptr && (sc.noise_scale != 0.f || a.step_per_sample != nullptr);
  const uint64_t seed = draw ? rng_seed(a.rng) : 0;
  // two elements per thread so one Philox call feeds both; pairs never straddle two chains
  const size_t ce = a.chain_elems ? a.chain_elems : a.n;
  const size_t ppc = (ce + 1) / 2, nchains = (a.n + ce - 1) / ce;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < ppc * nchains; p += (size_t)gridDim.x * blockDim.x) {
    const size_t ci = p / ppc, pl = p % ppc;
    const size_t i0 = ci * ce + 2 * pl, i1 = i0 + 1;
    const bool has1 = 2 * pl + 1 < ce && i1 < a.n;
    if (i0 >= a.n) continue;
    float2 nz = make_float2(0.f, 0.f);
    if (a.noise != nullptr) nz = make_float2(a.noise[i0], has1 ? a.noise[i1] : 0.f);
    else if (draw) nz = philox_normal2(seed, rng_chain(a.rng, ci), pl, rstep);
    float st0 = sc.step, ns0 = sc.noise_scale, st1 = sc.step, ns1 = sc.noise_scale;
    if (a.step_per_sample != nullptr) {
      st0 = a.step_per_sample[i0 / a.per_sample];
      ns0 = sqrtf(2.f * st0);
      if (has1) {
        st1 = a.step_per_sample[i1 / a.per_sample];
        ns1 = sqrtf(2.f * st1);
      }
    }
    const float m0 = a.x[i0] + st0 * a.grad[i0];
    if (a.x_mean) a.x_mean[i0] = m0;
    a.x[i0] = m0 + ns0 * nz.x;
    if (has1) {
      const float m1 = a.x[i1] + st1 * a.grad[i1];
      if (a.x_mean) a.x_mean[i1] = m1;
      a.x[i1] = m1 + ns1 * nz.y;
    }
  }
}

__global__ void k_advance(int* cursor, int64_t* labels, int batch, int n_steps_each) {
  const int cur = *cursor;
  if (labels != nullptr) {
    const int64_t lvl = (cur + 1) / n_steps_each;
    for (int i = threadIdx.x; i < batch; i += blockDim.x) labels[i] = lvl;
  }
  __syncthreads();
  if (threadIdx.x == 0) *cursor = cur + 1;
}

// x[p][b][t][i] += -lamda * (s[t-1] - s[t]),  s[t] = sign(x[t+1] - x[t]) (circular in t)
__global__ void k_temporal_tv(float* x, int T, size_t hw, float lamda, size_t nvol) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvol * hw; i += (size_t)gridDim.x * blockDim.x) {
    const size_t vol = i / hw, pix = i % hw;
    float* base = x + vol * T * hw + pix;
    auto sg = [](float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); };
    const float first = base[0];
    float prev_x = first;
    float s_prev = sg(first - base[(size_t)(T - 1) * hw]);  // s[T-1] = sign(x[0] - x[T-1])
    for (int t = 0; t < T; ++t) {
      const float nxt = (t + 1 < T) ? base[(size_t)(t + 1) * hw] : first;
      const float s_t = sg(nxt - prev_x);
      base[(size_t)t * hw] = prev_x - lamda * (s_prev - s_t);
      s_prev = s_t;
      prev_x = nxt;
    }
  }
}

__global__ void k_chain_stats(const cf32* x, double* acc, int chains, size_t hw) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int c = 0; c < chains; ++c) {
      const cf32 v = x[(size_t)c * hw + i];
      const float mag = sqrtf(v.x * v.x + v.y * v.y);
      const float ang = atan2f(v.y, v.x);
      s0 += mag;
      s1 += (double)mag * mag;
      s2 += ang;
      s3 += (double)ang * ang;
    }
    acc[i] += s0;
    acc[hw + i] += s1;
    acc[2 * hw + i] += s2;
    acc[3 * hw + i] += s3;
  }
}

static int grid_for(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ipdm

using namespace ipdm;

extern "C" size_t ipdm_sense_workspace_bytes(int ncoils, int batch, int H, int W) {
  return (size_t)ncoils * batch * H * W * sizeof(cf32);
}

extern "C" int ipdm_sense_forward(const void* x, const float* maps_re, const float* maps_im, const uint8_t* mask,
                                  int mask_frames, void* out, int ncoils, int batch, int H, int W, void* workspace,
                                  void* stream) {
  IPDM_REQUIRE(x && out && workspace, IPDM_E_BADARG, "sense_forward: null pointer");
  IPDM_REQUIRE(ncoils >= 1 && batch >= 1, IPDM_E_BADARG, "sense_forward: bad ncoils/batch");
  IPDM_REQUIRE(maps_re != nullptr || ncoils == 1, IPDM_E_BADARG, "sense_forward: maps NULL needs ncoils == 1");
  IPDM_REQUIRE(pow2_ok(H) && pow2_ok(W), IPDM_E_UNSUPPORTED, "sense_forward: H=%d W=%d must be powers of two in [8,512]", H, W);
  IPDM_REQUIRE(mask == nullptr || mask_frames >= 1, IPDM_E_BADARG, "sense_forward: mask_frames");
  SenseArgs a{};
  a.in = (const cf32*)x; a.out = (cf32*)out; a.ws = (cf32*)workspace;
  a.mre = maps_re; a.mim = maps_im; a.mask = mask; a.mask_frames = mask ? mask_frames : 1;
  a.ncoils = ncoils; a.batch = batch; a.H = H; a.W = W; a.ssos = 0;
  a.sparse = mask != nullptr ? 1 : 0;   // masked columns are written straight from registers (any density is correct)
  a.scale = (((H / 2 + W / 2) & 1) ? -1.f : 1.f) / sqrtf((float)H * (float)W);
  if (use_fast(H, W)) {
    if (int e = launch_rows_fast(true, a, as_stream(stream))) return e;
    return launch_cols_fast(true, a, as_stream(stream));
  }
  if (int e = launch_rows(true, a, as_stream(stream))) return e;
  return launch_cols(true, a, as_stream(stream));
}

extern "C" int ipdm_sense_adjoint(const void* S, const float* maps_re, const float* maps_im, const uint8_t* mask,
                                  int mask_frames, void* out, int ncoils, int batch, int H, int W, int ssos,
                                  void* workspace, void* stream) {
  IPDM_REQUIRE(S && out && workspace, IPDM_E_BADARG, "sense_adjoint: null pointer");
  IPDM_REQUIRE(ncoils >= 1 && batch >= 1, IPDM_E_BADARG, "sense_adjoint: bad ncoils/batch");
  IPDM_REQUIRE(pow2_ok(H) && pow2_ok(W), IPDM_E_UNSUPPORTED, "sense_adjoint: H=%d W=%d must be powers of two in [8,512]", H, W);
  SenseArgs a{};
  a.in = (const cf32*)S; a.out = (cf32*)out; a.ws = (cf32*)workspace;
  a.mre = ssos ? nullptr : maps_re; a.mim = ssos ? nullptr : maps_im;
  a.mask = mask; a.mask_frames = mask ? mask_frames : 1;
  a.ncoils = ncoils; a.batch = batch; a.H = H; a.W = W; a.ssos = ssos ? 1 : 0;
  a.scale = (((H / 2 + W / 2) & 1) ? -1.f : 1.f) / sqrtf((float)H * (float)W);
  if (use_fast(H, W)) {
    if (int e = launch_cols_fast(false, a, as_stream(stream))) return e;
    return launch_rows_fast(false, a, as_stream(stream));
  }
  if (int e = launch_cols(false, a, as_stream(stream))) return e;
  return launch_rows(false, a, as_stream(stream));
}

extern "C" int ipdm_kspace_combine(void* S, const void* Y, const uint8_t* mask, int mask_frames, float a, int mode,
                                   int batch, int H, int W, void* stream) {
  IPDM_REQUIRE(S && mask && mask_frames >= 1, IPDM_E_BADARG, "kspace_combine: null pointer");
  IPDM_REQUIRE(mode == 0 || mode == 2 || (mode == 1 && Y), IPDM_E_BADARG, "kspace_combine: mode");
  const size_t n = (size_t)batch * H * W;
  k_kspace_combine<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>((cf32*)S, (const cf32*)Y, mask, mask_frames, a, mode, H, W, n);
  return launched("k_kspace_combine");
}

extern "C" int ipdm_caxpy(void* out, const void* a, const void* b, float s, size_t n, void* stream) {
  IPDM_REQUIRE(out && a && b, IPDM_E_BADARG, "caxpy: null pointer");
  k_caxpy<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>((cf32*)out, (const cf32*)a, (const cf32*)b, s, n);
  return launched("k_caxpy");
}

extern "C" int ipdm_planar_to_c64(const float* planar, void* c64, size_t n, void* stream) {
  IPDM_REQUIRE(planar && c64, IPDM_E_BADARG, "planar_to_c64: null pointer");
  k_planar_to_c64<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(planar, (cf32*)c64, n);
  return launched("k_planar_to_c64");
}

extern "C" int ipdm_c64_to_planar(const void* c64, float* planar, size_t n, void* stream) {
  IPDM_REQUIRE(planar && c64, IPDM_E_BADARG, "c64_to_planar: null pointer");
  k_c64_to_planar<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>((const cf32*)c64, planar, n);
  return launched("k_c64_to_planar");
}

extern "C" int ipdm_langevin_update(float* x, const float* grad, const float* noise, float* x_mean, size_t n,
                                    const ipdm_ald_scalars* scalars_host, const ipdm_ald_scalars* sched,
                                    const int* cursor, const float* step_per_sample, size_t per_sample_elems,
                                    const ipdm_rng* rng_host, void* stream) {
  IPDM_REQUIRE(x && grad, IPDM_E_BADARG, "langevin_update: null pointer");
  IPDM_REQUIRE(scalars_host || (sched && cursor) || step_per_sample, IPDM_E_BADARG, "langevin_update: no step size given");
  IPDM_REQUIRE(!step_per_sample || per_sample_elems > 0, IPDM_E_BADARG, "langevin_update: per_sample_elems");
  if (n == 0) return 0;
  LangevinArgs a{};
  a.x = x; a.grad = grad; a.noise = noise; a.x_mean = x_mean; a.n = n;
  if (scalars_host) a.sc = *scalars_host;
  a.sched = sched; a.cursor = cursor; a.step_per_sample = step_per_sample; a.per_sample = per_sample_elems;
  a.rng = rng_args(rng_host);
  a.chain_elems = rng_host ? rng_host->chain_elems : 0;
  IPDM_REQUIRE(a.chain_elems == 0 || n % a.chain_elems == 0, IPDM_E_BADARG, "langevin_update: n is not a multiple of chain_elems");
  k_langevin<<<grid_for((n + 1) / 2, 256), 256, 0, as_stream(stream)>>>(a);
  return launched("k_langevin");
}

static int ald_sense_general(const AldArgs& a, cudaStream_t s) {
  const int H = a.H, W = a.W, batch = a.batch;
  if (use_fast(W, W) && H % 16 == 0) {
    const bool cplx = a.mim != nullptr;
#define ALD2_CASE(LL)                                                                   \
  {                                                                                     \
    using G = Geo<LL>;                                                                  \
    dim3 grid(batch, H / G::TPC);                                                       \
    if (cplx) k2_ald_sense<LL, true, TWREG_ALD><<<grid, G::NT, G::SMEM_ROWS, s>>>(a);   \
    else k2_ald_sense<LL, false, TWREG_ALD><<<grid, G::NT, G::SMEM_ROWS, s>>>(a);       \
  }
    IPDM_FOR_FAST_LEN(W, ALD2_CASE)
#undef ALD2_CASE
    return launched("k2_ald_sense");
  }
#define ALD_CASE(LL)                                                         \
  {                                                                          \
    using TL = Tile<LL>;                                                     \
    dim3 grid((H + TL::ROWS - 1) / TL::ROWS, batch);                         \
    if (int e = set_smem(k_ald_sense<LL>, TL::SMEM)) return e;               \
    k_ald_sense<LL><<<grid, TL::NT, TL::SMEM, s>>>(a);                       \
  }
  IPDM_FOR_LEN(W, ALD_CASE)
#undef ALD_CASE
  return launched("k_ald_sense");
}

extern "C" int ipdm_ald_sense_step(float* x, const float* grad, const float* noise, const float* b,
                                   const float* maps_re, const float* maps_im, const uint8_t* mask, int mask_frames,
                                   int ncoils, int batch, int H, int W, const ipdm_ald_scalars* scalars_host,
                                   const ipdm_ald_scalars* sched, const int* cursor, const ipdm_rng* rng_host,
                                   void* stream) {
  IPDM_REQUIRE(x && grad && b && maps_re, IPDM_E_BADARG, "ald_sense_step: null pointer");
  IPDM_REQUIRE(scalars_host || (sched && cursor), IPDM_E_BADARG, "ald_sense_step: no scalars given");
  IPDM_REQUIRE(pow2_ok(W), IPDM_E_UNSUPPORTED, "ald_sense_step: W=%d must be a power of two in [8,512]", W);
  IPDM_REQUIRE(ncoils >= 1 && batch >= 1 && H >= 1, IPDM_E_BADARG, "ald_sense_step: bad shape");
  AldArgs a{};
  a.x = x; a.grad = grad; a.noise = noise; a.bvec = b; a.mre = maps_re; a.mim = maps_im;
  a.mask = mask; a.mask_frames = mask ? mask_frames : 1;
  a.ncoils = ncoils; a.batch = batch; a.H = H; a.W = W;
  if (scalars_host) a.sc = *scalars_host;
  a.sched = sched; a.cursor = cursor; a.rng = rng_args(rng_host);
  return ald_sense_general(a, as_stream(stream));
}

extern "C" int ipdm_ald_sense_step_plan(const void* plan, float* x, const float* grad, const float* noise, const float* b,
                                        const float* maps_re, const float* maps_im, int ncoils, int batch, int H,
                                        const ipdm_ald_scalars* scalars_host, const ipdm_ald_scalars* sched,
                                        const int* cursor, const ipdm_rng* rng_host, void* stream) {
  const SensePlan* pl = as_plan(plan);
  IPDM_REQUIRE(pl, IPDM_E_BADARG, "ald_sense_step_plan: not a plan");
  IPDM_REQUIRE(x && grad && b && maps_re, IPDM_E_BADARG, "ald_sense_step_plan: null pointer");
  IPDM_REQUIRE(scalars_host || (sched && cursor), IPDM_E_BADARG, "ald_sense_step_plan: no scalars given");
  IPDM_REQUIRE(ncoils >= 1 && batch >= 1 && H >= 1, IPDM_E_BADARG, "ald_sense_step_plan: bad shape");
  AldArgs a{};
  a.x = x; a.grad = grad; a.noise = noise; a.bvec = b; a.mre = maps_re; a.mim = maps_im;
  a.mask = pl->mask_dev; a.mask_frames = pl->frames;
  a.ncoils = ncoils; a.batch = batch; a.H = H; a.W = pl->W;
  if (scalars_host) a.sc = *scalars_host;
  a.sched = sched; a.cursor = cursor; a.rng = rng_args(rng_host);
  cudaStream_t s = as_stream(stream);
  static const bool no_pruned = getenv("IPDM_SENSE_NO_PRUNED") != nullptr;   // A/B switch for profiling only
  if (pl->pruned_rows && !no_pruned && H % 16 == 0 && maps_im == nullptr) return launch_pruned_ald(a, pl, s);
  return ald_sense_general(a, s);
}

extern "C" int ipdm_ald_advance(int* cursor, int64_t* labels, int batch, int n_steps_each, void* stream) {
  IPDM_REQUIRE(cursor && n_steps_each >= 1, IPDM_E_BADARG, "ald_advance: bad argument");
  k_advance<<<1, 128, 0, as_stream(stream)>>>(cursor, labels, batch, n_steps_each);
  return launched("k_advance");
}

extern "C" int ipdm_temporal_tv_step(float* x, int B, int T, size_t hw, float lamda, void* stream) {
  IPDM_REQUIRE(x && B >= 1 && T >= 2, IPDM_E_BADARG, "temporal_tv_step: bad argument");
  const size_t nvol = (size_t)2 * B;
  k_temporal_tv<<<grid_for(nvol * hw, 256), 256, 0, as_stream(stream)>>>(x, T, hw, lamda, nvol);
  return launched("k_temporal_tv");
}

extern "C" int ipdm_chain_stats_accumulate(const void* x, double* acc, int chains, size_t hw, void* stream) {
  IPDM_REQUIRE(x && acc && chains >= 1, IPDM_E_BADARG, "chain_stats_accumulate: bad argument");
  k_chain_stats<<<grid_for(hw, 256), 256, 0, as_stream(stream)>>>((const cf32*)x, acc, chains, hw);
  return launched("k_chain_stats");
}

// ---- plans ------------------------------------------------------------------------------------------------------
// Releases whatever a (possibly half-built) plan owns; every member is zero until it has been created.
static void plan_release(SensePlan* pl) {
  if (pl == nullptr) return;
  pl->magic = 0;
  if (pl->buf) cudaFree(pl->buf);
  if (pl->side) cudaStreamDestroy(pl->side);
  if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
  if (pl->ev_join) cudaEventDestroy(pl->ev_join);
  for (int i = 0; i < PLAN_SPLIT_MAX; ++i)
    if (pl->ev_part[i]) cudaEventDestroy(pl->ev_part[i]);
  delete pl->mu;
  delete pl;
}

extern "C" int ipdm_sense_plan_create(const uint8_t* mask_host, int mask_frames, int H, int W, void** plan_out) {
  IPDM_REQUIRE(mask_host && plan_out && mask_frames >= 1, IPDM_E_BADARG, "sense_plan_create: bad argument");
  IPDM_REQUIRE(pow2_ok(W) && H >= 1, IPDM_E_UNSUPPORTED, "sense_plan_create: W=%d must be a power of two in [8,512]", W);
  PlanHost ph = build_plan_host(mask_host, mask_frames, W);
  SensePlan* pl = new SensePlan();
  memset(pl, 0, sizeof(*pl));
  pl->magic = PLAN_MAGIC;
  pl->frames = mask_frames; pl->H = H; pl->W = W;
  pl->ns_max = ph.ns_max; pl->ns_pad = ph.ns_pad; pl->ng_max = ph.ng_max;
  pl->pruned_rows = ph.pruned;
  pl->pruned_2d = ph.pruned && fast_len(H);
  pl->nout = ph.pruned ? (ph.ns_max + ph.R1 - 1) / ph.R1 : 0;
  pl->cmax = ph.cmax;
  pl->nchunks_max = ph.nchunks_max;
  if (pl->nout > 2) { pl->pruned_rows = pl->pruned_2d = false; }
  cudaError_t ce = cudaGetDevice(&pl->device);
  if (ce != cudaSuccess) { plan_release(pl); set_error("sense_plan_create: %s", cudaGetErrorString(ce)); return (int)ce; }
  std::vector<float> tws;
  if (pl->pruned_2d) {
    switch (H) {
      case 64: tws = tws_for<64>(); break;
      case 128: tws = tws_for<128>(); break;
      case 256: tws = tws_for<256>(); break;
      default: tws = tws_for<512>(); break;
    }
  }
  // one allocation, every table 256-byte aligned
  struct Piece { const void* src; size_t bytes; size_t off; };
  std::vector<Piece> pieces;
  size_t total = 0;
  auto add = [&](const void* src, size_t bytes) {
    pieces.push_back(Piece{src, bytes, total});
    total += (bytes + 255) & ~(size_t)255;
    return pieces.size() - 1;
  };
  const size_t i_mask = add(ph.mask.data(), ph.mask.size());
  size_t i_ns = 0, i_ng = 0, i_nc = 0, i_kcol = 0, i_nat = 0, i_k0c = 0, i_ppos = 0, i_tcw = 0, i_tw = 0, i_twh = 0, i_groups = 0, i_gslot = 0,
         i_chunks = 0, i_gbm = 0, i_big = 0, i_tws = 0, i_crec = 0;
  if (pl->pruned_rows) {
    i_ns = add(ph.ns.data(), ph.ns.size() * sizeof(int));
    i_ng = add(ph.ngroups.data(), ph.ngroups.size() * sizeof(int));
    i_nc = add(ph.nchunks.data(), ph.nchunks.size() * sizeof(int));
    i_kcol = add(ph.kcol.data(), ph.kcol.size() * sizeof(uint16_t));
    i_nat = add(ph.nat.data(), ph.nat.size());
    i_k0c = add(ph.k0c.data(), ph.k0c.size());
    i_ppos = add(ph.ppos.data(), ph.ppos.size());
    i_tcw = add(ph.tcw.data(), ph.tcw.size());
    i_tw = add(ph.tw.data(), ph.tw.size() * sizeof(float));
    i_twh = add(ph.twh.data(), ph.twh.size() * sizeof(float));
    i_chunks = add(ph.chunks.data(), ph.chunks.size());
    i_groups = add(ph.groups.data(), ph.groups.size());
    i_gslot = add(ph.gslot.data(), ph.gslot.size());
    i_gbm = add(ph.gbitmap.data(), ph.gbitmap.size() * sizeof(uint32_t));
    i_big = add(ph.big.data(), ph.big.size() * sizeof(uint32_t));
    i_crec = add(ph.crec.data(), ph.crec.size() * sizeof(ChunkRec));
    if (!tws.empty()) i_tws = add(tws.data(), tws.size() * sizeof(float));
  }
  pl->mu = new std::mutex();
  {
    // highest priority: when both are runnable the block scheduler hands free SM slots to the side stream's short,
    // latency-bound kernel first, so it runs INSIDE the next row kernel instead of queueing behind its thousands of CTAs
    int lo = 0, hi = 0;
    ce = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithPriority(&pl->side, cudaStreamNonBlocking, hi);
  }
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming);
  for (int i = 0; i < PLAN_SPLIT_MAX && ce == cudaSuccess; ++i) ce = cudaEventCreateWithFlags(&pl->ev_part[i], cudaEventDisableTiming);
  if (ce != cudaSuccess) { plan_release(pl); set_error("sense_plan_create: stream / events: %s", cudaGetErrorString(ce)); return (int)ce; }
  ce = cudaMalloc(reinterpret_cast<void**>(&pl->buf), total);
  if (ce != cudaSuccess) { pl->buf = nullptr; plan_release(pl); set_error("sense_plan_create: cudaMalloc: %s", cudaGetErrorString(ce)); return (int)ce; }
  for (const Piece& pc : pieces) {
    if (pc.bytes == 0) continue;
    ce = cudaMemcpy(pl->buf + pc.off, pc.src, pc.bytes, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
      plan_release(pl);
      set_error("sense_plan_create: cudaMemcpy: %s", cudaGetErrorString(ce));
      return (int)ce;
    }
  }
  auto at = [&](size_t i) { return pl->buf + pieces[i].off; };
  pl->mask_dev = at(i_mask);
  if (pl->pruned_rows) {
    PlanView& v = pl->view;
    v.frames = mask_frames; v.W = W; v.ns_pad = ph.ns_pad; v.ng_all = W / PlanHost::GW; v.cmax = ph.cmax;
    v.ns = reinterpret_cast<const int*>(at(i_ns));
    v.ngroups = reinterpret_cast<const int*>(at(i_ng));
    v.nchunks = reinterpret_cast<const int*>(at(i_nc));
    v.kcol = reinterpret_cast<const uint16_t*>(at(i_kcol));
    v.nat = at(i_nat);
    v.k0c = at(i_k0c);
    v.ppos = at(i_ppos);
    v.tcw = at(i_tcw);
    v.nch_max = ph.nchunks_max;
    v.tw = reinterpret_cast<const cf32*>(at(i_tw));
    v.twh = reinterpret_cast<const cf32*>(at(i_twh));
    v.chunks = at(i_chunks);
    v.groups = at(i_groups);
    v.gslot = at(i_gslot);
    v.gbitmap = reinterpret_cast<const uint32_t*>(at(i_gbm));
    v.big = reinterpret_cast<const uint32_t*>(at(i_big));
    v.tws_h = tws.empty() ? nullptr : reinterpret_cast<const cf32*>(at(i_tws));
    v.crec = reinterpret_cast<const ChunkRec*>(at(i_crec));
  }
  *plan_out = pl;
  return 0;
}

extern "C" int ipdm_sense_plan_destroy(void* plan) {
  SensePlan* pl = const_cast<SensePlan*>(as_plan(plan));
  IPDM_REQUIRE(pl, IPDM_E_BADARG, "sense_plan_destroy: not a plan");
  plan_release(pl);
  return 0;
}

extern "C" int ipdm_sense_plan_info(const void* plan, int* info) {
  const SensePlan* pl = as_plan(plan);
  IPDM_REQUIRE(pl && info, IPDM_E_BADARG, "sense_plan_info: bad argument");
  info[0] = pl->pruned_2d; info[1] = pl->ns_max; info[2] = pl->ns_pad; info[3] = pl->ng_max;
  info[4] = pl->frames; info[5] = pl->H; info[6] = pl->W; info[7] = pl->pruned_rows;
  return 0;
}

// The two kernels of a pruned forward (rows, then columns) or adjoint (columns, then rows), over the whole batch or --
// for a big batch -- over up to four image sub-ranges with the second kernel of each on the plan's side stream.
namespace ipdm { int g_sense_split = 0; }   // ipdm_debug_option key 5

static int pruned_pair(bool fwd, SenseArgs a, const SensePlan* pl, cudaStream_t s) {
  // Off by default: measured at 32 coils x 512^2 x 64 images the split is 2-7 % SLOWER than the two whole-batch launches
  // (forward 1.252 vs 1.225 ms, masked adjoint 1.186 vs 1.107 ms) -- the row and column kernels compete for the same
  // L1 / LSU pipe, so running them side by side buys nothing.  ipdm_debug_option(5, 1) / env IPDM_SENSE_SPLIT turn it on.
  static const bool env_split = getenv("IPDM_SENSE_SPLIT") != nullptr;
  const bool no_split = !(env_split || g_sense_split);
  auto first = [&](cudaStream_t st) { return fwd ? launch_pruned_rows_any(true, a, pl, st) : launch_pruned_cols(false, a, pl, st); };
  auto second = [&](cudaStream_t st) { return fwd ? launch_pruned_cols(true, a, pl, st) : launch_pruned_rows_any(false, a, pl, st); };
  const size_t kspace = (size_t)a.ncoils * a.batch * a.H * a.W * sizeof(cf32);
  int parts = 1;
  if (!no_split && kspace >= ((size_t)512 << 20) && a.batch >= 8) parts = a.batch >= 16 ? 4 : 2;
  a.b0 = 0;
  a.nb = a.batch;
  if (parts == 1) {
    if (int e = first(s)) return e;
    return second(s);
  }
  // The COLUMN kernel of every sub-range goes to the side stream (highest priority): it is the short, latency-bound one, and
  // with priority its CTAs take SM slots as the row kernel's CTAs retire, so the two kinds run together.
  std::lock_guard<std::mutex> lk(*pl->mu);
  IPDM_CUDA(cudaEventRecord(pl->ev_fork, s));
  IPDM_CUDA(cudaStreamWaitEvent(pl->side, pl->ev_fork, 0));
  for (int k = 0; k < parts; ++k) {
    a.b0 = (int)((long long)a.batch * k / parts);
    a.nb = (int)((long long)a.batch * (k + 1) / parts) - a.b0;
    if (fwd) {          // rows on the caller's stream, then columns on the side stream
      if (int e = first(s)) return e;
      IPDM_CUDA(cudaEventRecord(pl->ev_part[k], s));
      IPDM_CUDA(cudaStreamWaitEvent(pl->side, pl->ev_part[k], 0));
      if (int e = second(pl->side)) return e;
    } else {            // columns on the side stream, then rows on the caller's stream
      if (int e = first(pl->side)) return e;
      IPDM_CUDA(cudaEventRecord(pl->ev_part[k], pl->side));
      IPDM_CUDA(cudaStreamWaitEvent(s, pl->ev_part[k], 0));
      if (int e = second(s)) return e;
    }
  }
  if (fwd) {            // join (the adjoint's last wait already joined the side stream)
    IPDM_CUDA(cudaEventRecord(pl->ev_join, pl->side));
    IPDM_CUDA(cudaStreamWaitEvent(s, pl->ev_join, 0));
  }
  return 0;
}

static bool plan_pruned_2d(const SensePlan* pl) {
  static const bool no_pruned = getenv("IPDM_SENSE_NO_PRUNED") != nullptr;   // A/B switch for profiling only
  return pl->pruned_2d && !no_pruned;
}

extern "C" int ipdm_sense_forward_plan(const void* plan, const void* x, const float* maps_re, const float* maps_im, void* out,
                                       int ncoils, int batch, void* workspace, void* stream) {
  const SensePlan* pl = as_plan(plan);
  IPDM_REQUIRE(pl, IPDM_E_BADARG, "sense_forward_plan: not a plan");
  if (!plan_pruned_2d(pl) || maps_im != nullptr)   // the pruned kernels take real coil maps (the reference's, quirk Q4) or none
    return ipdm_sense_forward(x, maps_re, maps_im, pl->mask_dev, pl->frames, out, ncoils, batch, pl->H, pl->W, workspace, stream);
  IPDM_REQUIRE(x && out && workspace, IPDM_E_BADARG, "sense_forward_plan: null pointer");
  IPDM_REQUIRE(ncoils >= 1 && batch >= 1, IPDM_E_BADARG, "sense_forward_plan: bad ncoils/batch");
  IPDM_REQUIRE(maps_re != nullptr || ncoils == 1, IPDM_E_BADARG, "sense_forward_plan: maps NULL needs ncoils == 1");
  const int H = pl->H, W = pl->W;
  SenseArgs a{};
  a.in = (const cf32*)x; a.out = (cf32*)out; a.ws = (cf32*)workspace;
  a.mre = maps_re; a.mim = maps_im; a.mask = pl->mask_dev; a.mask_frames = pl->frames;
  a.ncoils = ncoils; a.batch = batch; a.H = H; a.W = W; a.ssos = 0; a.sparse = 1;
  a.scale = (((H / 2 + W / 2) & 1) ? -1.f : 1.f) / sqrtf((float)H * (float)W);
  return pruned_pair(true, a, pl, as_stream(stream));
}

extern "C" int ipdm_sense_adjoint_plan(const void* plan, const void* S, const float* maps_re, const float* maps_im, void* out,
                                       int ncoils, int batch, int ssos, void* workspace, void* stream) {
  const SensePlan* pl = as_plan(plan);
  IPDM_REQUIRE(pl, IPDM_E_BADARG, "sense_adjoint_plan: not a plan");
  if (!plan_pruned_2d(pl) || (maps_im != nullptr && !ssos))
    return ipdm_sense_adjoint(S, maps_re, maps_im, pl->mask_dev, pl->frames, out, ncoils, batch, pl->H, pl->W, ssos, workspace, stream);
  IPDM_REQUIRE(S && out && workspace, IPDM_E_BADARG, "sense_adjoint_plan: null pointer");
  IPDM_REQUIRE(ncoils >= 1 && batch >= 1, IPDM_E_BADARG, "sense_adjoint_plan: bad ncoils/batch");
  const int H = pl->H, W = pl->W;
  SenseArgs a{};
  a.in = (const cf32*)S; a.out = (cf32*)out; a.ws = (cf32*)workspace;
  a.mre = ssos ? nullptr : maps_re; a.mim = nullptr;
  a.mask = pl->mask_dev; a.mask_frames = pl->frames;
  a.ncoils = ncoils; a.batch = batch; a.H = H; a.W = W; a.ssos = ssos ? 1 : 0;
  a.scale = (((H / 2 + W / 2) & 1) ? -1.f : 1.f) / sqrtf((float)H * (float)W);
  return pruned_pair(false, a, pl, as_stream(stream));
}
