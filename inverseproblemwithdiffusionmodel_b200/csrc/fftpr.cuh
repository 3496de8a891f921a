// Pruned row transforms for column-masked k-space: L = 16 * R1, only the sampled columns are computed.
//
// The sampling mask of this path acts on k-space COLUMNS only (undersampling_fourier.py:77-82), and at the
// accelerations the path is used at (R = 8 ... 40) a row of W samples keeps ns = 8 ... 32 of them.  A row
// transform therefore needs ns outputs of W (forward) or has ns non-zero inputs of W (adjoint).  With
// n = R1*q + t (thread t of R1 holds the 16 values q = 0..15: layout A of fft2p.cuh) and k = k0 + 16*k1,
//
//     X[k] = sum_t w_L^(t*k) * A_t[k mod 16],      A_t[k0] = sum_q x[R1*q + t] * w_16^(q*k0)
//
// so the forward transform is one dense 16-point DFT per thread in registers, one exchange through shared memory
// (line s[k0][t]) and, for each sampled column k, ONE R1-term sum with the twiddle vector w_L^(t*k) -- instead of
// the 16 R1-point DFTs of the full second pass.  The adjoint runs the same graph backwards: thread t forms
// B_t[k0] = sum over the sampled k with k mod 16 == k0 of conj(w_L^(t*k)) * Y[k] (columns sorted by class k0, so
// the register index is static and only the trip count is dynamic), then one inverse 16-point DFT gives
// x[R1*q + t] in the registers the data-consistency update wants them in.
//
// Everything is __host__ __device__ and free of CUDA built-ins: tests/cpu/fft_core_test.cpp runs the threads of a
// transform in a loop and checks both directions against a double-precision DFT.
#pragma once
#include <stdint.h>
#include "fft2p.cuh"

namespace ipdm {

template <int L> struct PR {
  static constexpr int R0 = 16;          // values per thread = in-register radix
  static constexpr int R1 = L / R0;      // threads per transform (8, 16 or 32: a fraction of one warp)
  // exchange line s[k0][t]: rows of R1 values 16-byte aligned (PITCH even) and 8 consecutive rows 16 bytes apart
  // modulo 128, so the 128-bit row reads of a quarter warp hit different banks when their k0 differ modulo 8
  static constexpr int PITCH = R1 + 2;
  static constexpr int LINE = R0 * PITCH;
  static_assert(L == 128 || L == 256 || L == 512, "pruned transforms: L in {128, 256, 512}");
};

struct alignas(16) cf32x2 {
  cf32 a, b;
};

// class boundaries of the sampled columns (17 bytes) packed into 5 words; k0 is a compile-time value after unrolling
IPDM_HD int cls_at(const uint32_t* cw, int i) { return (int)((cw[i >> 2] >> (8 * (i & 3))) & 0xffu); }

// ---- forward: A -> sampled columns ---------------------------------------------------------------------------
// v[q] = x[R1*q + t] on entry; leaves A_t[k0] in the exchange line.
template <int L, int DIR>
IPDM_HD void pr_first(cf32* v, int t, cf32* s) {
  using P = PR<L>;
  dft_n<P::R0, DIR>(v);
#pragma unroll
  for (int k0 = 0; k0 < P::R0; ++k0) s[k0 * P::PITCH + t] = v[k0];
}
// X[k] for one sampled column: k0 = k mod 16, tw[t] = w_L^(t*k) with the forward sign (conjugated for DIR > 0).
template <int L, int DIR>
IPDM_HD cf32 pr_gather(const cf32* s, int k0, const cf32* tw) {
  using P = PR<L>;
  const cf32x2* row = reinterpret_cast<const cf32x2*>(s + k0 * P::PITCH);
  cf32 acc0{0.f, 0.f}, acc1{0.f, 0.f};
#pragma unroll
  for (int i = 0; i < P::R1 / 2; ++i) {
    const cf32x2 p = row[i];
    acc0 = cadd(acc0, twmul<DIR>(p.a, tw[2 * i]));
    acc1 = cadd(acc1, twmul<DIR>(p.b, tw[2 * i + 1]));
  }
  return cadd(acc0, acc1);
}

// ---- adjoint: sampled columns -> A ---------------------------------------------------------------------------
// Y[jj]: the sampled columns in class order (sorted by k mod 16); twc[jj*twp + t] = w_L^(t*k_jj) (forward sign);
// cw: packed class boundaries (class k0 = positions cls_at(cw,k0) .. cls_at(cw,k0+1)-1).
// Leaves x[R1*q + t] in v[q].
template <int L, int DIR>
IPDM_HD void pr_scatter(cf32* v, int t, const cf32* Y, const cf32* twc, int twp, const uint32_t* cw) {
  using P = PR<L>;
#pragma unroll
  for (int k0 = 0; k0 < P::R0; ++k0) {
    cf32 acc{0.f, 0.f};
    const int e = cls_at(cw, k0 + 1);
    for (int jj = cls_at(cw, k0); jj < e; ++jj) acc = cadd(acc, twmul<DIR>(Y[jj], twc[jj * twp + t]));
    v[k0] = acc;
  }
  dft_n<P::R0, DIR>(v);
}

}  // namespace ipdm
