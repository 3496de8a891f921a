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
// B_t[k0] = sum over the sampled k with k mod 16 == k0 of conj(w_L^(t*k)) * Y[k], then one inverse 16-point DFT gives
// x[R1*q + t] in the registers the data-consistency update wants them in.  The residue classes are laid out padded
// to the size of the largest one (zeros in the pads), so every index of that sum is static: a first version walked
// the classes with mask-dependent trip counts and spent its time on branches and instruction fetch.
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
  static constexpr int R1H = R1 / 4;     // the R1-term sums are evaluated as R1H groups of 4 terms
  static constexpr int NTWH = 3 + (R1H - 1);   // factored twiddle vector: w^k, w^2k, w^3k, w^(4m*k) for m = 1..R1H-1
  // exchange line s[k0][t]: rows of R1 values 16-byte aligned (PITCH even) and 8 consecutive rows 16 bytes apart
  // modulo 128, so the 128-bit row reads of a quarter warp hit different banks when their k0 differ modulo 8
  static constexpr int PITCH = R1 + 2;
  static constexpr int LINE = R0 * PITCH;
  static_assert(L == 128 || L == 256 || L == 512, "pruned transforms: L in {128, 256, 512}");
};

struct alignas(16) cf32x2 {
  cf32 a, b;
};

// ---- forward: A -> sampled columns ---------------------------------------------------------------------------
// v[q] = x[R1*q + t] on entry; leaves A_t[k0] in the exchange line.
template <int L, int DIR>
IPDM_HD void pr_first(cf32* v, int t, cf32* s) {
  using P = PR<L>;
  dft_n<P::R0, DIR>(v);
#pragma unroll
  for (int k0 = 0; k0 < P::R0; ++k0) s[k0 * P::PITCH + t] = v[k0];
}
// X[k] for one sampled column: k0 = k mod 16; twh = the factored twiddle vector of k (forward sign, conjugated for
// DIR > 0): sum_t s[t] w^(tk) = sum_m w^(4mk) * (s[4m] + w^k s[4m+1] + w^2k s[4m+2] + w^3k s[4m+3]) -- as many multiply-adds
// as the plain sum, a third of its twiddle registers.
template <int L, int DIR>
IPDM_HD cf32 pr_gather(const cf32* s, int k0, const cf32* twh) {
  using P = PR<L>;
  const cf32x2* row = reinterpret_cast<const cf32x2*>(s + k0 * P::PITCH);
  cf32 acc{0.f, 0.f};
#pragma unroll
  for (int m = 0; m < P::R1H; ++m) {
    const cf32x2 p01 = row[2 * m], p23 = row[2 * m + 1];
    cf32 in = cadd(p01.a, twmul<DIR>(p01.b, twh[0]));
    in = cadd(in, twmul<DIR>(p23.a, twh[1]));
    in = cadd(in, twmul<DIR>(p23.b, twh[2]));
    acc = m == 0 ? in : cadd(acc, twmul<DIR>(in, twh[2 + m]));
  }
  return acc;
}

// ---- adjoint: sampled columns -> A ---------------------------------------------------------------------------
// Y[k0*CMAX + e]: the sampled columns in the padded class layout (class k0 = k mod 16 holds at most CMAX columns, unused
// entries are zero); twp[(k0*CMAX + e)*pitch + t] = w_L^(t*k) of that entry (forward sign; zero in the pads).
// Entries e < 2 are always evaluated (static indices, no mask-dependent control flow); entries 2 and 3 exist in few
// classes (a keep-centre mask of 21 columns has one or two such classes), so they sit behind bit tests that are
// uniform over the whole grid: big[e-2] bit k0 = class k0 has an entry e.  Leaves x[R1*q + t] in v[q].
template <int L, int DIR, int CMAX>
IPDM_HD void pr_scatter(cf32* v, int t, const cf32* Y, const cf32* twp, int pitch, uint32_t big2, uint32_t big3) {
  using P = PR<L>;
  static_assert(CMAX == 2 || CMAX == 4, "padded class size 2 or 4");
#pragma unroll
  for (int k0 = 0; k0 < P::R0; ++k0) {
    const cf32x2 y01 = reinterpret_cast<const cf32x2*>(Y + k0 * CMAX)[0];   // class rows are 16-byte aligned
    cf32 acc = twmul<DIR>(y01.a, twp[(k0 * CMAX) * pitch + t]);
    v[k0] = cadd(acc, twmul<DIR>(y01.b, twp[(k0 * CMAX + 1) * pitch + t]));
  }
  if (CMAX == 4) {
    // third and fourth entries: one test per four classes skips the common case outright
#pragma unroll
    for (int g = 0; g < P::R0 / 4; ++g) {
      if ((big2 >> (4 * g)) & 0xFu) {
#pragma unroll
        for (int k0 = 4 * g; k0 < 4 * g + 4; ++k0) {
          if ((big2 >> k0) & 1u) {
            v[k0] = cadd(v[k0], twmul<DIR>(Y[k0 * CMAX + 2], twp[(k0 * CMAX + 2) * pitch + t]));
            if ((big3 >> k0) & 1u) v[k0] = cadd(v[k0], twmul<DIR>(Y[k0 * CMAX + 3], twp[(k0 * CMAX + 3) * pitch + t]));
          }
        }
      }
    }
  }
  dft_n<P::R0, DIR>(v);
}

}  // namespace ipdm
