// Stockham autosort FFT building blocks shared by the SENSE / centred-FFT kernels.
//
// Everything here is `__host__ __device__` and free of CUDA built-ins so the index arithmetic can be
// unit-tested on the CPU (tests/cpu/fft_core_test.cpp emulates the threads of one transform in a
// loop); the kernels in sense.cu supply the thread index, the loads/stores and the barriers.
//
// One length-L transform (L = 2^k, 8 <= L <= 1024) is done by TPF = L/8 (L/4 for L = 16) cooperating
// threads in 1-4 passes of radix 8/4; each pass, thread t owns butterflies j = t + i*TPF.  A pass
// reads v[r] = src[j + r*L/R], multiplies by w_L^(r*(j%Ns)*L/(Ns*R)), does an R-point DFT in
// registers and writes dst[(j/Ns)*Ns*R + j%Ns + r*Ns]; Ns = product of the radices already done.
// With Ns = 1 first and natural-order output last, first-pass loads and last-pass stores are both
// "thread j touches j + r*L/R", i.e. coalesced when they go straight to global memory.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define IPDM_HD __host__ __device__ __forceinline__
#else
#define IPDM_HD inline
#endif

namespace ipdm {

// 8-byte aligned: complex64 tensors, scratch and shared-memory lines all are, and it lets one 64-bit access move a value
struct alignas(8) cf32 {
  float x, y;
};

IPDM_HD cf32 cmul(cf32 a, cf32 b) { return cf32{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
IPDM_HD cf32 cmulc(cf32 a, cf32 b) { return cf32{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }  // a * conj(b)
IPDM_HD cf32 cadd(cf32 a, cf32 b) { return cf32{a.x + b.x, a.y + b.y}; }
IPDM_HD cf32 csub(cf32 a, cf32 b) { return cf32{a.x - b.x, a.y - b.y}; }
IPDM_HD cf32 cscale(cf32 a, float s) { return cf32{a.x * s, a.y * s}; }
// multiply by -i (forward, DIR = -1) or +i (inverse, DIR = +1)
template <int DIR>
IPDM_HD cf32 rot90(cf32 a) {
  return DIR < 0 ? cf32{a.y, -a.x} : cf32{-a.y, a.x};
}

// ---- in-register DFTs, natural-order output --------------------------------------------------
template <int DIR>
IPDM_HD void dft2(cf32& a, cf32& b) {
  cf32 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <int DIR>
IPDM_HD void dft4(cf32* v) {
  cf32 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
  cf32 b0 = cadd(v[1], v[3]), b1 = rot90<DIR>(csub(v[1], v[3]));
  v[0] = cadd(a0, b0);
  v[2] = csub(a0, b0);
  v[1] = cadd(a1, b1);
  v[3] = csub(a1, b1);
}

template <int DIR>
IPDM_HD void dft8(cf32* v) {
  const float h = 0.70710678118654752440f;
  // split into even / odd 4-point transforms
  cf32 e[4] = {v[0], v[2], v[4], v[6]};
  cf32 o[4] = {v[1], v[3], v[5], v[7]};
  dft4<DIR>(e);
  dft4<DIR>(o);
  // twiddles w8^k, k = 0..3:  1, (1 -/+ i)/sqrt2, -/+ i, (-1 -/+ i)/sqrt2
  cf32 t1 = DIR < 0 ? cf32{(o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h} : cf32{(o[1].x - o[1].y) * h, (o[1].y + o[1].x) * h};
  cf32 t2 = rot90<DIR>(o[2]);
  cf32 t3 = DIR < 0 ? cf32{(-o[3].x + o[3].y) * h, (-o[3].y - o[3].x) * h} : cf32{(-o[3].x - o[3].y) * h, (-o[3].y + o[3].x) * h};
  v[0] = cadd(e[0], o[0]);
  v[4] = csub(e[0], o[0]);
  v[1] = cadd(e[1], t1);
  v[5] = csub(e[1], t1);
  v[2] = cadd(e[2], t2);
  v[6] = csub(e[2], t2);
  v[3] = cadd(e[3], t3);
  v[7] = csub(e[3], t3);
}

template <int R, int DIR>
IPDM_HD void dftR(cf32* v) {
  if (R == 8) dft8<DIR>(v);
  else if (R == 4) dft4<DIR>(v);
  else dft2<DIR>(v[0], v[1]);
}

// ---- plan: radices per length -----------------------------------------------------------------
template <int L> struct FftPlan;
template <> struct FftPlan<8>    { static constexpr int NP = 1, R0 = 8, R1 = 1, R2 = 1, R3 = 1, TPF = 1; };
template <> struct FftPlan<16>   { static constexpr int NP = 2, R0 = 4, R1 = 4, R2 = 1, R3 = 1, TPF = 4; };
template <> struct FftPlan<32>   { static constexpr int NP = 2, R0 = 8, R1 = 4, R2 = 1, R3 = 1, TPF = 4; };
template <> struct FftPlan<64>   { static constexpr int NP = 2, R0 = 8, R1 = 8, R2 = 1, R3 = 1, TPF = 8; };
template <> struct FftPlan<128>  { static constexpr int NP = 3, R0 = 8, R1 = 4, R2 = 4, R3 = 1, TPF = 16; };
template <> struct FftPlan<256>  { static constexpr int NP = 3, R0 = 8, R1 = 8, R2 = 4, R3 = 1, TPF = 32; };
template <> struct FftPlan<512>  { static constexpr int NP = 3, R0 = 8, R1 = 8, R2 = 8, R3 = 1, TPF = 64; };
template <> struct FftPlan<1024> { static constexpr int NP = 4, R0 = 8, R1 = 8, R2 = 4, R3 = 4, TPF = 128; };

template <int L, int P> struct PassRadix {
  static constexpr int value = P == 0 ? FftPlan<L>::R0 : P == 1 ? FftPlan<L>::R1 : P == 2 ? FftPlan<L>::R2 : FftPlan<L>::R3;
};
template <int L, int P> struct PassNs {
  static constexpr int value = P == 0 ? 1 : PassNs<L, P - 1>::value * PassRadix<L, P - 1>::value;
};
template <int L> struct PassNs<L, 0> { static constexpr int value = 1; };

// elements each thread keeps in registers during a pass
template <int L> struct FftRegs { static constexpr int E = L / FftPlan<L>::TPF; };

// Load the operands of pass P for thread t into v[E]:  butterfly i, leg r  ->  v[i*R + r].
template <int L, int P, class LoadFn>
IPDM_HD void pass_load(int t, cf32* v, LoadFn load) {
  constexpr int R = PassRadix<L, P>::value, TPF = FftPlan<L>::TPF, NB = (L / R) / TPF;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int j = t + i * TPF;
#pragma unroll
    for (int r = 0; r < R; ++r) v[i * R + r] = load(j + r * (L / R));
  }
}

// Twiddle + R-point DFT on the registers of pass P.  tw[m] = exp(-2*pi*i*m/L) (forward table);
// the inverse direction conjugates on the fly.
template <int L, int P, int DIR>
IPDM_HD void pass_compute(int t, cf32* v, const cf32* tw) {
  constexpr int R = PassRadix<L, P>::value, TPF = FftPlan<L>::TPF, NB = (L / R) / TPF, NS = PassNs<L, P>::value;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int j = t + i * TPF;
    if (NS > 1) {
      const int m = (j % NS) * (L / (NS * R));
#pragma unroll
      for (int r = 1; r < R; ++r) {
        const cf32 w = tw[r * m];
        v[i * R + r] = DIR < 0 ? cmul(v[i * R + r], w) : cmulc(v[i * R + r], w);
      }
    }
    dftR<R, DIR>(v + i * R);
  }
}

// Per-thread twiddles of pass P: they depend only on (t, P), so kernels hoist them out of their coil / row loops.
// twr[i*(R-1) + r-1] = w_L^(r*m_i), m_i = ((t + i*TPF) % Ns) * L/(Ns*R); forward table, conjugated on use.
template <int L, int P> struct PassTw {
  static constexpr int R = PassRadix<L, P>::value, NB = (L / R) / FftPlan<L>::TPF;
  static constexpr int N = PassNs<L, P>::value > 1 ? NB * (R - 1) : 0;   // pass 0 has unit twiddles
};
template <int L, int P>
IPDM_HD void pass_twiddles(int t, cf32* twr, const cf32* tw) {
  constexpr int R = PassRadix<L, P>::value, TPF = FftPlan<L>::TPF, NB = (L / R) / TPF, NS = PassNs<L, P>::value;
  if (NS == 1) return;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int m = ((t + i * TPF) % NS) * (L / (NS * R));
#pragma unroll
    for (int r = 1; r < R; ++r) twr[i * (R - 1) + r - 1] = tw[r * m];
  }
}
template <int L, int P, int DIR>
IPDM_HD void pass_compute_regtw(cf32* v, const cf32* twr) {
  constexpr int R = PassRadix<L, P>::value, TPF = FftPlan<L>::TPF, NB = (L / R) / TPF, NS = PassNs<L, P>::value;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    if (NS > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) {
        const cf32 w = twr[i * (R - 1) + r - 1];
        v[i * R + r] = DIR < 0 ? cmul(v[i * R + r], w) : cmulc(v[i * R + r], w);
      }
    }
    dftR<R, DIR>(v + i * R);
  }
}

// Store the results of pass P:  dst index (j/Ns)*Ns*R + j%Ns + r*Ns.
template <int L, int P, class StoreFn>
IPDM_HD void pass_store(int t, const cf32* v, StoreFn store) {
  constexpr int R = PassRadix<L, P>::value, TPF = FftPlan<L>::TPF, NB = (L / R) / TPF, NS = PassNs<L, P>::value;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int j = t + i * TPF;
    const int d = (j / NS) * NS * R + (j % NS);
#pragma unroll
    for (int r = 0; r < R; ++r) store(d + r * NS, v[i * R + r]);
  }
}

}  // namespace ipdm
