// Two-pass ("four-step") FFT engine for the SENSE kernels: L = R0 * R1, ONE shared-memory exchange per transform.
//
// A length-L transform is done by R1 cooperating threads (a fraction of ONE warp, so the only barrier is
// __syncwarp), each holding R0 complex values in registers:
//
//   layout A  thread t in [0,R1), register q in [0,R0)            <->  position n = R1*q + t
//   layout B  thread u in [0,R1), register i = j*R1 + k1           <->  position k = (u + R1*j) + R0*k1
//             (j in [0,R0/R1), k1 in [0,R1))
//
// In both layouts the R1 threads touch R1 consecutive positions for a fixed register index, so loads and stores
// that go straight to global memory are contiguous runs of 8*R1 bytes.
//
//   a2b:  A -> [DFT_R0 over q] -> exchange -> [* w_L^(t*k0)] -> [DFT_R1 over t] -> B
//   b2a:  B -> [DFT_R1 over k1] -> [* w_L^(t*k0)] -> exchange -> [DFT_R0 over k0] -> A
//
// Either pipeline is a complete DFT with sign DIR (n*k = (R1 q + t)(k0 + R0 k1) splits into w_R0^(q k0) *
// w_L^(t k0) * w_R1^(t k1)).  A forward a2b followed by an inverse b2a returns to the original register layout
// with the spectrum available in between -- the fused data-consistency step masks it there -- and both use the
// SAME per-thread twiddle set, held in registers.
//
// Everything is __host__ __device__ and free of CUDA built-ins; tests/cpu/fft_core_test.cpp runs the threads of
// one transform in a loop with the exchange as the barrier.
#pragma once
#include "fft_core.cuh"

namespace ipdm {

// cos / sin of 2*pi*k/32 (k any integer), usable in constant expressions so unrolled butterflies get immediates.
IPDM_HD constexpr float cos32(int k) {
  constexpr float c[9] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                          0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                          0.19509032201612826785f, 0.0f};
  k &= 31;
  if (k > 16) k = 32 - k;
  return k <= 8 ? c[k] : -c[16 - k];
}
IPDM_HD constexpr float sin32(int k) { return cos32(k - 8); }

// In-register N-point DFT (N = 1,2,4,8,16,32), natural order in and out, sign DIR.
template <int N, int DIR>
IPDM_HD void dft_n(cf32* v) {
  if constexpr (N == 1) {
  } else if constexpr (N == 2) {
    dft2<DIR>(v[0], v[1]);
  } else if constexpr (N == 4) {
    dft4<DIR>(v);
  } else if constexpr (N == 8) {
    dft8<DIR>(v);
  } else {
    cf32 e[N / 2], o[N / 2];
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      e[k] = v[2 * k];
      o[k] = v[2 * k + 1];
    }
    dft_n<N / 2, DIR>(e);
    dft_n<N / 2, DIR>(o);
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      cf32 t;
      if (k == 0) {
        t = o[0];
      } else if (k == N / 4) {
        t = rot90<DIR>(o[k]);
      } else {
        const float c = cos32(k * (32 / N)), s = (DIR < 0 ? -1.f : 1.f) * sin32(k * (32 / N));
        t = cf32{o[k].x * c - o[k].y * s, o[k].x * s + o[k].y * c};
      }
      v[k] = cadd(e[k], t);
      v[k + N / 2] = csub(e[k], t);
    }
  }
}

template <int L> struct Plan2;
template <> struct Plan2<8>   { static constexpr int R0 = 8,  R1 = 1; };
template <> struct Plan2<16>  { static constexpr int R0 = 4,  R1 = 4; };
template <> struct Plan2<32>  { static constexpr int R0 = 8,  R1 = 4; };
template <> struct Plan2<64>  { static constexpr int R0 = 8,  R1 = 8; };
template <> struct Plan2<128> { static constexpr int R0 = 16, R1 = 8; };
template <> struct Plan2<256> { static constexpr int R0 = 16, R1 = 16; };
template <> struct Plan2<512> { static constexpr int R0 = 32, R1 = 16; };

template <int L> struct P2 {
  static constexpr int R0 = Plan2<L>::R0, R1 = Plan2<L>::R1;
  static constexpr int TPF = R1;          // threads per transform
  static constexpr int E = R0;            // complex values per thread
  static constexpr int G = R0 / R1;       // R1-point transforms per thread in layout B (1 when R1 == 1: see below)
  static constexpr int PITCH = R1 + 1;    // exchange line [R0][PITCH]: conflict-free both ways
  static constexpr int BASE = R0 * PITCH;
  // Several transforms share a warp; with 8 threads per transform each STS.64/LDS.64 covers 64 bytes, so
  // neighbouring transforms must sit 64 bytes (8 elements mod 16) apart to use the other 16 banks.
  static constexpr int STRIDE = (R1 == 8 && BASE % 16 != 8) ? BASE + ((8 - BASE % 16) + 16) % 16 : BASE;
  static constexpr int NTW = R1 > 1 ? G * (R1 - 1) : 0;   // register twiddles per thread
};

// position = thread index + a per-register compile-time offset (so global accesses get immediate offsets)
template <int L> IPDM_HD constexpr int a_off(int q) { return P2<L>::R1 * q; }
template <int L> IPDM_HD constexpr int b_off(int i) { return P2<L>::R1 * (i / P2<L>::R1) + P2<L>::R0 * (i % P2<L>::R1); }
template <int L> IPDM_HD constexpr int a_pos(int t, int q) { return t + a_off<L>(q); }
template <int L> IPDM_HD constexpr int b_pos(int u, int i) { return u + b_off<L>(i); }

// twr[j*(R1-1) + t-1] = w_L^(t*(u + R1*j)) from the forward table tw[m] = exp(-2*pi*i*m/L); conjugated on use.
template <int L>
IPDM_HD void p2_twiddles(int u, cf32* twr, const cf32* tw) {
  using P = P2<L>;
  if constexpr (P::R1 > 1) {
#pragma unroll
    for (int j = 0; j < P::G; ++j)
#pragma unroll
      for (int t = 1; t < P::R1; ++t) twr[j * (P::R1 - 1) + t - 1] = tw[(t * (u + P::R1 * j)) & (L - 1)];
  }
}

template <int DIR>
IPDM_HD cf32 twmul(cf32 x, cf32 w) { return DIR < 0 ? cmul(x, w) : cmulc(x, w); }

// ---- A -> B -----------------------------------------------------------------------------------------------
template <int L, int DIR>
IPDM_HD void a2b_first(cf32* v, int t, cf32* s) {
  using P = P2<L>;
  dft_n<P::R0, DIR>(v);
  if constexpr (P::R1 > 1) {
#pragma unroll
    for (int k0 = 0; k0 < P::R0; ++k0) s[k0 * P::PITCH + t] = v[k0];
  }
}
// `tw(n)` returns twiddle n of this thread (p2_twiddles order): a register array or a per-CTA table.
template <int L, int DIR, class TW>
IPDM_HD void a2b_second(cf32* v, int u, const cf32* s, TW tw) {
  using P = P2<L>;
  if constexpr (P::R1 > 1) {
#pragma unroll
    for (int j = 0; j < P::G; ++j) {
#pragma unroll
      for (int t = 0; t < P::R1; ++t) {
        cf32 x = s[(u + P::R1 * j) * P::PITCH + t];
        if (t > 0) x = twmul<DIR>(x, tw(j * (P::R1 - 1) + t - 1));
        v[j * P::R1 + t] = x;
      }
      dft_n<P::R1, DIR>(v + j * P::R1);
    }
  }
}

// ---- B -> A -----------------------------------------------------------------------------------------------
template <int L, int DIR, class TW>
IPDM_HD void b2a_first(cf32* v, int u, cf32* s, TW tw) {
  using P = P2<L>;
  if constexpr (P::R1 > 1) {
#pragma unroll
    for (int j = 0; j < P::G; ++j) {
      dft_n<P::R1, DIR>(v + j * P::R1);
#pragma unroll
      for (int t = 0; t < P::R1; ++t) {
        cf32 x = v[j * P::R1 + t];
        if (t > 0) x = twmul<DIR>(x, tw(j * (P::R1 - 1) + t - 1));
        s[(u + P::R1 * j) * P::PITCH + t] = x;
      }
    }
  }
}
template <int L, int DIR>
IPDM_HD void b2a_second(cf32* v, int t, const cf32* s) {
  using P = P2<L>;
  if constexpr (P::R1 > 1) {
#pragma unroll
    for (int k0 = 0; k0 < P::R0; ++k0) v[k0] = s[k0 * P::PITCH + t];
  }
  dft_n<P::R0, DIR>(v);
}

}  // namespace ipdm
