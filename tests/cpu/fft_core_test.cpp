// CPU emulation of the Stockham passes in csrc/fft_core.cuh: every "thread" of one transform is run
// in a loop with a full barrier between load / store phases, exactly as the kernels do with
// __syncthreads().  Checks all supported lengths, both directions, against a double-precision DFT.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include "../../inverseproblemwithdiffusionmodel_b200/csrc/fft_core.cuh"
#include "../../inverseproblemwithdiffusionmodel_b200/csrc/fft2p.cuh"
using namespace ipdm;

template <int L, int P, int DIR>
struct RunPasses {
  static void go(std::vector<cf32>& buf, const std::vector<cf32>& tw) {
    constexpr int TPF = FftPlan<L>::TPF, E = FftRegs<L>::E;
    std::vector<cf32> regs(TPF * E);
    for (int t = 0; t < TPF; ++t) pass_load<L, P>(t, &regs[t * E], [&](int i) { return buf[i]; });
    // odd "threads" take the table-lookup path, even ones the hoisted register-twiddle path: both must agree
    for (int t = 0; t < TPF; ++t) {
      if (t & 1) {
        pass_compute<L, P, DIR>(t, &regs[t * E], tw.data());
      } else {
        cf32 twr[64];
        pass_twiddles<L, P>(t, twr, tw.data());
        pass_compute_regtw<L, P, DIR>(&regs[t * E], twr);
      }
    }
    for (int t = 0; t < TPF; ++t) pass_store<L, P>(t, &regs[t * E], [&](int i, cf32 v) { buf[i] = v; });
    if constexpr (P + 1 < FftPlan<L>::NP) RunPasses<L, P + 1, DIR>::go(buf, tw);
  }
};

template <int L, int DIR>
double check() {
  std::vector<cf32> x(L), tw(L);
  std::vector<std::complex<double>> xd(L);
  for (int i = 0; i < L; ++i) {
    x[i] = cf32{(float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f};
    xd[i] = {x[i].x, x[i].y};
    tw[i] = cf32{(float)cos(-2.0 * M_PI * i / L), (float)sin(-2.0 * M_PI * i / L)};
  }
  RunPasses<L, 0, DIR>::go(x, tw);
  double err = 0, nrm = 0;
  for (int k = 0; k < L; ++k) {
    std::complex<double> s = 0;
    for (int n = 0; n < L; ++n) s += xd[n] * std::polar(1.0, DIR * 2.0 * M_PI * k * n / L);
    err += std::norm(s - std::complex<double>(x[k].x, x[k].y));
    nrm += std::norm(s);
  }
  return sqrt(err / nrm);
}

// Two-pass engine (csrc/fft2p.cuh): A->B and B->A pipelines, threads run in a loop with the exchange as barrier.
template <int L, int DIR, bool A2B>
double check2p() {
  using P = P2<L>;
  std::vector<cf32> x(L), tw(L), out(L), smem(P::STRIDE + 8);
  std::vector<std::complex<double>> xd(L);
  for (int i = 0; i < L; ++i) {
    x[i] = cf32{(float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f};
    xd[i] = {x[i].x, x[i].y};
    tw[i] = cf32{(float)cos(-2.0 * M_PI * i / L), (float)sin(-2.0 * M_PI * i / L)};
  }
  std::vector<cf32> regs(P::TPF * P::E);
  cf32 twr[64];
  if (A2B) {
    for (int t = 0; t < P::TPF; ++t) {
      for (int q = 0; q < P::E; ++q) regs[t * P::E + q] = x[a_pos<L>(t, q)];
      a2b_first<L, DIR>(&regs[t * P::E], t, smem.data());
    }
    for (int u = 0; u < P::TPF; ++u) {
      p2_twiddles<L>(u, twr, tw.data());
      a2b_second<L, DIR>(&regs[u * P::E], u, smem.data(), [&](int n) { return twr[n]; });
      for (int i = 0; i < P::E; ++i) out[b_pos<L>(u, i)] = regs[u * P::E + i];
    }
  } else {
    for (int u = 0; u < P::TPF; ++u) {
      p2_twiddles<L>(u, twr, tw.data());
      for (int i = 0; i < P::E; ++i) regs[u * P::E + i] = x[b_pos<L>(u, i)];
      b2a_first<L, DIR>(&regs[u * P::E], u, smem.data(), [&](int n) { return twr[n]; });
    }
    for (int t = 0; t < P::TPF; ++t) {
      b2a_second<L, DIR>(&regs[t * P::E], t, smem.data());
      for (int q = 0; q < P::E; ++q) out[a_pos<L>(t, q)] = regs[t * P::E + q];
    }
  }
  double err = 0, nrm = 0;
  for (int k = 0; k < L; ++k) {
    std::complex<double> s = 0;
    for (int n = 0; n < L; ++n) s += xd[n] * std::polar(1.0, DIR * 2.0 * M_PI * k * n / L);
    err += std::norm(s - std::complex<double>(out[k].x, out[k].y));
    nrm += std::norm(s);
  }
  return sqrt(err / nrm);
}

int main() {
  double worst = 0;
#define CHK(L)                                                       \
  {                                                                  \
    double a = check<L, -1>(), b = check<L, +1>();                   \
    printf("L=%d fwd %.3e inv %.3e\n", L, a, b);                     \
    worst = fmax(worst, fmax(a, b));                                 \
  }
  CHK(8) CHK(16) CHK(32) CHK(64) CHK(128) CHK(256) CHK(512) CHK(1024)
#define CHK2(L)                                                                               \
  {                                                                                           \
    double a = check2p<L, -1, true>(), b = check2p<L, +1, true>();                            \
    double c = check2p<L, -1, false>(), d = check2p<L, +1, false>();                          \
    printf("2-pass L=%d a2b fwd %.3e inv %.3e  b2a fwd %.3e inv %.3e\n", L, a, b, c, d);      \
    worst = fmax(worst, fmax(fmax(a, b), fmax(c, d)));                                        \
  }
  CHK2(8) CHK2(16) CHK2(32) CHK2(64) CHK2(128) CHK2(256) CHK2(512)
  printf("worst %.3e\n", worst);
  return worst < 2e-6 ? 0 : 1;
}
