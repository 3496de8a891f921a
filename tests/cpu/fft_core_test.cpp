// CPU emulation of the Stockham passes in csrc/fft_core.cuh: every "thread" of one transform is run
// in a loop with a full barrier between load / store phases, exactly as the kernels do with
// __syncthreads().  Checks all supported lengths, both directions, against a double-precision DFT.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include "../../inverseproblemwithdiffusionmodel_b200/csrc/fft_core.cuh"
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

int main() {
  double worst = 0;
#define CHK(L)                                                       \
  {                                                                  \
    double a = check<L, -1>(), b = check<L, +1>();                   \
    printf("L=%d fwd %.3e inv %.3e\n", L, a, b);                     \
    worst = fmax(worst, fmax(a, b));                                 \
  }
  CHK(8) CHK(16) CHK(32) CHK(64) CHK(128) CHK(256) CHK(512) CHK(1024)
  printf("worst %.3e\n", worst);
  return worst < 2e-6 ? 0 : 1;
}
