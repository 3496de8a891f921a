// CPU emulation of the Stockham passes in csrc/fft_core.cuh: every "thread" of one transform is run
// in a loop with a full barrier between load / store phases, exactly as the kernels do with
// __syncthreads().  Checks all supported lengths, both directions, against a double-precision DFT.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include "../../inverseproblemwithdiffusionmodel_b200/csrc/fft_core.cuh"
#include "../../inverseproblemwithdiffusionmodel_b200/csrc/fft2p.cuh"
#include "../../inverseproblemwithdiffusionmodel_b200/csrc/fftpr.cuh"
#include "../../inverseproblemwithdiffusionmodel_b200/csrc/sense_plan.h"
#include <cstring>
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

// Pruned row transforms (csrc/fftpr.cuh) driven by the tables of csrc/sense_plan.h: forward = the sampled columns of a
// full DFT, adjoint = the inverse DFT of a spectrum that is zero off the sampled columns.  Two mask frames with
// different column sets, a centre window plus scattered lines as the keep-centre masks have them.
template <int L>
double check_pruned(int ns_target, unsigned seed) {
  using P = PR<L>;
  srand(seed);
  const int frames = 2;
  std::vector<uint8_t> mask(frames * L, 0);
  for (int f = 0; f < frames; ++f) {
    int n = 0;
    const int win = ns_target / 3;
    for (int k = L / 2 - win / 2; k < L / 2 - win / 2 + win; ++k) { mask[f * L + k] = 1; ++n; }
    while (n < ns_target - f) { const int k = rand() % L; if (!mask[f * L + k]) { mask[f * L + k] = 1; ++n; } }
  }
  PlanHost pl = build_plan_host(mask.data(), frames, L);
  if (!pl.pruned) {   // legitimate: a residue class with more than 4 columns (the general kernels take over)
    int cm = 0;
    for (int f = 0; f < frames; ++f) { int cnt[16] = {0}; for (int k = 0; k < L; ++k) if (mask[f * L + k]) cm = std::max(cm, ++cnt[k & 15]); }
    printf("plan not pruned for L=%d ns=%d (largest class %d)\n", L, ns_target, cm);
    return cm > 4 ? 0.0 : 1.0;
  }
  if (pl.R1 != P::R1) return 1.0;
  double worst = 0;
  for (int f = 0; f < frames; ++f) {
    const int ns = pl.ns[f], NP = pl.ns_pad;
    // structure checks: natural order ascending and matching the mask, classes sorted, group slots consistent
    int cnt = 0;
    for (int k = 0; k < L; ++k) if (mask[f * L + k]) { if (pl.kcol[f * NP + cnt] != k) return 2.0; ++cnt; }
    if (cnt != ns) return 3.0;
    const uint8_t* cls = &pl.cls[f * PlanHost::CLS_PITCH];
    if (cls[0] != 0 || cls[16] != ns) return 4.0;
    for (int jj = 0; jj < ns; ++jj) {
      const int k = pl.kcol[f * NP + pl.nat[f * NP + jj]];
      if ((k & 15) != pl.k0c[f * NP + jj] || jj < cls[k & 15] || jj >= cls[(k & 15) + 1]) return 5.0;
    }
    {   // column-kernel work items: whole groups, consecutive, <= 16 sampled columns, covering everything once
      int g_next = 0, s_next = 0;
      for (int c = 0; c < pl.nchunks[f]; ++c) {
        const uint8_t* ch = &pl.chunks[(f * (L / PlanHost::GW) + c) * 4];
        if (ch[0] != g_next || ch[2] != s_next || ch[1] == 0 || ch[3] == 0 || ch[3] > PlanHost::CHUNK_SLOTS) return 11.0;
        int cnt2 = 0;
        for (int g = ch[0]; g < ch[0] + ch[1]; ++g)
          for (int i = 0; i < PlanHost::GW; ++i) cnt2 += pl.gslot[(f * (L / PlanHost::GW) + g) * PlanHost::GW + i] != 255;
        if (cnt2 != ch[3]) return 12.0;
        g_next += ch[1];
        s_next += ch[3];
      }
      if (g_next != pl.ngroups[f] || s_next != ns) return 13.0;
      // the one-read chunk records of the column kernels say the same as the tables they were built from
      for (int c = 0; c < L / PlanHost::GW; ++c) {
        const ipdm::ChunkRec& rec = pl.crec[f * (L / PlanHost::GW) + c];
        if ((rec.valid != 0) != (c < pl.nchunks[f])) return 16.0;
        if (!rec.valid) continue;
        const uint8_t* ch = &pl.chunks[(f * (L / PlanHost::GW) + c) * 4];
        if (rec.g_cnt != ch[1] || rec.s_lo != ch[2] || rec.s_cnt != ch[3]) return 17.0;
        for (int i = 0; i < rec.s_cnt; ++i)
          if (rec.kcol[i] != pl.kcol[f * NP + rec.s_lo + i]) return 18.0;
        for (int gi = 0; gi < rec.g_cnt; ++gi) {
          const int g = ch[0] + gi;
          if (rec.gcol[gi] != PlanHost::GW * pl.groups[f * (L / PlanHost::GW) + g]) return 19.0;
          for (int i = 0; i < PlanHost::GW; ++i) {
            const int sl = pl.gslot[(f * (L / PlanHost::GW) + g) * PlanHost::GW + i];
            if (rec.gline[gi][i] != (sl != 255 ? sl - rec.s_lo : -1)) return 20.0;
            if (sl != 255 && rec.kcol[rec.gline[gi][i]] != rec.gcol[gi] + i) return 21.0;
          }
        }
      }
      for (int jj = 0; jj < ns; ++jj) {      // scratch position of every class entry: (chunk, position) of its natural slot
        const int cw = pl.tcw[f * NP + jj], cc = cw >> 3, within = cw & 7;
        if (cc >= pl.nchunks[f]) return 14.0;
        const uint8_t* ch = &pl.chunks[(f * (L / PlanHost::GW) + cc) * 4];
        if (within >= ch[3] || ch[2] + within != pl.nat[f * NP + jj]) return 15.0;
      }
    }
    for (int g = 0; g < pl.ngroups[f]; ++g) {
      const int q = pl.groups[f * (L / PlanHost::GW) + g];
      if (!((pl.gbitmap[f * 4 + (q >> 5)] >> (q & 31)) & 1u)) return 6.0;
      for (int i = 0; i < PlanHost::GW; ++i) {
        const int s = pl.gslot[(f * (L / PlanHost::GW) + g) * PlanHost::GW + i];
        if ((s != 255) != (mask[f * L + PlanHost::GW * q + i] != 0)) return 7.0;
        if (s != 255 && pl.kcol[f * NP + s] != PlanHost::GW * q + i) return 8.0;
      }
    }
    const cf32* tw = reinterpret_cast<const cf32*>(pl.tw.data()) + (size_t)f * NP * P::R1;
    const cf32* twh = reinterpret_cast<const cf32*>(pl.twh.data()) + (size_t)f * NP * PlanHost::TWH;
    // padded class positions: inside the class block, distinct
    std::vector<int> seen(16 * pl.cmax, 0);
    for (int jj = 0; jj < ns; ++jj) {
      const int pp = pl.ppos[f * NP + jj];
      if (pp / pl.cmax != pl.k0c[f * NP + jj] || seen[pp]++) return 9.0;
    }
    // ---- forward
    std::vector<cf32> x(L), line(P::LINE + 8);
    std::vector<std::complex<double>> xd(L);
    for (int i = 0; i < L; ++i) {
      x[i] = cf32{(float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f};
      xd[i] = {x[i].x, x[i].y};
    }
    for (int t = 0; t < P::R1; ++t) {
      cf32 v[P::R0];
      for (int q = 0; q < P::R0; ++q) v[q] = x[P::R1 * q + t];
      pr_first<L, -1>(v, t, line.data());
    }
    double err = 0, nrm = 0;
    for (int jj = 0; jj < ns; ++jj) {
      const cf32 o = pr_gather<L, -1>(line.data(), pl.k0c[f * NP + jj], twh + jj * PlanHost::TWH);
      const int k = pl.kcol[f * NP + pl.nat[f * NP + jj]];
      std::complex<double> s = 0;
      for (int n = 0; n < L; ++n) s += xd[n] * std::polar(1.0, -2.0 * M_PI * k * n / L);
      err += std::norm(s - std::complex<double>(o.x, o.y));
      nrm += std::norm(s);
    }
    worst = fmax(worst, sqrt(err / nrm));
    // ---- adjoint: padded class layout, static indices
    std::vector<cf32> Y(16 * pl.cmax, cf32{0.f, 0.f}), twp((size_t)16 * pl.cmax * P::R1, cf32{0.f, 0.f}), out(L);
    std::vector<std::complex<double>> Yd(L, 0.0);
    for (int jj = 0; jj < ns; ++jj) {
      const int pp = pl.ppos[f * NP + jj];
      Y[pp] = cf32{(float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f};
      Yd[pl.kcol[f * NP + pl.nat[f * NP + jj]]] = {Y[pp].x, Y[pp].y};
      for (int t = 0; t < P::R1; ++t) twp[(size_t)pp * P::R1 + t] = tw[jj * P::R1 + t];
    }
    for (int t = 0; t < P::R1; ++t) {
      cf32 v[P::R0];
      const uint32_t b2 = pl.big[f * 2], b3 = pl.big[f * 2 + 1];
      if (pl.cmax == 2) pr_scatter<L, +1, 2>(v, t, Y.data(), twp.data(), P::R1, b2, b3);
      else if (pl.cmax == 4) pr_scatter<L, +1, 4>(v, t, Y.data(), twp.data(), P::R1, b2, b3);
      else return 10.0;
      for (int q = 0; q < P::R0; ++q) out[P::R1 * q + t] = v[q];
    }
    err = 0; nrm = 0;
    for (int n = 0; n < L; ++n) {
      std::complex<double> s = 0;
      for (int k = 0; k < L; ++k) s += Yd[k] * std::polar(1.0, 2.0 * M_PI * k * n / L);
      err += std::norm(s - std::complex<double>(out[n].x, out[n].y));
      nrm += std::norm(s);
    }
    worst = fmax(worst, sqrt(err / nrm));
  }
  return worst;
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
#define CHKP(L, NS)                                                   \
  {                                                                   \
    double a = check_pruned<L>(NS, 17 * L + NS);                      \
    printf("pruned L=%d ns=%d worst %.3e\n", L, NS, a);               \
    worst = fmax(worst, a);                                           \
  }
  CHKP(128, 8) CHKP(128, 16) CHKP(256, 11) CHKP(256, 20) CHKP(256, 32) CHKP(512, 21) CHKP(512, 32) CHKP(512, 3)
  printf("worst %.3e\n", worst);
  return worst < 2e-6 ? 0 : 1;
}
