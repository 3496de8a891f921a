// Host-side compilation of a k-space column mask into the tables the pruned SENSE kernels read (sense_pruned.cuh).
// Pure C++ (no CUDA): sense.cu uploads the result, tests/cpu/fft_core_test.cpp checks it and drives the CPU
// emulation of the pruned transforms with it.
//
// Per mask frame f (the reference's live mask has 24 frames, undersampling_fourier.py:63-75; the keep-centre masks 1):
//   ns[f]              number of sampled columns
//   kcol[f][s]         the sampled columns, ascending ("natural" slot order = the order of the compact scratch rows)
//   nat[f][jj], k0c[f][jj], cls[f][0..16], tw[f][jj][t]
//                      the same columns in CLASS order (sorted by k mod 16, then k): natural slot, class k mod 16,
//                      class boundaries, and the twiddle vectors w_W^(t*k) = exp(-2*pi*i*t*k/W), t < W/16
//   ngroups[f], groups[f][g], gslot[f][g][0..3], gbitmap[f][4]
//                      the active 4-column groups (one 32-byte sector of a k-space row each), for every group the
//                      natural slot of each of its columns (255 = not sampled), and the bitmap over all W/4 groups
#pragma once
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <vector>

namespace ipdm {

struct PlanHost {
  int frames = 0, W = 0, R1 = 0, ns_max = 0, ns_pad = 0, ng_max = 0;
  bool pruned = false;   // every frame keeps <= the limit the pruned kernels are built for
  std::vector<int> ns, ngroups;
  std::vector<uint16_t> kcol;
  std::vector<uint8_t> nat, k0c, cls, groups, gslot, mask;
  std::vector<uint32_t> gbitmap;
  std::vector<float> tw;   // interleaved (re, im)
  static constexpr int CLS_PITCH = 20;   // 17 boundaries padded to five 32-bit words
};

// Largest ns the pruned row kernels take for a row length W (0: W not served).  One register-resident twiddle vector
// per output and thread: W = 512 -> 32 threads x 1 output, 256 -> 16 x 2, 128 -> 8 x 2.
inline int pruned_ns_limit(int W) { return W == 512 ? 32 : W == 256 ? 32 : W == 128 ? 16 : 0; }

inline PlanHost build_plan_host(const uint8_t* mask, int frames, int W) {
  PlanHost p;
  p.frames = frames;
  p.W = W;
  p.R1 = W / 16;
  p.mask.assign(mask, mask + (size_t)frames * W);
  p.ns.resize(frames);
  p.ngroups.resize(frames);
  const int ng_all = W / 4;
  for (int f = 0; f < frames; ++f) {
    int n = 0, g = 0;
    for (int k = 0; k < W; ++k) n += mask[(size_t)f * W + k] != 0;
    for (int q = 0; q < ng_all; ++q) {
      const uint8_t* m = mask + (size_t)f * W + 4 * q;
      g += (m[0] | m[1] | m[2] | m[3]) != 0;
    }
    p.ns[f] = n;
    p.ngroups[f] = g;
    p.ns_max = std::max(p.ns_max, n);
    p.ng_max = std::max(p.ng_max, g);
  }
  const int limit = pruned_ns_limit(W);
  p.pruned = limit > 0 && p.ns_max >= 1 && p.ns_max <= limit;
  p.ns_pad = std::max(4, (p.ns_max + 3) & ~3);
  if (!p.pruned) return p;
  const int NP = p.ns_pad, R1 = p.R1;
  p.kcol.assign((size_t)frames * NP, 0);
  p.nat.assign((size_t)frames * NP, 0);
  p.k0c.assign((size_t)frames * NP, 0);
  p.cls.assign((size_t)frames * PlanHost::CLS_PITCH, 0);
  p.tw.assign((size_t)frames * NP * R1 * 2, 0.f);
  p.groups.assign((size_t)frames * ng_all, 0);
  p.gslot.assign((size_t)frames * ng_all * 4, 255);
  p.gbitmap.assign((size_t)frames * 4, 0u);
  for (int f = 0; f < frames; ++f) {
    const uint8_t* m = mask + (size_t)f * W;
    std::vector<int> cols;
    for (int k = 0; k < W; ++k)
      if (m[k]) cols.push_back(k);
    for (size_t s = 0; s < cols.size(); ++s) p.kcol[(size_t)f * NP + s] = (uint16_t)cols[s];
    // class order: by k mod 16, then k
    std::vector<int> order(cols.size());
    for (size_t s = 0; s < cols.size(); ++s) order[s] = (int)s;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (cols[a] & 15) < (cols[b] & 15); });
    uint8_t* cls = &p.cls[(size_t)f * PlanHost::CLS_PITCH];
    size_t jj = 0;
    for (int k0 = 0; k0 < 16; ++k0) {
      cls[k0] = (uint8_t)jj;
      while (jj < order.size() && (cols[order[jj]] & 15) == k0) ++jj;
    }
    cls[16] = (uint8_t)jj;
    for (size_t j = 0; j < order.size(); ++j) {
      const int s = order[j], k = cols[s];
      p.nat[(size_t)f * NP + j] = (uint8_t)s;
      p.k0c[(size_t)f * NP + j] = (uint8_t)(k & 15);
      for (int t = 0; t < R1; ++t) {
        const int e = (int)(((long long)t * k) % W);
        const double ang = -2.0 * M_PI * (double)e / (double)W;
        p.tw[(((size_t)f * NP + j) * R1 + t) * 2] = (float)cos(ang);
        p.tw[(((size_t)f * NP + j) * R1 + t) * 2 + 1] = (float)sin(ang);
      }
    }
    int g = 0;
    for (int q = 0; q < ng_all; ++q) {
      if (!(m[4 * q] | m[4 * q + 1] | m[4 * q + 2] | m[4 * q + 3])) continue;
      p.groups[(size_t)f * ng_all + g] = (uint8_t)q;
      p.gbitmap[(size_t)f * 4 + (q >> 5)] |= 1u << (q & 31);
      for (int i = 0; i < 4; ++i) {
        if (!m[4 * q + i]) continue;
        const int s = (int)(std::lower_bound(cols.begin(), cols.end(), 4 * q + i) - cols.begin());
        p.gslot[((size_t)f * ng_all + g) * 4 + i] = (uint8_t)s;
      }
      ++g;
    }
  }
  return p;
}

// Layout-B twiddles of the full two-pass engine (fft2p.cuh) for length L in the order the kernels keep them:
// tws[(j*(R1-1) + t-1)*R1 + u] = exp(-2*pi*i * (t*(u + R1*j) mod L) / L).
inline std::vector<float> build_tws_host(int L, int R0, int R1) {
  std::vector<float> out;
  if (R1 <= 1) return out;
  const int G = R0 / R1, n = G * (R1 - 1) * R1;
  out.resize((size_t)n * 2);
  for (int e = 0; e < n; ++e) {
    const int u = e % R1, m = e / R1, j = m / (R1 - 1), t = m % (R1 - 1) + 1;
    const double ang = -2.0 * M_PI * (double)((t * (u + R1 * j)) % L) / (double)L;
    out[2 * e] = (float)cos(ang);
    out[2 * e + 1] = (float)sin(ang);
  }
  return out;
}

}  // namespace ipdm
