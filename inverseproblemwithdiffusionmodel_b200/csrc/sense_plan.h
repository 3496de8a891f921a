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
//   ppos[f][jj]        position of class entry jj in the PADDED class layout k0*cmax + e (cmax = largest class of any
//                      frame): the inverse transforms read their classes with static indices, pads hold zeros
//   twh[f][jj][10]     the twiddle vector in factored form, w^(t*k) = w^((t&3)*k) * w^(4*(t>>2)*k): entries
//                      w^k, w^2k, w^3k, then w^(4m*k) for m = 1 .. W/64-1 (20 registers instead of 2*W/16)
//   ngroups[f], groups[f][g], gslot[f][g][0..GW-1], gbitmap[f][4]
//                      the active GW-column groups (GW = 4: one aligned 32-byte sector of a k-space row each), for every group the
//                      natural slot of each of its columns (255 = not sampled), and the bitmap over all W/GW groups
//   nchunks[f], chunks[f][c] = {first group, groups, first slot, slots}
//                      work items of the column kernels: runs of whole groups holding at most 8 sampled columns
//   crec[f][c]         ChunkRec: everything a column-kernel work item needs about chunk c in one 80-byte record (sampled
//                      columns, first column of each group, line of every group column)
//   tcw[f][jj]         chunk*8 + position inside the chunk of class entry jj: the compact scratch is laid out
//                      T[image][chunk][h][8], so that a column-kernel work item is one contiguous block
#pragma once
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <vector>

namespace ipdm {

#ifndef IPDM_PLAN_GW
#define IPDM_PLAN_GW 4
#endif

// Everything a column-kernel work item needs to know about its chunk, in ONE record (one 16-byte-aligned read instead of
// the chain nchunks -> chunks -> groups -> gslot -> kcol of dependent table reads: half of kp_fwd_cols' stall samples).
struct alignas(16) ChunkRec {
  uint8_t g_cnt, s_cnt, s_lo, valid;     // active groups, sampled columns, first natural slot; valid = chunk exists in this frame
  uint16_t kcol[8];                      // the sampled columns (natural order)
  uint16_t gcol[8];                      // first column of each active group
  int8_t gline[8][IPDM_PLAN_GW];         // per group and column: line (= slot - s_lo) or -1
};

struct PlanHost {
  int frames = 0, W = 0, R1 = 0, ns_max = 0, ns_pad = 0, ng_max = 0, cmax = 0, nchunks_max = 0;
  bool pruned = false;   // every frame keeps <= the limit the pruned kernels are built for and no residue class holds > 4 columns
  std::vector<int> ns, ngroups, nchunks;
  std::vector<uint16_t> kcol;
  std::vector<uint8_t> nat, k0c, cls, ppos, tcw, groups, gslot, chunks, mask;
  std::vector<uint32_t> gbitmap, big;   // big[f][2]: classes with >= 3 / == 4 entries, one bit per class
  std::vector<float> tw, twh;   // interleaved (re, im)
  std::vector<ChunkRec> crec;   // [frames][W/GW] (chunk index), zero (valid = 0) past nchunks[f]
  static constexpr int CLS_PITCH = 20;   // 17 boundaries padded to five 32-bit words
  static constexpr int TWH = 10;         // factored twiddle entries per column
  // Output columns are handled in groups of GW: the forward column kernel writes whole groups (GW * 8 bytes, aligned), the
  // forward row kernel zero-fills every group without a sampled column, so no GW*8-byte unit of the output is written by
  // both.  GW = 4 (one 32-byte sector) or 8 (64 bytes): measured the same within 2 % at every sweep point, GW = 4 slightly
  // ahead on the forward (1.200 vs 1.226 ms at 32 coils x 512^2 x 64; 51.5 vs 54 us at 4 x 256^2 x 64).
  static constexpr int GW = IPDM_PLAN_GW;
  static constexpr int CHUNK_SLOTS = GW > 8 ? GW : 8;  // sampled columns per work item of the column kernels (>= GW)
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
  const int ng_all = W / PlanHost::GW;
  constexpr int GW = PlanHost::GW;
  for (int f = 0; f < frames; ++f) {
    int n = 0, g = 0;
    for (int k = 0; k < W; ++k) n += mask[(size_t)f * W + k] != 0;
    for (int q = 0; q < ng_all; ++q) {
      const uint8_t* m = mask + (size_t)f * W + GW * q;
      int any = 0;
      for (int i = 0; i < GW; ++i) any |= m[i];
      g += any != 0;
    }
    p.ns[f] = n;
    p.ngroups[f] = g;
    p.ns_max = std::max(p.ns_max, n);
    p.ng_max = std::max(p.ng_max, g);
    int cnt[16] = {0};   // largest residue class (column mod 16) over all frames
    for (int k = 0; k < W; ++k)
      if (mask[(size_t)f * W + k]) p.cmax = std::max(p.cmax, ++cnt[k & 15]);
  }
  const int limit = pruned_ns_limit(W);
  p.pruned = limit > 0 && p.ns_max >= 1 && p.ns_max <= limit && p.cmax <= 4;
  p.cmax = p.cmax <= 2 ? 2 : 4;   // padded class size the kernels are built for
  p.ns_pad = std::max(4, (p.ns_max + 3) & ~3);
  if (!p.pruned) return p;
  const int NP = p.ns_pad, R1 = p.R1;
  p.kcol.assign((size_t)frames * NP, 0);
  p.nat.assign((size_t)frames * NP, 0);
  p.k0c.assign((size_t)frames * NP, 0);
  p.cls.assign((size_t)frames * PlanHost::CLS_PITCH, 0);
  p.ppos.assign((size_t)frames * NP, 0);
  p.tcw.assign((size_t)frames * NP, 0);
  p.tw.assign((size_t)frames * NP * R1 * 2, 0.f);
  p.twh.assign((size_t)frames * NP * PlanHost::TWH * 2, 0.f);
  p.nchunks.assign(frames, 0);
  p.chunks.assign((size_t)frames * ng_all * 4, 0);
  p.groups.assign((size_t)frames * ng_all, 0);
  p.gslot.assign((size_t)frames * ng_all * GW, 255);
  p.gbitmap.assign((size_t)frames * 4, 0u);
  p.big.assign((size_t)frames * 2, 0u);
  p.crec.assign((size_t)frames * ng_all, ChunkRec{});
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
    for (int k0 = 0; k0 < 16; ++k0) {
      if (cls[k0 + 1] - cls[k0] >= 3) p.big[(size_t)f * 2] |= 1u << k0;
      if (cls[k0 + 1] - cls[k0] >= 4) p.big[(size_t)f * 2 + 1] |= 1u << k0;
    }
    for (size_t j = 0; j < order.size(); ++j) {
      const int s = order[j], k = cols[s];
      p.nat[(size_t)f * NP + j] = (uint8_t)s;
      p.k0c[(size_t)f * NP + j] = (uint8_t)(k & 15);
      p.ppos[(size_t)f * NP + j] = (uint8_t)((k & 15) * p.cmax + ((int)j - cls[k & 15]));
      auto tw_of = [&](int t, float* out) {
        const int e = (int)(((long long)t * k) % W);
        const double ang = -2.0 * M_PI * (double)e / (double)W;
        out[0] = (float)cos(ang);
        out[1] = (float)sin(ang);
      };
      for (int t = 0; t < R1; ++t) tw_of(t, &p.tw[(((size_t)f * NP + j) * R1 + t) * 2]);
      float* th = &p.twh[((size_t)f * NP + j) * PlanHost::TWH * 2];
      for (int i = 1; i <= 3; ++i) tw_of(i, th + 2 * (i - 1));
      for (int m = 1; m < R1 / 4; ++m) tw_of(4 * m, th + 2 * (2 + m));
    }
    int g = 0;
    for (int q = 0; q < ng_all; ++q) {
      int any = 0;
      for (int i = 0; i < GW; ++i) any |= m[GW * q + i];
      if (!any) continue;
      p.groups[(size_t)f * ng_all + g] = (uint8_t)q;
      p.gbitmap[(size_t)f * 4 + (q >> 5)] |= 1u << (q & 31);
      for (int i = 0; i < GW; ++i) {
        if (!m[GW * q + i]) continue;
        const int s = (int)(std::lower_bound(cols.begin(), cols.end(), GW * q + i) - cols.begin());
        p.gslot[((size_t)f * ng_all + g) * GW + i] = (uint8_t)s;
      }
      ++g;
    }
    // column-kernel work items: consecutive whole groups, at most CHUNK_SLOTS sampled columns each
    int c = 0, g_lo = 0, s_lo = 0;
    while (g_lo < g) {
      int g_hi = g_lo, s_hi = s_lo;
      while (g_hi < g) {
        int in_group = 0;
        for (int i = 0; i < GW; ++i) in_group += p.gslot[((size_t)f * ng_all + g_hi) * GW + i] != 255;
        if (s_hi - s_lo + in_group > PlanHost::CHUNK_SLOTS) break;
        s_hi += in_group;
        ++g_hi;
      }
      uint8_t* ch = &p.chunks[((size_t)f * ng_all + c) * 4];
      ch[0] = (uint8_t)g_lo; ch[1] = (uint8_t)(g_hi - g_lo); ch[2] = (uint8_t)s_lo; ch[3] = (uint8_t)(s_hi - s_lo);
      ChunkRec& rec = p.crec[(size_t)f * ng_all + c];
      rec.g_cnt = ch[1]; rec.s_cnt = ch[3]; rec.s_lo = ch[2]; rec.valid = 1;
      for (int i = 0; i < s_hi - s_lo; ++i) rec.kcol[i] = (uint16_t)cols[s_lo + i];
      for (int gi = 0; gi < g_hi - g_lo; ++gi) {
        rec.gcol[gi] = (uint16_t)(GW * p.groups[(size_t)f * ng_all + g_lo + gi]);
        for (int i = 0; i < GW; ++i) {
          const uint8_t sl = p.gslot[((size_t)f * ng_all + g_lo + gi) * GW + i];
          rec.gline[gi][i] = sl != 255 ? (int8_t)(sl - s_lo) : (int8_t)-1;
        }
      }
      g_lo = g_hi;
      s_lo = s_hi;
      ++c;
    }
    p.nchunks[f] = c;
    p.nchunks_max = std::max(p.nchunks_max, c);
    for (size_t j = 0; j < order.size(); ++j) {      // natural slot -> (chunk, position)
      const int sl = p.nat[(size_t)f * NP + j];
      for (int cc = 0; cc < c; ++cc) {
        const uint8_t* ch = &p.chunks[((size_t)f * ng_all + cc) * 4];
        if (sl >= ch[2] && sl < ch[2] + ch[3]) p.tcw[(size_t)f * NP + j] = (uint8_t)(cc * PlanHost::CHUNK_SLOTS + (sl - ch[2]));
      }
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
