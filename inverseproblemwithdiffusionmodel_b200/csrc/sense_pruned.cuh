// SENSE kernels for column masks that keep few columns (ns <= 32 of W): pruned row transforms (fftpr.cuh) and a
// compact scratch.  Included by sense.cu; used whenever a plan (ipdm_sense_plan_create) says `pruned`.
//
// * Row kernels never run the second pass of a full transform: per coil a thread does one 16-point DFT in registers,
//   the transform's R1 threads (8, 16 or 32 lanes of ONE warp: __syncwarp only) exchange through a private line of
//   shared memory, and each sampled column is one R1-term sum with a register-resident twiddle vector (forward) or
//   is spread over its residue class (adjoint).  W = 512 takes R1 = 32: 16 values per thread instead of the 32 of the
//   full engine, which is what keeps these kernels near 128 registers.
// * The scratch is compact: T[c][b][h][slot], slot = rank of the column among the sampled ones (ns_pad per row), so a
//   row kernel reads / writes one contiguous run per row and coil, and the column kernels (full two-pass transforms
//   along H of the ns sampled columns only) see a few percent of k-space.
// * Everything that depends on the mask alone -- column lists, residue classes, twiddle vectors, active 32-byte
//   sectors, the zero-fill bitmap, the H-transform twiddles -- comes from the plan; no kernel compiles a mask or calls
//   sincos.
// * The forward row kernel also zero-fills every inactive sector of its rows of the output (those stores overlap
//   the arithmetic); the forward column kernel writes the active sectors whole.
#pragma once
#include "fftpr.cuh"

namespace ipdm {

struct PlanView {
  int frames, W, ns_pad, ng_all;
  const int* ns;
  const int* ngroups;
  const uint16_t* kcol;
  const uint8_t* nat;
  const uint8_t* k0c;
  const uint32_t* cls;      // [frames][5]
  const cf32* tw;           // [frames][ns_pad][W/16], class order
  const uint8_t* groups;    // [frames][W/4]
  const uint8_t* gslot;     // [frames][W/4][4]
  const uint32_t* gbitmap;  // [frames][4]
  const cf32* tws_h;        // layout-B twiddles of the H transform (Geo<H>::NTWS entries)
};

template <int L> struct PGeo {
  using P = PR<L>;
  static constexpr int RPW = 32 / P::R1, WARPS = 4, NT = 128, TPC = RPW * WARPS;   // rows per CTA: 4, 8, 16
  static constexpr int YP = 32;                                                    // per-row spectrum line (class order)
  static constexpr int ZP = TPC * (L / 2) / NT;                                    // 16-byte zero-fill pieces per thread and coil
};

// The outputs (forward) / inputs (adjoint) of thread t: class positions jj = t + R1*o, o < NOUT.
template <int L, int NOUT> struct MySlots {
  int k0[NOUT], slot[NOUT];
  __device__ __forceinline__ void init(const PlanView& p, int f, int ns, int t) {
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
      const int jj = t + PR<L>::R1 * o;
      const bool on = jj < ns;
      k0[o] = on ? p.k0c[f * p.ns_pad + jj] : 0;
      slot[o] = on ? p.nat[f * p.ns_pad + jj] : -1;
    }
  }
};

// twr[o][tt] = w_W^(tt * k) for the thread's outputs; ALT multiplies by (-1)^tt (the fused step works on the
// un-centred spectrum: column k of the mask sits at plain index k ^ (W/2)).
template <int L, int NOUT, bool ALT>
__device__ __forceinline__ void load_my_twiddles(cf32 (&twr)[NOUT][PR<L>::R1], const PlanView& p, int f, int ns, int t) {
  using P = PR<L>;
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    const int jj = t + P::R1 * o;
    const cf32x2* src = reinterpret_cast<const cf32x2*>(p.tw + ((size_t)f * p.ns_pad + (jj < ns ? jj : 0)) * P::R1);
#pragma unroll
    for (int i = 0; i < P::R1 / 2; ++i) {
      const cf32x2 w = src[i];
      twr[o][2 * i] = w.a;
      twr[o][2 * i + 1] = ALT ? cf32{-w.b.x, -w.b.y} : w.b;
    }
  }
}

// ---- forward, rows: coil multiply, pruned transform along W, compact scratch, zero-fill -------------------------
// grid (batch, H / TPC)
template <int L, int NOUT, bool CPLX>
__global__ void __launch_bounds__(128) kp_fwd_rows(SenseArgs a, PlanView p) {
  using G = PGeo<L>;
  using P = PR<L>;
  __shared__ __align__(16) cf32 xch[G::TPC * P::LINE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = lane % P::R1, r = warp * G::RPW + lane / P::R1;
  // batch index fastest: CTAs that run together share the same rows of the coil maps (L2 hits instead of DRAM re-reads)
  const int b = blockIdx.x, h0 = blockIdx.y * G::TPC, h = h0 + r;
  const int f = b % p.frames, ns = p.ns[f];
  cf32* sx = xch + r * P::LINE;
  MySlots<L, NOUT> my;
  my.init(p, f, ns, t);
  cf32 twr[NOUT][P::R1];
  load_my_twiddles<L, NOUT, false>(twr, p, f, ns, t);
  cf32 xq[P::R0];
  {
    const cf32* xp = a.in + ((size_t)b * a.H + h) * L + t;
    const float sg = sgn(h + t);   // R1 is even: the (-1)^(h+w) factor is one sign per thread
#pragma unroll
    for (int q = 0; q < P::R0; ++q) xq[q] = cscale(xp[P::R1 * q], sg);
  }
  const bool has_maps = a.mre != nullptr;
  const size_t map_img = (size_t)a.H * L;
  const float* mre = a.mre + (size_t)h * L + t;
  const float* mim = a.mim + (size_t)h * L + t;
  MapVal<CPLX> mnext[P::R0];
  auto fetch_maps = [&](int c) {
    if (has_maps && c < a.ncoils) {
#pragma unroll
      for (int q = 0; q < P::R0; ++q) mnext[q].load(mre, mim, P::R1 * q);
      mre += map_img;
      mim += map_img;
    }
  };
  fetch_maps(0);
  // this thread's zero-fill pieces are the same for every coil: which of them lie in inactive sectors
  uint32_t zmask = 0;
#pragma unroll
  for (int z = 0; z < G::ZP; ++z) {
    const int grp = ((tid + z * G::NT) % (L / 2)) >> 1;
    if (((p.gbitmap[f * 4 + (grp >> 5)] >> (grp & 31)) & 1u) == 0u) zmask |= 1u << z;
  }
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t img_stride = (size_t)a.batch * a.H * L;
  float4* zbase = reinterpret_cast<float4*>(a.out + ((size_t)b * a.H + h0) * L);
  cf32* wsp = a.ws + ((size_t)b * a.H + h) * p.ns_pad;
  const size_t ws_stride = (size_t)a.batch * a.H * p.ns_pad;
  for (int c = 0; c < a.ncoils; ++c) {
    cf32 u[P::R0];
#pragma unroll
    for (int q = 0; q < P::R0; ++q) u[q] = has_maps ? mnext[q].mul(xq[q]) : xq[q];
    fetch_maps(c + 1);
#pragma unroll
    for (int z = 0; z < G::ZP; ++z)
      if ((zmask >> z) & 1u) zbase[tid + z * G::NT] = zero4;
    zbase += img_stride / 2;
    __syncwarp();   // the previous coil's sums have read the line
    pr_first<L, -1>(u, t, sx);
    __syncwarp();
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
      if (my.slot[o] >= 0) wsp[my.slot[o]] = pr_gather<L, -1>(sx, my.k0[o], twr[o]);
    wsp += ws_stride;
  }
}

// ---- column kernels: full two-pass transforms along H of the sampled columns ------------------------------------
// grid (ncoils * batch, chunks); a chunk = 4 consecutive active groups = 16 column positions, one transform each
// (positions that are not sampled idle).
template <int LH>
__device__ __forceinline__ void copy_tws(cf32* tws, const cf32* src, int tid, int nt) {
  for (int e = tid; e < Geo<LH>::NTWS; e += nt) tws[e] = src[e];
}

template <int LH>
__global__ void __launch_bounds__(Geo<LH>::NT_COLS) kp_fwd_cols(SenseArgs a, PlanView p) {
  using G = Geo<LH>;
  using P = P2<LH>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tws = reinterpret_cast<cf32*>(smem_raw);
  cf32* xch = tws + G::NTWS;
  const int tid = threadIdx.x;
  const size_t img = blockIdx.x;
  const int b = (int)(img % a.batch), f = b % p.frames;
  const int ng = p.ngroups[f], g0 = blockIdx.y * 4;
  if (g0 >= ng) return;
  copy_tws<LH>(tws, p.tws_h, tid, G::NT_COLS);
  const int cs = tid / G::TPF, t = tid % G::TPF;
  cf32* sx = xch + cs * P::STRIDE;
  const bool gvalid = g0 + (cs >> 2) < ng;
  const int slot = gvalid ? p.gslot[((size_t)f * p.ng_all + g0 + (cs >> 2)) * 4 + (cs & 3)] : 255;
  cf32 v[G::E];
  {
    const cf32* wp = a.ws + (img * LH + t) * p.ns_pad + (slot != 255 ? slot : 0);
#pragma unroll
    for (int q = 0; q < G::E; ++q) v[q] = slot != 255 ? wp[(size_t)a_off<LH>(q) * p.ns_pad] : cf32{0.f, 0.f};
  }
  __syncthreads();   // twiddle table complete
  Twid<LH, (LH < 512)> tw;
  tw.init(tws, t);
  a2b_first<LH, -1>(v, t, sx);
  __syncwarp();
  a2b_second<LH, -1>(v, t, sx, tw);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < G::E; ++i) sx[b_pos<LH>(t, i)] = v[i];   // the exchange line doubles as the column's tile line
  __syncthreads();
  // 16-byte pieces: idx = ((gi*LH + h)*2 + half)
  for (int idx = tid; idx < 4 * LH * 2; idx += G::NT_COLS) {
    const int half = idx & 1, hh = (idx >> 1) % LH, gi = (idx >> 1) / LH;
    if (g0 + gi >= ng) break;
    const int kk = 4 * p.groups[f * p.ng_all + g0 + gi] + 2 * half;
    const cf32 p0 = xch[(4 * gi + 2 * half) * P::STRIDE + hh], p1 = xch[(4 * gi + 2 * half + 1) * P::STRIDE + hh];
    const float s0 = a.scale * sgn(hh + kk);
    *reinterpret_cast<float4*>(a.out + (img * LH + hh) * a.W + kk) = make_float4(p0.x * s0, p0.y * s0, -p1.x * s0, -p1.y * s0);
  }
}

// adjoint, columns: inverse transform of the sampled columns along H into the compact scratch
template <int LH>
__global__ void __launch_bounds__(Geo<LH>::NT_COLS) kp_adj_cols(SenseArgs a, PlanView p) {
  using G = Geo<LH>;
  using P = P2<LH>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tws = reinterpret_cast<cf32*>(smem_raw);
  cf32* xch = tws + G::NTWS;
  const int tid = threadIdx.x;
  const size_t img = blockIdx.x;
  const int b = (int)(img % a.batch), f = b % p.frames;
  const int ng = p.ngroups[f], g0 = blockIdx.y * 4;
  if (g0 >= ng) return;
  copy_tws<LH>(tws, p.tws_h, tid, G::NT_COLS);
  for (int idx = tid; idx < 4 * LH * 2; idx += G::NT_COLS) {
    const int half = idx & 1, hh = (idx >> 1) % LH, gi = (idx >> 1) / LH;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g0 + gi < ng) {
      const int g = g0 + gi, kk = 4 * p.groups[f * p.ng_all + g] + 2 * half;
      const uint8_t* gs = p.gslot + ((size_t)f * p.ng_all + g) * 4 + 2 * half;
      q = *reinterpret_cast<const float4*>(a.in + (img * LH + hh) * a.W + kk);
      const float s0 = sgn(hh + kk);
      const float m0 = gs[0] != 255 ? s0 : 0.f, m1 = gs[1] != 255 ? -s0 : 0.f;
      q = make_float4(q.x * m0, q.y * m0, q.z * m1, q.w * m1);
    }
    xch[(4 * gi + 2 * half) * P::STRIDE + hh] = cf32{q.x, q.y};
    xch[(4 * gi + 2 * half + 1) * P::STRIDE + hh] = cf32{q.z, q.w};
  }
  __syncthreads();
  const int cs = tid / G::TPF, t = tid % G::TPF;
  cf32* sx = xch + cs * P::STRIDE;
  const bool gvalid = g0 + (cs >> 2) < ng;
  const int slot = gvalid ? p.gslot[((size_t)f * p.ng_all + g0 + (cs >> 2)) * 4 + (cs & 3)] : 255;
  Twid<LH, (LH < 512)> tw;
  tw.init(tws, t);
  cf32 v[G::E];
#pragma unroll
  for (int q = 0; q < G::E; ++q) v[q] = sx[a_pos<LH>(t, q)];
  __syncwarp();
  a2b_first<LH, +1>(v, t, sx);
  __syncwarp();
  a2b_second<LH, +1>(v, t, sx, tw);
  if (slot != 255) {
    cf32* wp = a.ws + (img * LH + t) * p.ns_pad + slot;
#pragma unroll
    for (int i = 0; i < G::E; ++i) wp[(size_t)b_off<LH>(i) * p.ns_pad] = v[i];
  }
}

// ---- adjoint, rows: compact scratch -> pruned inverse transform along W -> conj-coil sum (or SSOS) ---------------
// grid (batch, H / TPC)
template <int L, int NOUT, bool CPLX>
__global__ void __launch_bounds__(128) kp_adj_rows(SenseArgs a, PlanView p) {
  using G = PGeo<L>;
  using P = PR<L>;
  __shared__ __align__(16) cf32 twc[32 * P::R1];
  __shared__ __align__(16) cf32 ysm[G::TPC * G::YP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = lane % P::R1, r = warp * G::RPW + lane / P::R1;
  const int b = blockIdx.x, h = blockIdx.y * G::TPC + r;
  const int f = b % p.frames, ns = p.ns[f];
  for (int e = tid; e < ns * P::R1; e += G::NT) twc[e] = p.tw[(size_t)f * p.ns_pad * P::R1 + e];
  uint32_t cw[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) cw[i] = p.cls[f * 5 + i];
  MySlots<L, NOUT> my;
  my.init(p, f, ns, t);
  cf32* yr = ysm + r * G::YP;
  const cf32* wsp = a.ws + ((size_t)b * a.H + h) * p.ns_pad;
  const size_t ws_stride = (size_t)a.batch * a.H * p.ns_pad;
  const bool has_maps = a.mre != nullptr && !a.ssos;
  const size_t map_img = (size_t)a.H * L;
  const float* mre = a.mre + (size_t)h * L + t;
  const float* mim = a.mim + (size_t)h * L + t;
  const float sct = a.scale * sgn(h + t);
  cf32 acc[P::R0], ynext[NOUT];
  MapVal<CPLX> mnext[P::R0];
#pragma unroll
  for (int q = 0; q < P::R0; ++q) acc[q] = cf32{0.f, 0.f};
  auto prefetch = [&](int c) {
    if (c < a.ncoils) {
#pragma unroll
      for (int o = 0; o < NOUT; ++o) ynext[o] = my.slot[o] >= 0 ? wsp[my.slot[o]] : cf32{0.f, 0.f};
      wsp += ws_stride;
      if (has_maps) {
#pragma unroll
        for (int q = 0; q < P::R0; ++q) mnext[q].load(mre, mim, P::R1 * q);
        mre += map_img;
        mim += map_img;
      }
    }
  };
  prefetch(0);
  __syncthreads();   // twiddle table complete
  for (int c = 0; c < a.ncoils; ++c) {
    MapVal<CPLX> m[P::R0];
#pragma unroll
    for (int q = 0; q < P::R0; ++q) m[q] = mnext[q];
    __syncwarp();   // the previous coil's sums have read the spectrum line
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
      if (my.slot[o] >= 0) yr[t + P::R1 * o] = ynext[o];
    prefetch(c + 1);
    __syncwarp();
    cf32 v[P::R0];
    pr_scatter<L, +1>(v, t, yr, twc, P::R1, cw);
#pragma unroll
    for (int q = 0; q < P::R0; ++q) {
      if (a.ssos) {
        acc[q].x += v[q].x * v[q].x + v[q].y * v[q].y;
      } else if (has_maps) {
        acc[q] = cadd(acc[q], m[q].mulc(v[q]));
      } else {
        acc[q] = cadd(acc[q], v[q]);
      }
    }
  }
  if (a.ssos) {
    float* op = reinterpret_cast<float*>(a.out) + ((size_t)b * a.H + h) * L + t;
#pragma unroll
    for (int q = 0; q < P::R0; ++q) op[P::R1 * q] = sqrtf(acc[q].x) * fabsf(sct);
  } else {
    cf32* op = a.out + ((size_t)b * a.H + h) * L + t;
#pragma unroll
    for (int q = 0; q < P::R0; ++q) op[P::R1 * q] = cscale(acc[q], sct);
  }
}

// ---- fused Langevin update + SENSE L2-penalty step, pruned.  grid (batch, H / TPC) ------------------------------
// z = x + step*g + noise_scale*n;  x <- z - kappa*(A^H A z - b).  The H-axis transforms cancel in A^H A (the mask
// acts on W only) and the (-1)^w factors of the centred transforms turn into the half-period shift k ^ (W/2) of the
// sampled columns, so per coil: multiply, pruned forward transform (the ns sampled columns), pruned inverse
// transform, conj multiply-accumulate -- all on the 16 values a thread holds.
template <int L, int NOUT, bool CPLX>
__global__ void __launch_bounds__(128) kp_ald_sense(AldArgs a, PlanView p) {
  using G = PGeo<L>;
  using P = PR<L>;
  __shared__ __align__(16) cf32 xch[G::TPC * P::LINE];
  __shared__ __align__(16) cf32 twc[32 * P::R1];
  __shared__ __align__(16) cf32 ysm[G::TPC * G::YP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = lane % P::R1, r = warp * G::RPW + lane / P::R1;
  const int b = blockIdx.x, h = blockIdx.y * G::TPC + r;
  const int f = b % p.frames, ns = p.ns[f];
  ipdm_ald_scalars sc = a.sc;
  uint32_t rstep = a.rng.step;
  if (a.sched != nullptr) {
    const int cur = *a.cursor;
    sc = a.sched[cur];
    rstep += (uint32_t)cur;
  }
  for (int e = tid; e < ns * P::R1; e += G::NT) {   // entry e belongs to thread e % R1: (-1)^t' folds the k ^ (W/2) shift in
    const cf32 w = p.tw[(size_t)f * p.ns_pad * P::R1 + e];
    twc[e] = (e & 1) ? cf32{-w.x, -w.y} : w;
  }
  uint32_t cw[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) cw[i] = p.cls[f * 5 + i];
  MySlots<L, NOUT> my;
  my.init(p, f, ns, t);
  cf32 twr[NOUT][P::R1];
  load_my_twiddles<L, NOUT, true>(twr, p, f, ns, t);
  cf32* sx = xch + r * P::LINE;
  cf32* yr = ysm + r * G::YP;
  const size_t plane = (size_t)a.batch * a.H * L;
  const size_t rowoff = ((size_t)b * a.H + h) * L + t;
  const size_t map_img = (size_t)a.H * L;
  const float* mre = a.mre + (size_t)h * L + t;
  const float* mim = a.mim + (size_t)h * L + t;
  MapVal<CPLX> mnext[P::R0];
  auto fetch_maps = [&](int c) {
    if (c < a.ncoils) {
#pragma unroll
      for (int q = 0; q < P::R0; ++q) mnext[q].load(mre, mim, P::R1 * q);
      mre += map_img;
      mim += map_img;
    }
  };
  fetch_maps(0);
  cf32 z[P::R0], acc[P::R0];
  {
    const float *xr = a.x + rowoff, *xi = a.x + plane + rowoff, *gr = a.grad + rowoff, *gi = a.grad + plane + rowoff;
#pragma unroll
    for (int q = 0; q < P::R0; ++q) {
      z[q].x = xr[P::R1 * q] + sc.step * gr[P::R1 * q];
      z[q].y = xi[P::R1 * q] + sc.step * gi[P::R1 * q];
      acc[q] = cf32{0.f, 0.f};
    }
    if (a.noise != nullptr) {
      const float *nr = a.noise + rowoff, *ni = a.noise + plane + rowoff;
#pragma unroll
      for (int q = 0; q < P::R0; ++q) {
        z[q].x += sc.noise_scale * nr[P::R1 * q];
        z[q].y += sc.noise_scale * ni[P::R1 * q];
      }
    } else if (sc.noise_scale != 0.f) {
      const uint64_t seed = rng_seed(a.rng);
      const uint32_t chain = rng_chain(a.rng, b);
#pragma unroll
      for (int q = 0; q < P::R0 / 2; ++q) {   // pixels w and w + W/2 share one Philox call (same pairing in every kernel family)
        float n[4];
        philox_chain_normal4(seed, chain, (uint32_t)(h * L + P::R1 * q + t), rstep, n);
        z[q].x += sc.noise_scale * n[0];
        z[q].y += sc.noise_scale * n[1];
        z[q + P::R0 / 2].x += sc.noise_scale * n[2];
        z[q + P::R0 / 2].y += sc.noise_scale * n[3];
      }
    }
  }
  __syncthreads();   // twiddle table complete
  for (int c = 0; c < a.ncoils; ++c) {
    MapVal<CPLX> m[P::R0];
    cf32 u[P::R0];
#pragma unroll
    for (int q = 0; q < P::R0; ++q) {
      m[q] = mnext[q];
      u[q] = m[q].mul(z[q]);
    }
    fetch_maps(c + 1);
    __syncwarp();   // line and spectrum of the previous coil are consumed
    pr_first<L, -1>(u, t, sx);
    __syncwarp();
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
      if (my.slot[o] >= 0) yr[t + P::R1 * o] = pr_gather<L, -1>(sx, my.k0[o], twr[o]);
    __syncwarp();
    pr_scatter<L, +1>(u, t, yr, twc, P::R1, cw);
#pragma unroll
    for (int q = 0; q < P::R0; ++q) acc[q] = cadd(acc[q], m[q].mulc(u[q]));
  }
  const float ks = sc.kappa / (float)L;
  {
    float *xr = a.x + rowoff, *xi = a.x + plane + rowoff;
    const float *br = a.bvec + rowoff, *bi = a.bvec + plane + rowoff;
    cf32 bv[P::R0];   // every b load is issued before the first store (x and b may alias as far as the compiler knows)
#pragma unroll
    for (int q = 0; q < P::R0; ++q) bv[q] = cf32{br[P::R1 * q], bi[P::R1 * q]};
#pragma unroll
    for (int q = 0; q < P::R0; ++q) {
      xr[P::R1 * q] = z[q].x - ks * acc[q].x + sc.kappa * bv[q].x;
      xi[P::R1 * q] = z[q].y - ks * acc[q].y + sc.kappa * bv[q].y;
    }
  }
}

}  // namespace ipdm
