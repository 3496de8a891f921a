// SENSE kernels on the two-pass FFT engine (fft2p.cuh) for transform lengths 64..512.  Included by sense.cu.
//
// * One transform = R1 (8 or 16) threads of ONE warp, so every exchange is guarded by __syncwarp only: the warps of
//   a CTA never wait for each other inside the coil loop and global-memory latency is hidden by the other warps.
// * The k-space column mask is compiled, per CTA, into the list of ACTIVE 4-column groups (one 32-byte sector of a
//   k-space row).  The forward row kernel zero-fills every inactive sector of the output itself -- the stores
//   overlap its FFT arithmetic -- and scatters the sampled columns into the transposed scratch T[c][b][k][h]; the
//   column kernels only ever touch active sectors (full-sector reads / writes).
// * The fused Langevin + data-consistency step runs forward (A->B), masks the spectrum where it lies, and comes
//   back (B->A) into the registers it started from: two exchanges per coil.
#pragma once
#include "fft2p.cuh"

namespace ipdm {

template <int L> struct Geo {
  using P = P2<L>;
  static constexpr int TPF = P::TPF, E = P::E;
  static constexpr int RPW = 32 / TPF;        // transforms per warp
  static constexpr int WARPS = 4, NT = 128;
  static constexpr int TPC = RPW * WARPS;     // transforms (image rows) per CTA of the row kernels
  static constexpr int NTWS = P::NTW * P::R1; // per-CTA twiddle table tws[n*R1 + u]
  static constexpr size_t SMEM_ROWS = (size_t)(NTWS + TPC * P::STRIDE) * sizeof(cf32);
  // column kernels: 16 columns (4 sectors) per CTA
  static constexpr int NT_COLS = 16 * TPF;
  static constexpr size_t SMEM_COLS = (size_t)(NTWS + 16 * P::STRIDE) * sizeof(cf32);
  static_assert(P::STRIDE >= L, "the exchange line doubles as a tile line of L values");
};

// tws[(j*(R1-1) + t-1)*R1 + u] = w_L^(t*(u + R1*j)) (forward sign), the layout-B twiddles of thread u.
template <int L>
__device__ __forceinline__ void fill_tws(cf32* tws, int tid, int nt) {
  using P = P2<L>;
  for (int e = tid; e < Geo<L>::NTWS; e += nt) {
    const int u = e % P::R1, n = e / P::R1, j = n / (P::R1 - 1), t = n % (P::R1 - 1) + 1;
    float s, c;
    sincospif(-2.0f * (float)((t * (u + P::R1 * j)) & (L - 1)) / (float)L, &s, &c);
    tws[e] = cf32{c, s};
  }
}

// A thread's twiddles: copied into registers (REG) or read from the per-CTA table on use (saves 2*NTW registers).
template <int L, bool REG> struct Twid;
template <int L> struct Twid<L, true> {
  cf32 r[P2<L>::NTW];
  __device__ __forceinline__ void init(const cf32* tws, int u) {
#pragma unroll
    for (int n = 0; n < P2<L>::NTW; ++n) r[n] = tws[n * P2<L>::R1 + u];
  }
  __device__ __forceinline__ cf32 operator()(int n) const { return r[n]; }
};
template <int L> struct Twid<L, false> {
  const cf32* p;
  __device__ __forceinline__ void init(const cf32* tws, int u) { p = tws + u; }
  __device__ __forceinline__ cf32 operator()(int n) const { return p[n * P2<L>::R1]; }
};

struct GroupList {
  uint32_t bitmap[4];   // bit g: 4-column group g holds at least one sampled column
  uint8_t list[128];    // active groups, ascending
  int count;
};

// Call from all threads of the CTA (ends with __syncthreads).  mrow == nullptr: every group is active.
__device__ __forceinline__ int build_group_list(const uint8_t* mrow, int W, GroupList* gl) {
  const int ng = W >> 2;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int running = 0;
    for (int base = 0; base < ng; base += 32) {
      const int g = base + lane;
      bool act = false;
      if (g < ng) act = mrow == nullptr || (mrow[4 * g] | mrow[4 * g + 1] | mrow[4 * g + 2] | mrow[4 * g + 3]) != 0;
      const unsigned bal = __ballot_sync(0xffffffffu, act);
      if (act) gl->list[running + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)g;
      if (lane == 0) gl->bitmap[base >> 5] = bal;
      running += __popc(bal);
    }
    if (lane == 0) gl->count = running;
  }
  __syncthreads();
  return gl->count;
}

template <bool CPLX> struct MapVal;
template <> struct MapVal<false> {
  float v;
  __device__ __forceinline__ void load(const float* re, const float*, size_t i) { v = re[i]; }
  __device__ __forceinline__ cf32 mul(cf32 x) const { return cscale(x, v); }
  __device__ __forceinline__ cf32 mulc(cf32 x) const { return cscale(x, v); }
};
template <> struct MapVal<true> {
  cf32 v;
  __device__ __forceinline__ void load(const float* re, const float* im, size_t i) { v = cf32{re[i], im[i]}; }
  __device__ __forceinline__ cf32 mul(cf32 x) const { return cmul(x, v); }
  __device__ __forceinline__ cf32 mulc(cf32 x) const { return cmulc(x, v); }
};

// ---- forward, rows: coil multiply, transform along W, scatter sampled columns, zero-fill inactive sectors ------
// grid (batch, H / TPC)
// DENSE (no mask): every column survives, so the scratch keeps the NATURAL layout T[c][b][h][k] -- the row kernels
// access it in runs of 8*R1 bytes and the column kernels in 16-column tiles through shared memory, every global
// access a full line -- instead of scattering 8-byte elements into the transposed layout.
template <int L, bool CPLX, bool TWREG, bool DENSE>
__global__ void __launch_bounds__(128) k2_fwd_rows(SenseArgs a) {
  using G = Geo<L>;
  using P = P2<L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tws = reinterpret_cast<cf32*>(smem_raw);
  cf32* xch = tws + G::NTWS;
  __shared__ GroupList gl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = lane % G::TPF, r = warp * G::RPW + lane / G::TPF;
  // batch index fastest: CTAs that run together share the same rows of the coil maps (L2 hits instead of DRAM re-reads)
  const int b = blockIdx.x, h0 = blockIdx.y * G::TPC, h = h0 + r;
  const uint8_t* mrow = a.mask ? a.mask + (size_t)(b % a.mask_frames) * L : nullptr;
  fill_tws<L>(tws, tid, G::NT);
  build_group_list(mrow, L, &gl);
  Twid<L, TWREG> tw;
  tw.init(tws, t);
  cf32* sx = xch + r * P::STRIDE;
  uint32_t keep = 0;   // layout B
#pragma unroll
  for (int i = 0; i < G::E; ++i)
    if (mrow == nullptr || mrow[b_pos<L>(t, i)] != 0) keep |= 1u << i;
  cf32 xq[G::E];
  {
    const cf32* xp = a.in + ((size_t)b * a.H + h) * L + t;
    const float sg = sgn(h + t);   // a_off and b_off are even: the (-1)^(h+w) factor is one sign per thread
#pragma unroll
    for (int q = 0; q < G::E; ++q) xq[q] = cscale(xp[a_off<L>(q)], sg);
  }
  const bool has_maps = a.mre != nullptr;
  const size_t map_img = (size_t)a.H * L;
  const float* mre = a.mre + (size_t)h * L + t;
  const float* mim = a.mim + (size_t)h * L + t;
  MapVal<CPLX> mnext[G::E];
  auto fetch_maps = [&](int c) {
    if (has_maps && c < a.ncoils) {
#pragma unroll
      for (int q = 0; q < G::E; ++q) mnext[q].load(mre, mim, a_off<L>(q));
      mre += map_img;
      mim += map_img;
    }
  };
  fetch_maps(0);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t img_stride = (size_t)a.batch * a.H * L;
  float4* zbase = reinterpret_cast<float4*>(a.out + ((size_t)b * a.H + h0) * L);
  cf32* wsp = DENSE ? a.ws + ((size_t)b * a.H + h) * L + t : a.ws + ((size_t)b * L + t) * a.H + h;
  // this thread's zero-fill pieces are the same for every coil: precompute which of them are inactive
  constexpr int ZP = G::TPC * (L / 2) / G::NT;
  uint32_t zmask = 0;
  if (mrow != nullptr) {
#pragma unroll
    for (int z = 0; z < ZP; ++z) {
      const int grp = ((tid + z * G::NT) % (L / 2)) >> 1;
      if (((gl.bitmap[grp >> 5] >> (grp & 31)) & 1u) == 0u) zmask |= 1u << z;
    }
  }
  for (int c = 0; c < a.ncoils; ++c) {
    cf32 u[G::E];
#pragma unroll
    for (int q = 0; q < G::E; ++q) u[q] = has_maps ? mnext[q].mul(xq[q]) : xq[q];
    fetch_maps(c + 1);
    // zero every inactive sector of this CTA's rows of coil image c (16-byte pieces, fully coalesced)
#pragma unroll
    for (int z = 0; z < ZP; ++z)
      if ((zmask >> z) & 1u) zbase[tid + z * G::NT] = zero4;
    zbase += img_stride / 2;
    __syncwarp();
    a2b_first<L, -1>(u, t, sx);
    __syncwarp();
    a2b_second<L, -1>(u, t, sx, tw);
    if (DENSE) {
#pragma unroll
      for (int i = 0; i < G::E; ++i) wsp[b_off<L>(i)] = u[i];
    } else {
#pragma unroll
      for (int i = 0; i < G::E; ++i)
        if ((keep >> i) & 1u) wsp[(size_t)b_off<L>(i) * a.H] = u[i];
    }
    wsp += img_stride;
  }
}

// ---- forward, columns: transform the active sectors along H and write them (scaled, centred) ------------------
// grid (ncoils * batch, <= W / 16); CTA (img, j) handles active groups 4j .. 4j+3 of its frame's list, then 4(j+gridDim.y) ..
template <int L, bool DENSE>
__global__ void __launch_bounds__(Geo<L>::NT_COLS) k2_fwd_cols(SenseArgs a) {
  using G = Geo<L>;
  using P = P2<L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tws = reinterpret_cast<cf32*>(smem_raw);
  cf32* xch = tws + G::NTWS;
  __shared__ GroupList gl;
  const int tid = threadIdx.x;
  const size_t img = blockIdx.x;
  const int b = (int)(img % a.batch);
  const uint8_t* mrow = a.mask ? a.mask + (size_t)(b % a.mask_frames) * a.W : nullptr;
  const int count = build_group_list(mrow, a.W, &gl);
  if ((int)blockIdx.y * 4 >= count) return;
  fill_tws<L>(tws, tid, G::NT_COLS);
  __syncthreads();   // twiddle table complete
  const int cs = tid / G::TPF, t = tid % G::TPF;
  Twid<L, (L < 512)> tw;
  tw.init(tws, t);
  cf32* sx = xch + cs * P::STRIDE;
  for (int g0 = blockIdx.y * 4; g0 < count; g0 += gridDim.y * 4) {
    const bool gvalid = g0 + (cs >> 2) < count;
    const int k = gvalid ? 4 * gl.list[g0 + (cs >> 2)] + (cs & 3) : 0;
    const bool active = gvalid && (mrow == nullptr || mrow[k] != 0);
    cf32 v[G::E];
    if (DENSE) {
      // natural-layout scratch: 16-column tile -> per-column lines (16-byte loads, one 128-byte run per row)
      for (int idx = tid; idx < 4 * L * 2; idx += G::NT_COLS) {
        const int half = idx & 1, hh = (idx >> 1) % L, gi = (idx >> 1) / L;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g0 + gi < count) p = *reinterpret_cast<const float4*>(a.ws + (img * L + hh) * a.W + 4 * gl.list[g0 + gi] + 2 * half);
        xch[(4 * gi + 2 * half) * P::STRIDE + hh] = cf32{p.x, p.y};
        xch[(4 * gi + 2 * half + 1) * P::STRIDE + hh] = cf32{p.z, p.w};
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < G::E; ++q) v[q] = sx[a_pos<L>(t, q)];
      __syncwarp();
    } else {
      const cf32* wp = a.ws + (img * a.W + k) * L + t;
#pragma unroll
      for (int q = 0; q < G::E; ++q) v[q] = active ? wp[a_off<L>(q)] : cf32{0.f, 0.f};
    }
    a2b_first<L, -1>(v, t, sx);
    __syncwarp();
    a2b_second<L, -1>(v, t, sx, tw);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < G::E; ++i) sx[b_pos<L>(t, i)] = v[i];   // the exchange line doubles as the column's tile line
    __syncthreads();
    // 16-byte pieces: idx = ((gi*L + h)*2 + half)
    for (int idx = tid; idx < 4 * L * 2; idx += G::NT_COLS) {
      const int half = idx & 1, hh = (idx >> 1) % L, gi = (idx >> 1) / L;
      if (g0 + gi >= count) break;
      const int kk = 4 * gl.list[g0 + gi] + 2 * half;
      const cf32 p0 = xch[(4 * gi + 2 * half) * P::STRIDE + hh], p1 = xch[(4 * gi + 2 * half + 1) * P::STRIDE + hh];
      const float s0 = a.scale * sgn(hh + kk);
      *reinterpret_cast<float4*>(a.out + (img * L + hh) * a.W + kk) = make_float4(p0.x * s0, p0.y * s0, -p1.x * s0, -p1.y * s0);
    }
    __syncthreads();   // the tile is reused by the next chunk
  }
}

// ---- adjoint, columns: inverse transform of the active sectors along H into the transposed scratch ------------
template <int L, bool DENSE>
__global__ void __launch_bounds__(Geo<L>::NT_COLS) k2_adj_cols(SenseArgs a) {
  using G = Geo<L>;
  using P = P2<L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tws = reinterpret_cast<cf32*>(smem_raw);
  cf32* xch = tws + G::NTWS;
  __shared__ GroupList gl;
  const int tid = threadIdx.x;
  const size_t img = blockIdx.x;
  const int b = (int)(img % a.batch);
  const uint8_t* mrow = a.mask ? a.mask + (size_t)(b % a.mask_frames) * a.W : nullptr;
  const int count = build_group_list(mrow, a.W, &gl);
  if ((int)blockIdx.y * 4 >= count) return;
  fill_tws<L>(tws, tid, G::NT_COLS);
  __syncthreads();   // twiddle table complete
  const int cs = tid / G::TPF, t = tid % G::TPF;
  Twid<L, (L < 512)> tw;
  tw.init(tws, t);
  cf32* sx = xch + cs * P::STRIDE;
  for (int g0 = blockIdx.y * 4; g0 < count; g0 += gridDim.y * 4) {
    for (int idx = tid; idx < 4 * L * 2; idx += G::NT_COLS) {
      const int half = idx & 1, hh = (idx >> 1) % L, gi = (idx >> 1) / L;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g0 + gi < count) {
        const int kk = 4 * gl.list[g0 + gi] + 2 * half;
        p = *reinterpret_cast<const float4*>(a.in + (img * L + hh) * a.W + kk);
        const float s0 = sgn(hh + kk);
        const float m0 = (mrow == nullptr || mrow[kk] != 0) ? s0 : 0.f, m1 = (mrow == nullptr || mrow[kk + 1] != 0) ? -s0 : 0.f;
        p = make_float4(p.x * m0, p.y * m0, p.z * m1, p.w * m1);
      }
      xch[(4 * gi + 2 * half) * P::STRIDE + hh] = cf32{p.x, p.y};
      xch[(4 * gi + 2 * half + 1) * P::STRIDE + hh] = cf32{p.z, p.w};
    }
    __syncthreads();
    const bool gvalid = g0 + (cs >> 2) < count;
    const int k = gvalid ? 4 * gl.list[g0 + (cs >> 2)] + (cs & 3) : 0;
    const bool active = gvalid && (mrow == nullptr || mrow[k] != 0);
    cf32 v[G::E];
#pragma unroll
    for (int q = 0; q < G::E; ++q) v[q] = sx[a_pos<L>(t, q)];
    __syncwarp();
    a2b_first<L, +1>(v, t, sx);
    __syncwarp();
    a2b_second<L, +1>(v, t, sx, tw);
    if (DENSE) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < G::E; ++i) sx[b_pos<L>(t, i)] = v[i];
      __syncthreads();
      for (int idx = tid; idx < 4 * L * 2; idx += G::NT_COLS) {
        const int half = idx & 1, hh = (idx >> 1) % L, gi = (idx >> 1) / L;
        if (g0 + gi >= count) break;
        const cf32 p0 = xch[(4 * gi + 2 * half) * P::STRIDE + hh], p1 = xch[(4 * gi + 2 * half + 1) * P::STRIDE + hh];
        *reinterpret_cast<float4*>(a.ws + (img * L + hh) * a.W + 4 * gl.list[g0 + gi] + 2 * half) = make_float4(p0.x, p0.y, p1.x, p1.y);
      }
    } else if (active) {
      cf32* wp = a.ws + (img * a.W + k) * L + t;
#pragma unroll
      for (int i = 0; i < G::E; ++i) wp[b_off<L>(i)] = v[i];
    }
    __syncthreads();   // the tile is refilled by the next chunk
  }
}

// ---- adjoint, rows: gather sampled columns, inverse transform along W, conj-coil sum (or SSOS) ----------------
// grid (batch, H / TPC)
template <int L, bool CPLX, bool TWREG, bool DENSE>
__global__ void __launch_bounds__(128, (L <= 256 ? 4 : 2)) k2_adj_rows(SenseArgs a) {
  using G = Geo<L>;
  using P = P2<L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tws = reinterpret_cast<cf32*>(smem_raw);
  cf32* xch = tws + G::NTWS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = lane % G::TPF, r = warp * G::RPW + lane / G::TPF;
  const int b = blockIdx.x, h = blockIdx.y * G::TPC + r;
  const uint8_t* mrow = a.mask ? a.mask + (size_t)(b % a.mask_frames) * L : nullptr;
  fill_tws<L>(tws, tid, G::NT);
  cf32* sx = xch + r * P::STRIDE;
  uint32_t keep = 0;   // layout A
#pragma unroll
  for (int q = 0; q < G::E; ++q)
    if (mrow == nullptr || mrow[a_pos<L>(t, q)] != 0) keep |= 1u << q;
  const size_t img_stride = (size_t)a.batch * a.H * L;
  const cf32* wsp = DENSE ? a.ws + ((size_t)b * a.H + h) * L + t : a.ws + ((size_t)b * L + t) * a.H + h;
  const bool has_maps = a.mre != nullptr && !a.ssos;
  const size_t map_img = (size_t)a.H * L;
  const float* mre = a.mre + (size_t)h * L + t;
  const float* mim = a.mim + (size_t)h * L + t;
  const float sct = a.scale * sgn(h + t);
  cf32 acc[G::E], nxt[G::E];
#pragma unroll
  for (int i = 0; i < G::E; ++i) acc[i] = cf32{0.f, 0.f};
  auto gather = [&](int c) {
#pragma unroll
    for (int q = 0; q < G::E; ++q) {
      nxt[q] = cf32{0.f, 0.f};
      if (DENSE) {
        if (c < a.ncoils) nxt[q] = wsp[a_off<L>(q)];
      } else if (c < a.ncoils && ((keep >> q) & 1u)) {
        nxt[q] = wsp[(size_t)a_off<L>(q) * a.H];
      }
    }
    wsp += img_stride;
  };
  gather(0);
  __syncthreads();   // twiddle table complete
  Twid<L, TWREG> tw;
  tw.init(tws, t);
  for (int c = 0; c < a.ncoils; ++c) {
    cf32 v[G::E];
    MapVal<CPLX> m[G::E];
#pragma unroll
    for (int q = 0; q < G::E; ++q) v[q] = nxt[q];
    if (has_maps) {   // layout-B positions; in flight during the transform
#pragma unroll
      for (int i = 0; i < G::E; ++i) m[i].load(mre, mim, b_off<L>(i));
      mre += map_img;
      mim += map_img;
    }
    gather(c + 1);
    __syncwarp();
    a2b_first<L, +1>(v, t, sx);
    __syncwarp();
    a2b_second<L, +1>(v, t, sx, tw);
#pragma unroll
    for (int i = 0; i < G::E; ++i) {
      if (a.ssos) {
        acc[i].x += v[i].x * v[i].x + v[i].y * v[i].y;
      } else if (has_maps) {
        acc[i] = cadd(acc[i], m[i].mulc(v[i]));
      } else {
        acc[i] = cadd(acc[i], v[i]);
      }
    }
  }
  if (a.ssos) {
    float* op = reinterpret_cast<float*>(a.out) + ((size_t)b * a.H + h) * L + t;
#pragma unroll
    for (int i = 0; i < G::E; ++i) op[b_off<L>(i)] = sqrtf(acc[i].x) * fabsf(sct);
  } else {
    cf32* op = a.out + ((size_t)b * a.H + h) * L + t;
#pragma unroll
    for (int i = 0; i < G::E; ++i) op[b_off<L>(i)] = cscale(acc[i], sct);
  }
}

// ---- fused Langevin update + SENSE L2-penalty step (row-only).  grid (H / TPC, batch) --------------------------
template <int L, bool CPLX, bool TWREG>
__global__ void __launch_bounds__(128, (L <= 256 ? 3 : 2)) k2_ald_sense(AldArgs a) {
  using G = Geo<L>;
  using P = P2<L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tws = reinterpret_cast<cf32*>(smem_raw);
  cf32* xch = tws + G::NTWS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = lane % G::TPF, r = warp * G::RPW + lane / G::TPF;
  const int b = blockIdx.x, h = blockIdx.y * G::TPC + r;
  ipdm_ald_scalars sc = a.sc;
  uint32_t rstep = a.rng.step;
  if (a.sched != nullptr) {
    const int cur = *a.cursor;
    sc = a.sched[cur];
    rstep += (uint32_t)cur;
  }
  fill_tws<L>(tws, tid, G::NT);
  cf32* sx = xch + r * P::STRIDE;
  const uint8_t* mrow = a.mask ? a.mask + (size_t)(b % a.mask_frames) * L : nullptr;
  // The (-1)^w factors of the centred transforms are a half-period shift of the spectrum and their (-1)^k
  // partners cancel in A^H A, so the plain spectrum is masked at k ^ (L/2) and no sign flips are needed.
  uint32_t keep = 0;   // layout B
#pragma unroll
  for (int i = 0; i < G::E; ++i)
    if (mrow == nullptr || mrow[b_pos<L>(t, i) ^ (L / 2)] != 0) keep |= 1u << i;
  const size_t plane = (size_t)a.batch * a.H * L;
  const size_t rowoff = ((size_t)b * a.H + h) * L + t;
  const size_t map_img = (size_t)a.H * L;
  const float* mre = a.mre + (size_t)h * L + t;
  const float* mim = a.mim + (size_t)h * L + t;
  MapVal<CPLX> mnext[G::E];
  auto fetch_maps = [&](int c) {
    if (c < a.ncoils) {
#pragma unroll
      for (int q = 0; q < G::E; ++q) mnext[q].load(mre, mim, a_off<L>(q));
      mre += map_img;
      mim += map_img;
    }
  };
  fetch_maps(0);
  cf32 z[G::E], acc[G::E];
  {
    const float *xr = a.x + rowoff, *xi = a.x + plane + rowoff, *gr = a.grad + rowoff, *gi = a.grad + plane + rowoff;
    // all loads first (64 independent requests per thread in flight), the noise arithmetic afterwards
#pragma unroll
    for (int q = 0; q < G::E; ++q) {
      z[q].x = xr[a_off<L>(q)] + sc.step * gr[a_off<L>(q)];
      z[q].y = xi[a_off<L>(q)] + sc.step * gi[a_off<L>(q)];
      acc[q] = cf32{0.f, 0.f};
    }
    if (a.noise != nullptr) {
      const float *nr = a.noise + rowoff, *ni = a.noise + plane + rowoff;
#pragma unroll
      for (int q = 0; q < G::E; ++q) {
        z[q].x += sc.noise_scale * nr[a_off<L>(q)];
        z[q].y += sc.noise_scale * ni[a_off<L>(q)];
      }
    } else if (sc.noise_scale != 0.f) {
      const uint64_t seed = rng_seed(a.rng);
      const uint32_t chain = rng_chain(a.rng, b);
#pragma unroll
      for (int q = 0; q < G::E / 2; ++q) {   // pixels w and w + W/2 share one Philox call, as in every kernel family
        float n[4];
        philox_chain_normal4(seed, chain, (uint32_t)(h * L + a_off<L>(q) + t), rstep, n);
        z[q].x += sc.noise_scale * n[0];
        z[q].y += sc.noise_scale * n[1];
        z[q + G::E / 2].x += sc.noise_scale * n[2];
        z[q + G::E / 2].y += sc.noise_scale * n[3];
      }
    }
  }
  __syncthreads();   // twiddle table complete
  Twid<L, TWREG> tw;
  tw.init(tws, t);
  for (int c = 0; c < a.ncoils; ++c) {
    MapVal<CPLX> m[G::E];
    cf32 u[G::E];
#pragma unroll
    for (int q = 0; q < G::E; ++q) {
      m[q] = mnext[q];
      u[q] = m[q].mul(z[q]);
    }
    fetch_maps(c + 1);
    __syncwarp();
    a2b_first<L, -1>(u, t, sx);
    __syncwarp();
    a2b_second<L, -1>(u, t, sx, tw);
#pragma unroll
    for (int i = 0; i < G::E; ++i)
      if (((keep >> i) & 1u) == 0u) u[i] = cf32{0.f, 0.f};
    __syncwarp();
    b2a_first<L, +1>(u, t, sx, tw);
    __syncwarp();
    b2a_second<L, +1>(u, t, sx);
#pragma unroll
    for (int q = 0; q < G::E; ++q) acc[q] = cadd(acc[q], m[q].mulc(u[q]));
  }
  const float ks = sc.kappa / (float)L;
  {
    float *xr = a.x + rowoff, *xi = a.x + plane + rowoff;
    const float *br = a.bvec + rowoff, *bi = a.bvec + plane + rowoff;
    cf32 bv[G::E];   // every b load is issued before the first store (x and b may alias as far as the compiler knows)
#pragma unroll
    for (int q = 0; q < G::E; ++q) bv[q] = cf32{br[a_off<L>(q)], bi[a_off<L>(q)]};
#pragma unroll
    for (int q = 0; q < G::E; ++q) {
      xr[a_off<L>(q)] = z[q].x - ks * acc[q].x + sc.kappa * bv[q].x;
      xi[a_off<L>(q)] = z[q].y - ks * acc[q].y + sc.kappa * bv[q].y;
    }
  }
}

}  // namespace ipdm
