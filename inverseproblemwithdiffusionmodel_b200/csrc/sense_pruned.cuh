// SENSE kernels for column masks that keep few columns (ns <= 32 of W): pruned row transforms (fftpr.cuh) and a
// compact scratch.  Included by sense.cu; used whenever a plan (ipdm_sense_plan_create) says `pruned`.
//
// * Row kernels never run the second pass of a full transform: per coil a thread does one 16-point DFT in registers,
//   the transform's R1 threads (8, 16 or 32 lanes of ONE warp: __syncwarp only) exchange through a private line of
//   shared memory, and each sampled column is one R1-term sum with a register-resident twiddle vector (forward) or
//   is spread over its residue class (adjoint).  W = 512 takes R1 = 32: 16 values per thread instead of the 32 of the
//   full engine, which is what keeps these kernels near 128 registers.
// * The scratch is compact: T[c][b][chunk][h][8] holds the sampled columns only, eight per chunk (a chunk = the work
//   item of a column kernel: one contiguous block), so a row kernel reads / writes a few 64-byte runs per row and
//   coil, and the column kernels (full two-pass transforms along H of the ns sampled columns only) see a few percent
//   of k-space.  (A first layout T[c][b][h][slot] made the column kernels fetch 64-byte pieces at a 192-byte pitch:
//   2.7x the bytes from DRAM.)
// * Everything that depends on the mask alone -- column lists, residue classes, twiddle vectors, active 32-byte
//   sectors, the zero-fill bitmap, the H-transform twiddles -- comes from the plan; no kernel compiles a mask or calls
//   sincos.
// * The forward row kernel also zero-fills every inactive sector of its rows of the output (those stores overlap
//   the arithmetic); the forward column kernel writes the active sectors whole.
#pragma once
#include "fftpr.cuh"
#include "sense_plan.h"

namespace ipdm {

struct PlanView {
  int frames, W, ns_pad, ng_all, cmax, nch_max;
  const int* ns;
  const int* ngroups;
  const int* nchunks;
  const uint16_t* kcol;
  const uint8_t* nat;       // [frames][ns_pad]  class position -> natural slot
  const uint8_t* k0c;       // [frames][ns_pad]  class position -> k mod 16
  const uint8_t* ppos;      // [frames][ns_pad]  class position -> position in the padded class layout (k0*cmax + e)
  const uint8_t* tcw;       // [frames][ns_pad]  class position -> chunk*8 + position inside the chunk (scratch layout)
  const cf32* tw;           // [frames][ns_pad][W/16], class order
  const cf32* twh;          // [frames][ns_pad][10], class order, factored form
  const uint8_t* groups;    // [frames][W/GW]
  const uint8_t* gslot;     // [frames][W/GW][GW]
  const uint8_t* chunks;    // [frames][W/GW][4] = {first group, groups, first slot, slots}
  const uint32_t* gbitmap;  // [frames][4]
  const uint32_t* big;      // [frames][2]  classes with >= 3 / 4 sampled columns
  const cf32* tws_h;        // layout-B twiddles of the H transform (Geo<H>::NTWS entries)
  const ChunkRec* crec;     // [frames][W/GW] one record per (frame, chunk)
};
constexpr int PLAN_TWH = 10;
constexpr int PLAN_GW = PlanHost::GW;      // columns per output group (see sense_plan.h)

template <int L> struct PGeo {
  using P = PR<L>;
  static constexpr int RPW = 32 / P::R1, WARPS = 4, NT = 128, TPC = RPW * WARPS;   // rows per CTA: 4, 8, 16
  static constexpr int ZP = TPC * (L / 2) / NT;                                    // 16-byte zero-fill pieces per thread and coil
};

// The outputs (forward) / inputs (adjoint) of thread t: class positions jj = t + R1*o, o < NOUT.
// slot[o] = offset of the column inside the image's scratch block T[chunk][h][8] at h = 0 (-1: no column)
template <int L, int NOUT> struct MySlots {
  int k0[NOUT], slot[NOUT], pp[NOUT];
  __device__ __forceinline__ void init(const PlanView& p, int f, int ns, int t, int H) {
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
      const int jj = t + PR<L>::R1 * o;
      const bool on = jj < ns;
      k0[o] = on ? p.k0c[f * p.ns_pad + jj] : 0;
      const int cw = on ? p.tcw[f * p.ns_pad + jj] : 0;
      slot[o] = on ? (cw >> 3) * H * 8 + (cw & 7) : -1;
      pp[o] = on ? p.ppos[f * p.ns_pad + jj] : 0;
    }
  }
};

// Factored twiddle vectors of the thread's outputs; ALT folds (-1)^t in (the fused step works on the un-centred spectrum:
// column k of the mask sits at plain index k ^ (W/2), i.e. w^(t*k) picks up (-1)^t = (-1)^(t&3): w^k and w^3k change sign).
template <int L, int NOUT, bool ALT>
__device__ __forceinline__ void load_my_twiddles(cf32 (&twh)[NOUT][PR<L>::NTWH], const PlanView& p, int f, int ns, int t) {
  using P = PR<L>;
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    const int jj = t + P::R1 * o;
    const cf32* src = p.twh + ((size_t)f * p.ns_pad + (jj < ns ? jj : 0)) * PLAN_TWH;
#pragma unroll
    for (int i = 0; i < P::NTWH; ++i) {
      const cf32 w = src[i];
      twh[o][i] = (ALT && (i == 0 || i == 2)) ? cf32{-w.x, -w.y} : w;
    }
  }
}

// ---- forward, rows: coil multiply, pruned transform along W, compact scratch, zero-fill -------------------------
// grid (batch, H / TPC).  Real coil maps (or none).
template <int L, int NOUT>
__global__ void __launch_bounds__(128, 3) kp_fwd_rows(SenseArgs a, PlanView p) {
  using G = PGeo<L>;
  using P = PR<L>;
  __shared__ __align__(16) cf32 xch[G::TPC * P::LINE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = lane % P::R1, r = warp * G::RPW + lane / P::R1;
  // batch index fastest: CTAs that run together share the same rows of the coil maps (L2 hits instead of DRAM re-reads)
  const int b = a.b0 + blockIdx.x, h0 = blockIdx.y * G::TPC, h = h0 + r;   // images [b0, b0 + nb) of the batch in this launch
  const int f = b % p.frames, ns = p.ns[f];
  cf32* sx = xch + r * P::LINE;
  MySlots<L, NOUT> my;
  my.init(p, f, ns, t, a.H);
  cf32 twh[NOUT][P::NTWH];
  load_my_twiddles<L, NOUT, false>(twh, p, f, ns, t);
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
  float mnext[P::R0];
  auto fetch_maps = [&](int c) {
    if (has_maps && c < a.ncoils) {
#pragma unroll
      for (int q = 0; q < P::R0; ++q) mnext[q] = mre[P::R1 * q];
      mre += map_img;
    }
  };
  fetch_maps(0);
  // this thread's zero-fill pieces are the same for every coil: which of them lie in inactive sectors
  uint32_t zmask = 0;
#pragma unroll
  for (int z = 0; z < G::ZP; ++z) {
    const int grp = ((tid + z * G::NT) % (L / 2)) / (PLAN_GW / 2);      // 16-byte piece -> group
    if (((p.gbitmap[f * 4 + (grp >> 5)] >> (grp & 31)) & 1u) == 0u) zmask |= 1u << z;
  }
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t img_stride = (size_t)a.batch * a.H * L;
  float4* zbase = reinterpret_cast<float4*>(a.out + ((size_t)b * a.H + h0) * L);
  const size_t ws_img = (size_t)p.nch_max * a.H * 8;   // scratch of one coil image: [chunk][h][8]
  cf32* wsp = a.ws + (size_t)b * ws_img + (size_t)h * 8;
  const size_t ws_stride = (size_t)a.batch * ws_img;
  for (int c = 0; c < a.ncoils; ++c) {
    // (re-reading the row per coil instead of keeping it would free 32 registers for a fourth resident CTA, but this
    // kernel is bound by the L1 / shared-memory data pipe -- ncu: 95 % -- and the re-reads land exactly there)
    cf32 u[P::R0];
#pragma unroll
    for (int q = 0; q < P::R0; ++q) u[q] = has_maps ? cscale(xq[q], mnext[q]) : xq[q];
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
      if (my.slot[o] >= 0) wsp[my.slot[o]] = pr_gather<L, -1>(sx, my.k0[o], twh[o]);
    wsp += ws_stride;
  }
}

// ---- column kernels: full two-pass transforms along H of the sampled columns ------------------------------------
// Work item = (coil image, chunk); a chunk = a run of whole active groups holding at most 8 sampled columns (plan table),
// one transform per sampled column, one warp-slice per line.
// These kernels move little data and used to wait on its latency (ncu: long-scoreboard stalls, 4 KB in flight per CTA).
// Every byte of an item is requested at once by asynchronous copies into a row-major staging tile stage[h][slot]
// (pitch 10): the scratch block of a chunk is contiguous ([h][8], 32 KB at H = 512, 16-byte copies), the adjoint's k-space
// side is 8-byte copies of the sampled columns.  The exchange lines of the transforms alias the staging tile once it has
// been read into registers.  The kernels are persistent with two tiles: the copies of the next item fly during the
// transform and the drain of the current one (masked adjoint at 32 coils x 512^2 x 64: 1.11 -> 0.98 ms).
template <int LH>
__device__ __forceinline__ void copy_tws(cf32* tws, const cf32* src, int tid, int nt) {
  for (int e = tid; e < Geo<LH>::NTWS; e += nt) tws[e] = src[e];
}
template <int LH> struct CGeo {
  static constexpr int CL = PlanHost::CHUNK_SLOTS;            // lines (sampled columns) per CTA
  static constexpr int NT = CL * Geo<LH>::TPF;
  static constexpr int SP = CL + 2;                          // staging pitch: rows 16-byte aligned, a column read is 2-way conflicted at worst
  static constexpr int CSTRIDE = P2<LH>::STRIDE | 1;         // line pitch: odd, the drain loops walk 8 lines at one row
  static constexpr int TILE = ((LH * SP > CL * CSTRIDE ? LH * SP : CL * CSTRIDE) + 1) & ~1;   // even: both tiles 16-byte aligned
  static constexpr size_t SMEM = (size_t)(Geo<LH>::NTWS + 2 * TILE) * sizeof(cf32);           // twiddles + two tiles (persistent CTAs)
  static constexpr size_t SMEM1 = (size_t)(Geo<LH>::NTWS + TILE) * sizeof(cf32);              // one item per CTA: one tile
};
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One work item of the column kernels = (coil image, chunk of <= 8 sampled columns).  The kernels are PERSISTENT: CTA i takes
// items i, i + gridDim.x, ... and keeps two tiles, so the (latency-bound) loads of the next item are in flight while the
// current one is transformed and drained.  Items are numbered (coil * nb + image) * nch_max + chunk = scratch order.
// What an item needs to know about its chunk is ONE ChunkRec (sense_plan.h); the records travel two items ahead through a
// ring of three in shared memory (the record of item k+1 addresses the copies issued during item k), so no thread ever
// waits on a table read: the first version chased nchunks -> chunks -> groups -> gslot / kcol per item, 50 % of its samples.
struct ColItem {
  int f, chunk;
  size_t img;
  bool in_range;
};
__device__ __forceinline__ ColItem col_item(const SenseArgs& a, const PlanView& p, int item, int n_items) {
  ColItem it;
  const int bx = item / p.nch_max;
  it.chunk = item - bx * p.nch_max;
  const int b = a.b0 + bx % a.nb;
  it.f = b % p.frames;
  it.img = (size_t)(bx / a.nb) * a.batch + b;
  it.in_range = item < n_items;
  return it;
}
constexpr int REC_PIECES = (int)(sizeof(ChunkRec) / 16);
// asynchronous: the record lands with the next cp.async wait + barrier
__device__ __forceinline__ void fetch_rec(const SenseArgs& a, const PlanView& p, int item, int n_items, ChunkRec* dst, int tid) {
  const ColItem it = col_item(a, p, item, n_items);
  if (it.in_range) {
    if (tid < REC_PIECES)
      cp_async16(reinterpret_cast<char*>(dst) + 16 * tid, reinterpret_cast<const char*>(p.crec + (size_t)it.f * p.ng_all + it.chunk) + 16 * tid);
  } else if (tid == 0) {
    dst->valid = 0;
  }
}

template <int LH>
__global__ void __launch_bounds__(CGeo<LH>::NT) kp_fwd_cols(SenseArgs a, PlanView p) {
  using G = Geo<LH>;
  using C = CGeo<LH>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tws = reinterpret_cast<cf32*>(smem_raw);
  cf32* tiles = tws + G::NTWS;       // per tile: staging [h][SP] first, then the lines [cs][CSTRIDE]
  __shared__ ChunkRec recs[3];
  const int tid = threadIdx.x;
  const int n_items = a.ncoils * a.nb * p.nch_max;
  auto issue = [&](int item, cf32* tile, const ChunkRec& rec) {
    if (item >= n_items || !rec.valid) return;
    const ColItem it = col_item(a, p, item, n_items);
    const cf32* wp = a.ws + (it.img * p.nch_max + it.chunk) * (size_t)(LH * 8);   // this chunk's block [h][8]
    for (int idx = tid; idx < 4 * LH; idx += C::NT) {
      const int pc = idx & 3, hh = idx >> 2;                                      // 16-byte piece pc of row hh
      cp_async16(tile + hh * C::SP + 2 * pc, wp + hh * 8 + 2 * pc);
    }
  };
  fetch_rec(a, p, blockIdx.x, n_items, &recs[0], tid);
  fetch_rec(a, p, blockIdx.x + gridDim.x, n_items, &recs[1], tid);
  copy_tws<LH>(tws, p.tws_h, tid, C::NT);
  cp_async_wait_all();
  __syncthreads();
  issue(blockIdx.x, tiles, recs[0]);
  const int cs = tid / G::TPF, t = tid % G::TPF;
  Twid<LH, (LH < 512)> tw;
  tw.init(tws, t);
  int par = 0, slot = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, par ^= 1, slot = slot == 2 ? 0 : slot + 1) {
    cf32* tile = tiles + par * C::TILE;
    const ChunkRec& rec = recs[slot];
    const ChunkRec& rec_next = recs[slot == 2 ? 0 : slot + 1];
    cp_async_wait_all();
    __syncthreads();      // this item's tile and the next item's record have landed; every thread is done with the previous item
    fetch_rec(a, p, item + 2 * gridDim.x, n_items, &recs[slot == 0 ? 2 : slot - 1], tid);
    if (!rec.valid) {     // CTA-uniform
      issue(item + gridDim.x, tiles + (par ^ 1) * C::TILE, rec_next);
      continue;
    }
    const ColItem it = col_item(a, p, item, n_items);
    const int g_cnt = rec.g_cnt, s_cnt = rec.s_cnt;
    cf32* sx = tile + cs * C::CSTRIDE;
    cf32 v[G::E];
#pragma unroll
    for (int q = 0; q < G::E; ++q) v[q] = cs < s_cnt ? tile[a_pos<LH>(t, q) * C::SP + cs] : cf32{0.f, 0.f};
    __syncthreads();   // the staging tile is in registers: the lines may overwrite it
    issue(item + gridDim.x, tiles + (par ^ 1) * C::TILE, rec_next);
    a2b_first<LH, -1>(v, t, sx);
    __syncwarp();
    a2b_second<LH, -1>(v, t, sx, tw);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < G::E; ++i) sx[b_pos<LH>(t, i)] = v[i];   // the exchange line doubles as the column's tile line
    __syncthreads();
    // Drain: 16-byte pieces of the active groups.  A warp writes RPI image rows per round, lane = (row in round, group,
    // piece); everything that does not depend on the row -- the two source lines, the column -- is fixed per lane up front,
    // so one piece costs two shared loads, four multiplies and the store (the first version divided by g_cnt per piece).
    // Group columns are even, so (-1)^(h + k) is (-1)^h for a piece's first column.
    {
      constexpr int PG = PLAN_GW / 2, NW = C::NT / 32;
      const int lane = tid & 31, warp = tid >> 5;
      const int npc = g_cnt * PG, rpi = 32 / npc;        // g_cnt <= CL = 8 groups -> npc <= 32
      const int rsub = lane / npc, pcs = lane - rsub * npc;
      const int gi = pcs / PG, pc = pcs - gi * PG;
      const bool on = rsub < rpi;
      const int kk = on ? rec.gcol[gi] + 2 * pc : 0;
      const int l0 = on ? rec.gline[gi][2 * pc] : -1, l1 = on ? rec.gline[gi][2 * pc + 1] : -1;
      const cf32* s0 = tile + (l0 >= 0 ? l0 : 0) * C::CSTRIDE;
      const cf32* s1 = tile + (l1 >= 0 ? l1 : 0) * C::CSTRIDE;
      const float m0 = l0 >= 0 ? a.scale : 0.f, m1 = l1 >= 0 ? -a.scale : 0.f;
      cf32* op = a.out + it.img * LH * a.W + kk;
      if (on)
        for (int hh = warp * rpi + rsub; hh < LH; hh += NW * rpi) {
          const cf32 p0 = s0[hh], p1 = s1[hh];
          const float sg = (hh & 1) ? -1.f : 1.f;
          const float c0 = m0 * sg, c1 = m1 * sg;
          *reinterpret_cast<float4*>(op + (size_t)hh * a.W) = make_float4(p0.x * c0, p0.y * c0, p1.x * c1, p1.y * c1);
        }
    }
  }
  cp_async_wait_all();
}

// adjoint, columns: inverse transform of the sampled columns along H into the compact scratch
template <int LH>
__global__ void __launch_bounds__(CGeo<LH>::NT) kp_adj_cols(SenseArgs a, PlanView p) {
  using G = Geo<LH>;
  using C = CGeo<LH>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf32* tws = reinterpret_cast<cf32*>(smem_raw);
  cf32* tiles = tws + G::NTWS;
  __shared__ ChunkRec recs[3];
  const int tid = threadIdx.x;
  const int n_items = a.ncoils * a.nb * p.nch_max;
  // the sampled columns themselves, 8 bytes each (the memory system fetches their sectors either way), as asynchronous
  // copies straight into the staging tile: LH * CL / NT of them per thread in flight, none of them holding a register
  auto issue = [&](int item, cf32* tile, const ChunkRec& rec) {
    if (item >= n_items || !rec.valid) return;
    const int si = tid % C::CL;
    if (si >= rec.s_cnt) return;
    const ColItem it = col_item(a, p, item, n_items);
    const cf32* sp = a.in + it.img * LH * a.W + rec.kcol[si];
    constexpr int RS = C::NT / C::CL;      // rows per round
    for (int h = tid / C::CL; h < LH; h += RS) cp_async8(tile + h * C::SP + si, sp + (size_t)h * a.W);
  };
  fetch_rec(a, p, blockIdx.x, n_items, &recs[0], tid);
  fetch_rec(a, p, blockIdx.x + gridDim.x, n_items, &recs[1], tid);
  copy_tws<LH>(tws, p.tws_h, tid, C::NT);
  cp_async_wait_all();
  __syncthreads();
  issue(blockIdx.x, tiles, recs[0]);
  const int cs = tid / G::TPF, t = tid % G::TPF;
  Twid<LH, (LH < 512)> tw;
  tw.init(tws, t);
  int par = 0, slot = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, par ^= 1, slot = slot == 2 ? 0 : slot + 1) {
    cf32* tile = tiles + par * C::TILE;
    const ChunkRec& rec = recs[slot];
    const ChunkRec& rec_next = recs[slot == 2 ? 0 : slot + 1];
    cp_async_wait_all();
    __syncthreads();      // this item's columns and the next item's record have landed; every thread is done with the previous item
    fetch_rec(a, p, item + 2 * gridDim.x, n_items, &recs[slot == 0 ? 2 : slot - 1], tid);
    if (!rec.valid) {     // CTA-uniform
      issue(item + gridDim.x, tiles + (par ^ 1) * C::TILE, rec_next);
      continue;
    }
    const ColItem it = col_item(a, p, item, n_items);
    const int s_cnt = rec.s_cnt;
    cf32* sx = tile + cs * C::CSTRIDE;
    const int kc = cs < s_cnt ? rec.kcol[cs] : 0;
    cf32 v[G::E];
    {
      const float sg = sgn(t + kc);   // a_off is even: (-1)^(h + k) is one sign per thread
#pragma unroll
      for (int q = 0; q < G::E; ++q) v[q] = cs < s_cnt ? cscale(tile[a_pos<LH>(t, q) * C::SP + cs], sg) : cf32{0.f, 0.f};
    }
    __syncthreads();   // the staging tile is in registers: the lines may overwrite it
    issue(item + gridDim.x, tiles + (par ^ 1) * C::TILE, rec_next);
    a2b_first<LH, +1>(v, t, sx);
    __syncwarp();
    a2b_second<LH, +1>(v, t, sx, tw);
    __syncthreads();   // every exchange is finished: the tile becomes the row-major staging of the results
#pragma unroll
    for (int i = 0; i < G::E; ++i) tile[b_pos<LH>(t, i) * C::SP + cs] = v[i];
    __syncthreads();
    {
      cf32* wp = a.ws + (it.img * p.nch_max + it.chunk) * (size_t)(LH * 8);   // this chunk's block [h][8], 16-byte pieces
      for (int idx = tid; idx < 4 * LH; idx += C::NT) {
        const int pc = idx & 3, hh = idx >> 2;
        *reinterpret_cast<float4*>(wp + hh * 8 + 2 * pc) = *reinterpret_cast<const float4*>(tile + hh * C::SP + 2 * pc);
      }
    }
  }
  cp_async_wait_all();
}

// Zero the padded class tables and install the frame's twiddle vectors at their padded positions (sign: ALT as above).
template <int L, int CMAX, bool ALT>
__device__ __forceinline__ void fill_padded_tables(cf32* twp, cf32* ysm, int ysm_elems, const PlanView& p, int f, int ns, int tid, int nt) {
  using P = PR<L>;
  for (int e = tid; e < 16 * CMAX * P::R1; e += nt) twp[e] = cf32{0.f, 0.f};
  for (int e = tid; e < ysm_elems; e += nt) ysm[e] = cf32{0.f, 0.f};
  __syncthreads();
  for (int e = tid; e < ns * P::R1; e += nt) {
    const int jj = e / P::R1, tt = e % P::R1;
    const cf32 w = p.tw[((size_t)f * p.ns_pad + jj) * P::R1 + tt];
    twp[p.ppos[f * p.ns_pad + jj] * P::R1 + tt] = (ALT && (tt & 1)) ? cf32{-w.x, -w.y} : w;
  }
}

// ---- adjoint, rows: compact scratch -> pruned inverse transform along W -> conj-coil sum (or SSOS) ---------------
// grid (batch, H / TPC).  Real coil maps (or none / SSOS).
template <int L, int NOUT, int CMAX>
__global__ void __launch_bounds__(128, 4) kp_adj_rows(SenseArgs a, PlanView p) {
  using G = PGeo<L>;
  using P = PR<L>;
  __shared__ __align__(16) cf32 twp[16 * CMAX * P::R1];
  __shared__ __align__(16) cf32 ysm[G::TPC * 16 * CMAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = lane % P::R1, r = warp * G::RPW + lane / P::R1;
  const int b = a.b0 + blockIdx.x, h = blockIdx.y * G::TPC + r;
  const int f = b % p.frames, ns = p.ns[f];
  fill_padded_tables<L, CMAX, false>(twp, ysm, G::TPC * 16 * CMAX, p, f, ns, tid, G::NT);
  const uint32_t big2 = p.big[f * 2], big3 = p.big[f * 2 + 1];
  MySlots<L, NOUT> my;
  my.init(p, f, ns, t, a.H);
  cf32* yr = ysm + r * 16 * CMAX;
  const size_t ws_img = (size_t)p.nch_max * a.H * 8;   // scratch of one coil image: [chunk][h][8]
  const cf32* wsp = a.ws + (size_t)b * ws_img + (size_t)h * 8;
  const size_t ws_stride = (size_t)a.batch * ws_img;
  const bool has_maps = a.mre != nullptr && !a.ssos;
  const size_t map_img = (size_t)a.H * L;
  const float* mre = a.mre + (size_t)h * L + t;
  const float sct = a.scale * sgn(h + t);
  cf32 acc[P::R0], ynext[NOUT];
  float mnext[P::R0];
#pragma unroll
  for (int q = 0; q < P::R0; ++q) acc[q] = cf32{0.f, 0.f};
  auto fetch_y = [&](int c) {
    if (c < a.ncoils) {
#pragma unroll
      for (int o = 0; o < NOUT; ++o) ynext[o] = my.slot[o] >= 0 ? wsp[my.slot[o]] : cf32{0.f, 0.f};
      wsp += ws_stride;
    }
  };
  auto fetch_maps = [&](int c) {
    if (has_maps && c < a.ncoils) {
#pragma unroll
      for (int q = 0; q < P::R0; ++q) mnext[q] = mre[P::R1 * q];
      mre += map_img;
    }
  };
  fetch_y(0);
  fetch_maps(0);
  __syncthreads();   // padded tables complete
  for (int c = 0; c < a.ncoils; ++c) {
    __syncwarp();   // the previous coil's sums have read the spectrum line
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
      if (my.slot[o] >= 0) yr[my.pp[o]] = ynext[o];
    fetch_y(c + 1);   // the next coil's spectrum travels during this coil's transform
    __syncwarp();
    cf32 v[P::R0];
    pr_scatter<L, +1, CMAX>(v, t, yr, twp, P::R1, big2, big3);
#pragma unroll
    for (int q = 0; q < P::R0; ++q) {
      if (a.ssos) {
        acc[q].x += v[q].x * v[q].x + v[q].y * v[q].y;
      } else if (has_maps) {
        acc[q].x += mnext[q] * v[q].x;
        acc[q].y += mnext[q] * v[q].y;
      } else {
        acc[q] = cadd(acc[q], v[q]);
      }
    }
    fetch_maps(c + 1);   // ... and its maps during the next one
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
// transform, conj multiply-accumulate -- all on the 16 values a thread holds.  Real coil maps.
template <int L, int NOUT, int CMAX>
__global__ void __launch_bounds__(128, 3) kp_ald_sense(AldArgs a, PlanView p) {
  using G = PGeo<L>;
  using P = PR<L>;
  __shared__ __align__(16) cf32 xch[G::TPC * P::LINE];
  __shared__ __align__(16) cf32 twp[16 * CMAX * P::R1];
  __shared__ __align__(16) cf32 ysm[G::TPC * 16 * CMAX];
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
  fill_padded_tables<L, CMAX, true>(twp, ysm, G::TPC * 16 * CMAX, p, f, ns, tid, G::NT);
  const uint32_t big2 = p.big[f * 2], big3 = p.big[f * 2 + 1];
  MySlots<L, NOUT> my;
  my.init(p, f, ns, t, a.H);
  cf32 twh[NOUT][P::NTWH];
  load_my_twiddles<L, NOUT, true>(twh, p, f, ns, t);
  cf32* sx = xch + r * P::LINE;
  cf32* yr = ysm + r * 16 * CMAX;
  const size_t plane = (size_t)a.batch * a.H * L;
  const size_t rowoff = ((size_t)b * a.H + h) * L + t;
  const size_t map_img = (size_t)a.H * L;
  const float* mre = a.mre + (size_t)h * L + t;
  float mnext[P::R0];
  auto fetch_maps = [&](int c) {
    if (c < a.ncoils) {
#pragma unroll
      for (int q = 0; q < P::R0; ++q) mnext[q] = mre[P::R1 * q];
      mre += map_img;
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
  __syncthreads();   // padded tables complete
  for (int c = 0; c < a.ncoils; ++c) {
    float m[P::R0];
    cf32 u[P::R0];
#pragma unroll
    for (int q = 0; q < P::R0; ++q) {
      m[q] = mnext[q];
      u[q] = cscale(z[q], m[q]);
    }
    fetch_maps(c + 1);
    __syncwarp();   // line and spectrum of the previous coil are consumed
    pr_first<L, -1>(u, t, sx);
    __syncwarp();
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
      if (my.slot[o] >= 0) yr[my.pp[o]] = pr_gather<L, -1>(sx, my.k0[o], twh[o]);
    __syncwarp();
    pr_scatter<L, +1, CMAX>(u, t, yr, twp, P::R1, big2, big3);
#pragma unroll
    for (int q = 0; q < P::R0; ++q) {
      acc[q].x += m[q] * u[q].x;
      acc[q].y += m[q] * u[q].y;
    }
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
