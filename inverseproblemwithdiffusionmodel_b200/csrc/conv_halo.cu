// 3x3 (dilation 1 or 2) convolution as a PERSISTENT tcgen05 implicit GEMM with halo-tile reuse.
//
// The per-tap kernel (conv_igemm.cu) re-fetches the 256-pixel activation tile for each of the 9 taps:
// 864 KB of L2->SM traffic per (256 pixel x 128 channel) accumulator at Cin = 128, which made it
// L2-bandwidth bound (~9 TB/s measured).  Here the activation tile is fetched ONCE per 64-channel chunk
// together with its halo, and the nine taps are nine shared-memory *views* of that one buffer:
//
//   pixel tile = 32 rows x 8 columns (UMMA N = 256: 32 groups of 8 pixels, one group per tile row)
//   halo buffer = (32 + 2d) rows x (8 + 2d) pixel slots x 64 ch f16, row pitch (8 + 2d) * 128 B, filled by one
//                 4-D TMA box {64, 8+2d, 32+2d, 1}
//   tap (ky,kx) = UMMA smem descriptor with start = halo + ky*d*pitch + kx*d*128, stride between 8-row
//                 groups (SBO) = pitch (the SWIZZLE_128B XOR follows the absolute smem address, so neither the
//                 pitch nor a tap's start needs to sit on a 1024-byte pattern boundary: measured)
//
// so L2->SM traffic per accumulator drops to 2 x 43 KB (activations) + 288 KB (weights).  Weights stream
// through their own 3-slot ring (one 128 x 64 tile per tap).  The CTA is persistent (one per SM): TMEM
// holds two 256-column accumulators so the epilogue of item i overlaps the MMAs of item i+1, and the TMA
// producers run ahead across item boundaries.
//
// Warps (384 threads = 3 warpgroups): 0 = halo TMA, 1 = MMA issuer + TMEM owner, 2 = weight TMA, 3 = idle; 4-7 = epilogue
// of the first half of the accumulator's 32-column chunks, 8-11 = of the second half (warp % 4 = TMEM lane quadrant);
// the epilogue warps are independent of each other (no shared memory, no barrier: conv_common.cuh).  The launch gives
// every thread 168 registers; warpgroup 0 hands most of its share to the epilogue warpgroups (`setmaxnreg` 64 / 216), whose
// residual variant keeps two 32-value residual tiles, the accumulator chunk and its addresses live at once.
#include <cstdlib>
#include "conv_common.cuh"

namespace ipdm {

constexpr int HT_W = 8;                       // pixel tile = TH rows x 8 columns, TH = 32 (UMMA N = 256), 24 (N = 192) or
                                              // 12 (N = 96): the short tiles fit the 24 x 8 / 12 x 8 slices of the 3-D network
constexpr int NH = 2;                         // halo ring
constexpr int W_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int HALO_THREADS = 384;            // three warpgroups: 0 = producers + MMA issuer, 1 and 2 = epilogue teams
constexpr int HALO_REGS_LOW = 64, HALO_REGS_HIGH = 216;   // setmaxnreg: 64*128 + 216*256 <= 168*384 (the launch allocation)

// NS = images (slices) per work item: 2 for the 12-row tile, so that one streamed weight tile feeds two N = 96 MMAs
// (a 96-pixel tile alone re-streams its 128 channels' whole weight slab: L2 -> SM bound)
template <int DIL, int TH, int NS> struct HaloCfg {
  static constexpr int SLOTS = HT_W + 2 * DIL;                // pixel slots per halo row: exactly the 8 + 2d that are used
  static constexpr int PITCH = SLOTS * 128;                   // 1280 / 1536 B (need not be a multiple of the 1024-byte swizzle pattern)
  static constexpr int ROWS = TH + 2 * DIL;
  static constexpr int TILE_BYTES = (ROWS * PITCH + 1023) / 1024 * 1024;   // every tile starts on a swizzle-pattern boundary
  static constexpr int HALO_BYTES = NS * TILE_BYTES;                        // one stage = the halo tiles of the item's NS images
  static constexpr int FIXED = NH * HALO_BYTES + 1024 /*align*/ + 256 /*barriers*/;   // the epilogue uses no shared memory
  // weight ring: as deep as shared memory allows, 3 to 6 stages of 16 KB (one 128 x 64 tile per tap)
  static constexpr int NW = (232448 - FIXED) / W_BYTES >= 6 ? 6 : (232448 - FIXED) / W_BYTES;
  static_assert(NW >= 3, "weight ring");
  static constexpr int SMEM = FIXED + NW * W_BYTES;
};

struct HaloParams {
  IgemmParams g;
  int items, mtiles;
  int exp_skip_weights;
  int res_prefetch;
  int pdl;
  // Per 64-channel chunk of Cin, the taps that are evaluated (bit ky*3+kx; 0x1FF = all nine).  A convolution whose weight
  // blocks are structurally zero -- the 4x4 stride-2 form of ConvMeanPool on space-to-depth operands keeps 16 of its 36
  // (tap, parity) blocks -- neither streams those weight tiles nor issues their MMAs.  2-D convolutions only.
  uint16_t tap_mask[16];
  int masked;      // some chunk skips taps (host-computed): the dense loops stay exactly as they were otherwise
};

// MASKED: the sparse-3x3 instantiation (tap masks).  A template parameter rather than a run-time branch: the dense kernels
// stay textually what they were (a run-time branch in the single-thread producer / issuer loops cost the dominant launch 2.4 %).
template <int MODE, int DIL, int TH, int NS, bool MASKED = false>
__global__ void __launch_bounds__(HALO_THREADS, 1)
k_conv_halo(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, HaloParams hp) {
  using CFG = HaloCfg<DIL, TH, NS>;
  constexpr int NW = CFG::NW;
  static_assert(NS * TH * HT_W <= BLOCK_N, "accumulator columns");
  constexpr int BN = TH * HT_W;                 // pixels per accumulator (UMMA N)
  constexpr int NCH = BN / 32;                  // 32-column epilogue chunks: team A takes the first (NCH+1)/2
  const IgemmParams& p = hp.g;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* halo = smem;
  unsigned char* wts = smem + NH * CFG::HALO_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wts + NW * W_BYTES);
  uint64_t* halo_full = bars;             // [NH]
  uint64_t* halo_empty = bars + NH;       // [NH]
  uint64_t* w_full = bars + 2 * NH;       // [NW]
  uint64_t* w_empty = w_full + NW;        // [NW]
  uint64_t* acc_full = w_empty + NW;      // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = p.Cin / BLOCK_K;
  const int planes = p.taps == 27 ? 3 : 1;      // 27 taps: a 3x3x3 convolution over slice volumes, K loop over the 3 kx-planes

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int i = 0; i < NH; ++i) { mbar_init(&halo_full[i], 1); mbar_init(&halo_empty[i], 1); }
    for (int i = 0; i < NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch (no-ops in a plain launch): let the NEXT kernel's CTAs be scheduled as ours retire, and
  // keep everything that reads or overwrites the PREVIOUS kernel's tensors (activation TMA, residual, outputs) behind
  // griddepcontrol.wait; the weight ring (static data) and the set-up above run ahead of it.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  auto decode = [&](int item, int& n, int& h0, int& w0, int& m0) {
    const int mt = item % hp.mtiles;
    int tile = item / hp.mtiles;
    const int tw = tile % p.tiles_w; tile /= p.tiles_w;
    const int th = tile % p.tiles_h; tile /= p.tiles_h;
    n = tile * NS; h0 = th * TH; w0 = tw * HT_W; m0 = mt * BLOCK_M;     // first of the item's NS consecutive images
  };

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HALO_REGS_LOW));   // (role code below keeps its indentation)
  if (warp == 0) {
    // ===== halo producer: one TMA box per (item, 64-channel chunk) =====
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      uint32_t cnt = 0;
      for (int item = blockIdx.x; item < hp.items; item += gridDim.x) {
        int n, h0, w0, m0;
        decode(item, n, h0, w0, m0);
        if ((MODE & 1) && hp.res_prefetch) {
          // the residual tile of this item is needed by the epilogue one to two items from now: pull it into L2 now
          // (bulk prefetch, no registers, no shared memory), so that the epilogue's loads see L2 latency, not DRAM latency
          constexpr bool pool = (MODE & 8) != 0;
          const int Ho = pool ? p.H / 2 : p.H, Wo = pool ? p.W / 2 : p.W;
          const int oy0 = pool ? h0 / 2 : h0, ox0 = pool ? w0 / 2 : w0;
          const int rows = min(pool ? TH / 2 : TH, Ho - oy0), cols = min(pool ? HT_W / 2 : HT_W, Wo - ox0);
          for (int sl = 0; sl < NS; ++sl)
            for (int r = 0; r < rows; ++r) {
              constexpr int EB = (MODE & 16) ? 2 : 4;     // bytes per residual element
              const char* base = reinterpret_cast<const char*>(p.residual) + ((((size_t)(n + sl) * Ho + oy0 + r) * Wo + ox0) * p.Cout + m0) * EB;
              if (p.Cout == BLOCK_M) {
                prefetch_l2_bulk(base, cols * BLOCK_M * EB);
              } else {
                for (int c = 0; c < cols; ++c) prefetch_l2_bulk(base + (size_t)c * p.Cout * EB, BLOCK_M * EB);
              }
            }
        }
        for (int kc = 0; kc < kchunks; ++kc) {
          for (int pl = 0; pl < planes; ++pl, ++cnt) {      // 3x3x3: one halo tile per kx-plane, from slice x + (kx-1)*d
            const int s = cnt % NH;
            mbar_wait(&halo_empty[s], ((cnt / NH) & 1) ^ 1);
            mbar_expect_tx(&halo_full[s], NS * CFG::ROWS * CFG::PITCH);
#pragma unroll
            for (int sl = 0; sl < NS; ++sl)
              tma_load_5d(halo + s * CFG::HALO_BYTES + sl * CFG::TILE_BYTES, &tmap_x, &halo_full[s], kc * BLOCK_K, w0 - DIL, h0 - DIL,
                          (n + sl) % p.slices + p.slice_shift + (planes == 3 ? (pl - 1) * DIL : 0), (n + sl) / p.slices);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===== weight producer: one 128 x 64 tile per (item, chunk, tap) =====
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int item = blockIdx.x; item < hp.items; item += gridDim.x) {
        int n, h0, w0, m0;
        decode(item, n, h0, w0, m0);
        for (int kc = 0; kc < kchunks; ++kc) {
          if constexpr (!MASKED) {
            for (int tap = 0; tap < 9 * planes; ++tap, ++cnt) {     // weights [Cout][(kx,) ky, kx taps][Cin]
              const int s = cnt % NW;
              mbar_wait(&w_empty[s], ((cnt / NW) & 1) ^ 1);
              if (hp.exp_skip_weights && cnt >= (uint32_t)NW) {      // TIMING EXPERIMENT ONLY: stale weights, no L2 traffic
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&w_full[s])) : "memory");
                continue;
              }
              mbar_expect_tx(&w_full[s], W_BYTES);
              tma_load_2d(wts + s * W_BYTES, &tmap_w, &w_full[s], tap * p.Cin + kc * BLOCK_K, m0);
            }
          } else {                                                  // sparse 3x3: only the chunk's listed taps
            const uint32_t tmask = hp.tap_mask[kc & 15];
            for (int tap = 0; tap < 9; ++tap) {
              if (!((tmask >> tap) & 1u)) continue;
              const int s = cnt % NW;
              mbar_wait(&w_empty[s], ((cnt / NW) & 1) ^ 1);
              ++cnt;
              mbar_expect_tx(&w_full[s], W_BYTES);
              tma_load_2d(wts + s * W_BYTES, &tmap_w, &w_full[s], tap * p.Cin + kc * BLOCK_K, m0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, BN);
      uint32_t hcnt = 0, wcnt = 0, acnt = 0;
      for (int item = blockIdx.x; item < hp.items; item += gridDim.x, ++acnt) {
        const int as = acnt & 1;
        mbar_wait(&acc_empty[as], ((acnt >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + as * BLOCK_N;
        // one tap = BLOCK_K / UMMA_K MMAs per slice on the halo tile `hbase` with the next weight tile of the ring
        auto issue_tap = [&](uint32_t hbase, int tap, bool accumulate_first) {
          const int ws = wcnt % NW;
          mbar_wait(&w_full[ws], (wcnt / NW) & 1);
          ++wcnt;
          tcgen05_fence_after();
          const int dy = (tap / 3) * DIL, dx = (tap % 3) * DIL;
          const uint64_t adesc = make_smem_desc(smem_u32(wts + ws * W_BYTES));
#pragma unroll
          for (int sl = 0; sl < NS; ++sl) {
            const uint64_t bdesc = make_smem_desc(hbase + sl * CFG::TILE_BYTES + dy * CFG::PITCH + dx * 128, CFG::PITCH);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_f16(tacc + sl * BN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, accumulate_first || k != 0);
          }
          umma_commit(&w_empty[ws]);
        };
        if constexpr (!MASKED) {
          for (int kc = 0; kc < kchunks; ++kc) {
            for (int pl = 0; pl < planes; ++pl, ++hcnt) {
              const int hs = hcnt % NH;
              mbar_wait(&halo_full[hs], (hcnt / NH) & 1);
              const uint32_t hbase = smem_u32(halo + hs * CFG::HALO_BYTES);
              for (int tap = 0; tap < 9; ++tap) issue_tap(hbase, tap, (kc | pl | tap) != 0);
              umma_commit(&halo_empty[hs]);
            }
          }
        } else {           // sparse 3x3 (2-D): the chunk's listed taps only; the item's first MMA overwrites the accumulator
          bool fresh = true;
          for (int kc = 0; kc < kchunks; ++kc, ++hcnt) {
            const uint32_t tmask = hp.tap_mask[kc & 15];
            const int hs = hcnt % NH;
            mbar_wait(&halo_full[hs], (hcnt / NH) & 1);
            const uint32_t hbase = smem_u32(halo + hs * CFG::HALO_BYTES);
            for (int tap = 0; tap < 9; ++tap) {
              if (!((tmask >> tap) & 1u)) continue;
              issue_tap(hbase, tap, !fresh);
              fresh = false;
            }
            umma_commit(&halo_empty[hs]);
          }
        }
        umma_commit(&acc_full[as]);
      }
    }
  }
  } else {
    // ===== epilogue teams (TMEM lane quadrant = warp % 4): A = warps 4-7, B = warps 8-11 =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HALO_REGS_HIGH));
    const int quad = warp & 3;
    const int team = warp >= 8 ? 1 : 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    uint32_t acnt = 0;
    for (int item = blockIdx.x; item < hp.items; item += gridDim.x, ++acnt) {
      int n, h0, w0, m0;
      decode(item, n, h0, w0, m0);
      const int as = acnt & 1;
      // NS = 1: the two warp groups split the chunks of one image; NS = 2: one image (accumulator half) per group
      constexpr int NA = (NCH + 1) / 2;
      conv_epilogue_shfl<MODE, HT_W>(p, tmem_base + as * BLOCK_N + (NS == 2 ? team * BN : 0), quad, lane, NS == 2 ? n + team : n, h0, w0, m0,
                                     NS == 2 ? 0 : (team ? NA : 0), NS == 2 ? NCH : (team ? NCH - NA : NA), [&]() {
        mbar_wait(&acc_full[as], (acnt >> 1) & 1);
        tcgen05_fence_after();
      });
      // all of this warp's TMEM reads are complete (tcgen05.wait::ld inside): hand the accumulator back
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[as])) : "memory");
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

template <int MODE, int DIL, int TH, int NS = 1, bool MASKED = false>
static int launch_variant(const CUtensorMap& mw, const CUtensorMap& mx, const HaloParams& hp, int grid, cudaStream_t s) {
  static std::atomic<unsigned long long> attr_done{0};      // per device, any host thread
  if (device_needs_setup(attr_done)) {
    IPDM_CUDA(cudaFuncSetAttribute(k_conv_halo<MODE, DIL, TH, NS, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloCfg<DIL, TH, NS>::SMEM));
    device_setup_done(attr_done);
  }
  if (hp.pdl) {
    // programmatic dependent launch: this grid's CTAs may start (set-up, weight ring fill) as the previous kernel's CTAs
    // retire; everything that touches the previous kernel's data sits behind griddepcontrol.wait in the kernel
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(HALO_THREADS); cfg.dynamicSmemBytes = HaloCfg<DIL, TH, NS>::SMEM; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    IPDM_CUDA(cudaLaunchKernelEx(&cfg, k_conv_halo<MODE, DIL, TH, NS, MASKED>, mw, mx, hp));
    return 0;
  }
  k_conv_halo<MODE, DIL, TH, NS, MASKED><<<grid, HALO_THREADS, HaloCfg<DIL, TH, NS>::SMEM, s>>>(mw, mx, hp);
  return 0;
}

// tile rows: the candidate that pads the image height least (ties -> the taller tile); pooled outputs need whole 4-row chunks of an even height
static int pick_tile_rows(int H, bool pool) {
  if (pool) return 32;
  int best = 32, waste = (H + 31) / 32 * 32 - H;
  const int cand[2] = {24, 12};
  for (int c : cand) {
    const int w = (H + c - 1) / c * c - H;
    if (w < waste) { best = c; waste = w; }
  }
  return best;
}

// dilation 1 and 2 with every tile height; dilation 4 (halo of 16 slots x TH+8 rows per stage) only fits shared memory
// with the 12-row tile, i.e. for the 12 x 8 slices of the 3-D network's deepest stage
bool conv_halo_supports(const ipdm_conv_desc& d) {
  if (d.taps < 9) return false;
  if (d.dilation <= 2) return true;
  return d.dilation == 4 && pick_tile_rows(d.H, (d.flags & IPDM_CONV_POOL2) != 0) == 12;
}

int launch_conv_halo(const ipdm_conv_desc& d, cudaStream_t s) {
  const bool pool = (d.flags & IPDM_CONV_POOL2) != 0;
  const int th = pick_tile_rows(d.H, pool);
  CUtensorMap mw, mx;
  if (int e = get_weight_map(d.w_f16, d.Cout, d.taps * d.Cin, &mw)) return e;
  if (int e = get_act_map(d.in_f16, d.N, d.H, d.W, d.Cin, HT_W + 2 * d.dilation, th + 2 * d.dilation, d.slices, &mx)) return e;
  if (d.stats) IPDM_CUDA(cudaMemsetAsync(d.stats, 0, (size_t)(d.N / d.slices) * d.Cout * 2 * sizeof(double), s));
  HaloParams hp{};
  IgemmParams& p = hp.g;
  const bool t16 = d.residual_f16 != nullptr || d.out_raw_f16 != nullptr;     // 16-bit residual stream
  p.bias = d.bias; p.out_f16 = reinterpret_cast<__half*>(d.out_f16);
  p.residual = t16 ? reinterpret_cast<const float*>(d.residual_f16) : d.residual;
  p.out_f32 = t16 ? reinterpret_cast<float*>(d.out_raw_f16) : d.out_f32;
  for (int i = 0; i < 16; ++i) hp.tap_mask[i] = (d.taps == 9 && d.tap_mask[i] != 0) ? (uint16_t)(d.tap_mask[i] & 0x1FF) : (uint16_t)0x1FF;
  hp.masked = 0;
  for (int i = 0; i < 16 && i < d.Cin / 64; ++i) hp.masked |= hp.tap_mask[i] != 0x1FF;
  if (d.taps == 9) {
    for (int i = 0; i < 16 && i < d.Cin / 64; ++i)
      IPDM_REQUIRE(hp.tap_mask[i] != 0, IPDM_E_BADARG, "conv: tap_mask[%d] selects no tap", i);
    IPDM_REQUIRE(d.Cin <= 1024 || d.tap_mask[0] == 0, IPDM_E_UNSUPPORTED, "conv: tap masks need Cin <= 1024");
  }
  p.acc_scale = d.acc_scale != 0.f ? d.acc_scale : 1.f;
  p.out16_scale = d.out_f16_scale != 0.f ? d.out_f16_scale : 1.f;
  p.stats = d.stats;
  p.N = d.N; p.H = d.H; p.W = d.W; p.Cin = d.Cin; p.Cout = d.Cout; p.taps = d.taps; p.dilation = d.dilation; p.flags = d.flags;
  p.slices = d.slices; p.slice_shift = d.slice_shift;
  p.tiles_w = (d.W + HT_W - 1) / HT_W;
  p.tiles_h = (d.H + th - 1) / th;
  hp.mtiles = d.Cout / BLOCK_M;
#ifdef IPDM_EXPERIMENTS
  hp.exp_skip_weights = g_conv_variant == 2;      // timing experiment (wrong results): experimental builds only
#else
  hp.exp_skip_weights = 0;
#endif
  hp.res_prefetch = g_conv_res_prefetch;
  static const int env_pdl = getenv("IPDM_CONV_PDL") ? atoi(getenv("IPDM_CONV_PDL")) : -1;   // A/B runs of whole programs
  hp.pdl = env_pdl >= 0 ? env_pdl : g_conv_pdl;
  // two images per work item with the 12-row tile (dilation <= 2; pairs never straddle a volume: slices is even or 1 with even N)
  const bool pair = th == 12 && d.dilation <= 2 && d.N % 2 == 0 && (d.slices == 1 || d.slices % 2 == 0);
  hp.items = p.tiles_w * p.tiles_h * (pair ? d.N / 2 : d.N) * hp.mtiles;
  static std::atomic<int> sm_count[64];                       // per device
  int dev = 0;
  IPDM_CUDA(cudaGetDevice(&dev));
  int sms = sm_count[dev & 63].load(std::memory_order_relaxed);
  if (sms == 0) {
    IPDM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sm_count[dev & 63].store(sms, std::memory_order_relaxed);
  }
  const int grid = hp.items < sms ? hp.items : sms;
  const int mode = (p.residual ? 1 : 0) | (p.out_f32 ? 2 : 0) | (d.out_f16 ? 4 : 0) | (pool ? 8 : 0) | (t16 ? 16 : 0);
  int e = 0;
#define HALO_TH(M, D)                                                                                     \
  (th == 32 ? launch_variant<M, D, 32>(mw, mx, hp, grid, s)                                               \
            : th == 24 ? launch_variant<M, D, 24>(mw, mx, hp, grid, s)                                    \
                       : pair ? launch_variant<M, D, 12, 2>(mw, mx, hp, grid, s) : launch_variant<M, D, 12>(mw, mx, hp, grid, s))
#define HALO_CASE(M)                                                                 \
  case M:                                                                            \
    e = d.dilation == 1 ? HALO_TH(M, 1) : d.dilation == 2 ? HALO_TH(M, 2) : launch_variant<M, 4, 12>(mw, mx, hp, grid, s); \
    break;
#define HALO_CASE_POOL(M)                                                            \
  case M:                                                                            \
    e = d.dilation == 1 ? launch_variant<M, 1, 32>(mw, mx, hp, grid, s) : launch_variant<M, 2, 32>(mw, mx, hp, grid, s); \
    break;
  // sparse 3x3 (ConvMeanPool on space-to-depth operands): its own instantiations for the output modes the score network
  // uses there; anything else runs the dense kernel, which multiplies the zero blocks (same result)
  const bool sparse = hp.masked && d.dilation == 1 && (mode == 3 || mode == 7 || mode == 19 || mode == 23) && (th == 32 || th == 24);
  if (!sparse) hp.masked = 0;
#define HALO_SPARSE(M) (th == 32 ? launch_variant<M, 1, 32, 1, true>(mw, mx, hp, grid, s) : launch_variant<M, 1, 24, 1, true>(mw, mx, hp, grid, s))
  if (sparse) {      // residual + result, with or without the f16 operand copy, on either stream
    e = mode == 23 ? HALO_SPARSE(23) : mode == 19 ? HALO_SPARSE(19) : mode == 7 ? HALO_SPARSE(7) : HALO_SPARSE(3);
  } else
  switch (mode) {
    HALO_CASE(2) HALO_CASE(3) HALO_CASE(4) HALO_CASE(5) HALO_CASE(6) HALO_CASE(7)
    HALO_CASE_POOL(10) HALO_CASE_POOL(11) HALO_CASE_POOL(12) HALO_CASE_POOL(13) HALO_CASE_POOL(14) HALO_CASE_POOL(15)
    // 16-bit residual stream: result (+ residual) (+ operand copy)
    HALO_CASE(18) HALO_CASE(19) HALO_CASE(22) HALO_CASE(23)
    HALO_CASE_POOL(26) HALO_CASE_POOL(27) HALO_CASE_POOL(30) HALO_CASE_POOL(31)
    default:
      set_error("conv_halo: unsupported output combination %d", mode);
      return IPDM_E_BADARG;
  }
#undef HALO_SPARSE
#undef HALO_TH
#undef HALO_CASE_POOL
#undef HALO_CASE
  if (e) return e;
  return launched("k_conv_halo");
}

}  // namespace ipdm
