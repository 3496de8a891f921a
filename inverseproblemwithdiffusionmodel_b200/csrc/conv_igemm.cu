// 3x3 (dilated) / 1x1 stride-1 convolution as a tcgen05 implicit GEMM for sm_100a.
//
// GEMM view ("swap-AB"): D[co][pixel] = sum_k Wt[co][k] * X[pixel][k],  k = tap*Cin + ci
//   A operand = weights  f16 [Cout][taps*Cin]      (K-major), UMMA M = 128 output channels
//   B operand = pixels   f16 NHWC activations      (K-major), UMMA N = 256 pixels (16x16 tile)
//   D         = fp32 accumulators in TMEM: lane = output channel, column = pixel of the tile
// Putting the channels on the TMEM lanes makes the epilogue a pure per-thread affair: thread =
// one output channel, so bias is a scalar, the InstanceNorm++ sum / sum-of-squares are two
// registers, 2x2 mean-pooling combines four registers, and every warp-wide global access touches
// 32 consecutive channels of one pixel (one 128-byte line of the NHWC tensor).  It also keeps the
// shared-memory operand traffic at 96 B/clk for Cout = 128 (A 4 KB + B 8 KB per 128-cycle MMA),
// which a 128x128 tile cannot do in cta_group::1.
//
// Zero padding and dilation come for free from TMA: one 4-D tiled tensor map over [N][H][W][C],
// box {64 ch, 16, 16, 1}; tap (ky,kx) is the same box shifted by ((kx-1)*dil, (ky-1)*dil) and
// out-of-bounds elements are zero-filled by the copy engine.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2-5 = epilogue (TMEM lane quadrant = warp_id % 4).  Two CTAs are co-resident per SM
// (2 x 256 TMEM columns, 2 x ~97 KB smem) so one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include <map>
#include <mutex>
#include <tuple>
#include "common.cuh"

namespace ipdm {

constexpr int BLOCK_M = 128;   // output channels per CTA
constexpr int TILE_H = 16, TILE_W = 16;
constexpr int BLOCK_N = TILE_H * TILE_W;  // 256 pixels
constexpr int BLOCK_K = 64;    // f16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 2;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 128 /*barriers*/;
constexpr int TMEM_COLS = 256;
constexpr int NTHREADS = 192;
constexpr int EPI_PITCH = BLOCK_M + 4;   // floats per slab row: +4 keeps float4 alignment and staggers banks

__device__ unsigned int g_igemm_timeout = 0;

// ELU for the epilogue: exp via the SFU (absolute error ~1e-7 near 0, far below the f16 rounding that
// follows); keeps the unrolled epilogue small enough to stay in the instruction cache.
__device__ __forceinline__ float elu_fast(float v) { return v > 0.f ? v : __expf(v) - 1.0f; }

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a wrong descriptor must surface as an error, not as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  atomicAdd(&g_igemm_timeout, 1u);
  __trap();
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor: 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: f16 x f16 -> f32, both operands K-major, M = 128, N = 256.
__device__ __forceinline__ uint32_t make_idesc() {
  uint32_t d = 0;
  d |= 1u << 4;                      // D format f32
  d |= 0u << 7;                      // A format f16
  d |= 0u << 10;                     // B format f16
  d |= (uint32_t)(BLOCK_N >> 3) << 17;
  d |= (uint32_t)(BLOCK_M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct IgemmParams {
  const float* bias;
  const float* residual;
  float* out_f32;
  __half* out_f16;
  float* stats;
  int N, H, W, Cin, Cout, taps, dilation, flags;
  int tiles_w, tiles_h;
};

// MODE bits: 1 = residual, 2 = fp32 output, 4 = f16 output, 8 = 2x2 mean-pool (compile-time so the
// epilogue of each variant stays small; the ELU / pre-residual choices are cheap runtime selects).
template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 2)
k_conv_igemm(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, IgemmParams p) {
  constexpr bool kRes = (MODE & 1) != 0, kOut32 = (MODE & 2) != 0, kOut16 = (MODE & 4) != 0, kPool = (MODE & 8) != 0;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = blockIdx.x;
  const int tw = tile % p.tiles_w; tile /= p.tiles_w;
  const int th = tile % p.tiles_h; tile /= p.tiles_h;
  const int n = tile;
  const int h0 = th * TILE_H, w0 = tw * TILE_W;
  const int m0 = blockIdx.y * BLOCK_M;
  const int kchunks = p.Cin / BLOCK_K;
  const int num_kb = p.taps * kchunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int tap = kb / kchunks, kc = kb % kchunks;
        const int dy = p.taps == 9 ? (tap / 3 - 1) * p.dilation : 0;
        const int dx = p.taps == 9 ? (tap % 3 - 1) * p.dilation : 0;
        unsigned char* sa = smem + s * STAGE_BYTES;
        unsigned char* sb = sa + A_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_2d(sa, &tmap_w, &full_bar[s], tap * p.Cin + kc * BLOCK_K, m0);
        tma_load_4d(sb, &tmap_x, &full_bar[s], kc * BLOCK_K, w0 + dx, h0 + dy, n);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc();
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(sa);
        const uint64_t bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // advance 32 bytes (16 f16) inside the 128-byte swizzle row: +2 in 16-byte units
          umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // ===== epilogue =====
    // TMEM lane = output channel, column = pixel.  Each warp pulls 32 columns (two pixel rows of the
    // tile) for its 32 channels, transposes them through shared memory (the pipeline stages are idle
    // once the accumulator is complete), and the four warps then stream the [pixels][128 ch] slab
    // with 16-byte accesses: thread = 4 consecutive channels of one pixel, so a warp touches one
    // whole 512-byte pixel row of the NHWC tensor per instruction.  Double-buffered slab, one named
    // barrier per chunk.  All residual loads of a chunk are issued before the first dependent use.
    const int quad = warp & 3;
    const int te = quad * 32 + lane;          // 0..127 within the epilogue group
    const int c4 = (te & 31) * 4;             // first of this thread's 4 channels (within the 128)
    const int prow = te >> 5;                 // pixel sub-row 0..3
    constexpr bool pool = kPool;
    const int Ho = pool ? p.H / 2 : p.H, Wo = pool ? p.W / 2 : p.W;
    const int oy0 = pool ? h0 / 2 : h0, ox0 = pool ? w0 / 2 : w0;
    const int tw_out = pool ? TILE_W / 2 : TILE_W;          // output pixels per tile row
    const int npix = pool ? 8 : 32;                         // output pixels per chunk
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bias4 = *reinterpret_cast<const float4*>(p.bias + m0 + c4);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    float* slab = reinterpret_cast<float*>(smem);           // [2][32][EPI_PITCH]
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
      float v[32];
      tmem_ld32(taddr + chunk * 32, v);
      float* buf = slab + (chunk & 1) * (32 * EPI_PITCH);
      if (!pool) {
#pragma unroll
        for (int j = 0; j < 32; ++j) buf[j * EPI_PITCH + te] = v[j];
      } else {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          buf[jj * EPI_PITCH + te] = (((v[2 * jj] + v[16 + 2 * jj]) + v[2 * jj + 1]) + v[16 + 2 * jj + 1]) * 0.25f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // consume: pixel q = prow + 4*i of the chunk, four pixels per (rolled) half
      const int halves = pool ? 1 : 2;
#pragma unroll 1
      for (int half = 0; half < halves; ++half) {
        float4 res[4];
        size_t off[4];
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int q = prow + 4 * (half * 4 + i);
          const int yy = oy0 + (pool ? chunk : 2 * chunk + (q >> 4));
          const int xx = ox0 + (pool ? q : (q & 15));
          ok[i] = q < npix && yy < Ho && xx < Wo;
          off[i] = (((size_t)n * Ho + yy) * Wo + xx) * p.Cout + m0 + c4;
          if (kRes && ok[i]) res[i] = *reinterpret_cast<const float4*>(p.residual + off[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (ok[i]) {
            const int q = prow + 4 * (half * 4 + i);
            float4 a = *reinterpret_cast<const float4*>(buf + q * EPI_PITCH + c4);
            a.x += bias4.x; a.y += bias4.y; a.z += bias4.z; a.w += bias4.w;
            const float4 pre = a;
            if (kRes) {
              float4 r = res[i];
              if (p.flags & IPDM_CONV_RES_ELU) { r.x = elu_fast(r.x); r.y = elu_fast(r.y); r.z = elu_fast(r.z); r.w = elu_fast(r.w); }
              a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
            }
            if (kOut32) *reinterpret_cast<float4*>(p.out_f32 + off[i]) = a;
            if (kOut16) {
              float4 h = (p.flags & IPDM_CONV_F16_PRE_RES) ? pre : a;
              if (p.flags & IPDM_CONV_F16_ELU) { h.x = elu_fast(h.x); h.y = elu_fast(h.y); h.z = elu_fast(h.z); h.w = elu_fast(h.w); }
              __half2 lo = __floats2half2_rn(h.x, h.y), hi = __floats2half2_rn(h.z, h.w);
              uint2 pk;
              pk.x = *reinterpret_cast<unsigned*>(&lo);
              pk.y = *reinterpret_cast<unsigned*>(&hi);
              *reinterpret_cast<uint2*>(p.out_f16 + off[i]) = pk;
            }
            s1[0] += a.x; s1[1] += a.y; s1[2] += a.z; s1[3] += a.w;
            s2[0] += a.x * a.x; s2[1] += a.y * a.y; s2[2] += a.z * a.z; s2[3] += a.w * a.w;
          }
        }
      }
    }
    if (p.stats) {
      // combine the four pixel sub-rows that share a channel group, then 2 atomics per channel
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float* red = slab;                                    // [4][128][2]
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        red[(prow * 128 + c4 + k) * 2] = s1[k];
        red[(prow * 128 + c4 + k) * 2 + 1] = s2[k];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        t1 += red[(r * 128 + te) * 2];
        t2 += red[(r * 128 + te) * 2 + 1];
      }
      atomicAdd(&p.stats[((size_t)n * p.Cout + m0 + te) * 2], t1);
      atomicAdd(&p.stats[((size_t)n * p.Cout + m0 + te) * 2 + 1], t2);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

using MapKey = std::tuple<const void*, long long, long long, long long, long long>;
static std::map<MapKey, CUtensorMap> g_maps;
static std::mutex g_maps_mu;

static int weight_map(const void* w, int Cout, int K, CUtensorMap* out) {
  MapKey key{w, 2, Cout, K, 0};
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return 0; }
  PFN_encodeTiled enc = get_encode();
  IPDM_REQUIRE(enc, IPDM_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {BLOCK_K, BLOCK_M};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IPDM_REQUIRE(r == CUDA_SUCCESS, IPDM_E_DRIVER, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  g_maps[key] = m;
  *out = m;
  return 0;
}

static int act_map(const void* x, int N, int H, int W, int C, CUtensorMap* out) {
  MapKey key{x, 4, ((long long)N << 32) | H, ((long long)W << 32) | C, 0};
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return 0; }
  PFN_encodeTiled enc = get_encode();
  IPDM_REQUIRE(enc, IPDM_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {BLOCK_K, TILE_W, TILE_H, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IPDM_REQUIRE(r == CUDA_SUCCESS, IPDM_E_DRIVER, "cuTensorMapEncodeTiled(activations) failed: %d", (int)r);
  g_maps[key] = m;
  *out = m;
  return 0;
}

}  // namespace ipdm

using namespace ipdm;

extern "C" int ipdm_conv_igemm(const ipdm_conv_desc* dh, void* stream) {
  IPDM_REQUIRE(dh && dh->in_f16 && dh->w_f16, IPDM_E_BADARG, "conv_igemm: null pointer");
  const ipdm_conv_desc d = *dh;
  IPDM_REQUIRE(d.taps == 9 || d.taps == 1, IPDM_E_BADARG, "conv_igemm: taps must be 9 or 1");
  IPDM_REQUIRE(d.Cin % BLOCK_K == 0 && d.Cin >= BLOCK_K, IPDM_E_UNSUPPORTED, "conv_igemm: Cin=%d must be a multiple of 64", d.Cin);
  IPDM_REQUIRE(d.Cout % BLOCK_M == 0, IPDM_E_UNSUPPORTED, "conv_igemm: Cout=%d must be a multiple of 128", d.Cout);
  IPDM_REQUIRE(d.out_f32 || d.out_f16, IPDM_E_BADARG, "conv_igemm: no output");
  IPDM_REQUIRE(d.N >= 1 && d.H >= 1 && d.W >= 1 && d.dilation >= 1, IPDM_E_BADARG, "conv_igemm: bad shape");
  const bool pool = (d.flags & IPDM_CONV_POOL2) != 0;
  IPDM_REQUIRE(!pool || (d.H % 2 == 0 && d.W % 2 == 0), IPDM_E_BADARG, "conv_igemm: pooling needs even H, W");
  IPDM_REQUIRE((reinterpret_cast<uintptr_t>(d.in_f16) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.w_f16) & 15) == 0,
               IPDM_E_BADARG, "conv_igemm: operands must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  CUtensorMap mw, mx;
  if (int e = weight_map(d.w_f16, d.Cout, d.taps * d.Cin, &mw)) return e;
  if (int e = act_map(d.in_f16, d.N, d.H, d.W, d.Cin, &mx)) return e;
  const int mode = (d.residual ? 1 : 0) | (d.out_f32 ? 2 : 0) | (d.out_f16 ? 4 : 0) | (pool ? 8 : 0);
  if (d.stats) {
    IPDM_CUDA(cudaMemsetAsync(d.stats, 0, (size_t)d.N * d.Cout * 2 * sizeof(float), s));
  }
  IgemmParams p{};
  p.bias = d.bias; p.residual = d.residual; p.out_f32 = d.out_f32; p.out_f16 = reinterpret_cast<__half*>(d.out_f16);
  p.stats = d.stats;
  p.N = d.N; p.H = d.H; p.W = d.W; p.Cin = d.Cin; p.Cout = d.Cout; p.taps = d.taps; p.dilation = d.dilation; p.flags = d.flags;
  p.tiles_w = (d.W + TILE_W - 1) / TILE_W;
  p.tiles_h = (d.H + TILE_H - 1) / TILE_H;
  dim3 grid(p.tiles_w * p.tiles_h * d.N, d.Cout / BLOCK_M);
  static bool attr_set[16] = {};
#define IGEMM_CASE(M)                                                                                           \
  case M:                                                                                                       \
    if (!attr_set[M]) {                                                                                         \
      IPDM_CUDA(cudaFuncSetAttribute(k_conv_igemm<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); \
      attr_set[M] = true;                                                                                       \
    }                                                                                                           \
    k_conv_igemm<M><<<grid, NTHREADS, SMEM_BYTES, s>>>(mw, mx, p);                                              \
    break;
  switch (mode) {
    IGEMM_CASE(2) IGEMM_CASE(3) IGEMM_CASE(4) IGEMM_CASE(5) IGEMM_CASE(6) IGEMM_CASE(7)
    IGEMM_CASE(10) IGEMM_CASE(11) IGEMM_CASE(12) IGEMM_CASE(13) IGEMM_CASE(14) IGEMM_CASE(15)
    default:
      set_error("conv_igemm: unsupported output combination %d", mode);
      return IPDM_E_BADARG;
  }
#undef IGEMM_CASE
  return launched("k_conv_igemm");
}
