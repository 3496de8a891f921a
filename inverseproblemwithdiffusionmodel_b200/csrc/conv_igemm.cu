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

__device__ unsigned int g_igemm_timeout = 0;

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

__global__ void __launch_bounds__(NTHREADS, 2)
k_conv_igemm(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, IgemmParams p) {
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
    // ===== epilogue: thread = output channel, columns = the 256 pixels of the tile =====
    const int quad = warp & 3;
    const int co = m0 + quad * 32 + lane;
    const bool pool = (p.flags & IPDM_CONV_POOL2) != 0;
    const int Ho = pool ? p.H / 2 : p.H, Wo = pool ? p.W / 2 : p.W;
    const float bias_v = p.bias ? p.bias[co] : 0.f;
    float s1 = 0.f, s2 = 0.f;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
      float v[32];
      tmem_ld32(taddr + chunk * 32, v);
      const int py = 2 * chunk;  // two pixel rows of the tile per chunk
      if (!pool) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int y = h0 + py + (j >> 4), x = w0 + (j & 15);
          if (y < p.H && x < p.W) {
            const size_t o = (((size_t)n * p.H + y) * p.W + x) * p.Cout + co;
            float val = v[j] + bias_v;
            const float pre = val;
            if (p.residual) {
              float r = p.residual[o];
              if (p.flags & IPDM_CONV_RES_ELU) r = elu1(r);
              val += r;
            }
            if (p.out_f32) p.out_f32[o] = val;
            if (p.out_f16) {
              float s = (p.flags & IPDM_CONV_F16_PRE_RES) ? pre : val;
              if (p.flags & IPDM_CONV_F16_ELU) s = elu1(s);
              p.out_f16[o] = __float2half_rn(s);
            }
            s1 += val;
            s2 += val * val;
          }
        }
      } else {
        const int y = (h0 + py) >> 1;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int x = (w0 >> 1) + jj;
          if (y < Ho && x < Wo) {
            const size_t o = (((size_t)n * Ho + y) * Wo + x) * p.Cout + co;
            float val = (((v[2 * jj] + v[16 + 2 * jj]) + v[2 * jj + 1]) + v[16 + 2 * jj + 1]) * 0.25f + bias_v;
            const float pre = val;
            if (p.residual) {
              float r = p.residual[o];
              if (p.flags & IPDM_CONV_RES_ELU) r = elu1(r);
              val += r;
            }
            if (p.out_f32) p.out_f32[o] = val;
            if (p.out_f16) {
              float s = (p.flags & IPDM_CONV_F16_PRE_RES) ? pre : val;
              if (p.flags & IPDM_CONV_F16_ELU) s = elu1(s);
              p.out_f16[o] = __float2half_rn(s);
            }
            s1 += val;
            s2 += val * val;
          }
        }
      }
    }
    if (p.stats) {
      atomicAdd(&p.stats[((size_t)n * p.Cout + co) * 2], s1);
      atomicAdd(&p.stats[((size_t)n * p.Cout + co) * 2 + 1], s2);
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
  static bool attr_set = false;
  if (!attr_set) {
    IPDM_CUDA(cudaFuncSetAttribute(k_conv_igemm, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
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
  k_conv_igemm<<<grid, NTHREADS, SMEM_BYTES, s>>>(mw, mx, p);
  return launched("k_conv_igemm");
}
