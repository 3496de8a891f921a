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
#include "conv_common.cuh"

namespace ipdm {

constexpr int TILE_H = 16, TILE_W = 16;       // BLOCK_N = 256 pixels
constexpr int STAGES = 2;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 128 /*barriers*/;
constexpr int TMEM_COLS = 256;
constexpr int NTHREADS = 192;


// MODE bits: 1 = residual, 2 = fp32 output, 4 = f16 output, 8 = 2x2 mean-pool (compile-time so the
// epilogue of each variant stays small; the ELU / pre-residual choices are cheap runtime selects).
template <int MODE>
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
        const int dy = p.taps >= 9 ? ((tap % 9) / 3 - 1) * p.dilation : 0;
        const int dx = p.taps >= 9 ? (tap % 3 - 1) * p.dilation : 0;
        const int dz = p.taps == 27 ? (tap / 9 - 1) * p.dilation : 0;      // 3x3x3: kx-plane = slice offset
        unsigned char* sa = smem + s * STAGE_BYTES;
        unsigned char* sb = sa + A_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_2d(sa, &tmap_w, &full_bar[s], tap * p.Cin + kc * BLOCK_K, m0);
        tma_load_5d(sb, &tmap_x, &full_bar[s], kc * BLOCK_K, w0 + dx, h0 + dy, n % p.slices + p.slice_shift + dz, n / p.slices);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
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
    conv_epilogue_shfl<MODE, TILE_W>(p, tmem_base, warp & 3, lane, n, h0, w0, m0, 0, 8, [&]() {
      mbar_wait(tmem_full_bar, 0);
      tcgen05_fence_after();
    });
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

using MapKey = std::tuple<const void*, long long, long long, long long, long long>;   // device pointers are unique across devices (UVA)
static std::map<MapKey, CUtensorMap> g_maps;
static std::mutex g_maps_mu;

int g_conv_variant = 0;
int g_conv_res_prefetch = 0;
int g_conv_pdl = 0;
extern int g_sense_split;   // sense.cu

int get_weight_map(const void* w, int Cout, int K, CUtensorMap* out) {
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

// 5-D tiled map {C, W, H, X, P} over f16 activations [P][X][H][W][C] (X = slices per volume, 1 for plain NHWC) with
// box {64, box_w, box_h, 1, 1}, zero OOB fill -- in X too, which is what pads a 3-D convolution across slices.
int get_act_map(const void* x, int N, int H, int W, int C, int box_w, int box_h, int slices, CUtensorMap* out) {
  MapKey key{x, 4 + 16 * (long long)slices, ((long long)N << 32) | H, ((long long)W << 32) | C, ((long long)box_w << 32) | box_h};
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return 0; }
  PFN_encodeTiled enc = get_encode();
  IPDM_REQUIRE(enc, IPDM_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)slices, (cuuint64_t)(N / slices)};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)slices * H * W * C * 2};
  cuuint32_t box[5] = {BLOCK_K, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr,
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
  ipdm_conv_desc d = *dh;
  if (d.slices < 1) d.slices = 1;
  IPDM_REQUIRE(d.N % d.slices == 0 && (d.slices > 1 || d.slice_shift == 0) && abs(d.slice_shift) < d.slices + (d.slices == 1),
               IPDM_E_BADARG, "conv_igemm: N=%d slices=%d slice_shift=%d", d.N, d.slices, d.slice_shift);
  IPDM_REQUIRE(d.taps == 9 || d.taps == 1 || (d.taps == 27 && d.slices > 1), IPDM_E_BADARG,
               "conv_igemm: taps must be 9, 1, or 27 (3x3x3 over slice volumes, slices > 1)");
  IPDM_REQUIRE(d.Cin % BLOCK_K == 0 && d.Cin >= BLOCK_K, IPDM_E_UNSUPPORTED, "conv_igemm: Cin=%d must be a multiple of 64", d.Cin);
  IPDM_REQUIRE(d.Cout % BLOCK_M == 0, IPDM_E_UNSUPPORTED, "conv_igemm: Cout=%d must be a multiple of 128", d.Cout);
  const bool t16 = d.residual_f16 != nullptr || d.out_raw_f16 != nullptr;     // 16-bit residual stream
  IPDM_REQUIRE(!t16 || (d.residual == nullptr && d.out_f32 == nullptr), IPDM_E_BADARG,
               "conv_igemm: the residual stream is either f32 (residual / out_f32) or f16 (residual_f16 / out_raw_f16)");
  IPDM_REQUIRE(d.out_f32 || d.out_f16 || d.out_raw_f16, IPDM_E_BADARG, "conv_igemm: no output");
  IPDM_REQUIRE(!d.stats || d.out_f32 || d.out_raw_f16, IPDM_E_BADARG, "conv_igemm: stats need the result output");
  IPDM_REQUIRE(d.N >= 1 && d.H >= 1 && d.W >= 1 && d.dilation >= 1, IPDM_E_BADARG, "conv_igemm: bad shape");
  const bool pool = (d.flags & IPDM_CONV_POOL2) != 0;
  IPDM_REQUIRE(!pool || (d.H % 2 == 0 && d.W % 2 == 0), IPDM_E_BADARG, "conv_igemm: pooling needs even H, W");
  IPDM_REQUIRE((reinterpret_cast<uintptr_t>(d.in_f16) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.w_f16) & 15) == 0,
               IPDM_E_BADARG, "conv_igemm: operands must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  // 3x3 with dilation 1 or 2: persistent halo-tile kernel (each activation tile is fetched once per 64
  // input channels instead of once per tap); everything else: the per-tap tile kernel below.
  if (g_conv_variant != 1 && conv_halo_supports(d)) return launch_conv_halo(d, s);
  CUtensorMap mw, mx;
  if (int e = get_weight_map(d.w_f16, d.Cout, d.taps * d.Cin, &mw)) return e;
  if (int e = get_act_map(d.in_f16, d.N, d.H, d.W, d.Cin, TILE_W, TILE_H, d.slices, &mx)) return e;
  const int mode = ((d.residual || d.residual_f16) ? 1 : 0) | ((d.out_f32 || d.out_raw_f16) ? 2 : 0) | (d.out_f16 ? 4 : 0) | (pool ? 8 : 0) |
                   (t16 ? 16 : 0);
  if (d.stats) {
    IPDM_CUDA(cudaMemsetAsync(d.stats, 0, (size_t)(d.N / d.slices) * d.Cout * 2 * sizeof(double), s));
  }
  IgemmParams p{};
  p.slices = d.slices; p.slice_shift = d.slice_shift;
  p.bias = d.bias; p.out_f16 = reinterpret_cast<__half*>(d.out_f16);
  p.residual = t16 ? reinterpret_cast<const float*>(d.residual_f16) : d.residual;
  p.out_f32 = t16 ? reinterpret_cast<float*>(d.out_raw_f16) : d.out_f32;
  p.acc_scale = d.acc_scale != 0.f ? d.acc_scale : 1.f;
  p.out16_scale = d.out_f16_scale != 0.f ? d.out_f16_scale : 1.f;
  p.stats = d.stats;
  p.N = d.N; p.H = d.H; p.W = d.W; p.Cin = d.Cin; p.Cout = d.Cout; p.taps = d.taps; p.dilation = d.dilation; p.flags = d.flags;
  p.tiles_w = (d.W + TILE_W - 1) / TILE_W;
  p.tiles_h = (d.H + TILE_H - 1) / TILE_H;
  dim3 grid(p.tiles_w * p.tiles_h * d.N, d.Cout / BLOCK_M);
  static std::atomic<unsigned long long> attr_done[32];      // per mode: bit d = set up on device d
#define IGEMM_CASE(M)                                                                                           \
  case M:                                                                                                       \
    if (device_needs_setup(attr_done[M])) {                                                                     \
      IPDM_CUDA(cudaFuncSetAttribute(k_conv_igemm<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); \
      device_setup_done(attr_done[M]);                                                                          \
    }                                                                                                           \
    k_conv_igemm<M><<<grid, NTHREADS, SMEM_BYTES, s>>>(mw, mx, p);                                              \
    break;
  switch (mode) {
    IGEMM_CASE(2) IGEMM_CASE(3) IGEMM_CASE(4) IGEMM_CASE(5) IGEMM_CASE(6) IGEMM_CASE(7)
    IGEMM_CASE(10) IGEMM_CASE(11) IGEMM_CASE(12) IGEMM_CASE(13) IGEMM_CASE(14) IGEMM_CASE(15)
    IGEMM_CASE(18) IGEMM_CASE(19) IGEMM_CASE(22) IGEMM_CASE(23) IGEMM_CASE(26) IGEMM_CASE(27) IGEMM_CASE(30) IGEMM_CASE(31)
    default:
      set_error("conv_igemm: unsupported output combination %d", mode);
      return IPDM_E_BADARG;
  }
#undef IGEMM_CASE
  return launched("k_conv_igemm");
}

extern "C" int ipdm_debug_option(int key, int value) {
  switch (key) {
    case 1:
#ifndef IPDM_EXPERIMENTS
      IPDM_REQUIRE(value == 0 || value == 1, IPDM_E_BADARG, "debug_option(1, %d): experiment modes need a build with -DIPDM_EXPERIMENTS", value);
#endif
      g_conv_variant = value;
      return 0;
    case 2: g_conv_res_prefetch = value; return 0;
    case 3: g_conv_pdl = value; return 0;
    case 5: g_sense_split = value; return 0;
    case 4: IPDM_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value)); return 0;
    default: set_error("debug_option: unknown key %d", key); return IPDM_E_BADARG;
  }
}
