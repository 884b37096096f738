// Weight gradient of the NCSN 'same' stride-1 convolutions (1x1 / 3x3, dilation 1/2/4) on tcgen05 (sm_100a) -- the
// third GEMM of the denoising-score-matching train step (reference: train_ncsn.py:46-54, tape.gradient through
// ncsn/score_network.py:131-159 Conv2D layers):
//
//     dk[tap][ci][co] += sum_p x[p + offset(tap)][ci] * g[p][co]       x [N,H,W,Cin] bf16, g [N,H,W,Cout] bf16
//
// The reduction runs over pixels, so both operands are MN-major for the tensor core: a TMA box of [64 channels x 64
// pixels] lands as 64 rows of 128 B with the SWIZZLE_128B pattern = the canonical MN-major UMMA layout (same scheme as
// wgrad_tc.cu).  The x box is fetched through a 4-D tensor map (C, W, H, N) at the tap's shifted coordinates: rows
// and columns outside the image come back as zeros ('same' padding), so no im2col tensor is materialised.  One CTA owns
// one (tap, 128 input channels, n_mma output channels) tile and a slice of the pixel range (split-K over the grid),
// accumulates in TMEM and adds its partial result to dk with fp32 reductions.
#include <cuda.h>

#include <map>
#include <tuple>

#include "ncsn_train_kernels.h"
#include "tc_ptx.cuh"

namespace asep {

namespace {

constexpr int kKT = 64;                 // pixels per pipeline stage
constexpr int kBox = kKT * 128;         // one [64 ch x 64 px] bf16 box = 8 KB
constexpr int kThreads = 192;

struct CwParams {
  float* dk;
  int Cin, Cout, n_mma, nboxes_b;
  int taps, ksize, dil;
  int H, W, rows_per_kt, kt_per_img;
  int k_tiles_total, k_tiles_per_split, splits;
  int stages, stage_bytes;
};

// MN-major SWIZZLE_128B shared-memory descriptor: LBO = byte distance between 64-element MN blocks, SBO = 1024 B
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) k_conv_wgrad_tc(const __grid_constant__ CUtensorMap mapX,
                                                               const __grid_constant__ CUtensorMap mapG, const CwParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int S = prm.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * prm.stage_bytes);
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]), acc_ready = smem_u32(&bars[2 * S]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[2 * S + 1]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * 128, n0 = blockIdx.y * prm.n_mma;
  const int tap = blockIdx.z / prm.splits, split = blockIdx.z % prm.splits;
  const int kt0 = split * prm.k_tiles_per_split;
  const int kt1 = min(prm.k_tiles_total, kt0 + prm.k_tiles_per_split);
  const int nk = kt1 - kt0;
  const int half = prm.ksize / 2;
  const int dy = (tap / prm.ksize - half) * prm.dil, dx = (tap % prm.ksize - half) * prm.dil;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    mbar_init(acc_ready, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapG);
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        uint32_t stage = 0, phase = 0;
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t fb = full0 + 8 * stage;
          const uint32_t sa = smem_u32(smem + stage * prm.stage_bytes);
          mbar_expect_tx(fb, (uint32_t)((2 + prm.nboxes_b) * kBox));
          const int n = kt / prm.kt_per_img, h0 = (kt % prm.kt_per_img) * prm.rows_per_kt;
          // channels beyond Cin (second M tile of Cin = 192) and shifted rows / columns outside the image are zero-filled
          tma_load_4d(sa, &mapX, i0, dx, h0 + dy, n, fb);
          tma_load_4d(sa + kBox, &mapX, i0 + 64, dx, h0 + dy, n, fb);
          for (int b = 0; b < prm.nboxes_b; ++b) tma_load_2d(sa + (2 + b) * kBox, &mapG, n0 + 64 * b, kt * kKT, fb);
          if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        uint32_t stage = 0, phase = 0;
        const uint32_t idesc = make_idesc_mn(prm.n_mma);
        for (int kt = 0; kt < nk; ++kt) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * prm.stage_bytes);
#pragma unroll
          for (int k = 0; k < kKT / 16; ++k) {
            const uint64_t da = make_desc_mn(sa + k * 2048, kBox);
            const uint64_t db = make_desc_mn(sa + 2 * kBox + k * 2048, kBox);
            umma_bf16(tmem_base, da, db, idesc, (kt | k) != 0);
          }
          umma_commit(empty0 + 8 * stage);
          if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_ready);
      }
    } else {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      mbar_wait(acc_ready, 0);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const bool row_ok = i0 + row < prm.Cin;
      float* orow = prm.dk + ((size_t)tap * prm.Cin + (size_t)(i0 + row)) * prm.Cout + n0;
      for (int j = 0; j < prm.n_mma / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(t_lane + (uint32_t)(j * 32), v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (n0 + j * 32 + c < prm.Cout) atomicAdd(orow + j * 32 + c, __uint_as_float(v[c]));
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn3() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    ASEP_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, ASEP_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

CUtensorMap map_g(const __nv_bfloat16* g, long long rows, int ld) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)kKT};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn3()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(g), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ASEP_CHECK(r == CUDA_SUCCESS, ASEP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%lld, %d]", (int)r, rows, ld);
  return m;
}

CUtensorMap map_x(const __nv_bfloat16* x, int N, int H, int W, int C) {
  CUtensorMap m;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)(kKT / W), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_fn3()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ASEP_CHECK(r == CUDA_SUCCESS, ASEP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d]", (int)r, N, H, W, C);
  return m;
}

int g_sms3 = 0;

}  // namespace

bool conv_wgrad_tc_supported(int Cin, int Cout, int H, int W) {
  if (Cin % 64 != 0 || Cout % 64 != 0 || Cin <= 0 || Cout <= 0) return false;
  if (W <= 0 || W > kKT || kKT % W != 0) return false;
  if (Cout > 256 && (Cout % 2 != 0 || (Cout / 2) % 64 != 0 || Cout / 2 > 256)) return false;
  return H % (kKT / W) == 0;
}

void conv_wgrad_tc(const __nv_bfloat16* x, const __nv_bfloat16* g, float* dk, int N, int H, int W, int Cin, int Cout, int ksize,
                   int dil, cudaStream_t s) {
  if (N == 0) return;
  ASEP_CHECK(conv_wgrad_tc_supported(Cin, Cout, H, W) && (ksize == 1 || ksize == 3), ASEP_ERR_UNSUPPORTED,
             "conv_wgrad_tc: unsupported shape Cin=%d Cout=%d H=%d W=%d k=%d", Cin, Cout, H, W, ksize);
  if (g_sms3 == 0) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&g_sms3, cudaDevAttrMultiProcessorCount, dev));
  }
  CwParams prm{};
  prm.dk = dk; prm.Cin = Cin; prm.Cout = Cout;
  prm.n_mma = Cout <= 256 ? Cout : Cout / 2;
  prm.nboxes_b = prm.n_mma / 64;
  prm.taps = ksize * ksize; prm.ksize = ksize; prm.dil = dil;
  prm.H = H; prm.W = W; prm.rows_per_kt = kKT / W; prm.kt_per_img = H / prm.rows_per_kt;
  const int mtiles = (Cin + 127) / 128, ntiles = Cout / prm.n_mma;
  prm.k_tiles_total = N * prm.kt_per_img;
  int splits = std::max(1, (2 * g_sms3) / (prm.taps * mtiles * ntiles));
  splits = std::min(splits, prm.k_tiles_total);
  prm.k_tiles_per_split = (prm.k_tiles_total + splits - 1) / splits;
  splits = (prm.k_tiles_total + prm.k_tiles_per_split - 1) / prm.k_tiles_per_split;
  prm.splits = splits;
  prm.stage_bytes = (2 + prm.nboxes_b) * kBox;
  prm.stages = std::min(6, (227 * 1024 - 1024) / prm.stage_bytes);
  const int smem_bytes = prm.stages * prm.stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(k_conv_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const CUtensorMap mX = map_x(x, N, H, W, Cin);
  const CUtensorMap mG = map_g(g, (long long)N * H * W, Cout);
  dim3 grid(mtiles, ntiles, prm.taps * splits);
  k_conv_wgrad_tc<<<grid, kThreads, smem_bytes, s>>>(mX, mG, prm);
  ASEP_LAUNCH_CHECK();
}

}  // namespace asep
