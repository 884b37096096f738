// Weight-gradient GEMM of the Glow training step on tcgen05 (sm_100a):
//
//     out[i][n] += sum_p A[p][i] * B[p][n]        A [P, 512] bf16,  B [P, ldb] bf16,  out [512, ldo] fp32
//
// (conv2: A = relu(p1), B = dL/dp2;  conv3: A = relu(p2), B = im2col of dL/dr;  conv1: A = dL/dp1, B = im2col of
// the coupling input -- SURVEY.md App. A "backward").  The reduction runs over pixels, so both operands are
// MN-major for the tensor core (the contiguous index is the output row / column, not K): TMA boxes of
// [64 channels x 64 pixels] land as 128-byte rows with the SWIZZLE_128B pattern, which is exactly the canonical
// MN-major UMMA layout (8 pixel rows x 128 B atoms, SBO = 1024 B, LBO = one box).  One CTA owns a 128 x n_mma
// output tile and a slice of the pixel range (split-K over the grid), accumulates in TMEM and adds its partial
// result to `out` with fp32 reductions.
#include "wgrad_tc.h"

#include <cuda.h>

#include "tc_ptx.cuh"

namespace asep {

namespace {

constexpr int kKT = 64;                 // pixels per pipeline stage
constexpr int kBox = kKT * 128;         // one [64 ch x 64 px] bf16 box = 8 KB
constexpr int kThreads = 192;

struct WgradParams {
  float* out;
  int ldo, n_valid, n_mma, nboxes_b;
  int k_tiles_total, k_tiles_per_split;
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

__global__ void __launch_bounds__(kThreads, 1) k_wgrad_tc(const __grid_constant__ CUtensorMap mapA,
                                                          const __grid_constant__ CUtensorMap mapB, const WgradParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int S = prm.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * prm.stage_bytes);
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]), acc_ready = smem_u32(&bars[2 * S]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[2 * S + 1]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * 128, n0 = blockIdx.y * prm.n_mma;
  const int kt0 = blockIdx.z * prm.k_tiles_per_split;
  const int kt1 = min(prm.k_tiles_total, kt0 + prm.k_tiles_per_split);
  const int nk = kt1 - kt0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    mbar_init(acc_ready, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
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
          tma_load_2d(sa, &mapA, i0, kt * kKT, fb);
          tma_load_2d(sa + kBox, &mapA, i0 + 64, kt * kKT, fb);
          for (int b = 0; b < prm.nboxes_b; ++b) tma_load_2d(sa + (2 + b) * kBox, &mapB, n0 + 64 * b, kt * kKT, fb);
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
      float* orow = prm.out + (size_t)(i0 + row) * prm.ldo + n0;
      for (int j = 0; j < prm.n_mma / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(t_lane + (uint32_t)(j * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (n0 + j * 32 + c < prm.n_valid) atomicAdd(orow + j * 32 + c, __uint_as_float(v[c]));
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
EncodeTiledFn encode_fn2() {
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

CUtensorMap map_2d(const __nv_bfloat16* x, long long rows, int ld) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)kKT};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn2()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ASEP_CHECK(r == CUDA_SUCCESS, ASEP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%lld, %d]", (int)r, rows, ld);
  return m;
}

int g_sms2 = 0;

}  // namespace

void wgrad_tc(const __nv_bfloat16* A, const __nv_bfloat16* B, int ldb, int n_valid, float* out, int ldo, long long P,
              cudaStream_t s) {
  if (P == 0) return;
  ASEP_CHECK(ldb % 64 == 0 && n_valid <= ldb, ASEP_ERR_BAD_ARG, "wgrad_tc: ldb must be a multiple of 64 and >= n_valid");
  if (g_sms2 == 0) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&g_sms2, cudaDevAttrMultiProcessorCount, dev));
  }
  WgradParams prm{};
  prm.out = out; prm.ldo = ldo; prm.n_valid = n_valid;
  prm.n_mma = std::min(256, ldb);
  if (ldb > 256) ASEP_CHECK(ldb % 256 == 0, ASEP_ERR_BAD_ARG, "wgrad_tc: ldb > 256 must be a multiple of 256");
  prm.nboxes_b = prm.n_mma / 64;
  const int ntiles = (ldb + prm.n_mma - 1) / prm.n_mma;
  prm.k_tiles_total = (int)((P + kKT - 1) / kKT);
  int splits = std::max(1, g_sms2 / (4 * ntiles));
  splits = std::min(splits, prm.k_tiles_total);
  prm.k_tiles_per_split = (prm.k_tiles_total + splits - 1) / splits;
  splits = (prm.k_tiles_total + prm.k_tiles_per_split - 1) / prm.k_tiles_per_split;
  prm.stage_bytes = (2 + prm.nboxes_b) * kBox;
  prm.stages = std::min(6, (227 * 1024 - 1024) / prm.stage_bytes);
  const int smem_bytes = prm.stages * prm.stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const CUtensorMap mA = map_2d(A, P, 512);
  const CUtensorMap mB = map_2d(B, P, ldb);
  dim3 grid(4, ntiles, splits);
  k_wgrad_tc<<<grid, kThreads, smem_bytes, s>>>(mA, mB, prm);
  ASEP_LAUNCH_CHECK();
}

}  // namespace asep
