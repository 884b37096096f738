// Implicit-GEMM convolution of the NCSN score networks on tcgen05 / TMEM, fed by TMA (sm_100a).
//
//   out[n,h,w,co] = bias[co] + add[n,h,w,co] + sum_{tap,ci} xin[n, h+dy(tap)*dil, w+dx(tap)*dil, ci] * K[tap][ci][co]
//
// GEMM view per CTA tile: M = 128 consecutive pixels of one image (= 128/W full image rows, so the tile is a
// rectangle), N = Cout (128..384, all output channels at once: the activation tile is read once), K = taps*Cin
// walked as (tap, 64-channel panel) steps.
//   * A operand: the bf16 NHWC activation tensor is described by ONE 4-D tensor map (C, W, H, N); the tile of
//     tap (dy,dx) is the box [64 ch, W, 128/W rows, 1 image] whose start coordinate is shifted by (dx*dil,
//     dy*dil): TMA zero-fills everything outside the image, which is exactly Keras 'same' padding, for any
//     dilation, with no im2col buffer and no index arithmetic on the SMs (cp.async.bulk.tensor -> UTMALDG).
//   * B operand: host-pre-swizzled bf16 weight tile images [Cout x 64] per (tap, panel), streamed with bulk
//     copies (UBLKCP) into the same ring stage.
//   * D: fp32 accumulators in TMEM (Cout columns; two buffers when Cout <= 256 so the epilogue of tile i
//     overlaps the MMAs of tile i+1).
// Warp roles (192 threads, persistent, 1 CTA/SM): warp 0 = TMA producer, warp 1 = MMA issuer (one thread,
// tcgen05.mma M=128 N<=256 K=16, SWIZZLE_128B descriptors), warps 2-5 = epilogue (tcgen05.ld -> + bias + residual
// -> fp32 NHWC stores).
#include "conv_tc.h"

#include <cuda.h>
#include <map>
#include <tuple>

#include "tc_ptx.cuh"

namespace asep {

namespace {

constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;      // one 64-channel bf16 panel of 128 pixels
constexpr int kThreads = 192;
constexpr int kMaxStages = 6;
constexpr int kSmemBudget = 227 * 1024 - 1024;

struct ConvParams {
  const __nv_bfloat16* wimg;
  const float* bias;
  const float* add;
  const float* add2;         // optional second tensor summed in the epilogue (CRP: acc + path_1 + conv_2)
  float* out;
  __nv_bfloat16* out_bf16;   // optional bf16 copy of `out` (the next convolution's operand when no norm intervenes)
  double* stats;    // optional [N, Cout, 2]: per-(image, channel) sum and sum of squares of `out` (instance-norm statistics)
  int Cout, kpanels, taps, dil, H, W, rows_per_tile, tiles_per_img, num_units;
  int nunit;        // work units per pixel tile: Cout is processed as `nunit` column groups of `ncols` (<= 256) channels
  int ncols, stages, stage_bytes, b_bytes;
};

constexpr int kEpiBytes = 4 * 32 * 128 + 2048;   // per epilogue warp: a 32 x 32 fp32 transpose tile; + statistics exchange

__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const __grid_constant__ CUtensorMap tmap, const ConvParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* epi = smem + prm.stages * prm.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + kEpiBytes);
  // bars: [0,S) full  [S,2S) empty  [2S,2S+2) acc_ready  [2S+2,2S+4) acc_empty  [2S+4] tmem slot
  const int S = prm.stages;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]);
  const uint32_t accr0 = smem_u32(&bars[2 * S]), acce0 = smem_u32(&bars[2 * S + 2]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[2 * S + 4]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(accr0 + 8 * i, 1); mbar_init(acce0 + 8 * i, 4); }
    fence_barrier_init();
    tma_prefetch_desc(&tmap);
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kiters = prm.taps * prm.kpanels;

  // A work unit = (128-pixel tile, column group).  Consecutive units share the activation tile (L2 locality); the
  // two 256-column TMEM buffers alternate, so the epilogue of unit i overlaps the MMAs of unit i+1.
  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int unit = blockIdx.x; unit < prm.num_units; unit += gridDim.x) {
        const int tile = unit / prm.nunit, nh = unit - tile * prm.nunit;
        const int n = tile / prm.tiles_per_img;
        const int h0 = (tile % prm.tiles_per_img) * prm.rows_per_tile;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(prm.wimg) + (size_t)nh * prm.b_bytes;
        for (int tap = 0; tap < prm.taps; ++tap) {
          const int dy = prm.taps == 9 ? (tap / 3 - 1) * prm.dil : 0;
          const int dx = prm.taps == 9 ? (tap % 3 - 1) * prm.dil : 0;
          for (int kp = 0; kp < prm.kpanels; ++kp) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t fb = full0 + 8 * stage;
            const uint32_t sa = smem_u32(smem + stage * prm.stage_bytes);
            mbar_expect_tx(fb, (uint32_t)(kABytes + prm.b_bytes));
            tma_load_4d(sa, &tmap, kp * 64, dx, h0 + dy, n, fb);
            bulk_g2s(sa + kABytes, wsrc, (uint32_t)prm.b_bytes, fb);
            wsrc += (size_t)prm.Cout * 128;
            if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t idesc = make_idesc(prm.ncols);
      int it = 0;
      for (int unit = blockIdx.x; unit < prm.num_units; unit += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(acce0 + 8 * buf, ((uint32_t)(it >> 1) & 1u) ^ 1u);       // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
        for (int ki = 0; ki < kiters; ++ki) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * prm.stage_bytes);
          const uint64_t da = make_desc(sa);
          const uint64_t db = make_desc(sa + kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (ki | k) != 0);
          umma_commit(empty0 + 8 * stage);
          if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; }
        }
        umma_commit(accr0 + 8 * buf);
      }
    }
  } else {
    // Epilogue.  A thread owns one accumulator row (pixel), so storing it directly would touch 32 different
    // 128-byte lines per instruction (one LSU wavefront each).  Every 32 x 32 chunk is therefore transposed through
    // a swizzled shared-memory tile: afterwards 8 consecutive lanes hold the 8 float4 of one pixel's 32 channels, and
    // each load of the residual / store of the result covers four full lines.
    const int quarter = warp & 3;
    uint8_t* tbuf = epi + quarter * (32 * 128);
    const int c4 = lane & 7, g = lane >> 3;
    int it = 0;
    for (int unit = blockIdx.x; unit < prm.num_units; unit += gridDim.x, ++it) {
      const int tile = unit / prm.nunit, nh = unit - tile * prm.nunit;
      const int buf = it & 1;
      mbar_wait(accr0 + 8 * buf, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const long long p0 = (long long)tile * kTileM + quarter * 32;
      const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256);
      const int col0 = nh * prm.ncols;
      for (int j = 0; j < prm.ncols / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(t_lane + (uint32_t)(j * 32), v);
        const int col = col0 + j * 32 + c4 * 4;
        float4 a4[8];
        if (prm.add) {
#pragma unroll
          for (int i = 0; i < 8; ++i) a4[i] = *reinterpret_cast<const float4*>(prm.add + (p0 + 4 * i + g) * prm.Cout + col);
        }
        if (prm.add2) {           // both residual streams are in flight before the accumulator load is awaited
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(prm.add2 + (p0 + 4 * i + g) * prm.Cout + col);
            if (prm.add) { a4[i].x += t.x; a4[i].y += t.y; a4[i].z += t.z; a4[i].w += t.w; }
            else a4[i] = t;
          }
        }
        const bool has_add = prm.add != nullptr || prm.add2 != nullptr;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (prm.bias) b4 = __ldg(reinterpret_cast<const float4*>(prm.bias + col));
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(tbuf + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + g;
          float4 o = *reinterpret_cast<const float4*>(tbuf + r * 128 + ((c4 ^ (r & 7)) << 4));
          o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
          if (has_add) { o.x += a4[i].x; o.y += a4[i].y; o.z += a4[i].z; o.w += a4[i].w; }
          *reinterpret_cast<float4*>(prm.out + (p0 + r) * prm.Cout + col) = o;
          if (prm.out_bf16)
            *reinterpret_cast<uint2*>(prm.out_bf16 + (p0 + r) * prm.Cout + col) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
          s1.x += o.x; s1.y += o.y; s1.z += o.z; s1.w += o.w;
          s2.x = fmaf(o.x, o.x, s2.x); s2.y = fmaf(o.y, o.y, s2.y); s2.z = fmaf(o.z, o.z, s2.z); s2.w = fmaf(o.w, o.w, s2.w);
        }
        if (prm.stats) {
          // the next layer's instance-norm statistics ride along: 32 pixels x 4 channels per lane group -> one
          // double atomic per (channel, moment) and warp
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
            s1.z += __shfl_xor_sync(0xffffffffu, s1.z, o); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, o);
            s2.x += __shfl_xor_sync(0xffffffffu, s2.x, o); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, o);
            s2.z += __shfl_xor_sync(0xffffffffu, s2.z, o); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, o);
          }
          // combine the four warps (128 pixels) in shared memory: one double atomic per (channel, moment) and tile
          float* red = reinterpret_cast<float*>(epi + 4 * 32 * 128) + (j & 1) * 256;
          if (g == 0) {
            float4* rr = reinterpret_cast<float4*>(red + quarter * 64 + c4 * 8);
            rr[0] = make_float4(s1.x, s2.x, s1.y, s2.y);
            rr[1] = make_float4(s1.z, s2.z, s1.w, s2.w);
          }
          named_bar_sync(1, 128);
          if (quarter == 0) {
            float2 t = reinterpret_cast<const float2*>(red)[lane];
#pragma unroll
            for (int w = 1; w < 4; ++w) {
              const float2 u = reinterpret_cast<const float2*>(red + w * 64)[lane];
              t.x += u.x; t.y += u.y;
            }
            double* st = prm.stats + ((size_t)(tile / prm.tiles_per_img) * prm.Cout + col0 + j * 32 + lane) * 2;
            atomicAdd(st, (double)t.x);
            atomicAdd(st + 1, (double)t.y);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce0 + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// Swapped-operand form for Cout = 128.  A tcgen05.mma with M = 128 costs max(~100, N/2) cycles, so an N = 128
// instruction runs the tensor pipe at 64 %.  (Measured gain over the plain form is small, ~2 %: these layers are bound by
// the ~64 B/cycle at which an SM ingests operands from L2, not by the tensor pipe.  A cta_group::2 variant that halves
// the weight ingest per SM was bit-exact but 1.8-2.3x slower and is not kept.)  Here the WEIGHTS are the M = 128 operand (A = [Cout x 64] tile image) and
// 256 PIXELS are N (B = the TMA activation box, K-major rows of 64 channels): D[co][pixel] with lanes = output
// channels.  An epilogue thread then owns one channel: stores / residual loads of one pixel are 128 contiguous bytes
// across the warp (no transpose), the bias is a per-thread scalar and the instance-norm statistics are per-thread sums.
constexpr int kSwPixels = 256;
constexpr int kSwStageBytes = 128 * 128 + kSwPixels * 128;       // weights 16 KB + activations 32 KB

__global__ void __launch_bounds__(kThreads, 1) k_conv_tc_sw(const __grid_constant__ CUtensorMap tmap, const ConvParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + prm.stages * kSwStageBytes);
  const int S = prm.stages;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]);
  const uint32_t accr0 = smem_u32(&bars[2 * S]), acce0 = smem_u32(&bars[2 * S + 2]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[2 * S + 4]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(accr0 + 8 * i, 1); mbar_init(acce0 + 8 * i, 4); }
    fence_barrier_init();
    tma_prefetch_desc(&tmap);
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kiters = prm.taps * prm.kpanels;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int unit = blockIdx.x; unit < prm.num_units; unit += gridDim.x) {
        const int n = unit / prm.tiles_per_img;
        const int h0 = (unit % prm.tiles_per_img) * prm.rows_per_tile;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(prm.wimg);
        for (int tap = 0; tap < prm.taps; ++tap) {
          const int dy = prm.taps == 9 ? (tap / 3 - 1) * prm.dil : 0;
          const int dx = prm.taps == 9 ? (tap % 3 - 1) * prm.dil : 0;
          for (int kp = 0; kp < prm.kpanels; ++kp) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t fb = full0 + 8 * stage;
            const uint32_t sa = smem_u32(smem + stage * kSwStageBytes);
            mbar_expect_tx(fb, (uint32_t)kSwStageBytes);
            bulk_g2s(sa, wsrc, 128 * 128, fb);
            tma_load_4d(sa + 128 * 128, &tmap, kp * 64, dx, h0 + dy, n, fb);
            wsrc += 128 * 128;
            if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      constexpr uint32_t idesc = make_idesc(kSwPixels);
      int it = 0;
      for (int unit = blockIdx.x; unit < prm.num_units; unit += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(acce0 + 8 * buf, ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
        for (int ki = 0; ki < kiters; ++ki) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kSwStageBytes);
          const uint64_t da = make_desc(sa);                    // weights: 128 rows (output channels) x 64 k
          const uint64_t db = make_desc(sa + 128 * 128);        // activations: 256 rows (pixels) x 64 k
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (ki | k) != 0);
          umma_commit(empty0 + 8 * stage);
          if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; }
        }
        umma_commit(accr0 + 8 * buf);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int co = quarter * 32 + lane;                         // this thread's output channel
    const float bias = prm.bias ? __ldg(prm.bias + co) : 0.f;
    int it = 0;
    for (int unit = blockIdx.x; unit < prm.num_units; unit += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(accr0 + 8 * buf, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const long long p0 = (long long)unit * kSwPixels;
      const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256);
      float s1 = 0.f, s2 = 0.f;
      for (int j = 0; j < kSwPixels / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(t_lane + (uint32_t)(j * 32), v);
        const long long pj = p0 + j * 32;
        float a[32];
        if (prm.add) {
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = prm.add[(pj + i) * 128 + co];
        }
        if (prm.add2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = (prm.add ? a[i] : 0.f) + prm.add2[(pj + i) * 128 + co];
        }
        const bool has_add = prm.add != nullptr || prm.add2 != nullptr;
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float o = __uint_as_float(v[i]) + bias;
          if (has_add) o += a[i];
          prm.out[(pj + i) * 128 + co] = o;
          s1 += o;
          s2 = fmaf(o, o, s2);
          v[i] = __float_as_uint(o);
        }
        if (prm.out_bf16) {
          // pack channel pairs: the even lane stores (co, co+1) of pixel i, the odd lane those of pixel i+1
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const bool odd = lane & 1;
            const float mine = __uint_as_float(odd ? v[i + 1] : v[i]);          // even: pixel i, odd: pixel i+1
            const float send = __uint_as_float(odd ? v[i] : v[i + 1]);          // what the partner lane needs
            const float got = __shfl_xor_sync(0xffffffffu, send, 1);
            const uint32_t pk = (lane & 1) ? pack_bf16(got, mine) : pack_bf16(mine, got);
            *reinterpret_cast<uint32_t*>(prm.out_bf16 + (pj + i + (lane & 1)) * 128 + (co & ~1)) = pk;
          }
        }
      }
      if (prm.stats) {
        double* st = prm.stats + ((size_t)(unit / prm.tiles_per_img) * 128 + co) * 2;
        atomicAdd(st, (double)s1);
        atomicAdd(st + 1, (double)s2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce0 + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    ASEP_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, ASEP_ERR_CUDA,
               "cuTensorMapEncodeTiled is not available from the driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

std::map<std::tuple<const void*, int, int, int, int, int>, CUtensorMap> g_maps;

const CUtensorMap& activation_map(const __nv_bfloat16* x, int N, int H, int W, int C, int pixels = kTileM) {
  auto key = std::make_tuple((const void*)x, N, H, W, C, pixels);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) return it->second;
  if (g_maps.size() > 4096) g_maps.clear();
  CUtensorMap m;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)(pixels / W), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ASEP_CHECK(r == CUDA_SUCCESS, ASEP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d]", (int)r, N, H, W, C);
  return g_maps.emplace(key, m).first->second;
}

int g_sms = 0;
struct ProfRec { cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfRec> g_recs, g_pool;
double g_flops = 0.0;

}  // namespace

bool conv_tc_supported(int Cin, int Cout, int H, int W) {
  if (Cin % 64 != 0 || Cin <= 0) return false;
  if (!(Cout == 128 || Cout == 192 || Cout == 256 || Cout == 384)) return false;
  if (W <= 0 || W > 128 || kTileM % W != 0) return false;
  return H % (kTileM / W) == 0;
}

void conv_tc_prepare(ConvWeightsTC& w, const float* kernel, const float* bias, int ksize, int Cin, int Cout,
                     int dilation) {
  conv_tc_release(w);
  ASEP_CHECK(ksize == 1 || ksize == 3, ASEP_ERR_UNSUPPORTED, "conv kernel size %d (1 or 3 expected)", ksize);
  ASEP_CHECK(Cin % 64 == 0, ASEP_ERR_UNSUPPORTED, "tcgen05 conv needs Cin %% 64 == 0 (got %d)", Cin);
  const int taps = ksize * ksize, kpanels = Cin / 64;
  std::vector<__nv_bfloat16> img((size_t)taps * kpanels * Cout * 64);
  size_t base = 0;
  for (int tap = 0; tap < taps; ++tap)
    for (int kp = 0; kp < kpanels; ++kp) {
      for (int r = 0; r < Cout; ++r)
        for (int k = 0; k < 64; ++k) {
          const float v = kernel[((size_t)tap * Cin + kp * 64 + k) * Cout + r];       // HWIO
          const size_t off = (size_t)r * 64 + (size_t)((((k >> 3) ^ (r & 7)) << 3) + (k & 7));
          img[base + off] = __float2bfloat16(v);
        }
      base += (size_t)Cout * 64;
    }
  w.bytes = img.size() * sizeof(__nv_bfloat16);
  CUDA_CHECK(cudaMalloc(&w.img, w.bytes));
  CUDA_CHECK(cudaMemcpy(w.img, img.data(), w.bytes, cudaMemcpyHostToDevice));
  if (bias) {
    CUDA_CHECK(cudaMalloc(&w.bias, (size_t)Cout * sizeof(float)));
    CUDA_CHECK(cudaMemcpy(w.bias, bias, (size_t)Cout * sizeof(float), cudaMemcpyHostToDevice));
  }
  w.Cin = Cin; w.Cout = Cout; w.ksize = ksize; w.dil = dilation;
}

void conv_tc_alloc(ConvWeightsTC& w, int ksize, int Cin, int Cout, int dilation) {
  conv_tc_release(w);
  ASEP_CHECK((ksize == 1 || ksize == 3) && Cin % 64 == 0, ASEP_ERR_UNSUPPORTED, "conv image %dx%d Cin=%d", ksize, ksize, Cin);
  w.bytes = (size_t)ksize * ksize * Cin * Cout * sizeof(__nv_bfloat16);
  CUDA_CHECK(cudaMalloc(&w.img, w.bytes));
  w.Cin = Cin; w.Cout = Cout; w.ksize = ksize; w.dil = dilation;
}

void conv_tc_release(ConvWeightsTC& w) {
  if (w.img) cudaFree(w.img);
  if (w.bias && w.bias_owned) cudaFree(w.bias);
  w = ConvWeightsTC{};
}

void conv_tc_forward(const ConvWeightsTC& w, const __nv_bfloat16* xin, const float* add, float* out, int N, int H,
                     int W, cudaStream_t s, double* stats, __nv_bfloat16* out_bf16, const float* add2) {
  if (N == 0) return;
  ASEP_CHECK(conv_tc_supported(w.Cin, w.Cout, H, W), ASEP_ERR_UNSUPPORTED,
             "tcgen05 conv: unsupported shape Cin=%d Cout=%d H=%d W=%d", w.Cin, w.Cout, H, W);
  if (g_sms == 0) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const bool swapped = w.Cout == 128 && kSwPixels % W == 0 && H % (kSwPixels / W) == 0 && getenv("ASEP_CONV_NO_SWAP") == nullptr;
  ConvParams prm{};
  prm.wimg = w.img; prm.bias = w.bias; prm.add = add; prm.add2 = add2; prm.out = out; prm.stats = stats; prm.out_bf16 = out_bf16;
  if (stats) CUDA_CHECK(cudaMemsetAsync(stats, 0, (size_t)N * w.Cout * 2 * sizeof(double), s));
  prm.Cout = w.Cout; prm.kpanels = w.Cin / 64; prm.taps = w.ksize * w.ksize; prm.dil = w.dil;
  prm.H = H; prm.W = W; prm.rows_per_tile = kTileM / W; prm.tiles_per_img = H / prm.rows_per_tile;
  const int num_tiles = N * prm.tiles_per_img;
  prm.nunit = w.Cout > 256 ? 2 : 1;
  prm.ncols = w.Cout / prm.nunit;
  prm.num_units = num_tiles * prm.nunit;
  prm.b_bytes = prm.ncols * 128;
  prm.stage_bytes = kABytes + prm.b_bytes;
  prm.stages = std::min(kMaxStages, (kSmemBudget - kEpiBytes) / prm.stage_bytes);
  const int smem_bytes = prm.stages * prm.stage_bytes + kEpiBytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(k_conv_tc_sw, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (swapped) {
    prm.rows_per_tile = kSwPixels / W;
    prm.tiles_per_img = H / prm.rows_per_tile;
    prm.num_units = N * prm.tiles_per_img;
    prm.stages = 4;
  }
  const CUtensorMap& tmap = activation_map(xin, N, H, W, w.Cin, swapped ? kSwPixels : kTileM);
  const int grid = std::min(prm.num_units, g_sms);
  ProfRec rec{};
  if (g_prof_on) {
    if (!g_pool.empty()) { rec = g_pool.back(); g_pool.pop_back(); }
    else { CUDA_CHECK(cudaEventCreate(&rec.a)); CUDA_CHECK(cudaEventCreate(&rec.b)); }
    CUDA_CHECK(cudaEventRecord(rec.a, s));
  }
  if (swapped) k_conv_tc_sw<<<grid, kThreads, 4 * kSwStageBytes + 1024, s>>>(tmap, prm);
  else k_conv_tc<<<grid, kThreads, smem_bytes, s>>>(tmap, prm);
  ASEP_LAUNCH_CHECK();
  if (g_prof_on) {
    CUDA_CHECK(cudaEventRecord(rec.b, s));
    g_recs.push_back(rec);
    g_flops += 2.0 * (double)N * H * W * prm.taps * w.Cin * w.Cout;
  }
}

bool conv_tc_profile_enabled() { return g_prof_on; }

void conv_tc_profile(int on) {
  g_prof_on = on != 0;
  if (g_prof_on) {
    for (auto& r : g_recs) g_pool.push_back(r);
    g_recs.clear();
    g_flops = 0.0;
  }
}

void conv_tc_profile_read(double* total_ms, long long* launches, double* flops) {
  double ms = 0.0;
  for (auto& r : g_recs) {
    CUDA_CHECK(cudaEventSynchronize(r.b));
    float t = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = (long long)g_recs.size();
  if (flops) *flops = g_flops;
}

}  // namespace asep
