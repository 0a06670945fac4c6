// Tensor-core engine for the dense contractions of the path (nn.Linear layers of the PLM / ADM
// / MRTE encoders): fp32-grade GEMM on tcgen05 via a 3-way bf16 split ("bf16x3").
//
//   x = x1 + x2 + x3,  x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2)   (exact to 2^-24)
//   x.w ~= x1w1 + (x1w2 + x2w1) + (x2w2 + x1w3 + x3w1)       6 bf16 MMAs, fp32 accumulate in TMEM
//   dropped terms are O(2^-24) relative, i.e. fp32 rounding level: VQ / PLM ids keep matching the
//   fp32 oracle, which single-pass TF32/BF16 cannot promise.
//
// Kernel anatomy (sm_100a): persistent CTAs, one per SM; warp 0 = TMA producer
// (cp.async.bulk.tensor -> 128B-swizzled smem, mbarrier complete_tx), warp 1 = single-thread
// tcgen05.mma issuer (UMMA 128x128x16, kind::f16, accumulators in TMEM, double-buffered),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> bias / activation / residual ->
// coalesced fp32 stores).  Operands are K-major: A planes (M,K) bf16, W planes (N,K) bf16.
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>
#include <unordered_map>

#include "kernels.h"
#include "tc_ptx.cuh"

namespace mtts {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 2;
constexpr int TC_PLANE_BYTES = 128 * 128;                 // 128 rows x 128 B (64 bf16)
constexpr int TC_STAGE_BYTES = 6 * TC_PLANE_BYTES;        // A1 A2 A3 B1 B2 B3
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TC_TMEM_COLS = 512;                         // 2 buffers x (main + correction accumulator) x 128 fp32 columns
constexpr int TC_ACC_STRIDE = 2 * TC_BN;                  // columns per accumulator buffer

struct TcMaps {
  CUtensorMap a[3];
  CUtensorMap b[3];
};

struct TcArgs {
  int64_t M;
  int32_t N, K;
  const float* bias;
  const float* res; int32_t ldr;
  float* y; int32_t ldy;
  int32_t post_act;
  float out_scale;
};

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
gemm_bf16x3_kernel(const __grid_constant__ TcMaps maps, const TcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SWIZZLE_128B needs 1024 B alignment
  const uint32_t bars = smem_base + TC_STAGES * TC_STAGE_BYTES;              // 8-byte barriers
  const uint32_t full_bar = bars, empty_bar = bars + 8 * TC_STAGES;
  const uint32_t tfull_bar = bars + 16 * TC_STAGES, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (int)((g.M + TC_BM - 1) / TC_BM), num_n = (g.N + TC_BN - 1) / TC_BN;
  const int num_tiles = num_m * num_n, num_k = (g.K + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[i]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[i]) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + 8 * s, 1);
      mbar_init(tempty_bar + 8 * s, 4);     // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          mbar_expect_tx(fb, TC_STAGE_BYTES);
          const uint32_t sbase = smem_base + stage * TC_STAGE_BYTES;
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            tma_load_2d(sbase + p * TC_PLANE_BYTES, &maps.a[p], fb, kb * TC_BK, m_blk * TC_BM);
            tma_load_2d(sbase + (3 + p) * TC_PLANE_BYTES, &maps.b[p], fb, kb * TC_BK, n_blk * TC_BN);
          }
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(tempty_bar + 8 * as, aphase ^ 1);        // epilogue has drained this accumulator buffer
        tc_fence_after();
        // Two accumulators per tile: the tensor core accumulates with truncation, so the error grows with the
        // number of accumulations into a LARGE accumulator.  x1w1 (the O(1) term) gets its own accumulator
        // (K/16 accumulations); the five O(2^-8)..O(2^-16) correction products share a second, small one.
        const uint32_t d_main = tmem_base + as * TC_ACC_STRIDE;
        const uint32_t d_corr = d_main + TC_BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sbase = smem_base + stage * TC_STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < TC_BK / 16; ++ks) {
            uint64_t ad[3], bd[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
              ad[p] = umma_desc_sw128(sbase + p * TC_PLANE_BYTES + ks * 32);
              bd[p] = umma_desc_sw128(sbase + (3 + p) * TC_PLANE_BYTES + ks * 32);
            }
            const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
            tc_mma_bf16(d_corr, ad[1], bd[1], idesc, first);   // x2 w2   (smallest terms first)
            tc_mma_bf16(d_corr, ad[0], bd[2], idesc, 1u);      // x1 w3
            tc_mma_bf16(d_corr, ad[2], bd[0], idesc, 1u);      // x3 w1
            tc_mma_bf16(d_corr, ad[0], bd[1], idesc, 1u);      // x1 w2
            tc_mma_bf16(d_corr, ad[1], bd[0], idesc, 1u);      // x2 w1
            tc_mma_bf16(d_main, ad[0], bd[0], idesc, first);   // x1 w1
          }
          tc_commit(empty_bar + 8 * stage);                    // smem slot free once these MMAs retire
          if (kb == num_k - 1) tc_commit(tfull_bar + 8 * as);  // accumulator complete
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> global =================
    const int q = warp - 4;                       // == warp % 4: the TMEM lane quarter this warp may read
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
      const int as = it & 1, aphase = (it >> 1) & 1;
      mbar_wait(tfull_bar + 8 * as, aphase);
      tc_fence_after();
      const int64_t row = (int64_t)m_blk * TC_BM + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < TC_BN / 32; ++c) {
        uint32_t r[32], rc[32];
        tmem_ld32(tmem_base + as * TC_ACC_STRIDE + c * 32 + ((uint32_t)(q * 32) << 16), r);
        tmem_ld32(tmem_base + as * TC_ACC_STRIDE + TC_BN + c * 32 + ((uint32_t)(q * 32) << 16), rc);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(rc[j]));
        const int n0 = n_blk * TC_BN + c * 32;
        if (row < g.M && n0 < g.N) {
          float* yr = g.y + row * g.ldy + n0;
          const float* rr = g.res ? g.res + row * g.ldr + n0 : nullptr;
          if (n0 + 32 <= g.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float x = __uint_as_float(r[j + e]);
                if (g.bias) x += __ldg(g.bias + n0 + j + e);
                v[e] = act_apply(x, g.post_act, 0.f);
              }
              if (rr) {
                const float4 t = *reinterpret_cast<const float4*>(rr + j);
                v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
              }
              *reinterpret_cast<float4*>(yr + j) = make_float4(v[0] * g.out_scale, v[1] * g.out_scale, v[2] * g.out_scale, v[3] * g.out_scale);
            }
          } else {
            for (int j = 0; j < 32 && n0 + j < g.N; ++j) {
              float x = __uint_as_float(r[j]);
              if (g.bias) x += __ldg(g.bias + n0 + j);
              x = act_apply(x, g.post_act, 0.f);
              if (rr) x += rr[j];
              yr[j] = x * g.out_scale;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * as);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
  }
}

// fp32 -> three bf16 planes (optionally through a pre-activation); 4 elements per thread
__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const float* __restrict__ x, int ldx, int64_t rows, int K, int pre_act, float slope,
                    __nv_bfloat16* __restrict__ planes, int64_t plane_stride, int ldp) {
  const int kq = K >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * kq) return;
  const int64_t r = i / kq;
  const int c = (int)(i - r * kq) * 4;
  const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
  float f[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 p[3][4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float a = act_apply(f[e], pre_act, slope);
    p[0][e] = __float2bfloat16_rn(a);
    a -= __bfloat162float(p[0][e]);
    p[1][e] = __float2bfloat16_rn(a);
    a -= __bfloat162float(p[1][e]);
    p[2][e] = __float2bfloat16_rn(a);
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    uint2 o;
    o.x = (uint32_t)__bfloat16_as_ushort(p[q][0]) | ((uint32_t)__bfloat16_as_ushort(p[q][1]) << 16);
    o.y = (uint32_t)__bfloat16_as_ushort(p[q][2]) | ((uint32_t)__bfloat16_as_ushort(p[q][3]) << 16);
    *reinterpret_cast<uint2*>(planes + q * plane_stride + r * ldp + c) = o;
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::mutex g_tc_mu;
static bool g_tc_attr = false;
static int g_sm_count = 0;

struct MapKey {
  const void* p; uint64_t rows, cols, ld;
  bool operator==(const MapKey& o) const { return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<uint64_t>()((uint64_t)k.p ^ (k.rows * 0x9E3779B97F4A7C15ull) ^ (k.cols << 20) ^ (k.ld << 40));
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

static int tc_init() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
    return fail(MTTS_ERR_CUDA, "%s: cuTensorMapEncodeTiled not available", "gemm_tc");
  g_encode = (EncodeTiledFn)fn;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  return 0;
}

// (rows, cols) bf16, row stride ld elements, box 128 rows x 64 cols, 128B swizzle, zero OOB fill
static int get_map(const void* p, uint64_t rows, uint64_t cols, uint64_t ld, CUtensorMap* out) {
  MapKey key{p, rows, cols, ld};
  auto itf = g_maps.find(key);
  if (itf != g_maps.end()) { *out = itf->second; return 0; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {TC_BK, 128};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MTTS_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed: %lld", "gemm_tc", (long long)r);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps[key] = m;
  *out = m;
  return 0;
}

int64_t linear_tc_scratch_bytes(int64_t rows_cap, int K) { return 3 * rows_cap * (int64_t)K * 2 + 1024; }

// Y[M,N] = post(X[M,K] W^T + bias) (+res) * scale, W given as 3 bf16 planes (N,K).
// scratch holds the 3 activation planes with capacity rows_cap >= M (tensor maps are cached per capacity).
int linear_tc(const float* x, int ldx, int64_t M, int K, const void* w_planes, int N, const float* bias,
              const float* res, int ldr, float* y, int ldy, int pre_act, float pre_slope, int post_act,
              float out_scale, void* scratch, int64_t scratch_bytes, int64_t rows_cap, cudaStream_t st) {
  MTTS_REQUIRE(x && w_planes && y && scratch, "null pointer");
  MTTS_REQUIRE(K % 8 == 0 && ldx % 4 == 0 && (((uintptr_t)x) & 15) == 0, "K %% 8 / alignment");
  MTTS_REQUIRE(ldy % 4 == 0 && (((uintptr_t)y) & 15) == 0 && (!res || (ldr % 4 == 0 && (((uintptr_t)res) & 15) == 0)),
               "output alignment");
  MTTS_REQUIRE(rows_cap >= M && linear_tc_scratch_bytes(rows_cap, K) <= scratch_bytes, "scratch too small");
  if (M <= 0) return 0;
  std::lock_guard<std::mutex> lk(g_tc_mu);
  MTTS_TRY(tc_init());
  if (!g_tc_attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) return fail(MTTS_ERR_CUDA, "%s: cudaFuncSetAttribute failed: %lld", "gemm_tc", (long long)e);
    g_tc_attr = true;
  }
  __nv_bfloat16* planes = reinterpret_cast<__nv_bfloat16*>((((uintptr_t)scratch) + 1023) & ~(uintptr_t)1023);
  const int64_t plane_stride = rows_cap * (int64_t)K;
  {
    const int64_t n4 = M * (K / 4);
    split_bf16x3_kernel<<<(unsigned)cdiv64(n4, 256), 256, 0, st>>>(x, ldx, M, K, pre_act, pre_slope, planes, plane_stride, K);
    MTTS_CHECK_LAUNCH();
  }
  TcMaps maps;
  for (int p = 0; p < 3; ++p) {
    MTTS_TRY(get_map(planes + p * plane_stride, (uint64_t)rows_cap, (uint64_t)K, (uint64_t)K, &maps.a[p]));
    MTTS_TRY(get_map((const __nv_bfloat16*)w_planes + (int64_t)p * N * K, (uint64_t)N, (uint64_t)K, (uint64_t)K, &maps.b[p]));
  }
  TcArgs a;
  a.M = M; a.N = N; a.K = K; a.bias = bias; a.res = res; a.ldr = ldr; a.y = y; a.ldy = ldy; a.post_act = post_act;
  a.out_scale = out_scale;
  const int64_t tiles = cdiv64(M, TC_BM) * cdiv64(N, TC_BN);
  const int grid = (int)(tiles < g_sm_count ? tiles : g_sm_count);
  gemm_bf16x3_kernel<<<grid, 256, TC_SMEM_BYTES, st>>>(maps, a);
  MTTS_CHECK_LAUNCH();
  return 0;
}

}  // namespace mtts
