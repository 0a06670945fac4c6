// Training-only pieces of the VQ codebook (second half): what EuclideanCodebook.forward does in
// train mode around the search (modules/quantization/core_vq.py:74-96 k-means, :151-169 dead-code expiry, :217-229
// EMA update) and the straight-through / commitment-loss step of VectorQuantization.forward (:294-316).
//   kmeans_assign     idx[n] = first argmin_k sum_d (x[n,d] - means[k,d])^2   (the reference's direct-difference
//                     form, NOT the expanded |x|^2 - 2x.e + |e|^2 of quantize)
//   cluster_sum       sum[k,:] = sum_{n: idx[n]=k} x[n,:] in sample order, cnt[k] = #{n: idx[n]=k}   (one CTA per
//                     code walks the samples in order: the reference's scatter_add_ / one-hot matmul, deterministic)
//   kmeans_update     means[k] = cnt[k] ? sum[k] / cnt[k] : means[k]
//   ema_update        cluster_size, embed_avg <- decay * old + (1 - decay) * new;  embed = embed_avg / smoothed size
//   replace_rows      embed[k] = samples[pick[k]] where cluster_size[k] < threshold
//   ste_commit(_bwd)  out = x + (q - x);  loss = mean((out - x)^2);  dx = g_out + g_loss * w * 2 (x - out) / numel
// All fp32, HBM/latency-bound and tiny next to the encoder convolutions that feed them.
#include <float.h>
#include <math.h>

#include "kernels.h"

namespace mtts {

constexpr int KM_SAMPLES = 16, KM_CODES = 64;

__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const float* __restrict__ x, const float* __restrict__ means, int N, int K, int D,
                     int64_t* __restrict__ idx) {
  pdl_entry();
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                       // [KM_SAMPLES][D]
  float* ms = xs + KM_SAMPLES * D;        // [KM_CODES][D + 1]
  const int DP = D + 1;
  const int tid = threadIdx.x, cl = tid & (KM_CODES - 1), sg = tid >> 6;   // 4 sample groups of 4
  const int n0 = blockIdx.x * KM_SAMPLES;
  for (int e = tid; e < KM_SAMPLES * D; e += 256) {
    const int s = e / D, d = e - s * D;
    xs[e] = (n0 + s < N) ? x[(int64_t)(n0 + s) * D + d] : 0.f;
  }
  float best[4];
  int bidx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best[i] = FLT_MAX; bidx[i] = 0x7fffffff; }
  for (int k0 = 0; k0 < K; k0 += KM_CODES) {
    __syncthreads();   // previous chunk consumed (and, first time, xs staged)
    for (int e = tid; e < KM_CODES * D; e += 256) {
      const int c = e / D, d = e - c * D;
      ms[c * DP + d] = (k0 + c < K) ? means[(int64_t)(k0 + c) * D + d] : 0.f;
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* mr = ms + cl * DP;
    const float* xr = xs + sg * 4 * D;
    for (int d = 0; d < D; ++d) {
      const float m = mr[d];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float df = xr[i * D + d] - m;
        acc[i] = fmaf(df, df, acc[i]);
      }
    }
    if (k0 + cl < K) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (acc[i] < best[i]) { best[i] = acc[i]; bidx[i] = k0 + cl; }   // strict: the earliest code wins a tie
    }
  }
  __syncthreads();
  // per sample: minimum over the 64 partial winners, ties to the lowest code index (torch.max -> first index)
  float* rv = ms;                                       // [KM_SAMPLES][KM_CODES]
  int* ri = reinterpret_cast<int*>(ms + KM_SAMPLES * KM_CODES);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    rv[(sg * 4 + i) * KM_CODES + cl] = best[i];
    ri[(sg * 4 + i) * KM_CODES + cl] = bidx[i];
  }
  __syncthreads();
  if (tid < KM_SAMPLES && n0 + tid < N) {
    float b = FLT_MAX;
    int bi = 0x7fffffff;
    for (int c = 0; c < KM_CODES; ++c) {
      const float v = rv[tid * KM_CODES + c];
      const int vi = ri[tid * KM_CODES + c];
      if (v < b || (v == b && vi < bi)) { b = v; bi = vi; }
    }
    idx[n0 + tid] = bi;
  }
}

__global__ void __launch_bounds__(256)
cluster_sum_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, int N, int D,
                   float* __restrict__ sum, float* __restrict__ cnt) {
  pdl_entry();
  __shared__ int ids[1024];
  const int c = blockIdx.x, tid = threadIdx.x;
  float acc[2] = {0.f, 0.f};          // D <= 512: dims tid and tid + 256
  int count = 0;
  for (int n0 = 0; n0 < N; n0 += 1024) {
    __syncthreads();
    for (int j = tid; j < 1024; j += 256) ids[j] = (n0 + j < N) ? (int)idx[n0 + j] : -1;
    __syncthreads();
    const int lim = min(1024, N - n0);
    for (int j = 0; j < lim; ++j) {
      if (ids[j] == c) {              // CTA-uniform
        ++count;
        const float* xr = x + (int64_t)(n0 + j) * D;
        if (tid < D) acc[0] += xr[tid];
        if (tid + 256 < D) acc[1] += xr[tid + 256];
      }
    }
  }
  if (tid < D) sum[(int64_t)c * D + tid] = acc[0];
  if (tid + 256 < D) sum[(int64_t)c * D + tid + 256] = acc[1];
  if (tid == 0) cnt[c] = (float)count;
}

__global__ void kmeans_update_kernel(float* __restrict__ means, const float* __restrict__ sum,
                                     const float* __restrict__ cnt, int K, int D) {
  pdl_entry();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)K * D) return;
  const float n = cnt[i / D];
  if (n > 0.f) means[i] = sum[i] / n;
}

// one CTA: cluster_size <- ema; total = sum_k cluster_size (fixed-order tree); smoothed[k] = (cs + eps) / (total + K eps) * total
__global__ void __launch_bounds__(1024)
ema_cluster_kernel(float* __restrict__ cs, const float* __restrict__ cnt, int K, float decay, float one_minus, float eps,
                   float* __restrict__ smoothed) {
  pdl_entry();
  __shared__ float red[1024];
  const int tid = threadIdx.x;
  float part = 0.f;
  for (int k = tid; k < K; k += 1024) {
    const float v = fmaf(cnt[k], one_minus, __fmul_rn(cs[k], decay));
    cs[k] = v;
    part += v;
  }
  red[tid] = part;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  const float total = red[0];
  const float denom = total + (float)K * eps;
  for (int k = tid; k < K; k += 1024) smoothed[k] = (cs[k] + eps) / denom * total;
}

__global__ void ema_embed_kernel(float* __restrict__ embed_avg, float* __restrict__ embed, const float* __restrict__ sum,
                                 const float* __restrict__ smoothed, int K, int D, float decay, float one_minus) {
  pdl_entry();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)K * D) return;
  const float v = fmaf(sum[i], one_minus, __fmul_rn(embed_avg[i], decay));
  embed_avg[i] = v;
  embed[i] = v / smoothed[i / D];
}

__global__ void replace_rows_kernel(float* __restrict__ embed, const float* __restrict__ samples,
                                    const int64_t* __restrict__ pick, const float* __restrict__ cs, float thr, int K, int D,
                                    int N) {
  pdl_entry();
  const int k = blockIdx.x;
  if (!(cs[k] < thr)) return;
  int64_t p = pick[k];
  p = p < 0 ? 0 : (p >= N ? N - 1 : p);
  for (int d = threadIdx.x; d < D; d += blockDim.x) embed[(int64_t)k * D + d] = samples[p * D + d];
}

constexpr int STE_BLOCKS = 256;

__global__ void __launch_bounds__(256)
ste_commit_kernel(const float* __restrict__ x, const float* __restrict__ q, int64_t n, float* __restrict__ out,
                  float* __restrict__ partials) {
  pdl_entry();
  __shared__ float red[256];
  float part = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)STE_BLOCKS * 256) {
    const float xv = x[i];
    const float o = xv + (q[i] - xv);      // the straight-through value, rounded as the reference rounds it
    out[i] = o;
    const float d = o - xv;
    part = fmaf(d, d, part);
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256)
ste_loss_kernel(const float* __restrict__ partials, float inv_n, float* __restrict__ loss) {
  pdl_entry();
  __shared__ float red[256];
  red[threadIdx.x] = partials[threadIdx.x];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] * inv_n;
}

__global__ void ste_commit_bwd_kernel(const float* __restrict__ x, const float* __restrict__ out,
                                      const float* __restrict__ g_out, const float* __restrict__ g_loss, float scale,
                                      int64_t n, float* __restrict__ dx) {
  pdl_entry();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gl = g_loss ? g_loss[0] * scale : 0.f;   // scale = commitment_weight * 2 / numel
  const float go = g_out ? g_out[i] : 0.f;
  dx[i] = fmaf(gl, x[i] - out[i], go);
}

}  // namespace mtts

using namespace mtts;

extern "C" {

int mtts_kmeans_assign_f32(const float* x, const float* means, int32_t N, int32_t K, int32_t D, int64_t* idx, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MTTS_REQUIRE(x && means && idx, "null pointer");
  MTTS_REQUIRE(N >= 0 && K > 0 && D > 0 && D <= 512, "bad shape (D <= 512)");
  if (N == 0) return 0;
  size_t ms_floats = (size_t)KM_CODES * (D + 1);
  if (ms_floats < 2u * KM_SAMPLES * KM_CODES) ms_floats = 2u * KM_SAMPLES * KM_CODES;   // the final (value, index) exchange reuses it
  const size_t smem = sizeof(float) * ((size_t)KM_SAMPLES * D + ms_floats);
  static bool attr[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(MTTS_ERR_CUDA, "%s: cudaFuncSetAttribute failed: %lld", "kmeans_assign", (long long)e); }
    attr[dev] = true;
  }
  launch_k(kmeans_assign_kernel, (unsigned)cdiv64(N, KM_SAMPLES), 256, smem, st, x, means, (int)N, (int)K, (int)D, idx);
  MTTS_CHECK_LAUNCH();
  return 0;
}

int mtts_vq_cluster_sum_f32(const float* x, const int64_t* idx, int32_t N, int32_t K, int32_t D, float* sum, float* cnt,
                            void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MTTS_REQUIRE(x && idx && sum && cnt, "null pointer");
  MTTS_REQUIRE(N >= 0 && K > 0 && D > 0 && D <= 512, "bad shape (D <= 512)");
  launch_k(cluster_sum_kernel, (unsigned)K, 256, 0, st, x, idx, (int)N, (int)D, sum, cnt);
  MTTS_CHECK_LAUNCH();
  return 0;
}

int mtts_kmeans_update_f32(float* means, const float* sum, const float* cnt, int32_t K, int32_t D, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MTTS_REQUIRE(means && sum && cnt, "null pointer");
  MTTS_REQUIRE(K > 0 && D > 0, "bad shape");
  launch_k(kmeans_update_kernel, (unsigned)cdiv64((int64_t)K * D, 256), 256, 0, st, means, sum, cnt, (int)K, (int)D);
  MTTS_CHECK_LAUNCH();
  return 0;
}

int mtts_vq_ema_update_f32(float* cluster_size, float* embed_avg, float* embed, const float* sum, const float* cnt,
                           int32_t K, int32_t D, float decay, float eps, float* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MTTS_REQUIRE(cluster_size && embed_avg && embed && sum && cnt && scratch, "null pointer");
  MTTS_REQUIRE(K > 0 && D > 0, "bad shape");
  const float one_minus = (float)(1.0 - (double)decay);
  launch_k(ema_cluster_kernel, 1, 1024, 0, st, cluster_size, cnt, (int)K, decay, one_minus, eps, scratch);
  MTTS_CHECK_LAUNCH();
  launch_k(ema_embed_kernel, (unsigned)cdiv64((int64_t)K * D, 256), 256, 0, st, embed_avg, embed, sum,
           (const float*)scratch, (int)K, (int)D, decay, one_minus);
  MTTS_CHECK_LAUNCH();
  return 0;
}

int mtts_vq_replace_rows_f32(float* embed, const float* samples, const int64_t* pick, const float* cluster_size, float thr,
                             int32_t K, int32_t D, int32_t N, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MTTS_REQUIRE(embed && samples && pick && cluster_size, "null pointer");
  MTTS_REQUIRE(K > 0 && D > 0 && N > 0, "bad shape");
  launch_k(replace_rows_kernel, (unsigned)K, 128, 0, st, embed, samples, pick, cluster_size, thr, (int)K, (int)D, (int)N);
  MTTS_CHECK_LAUNCH();
  return 0;
}

int mtts_vq_ste_commit_f32(const float* x, const float* q, int64_t n, float* out, float* partials, float* loss, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MTTS_REQUIRE(x && q && out && partials && loss, "null pointer");
  MTTS_REQUIRE(n > 0, "empty input");
  launch_k(ste_commit_kernel, STE_BLOCKS, 256, 0, st, x, q, n, out, partials);
  MTTS_CHECK_LAUNCH();
  launch_k(ste_loss_kernel, 1, 256, 0, st, (const float*)partials, (float)(1.0 / (double)n), loss);
  MTTS_CHECK_LAUNCH();
  return 0;
}

int mtts_vq_ste_commit_bwd_f32(const float* x, const float* out, const float* g_out, const float* g_loss, float scale,
                               int64_t n, float* dx, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MTTS_REQUIRE(x && out && dx, "null pointer");
  MTTS_REQUIRE(n > 0, "empty input");
  launch_k(ste_commit_bwd_kernel, (unsigned)cdiv64(n, 256), 256, 0, st, x, out, g_out, g_loss, scale, n, dx);
  MTTS_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
