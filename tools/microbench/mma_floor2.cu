// Microbenchmark 2: cost of one tcgen05.mma (kind::f16, K = 16, SS operands) when it is issued the way conv_tc_kernel
// issues it since round 2 - warp-uniformly, one elected lane, operands in uniform registers - for
//   cta_group::1, M = 128          (single CTA)
//   cta_group::2, M = 256          (a CTA pair; the leader CTA issues, each CTA holds its 128 rows of A and N/2 rows of B)
// and N = 32 / 64 / 128.  One CTA (pair) per SM (pair), operands resident in shared memory, a commit every 24 MMAs.
// The round-1 figure (55 cycles for N <= 64, tools/microbench/mma_floor.cu) was taken with the MMAs inside a single-thread
// branch, i.e. including ptxas' ELECT / R2UR waterfall - this one measures the pipe itself.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mma_floor2 tools/microbench/mma_floor2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../megatts2_b200/csrc/tc_ptx.cuh"
using namespace mtts;

template <int N, int PAIR>
__global__ void __launch_bounds__(128, 1) mma_loop(int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 96 * 1024, slot = bar + 64;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(bar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(256) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(256) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *slot_ptr, 0);
  if (warp == 0 && crank == 0) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr int M = PAIR ? 256 : 128;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // f16 x f16 -> f32
    const uint64_t db = umma_desc_kmajor<128>(0u);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (it >= 8) mbar_wait(bar + 8 * (it & 7), ((it >> 3) - 1) & 1);
      const uint64_t a0 = db | (uint64_t)((base >> 4) & 0x3FFF), b0 = a0 + (2 * 16384 >> 4);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {     // f16x2: three MMAs per k-step
        const uint64_t a1 = a0 + 2 * (ks & 3), a2 = a1 + 1024;
        const uint64_t b1 = b0 + 2 * (ks & 3), b2 = b1 + 512;
        if (PAIR) {
          tc_mma_2sm_l(tmem + N, a1, b2, idesc, 1u, leader);
          tc_mma_2sm_l(tmem + N, a2, b1, idesc, 1u, leader);
          tc_mma_2sm_l(tmem, a1, b1, idesc, 1u, leader);
        } else {
          tc_mma_l(tmem + N, a1, b2, idesc, 1u, leader);
          tc_mma_l(tmem + N, a2, b1, idesc, 1u, leader);
          tc_mma_l(tmem, a1, b1, idesc, 1u, leader);
        }
      }
      if (PAIR) tc_commit_2sm_l(bar + 8 * (it & 7), leader); else tc_commit_l(bar + 8 * (it & 7), leader);
    }
    for (int it = iters > 8 ? iters - 8 : 0; it < iters; ++it) mbar_wait(bar + 8 * (it & 7), (it >> 3) & 1);
    const long long t1 = clock64();
    if (leader) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
  }
}

template <int N, int PAIR>
static void run() {
  const int smem = 100 * 1024, iters = 4000, ctas = 148;
  cudaFuncSetAttribute(mma_loop<N, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  cudaMalloc(&d, ctas * sizeof(long long));
  cudaMemset(d, 0, ctas * sizeof(long long));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, mma_loop<N, PAIR>, 200, d);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, mma_loop<N, PAIR>, iters, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; int n = 0;
  for (int i = 0; i < ctas; ++i) if (h[i]) { avg += h[i]; ++n; }
  avg /= (n ? n : 1);
  const double mmas = (double)iters * 24;
  const int M = PAIR ? 256 : 128;
  printf("cta_group::%d M=%3d N=%3d: %6.1f SM-cycles per MMA (pipe floor %3d per SM), %.3f ms, %7.1f dense 16-bit TFLOP/s chip-wide, err=%d\n",
         PAIR ? 2 : 1, M, N, avg / mmas, N / 2, ms, n * mmas * 2.0 * M * N * 16 / (ms * 1e-3) / 1e12, (int)e);
  cudaFree(d);
}

int main() {
  run<32, 0>(); run<64, 0>(); run<128, 0>();
  run<32, 1>(); run<64, 1>(); run<128, 1>();
  return 0;
}
