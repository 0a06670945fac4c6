// Microbenchmark: what does one tcgen05.mma (kind::f16, M = 128, K = 16, SS operands) cost on this part?
// One CTA per SM, one issuing thread, operands resident in shared memory (no loads), commits every 24 MMAs like
// conv_bf16x3_kernel.  Optional: other warps stream stores into a disjoint shared-memory region, to see whether
// operand reads and TMA-style writes contend for the same shared-memory bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mma_floor tools/microbench/mma_floor.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../megatts2_b200/csrc/tc_ptx.cuh"
using namespace mtts;

template <int N>
__global__ void __launch_bounds__(256, 1) mma_loop(int iters, int hammer, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 160 * 1024, slot = bar + 64;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(bar + 8 * i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  __shared__ volatile int done;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t db = umma_desc_kmajor<128>(0u);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // ring of 8 commit barriers: at most 8 commit groups in flight, like a deep pipeline would allow
      if (it >= 8) mbar_wait(bar + 8 * (it & 7), ((it >> 3) - 1) & 1);
      // 2 "stages" x 3 planes of A (16 KB each) and B, like the real kernel
      const uint32_t sa = base + (it & 1) * 96 * 1024 * 0;     // same stage: operands stay put
      const uint64_t a0 = db | (uint64_t)((sa >> 4) & 0x3FFF), b0 = a0 + (3 * 16384 >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t a1 = a0 + 2 * ks, a2 = a1 + 1024, a3 = a2 + 1024;
        const uint64_t b1 = b0 + 2 * ks, b2 = b1 + (N * 128 >> 4), b3 = b2 + (N * 128 >> 4);
        tc_mma_bf16(tmem + N, a2, b2, idesc, 1u);
        tc_mma_bf16(tmem + N, a1, b3, idesc, 1u);
        tc_mma_bf16(tmem + N, a3, b1, idesc, 1u);
        tc_mma_bf16(tmem + N, a1, b2, idesc, 1u);
        tc_mma_bf16(tmem + N, a2, b1, idesc, 1u);
        tc_mma_bf16(tmem, a1, b1, idesc, 1u);
      }
      tc_commit(bar + 8 * (it & 7));
    }
    for (int it = iters > 8 ? iters - 8 : 0; it < iters; ++it) mbar_wait(bar + 8 * (it & 7), (it >> 3) & 1);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    done = 1;
  } else if (hammer && warp >= 4) {
    // stream 16-byte stores into the upper part of shared memory (disjoint from the operands)
    uint4* dst = reinterpret_cast<uint4*>(smem_raw + (base + 112 * 1024 - smem_u32(smem_raw)));
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    int i = threadIdx.x - 128;
    while (!done) {
#pragma unroll
      for (int r = 0; r < 16; ++r) { dst[(i + r * 128) & 2047] = v; }
      v.x++;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

template <int N>
static void run(int hammer) {
  const int smem = 164 * 1024, iters = 4000, sms = 148;
  cudaFuncSetAttribute(mma_loop<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  cudaMalloc(&d, sms * sizeof(long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mma_loop<N><<<sms, 256, smem>>>(200, hammer, d);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  mma_loop<N><<<sms, 256, smem>>>(iters, hammer, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
  const double mmas = (double)iters * 24;
  printf("N=%3d hammer=%d: %7.1f SM-cycles per MMA (floor %3d), %.3f ms, %.1f dense bf16 TFLOP/s chip-wide, err=%d\n", N, hammer,
         avg / mmas, N / 2, ms, sms * mmas * 2.0 * 128 * N * 16 / (ms * 1e-3) / 1e12, (int)e);
  cudaFree(d);
}

int main() {
  for (int h = 0; h < 2; ++h) { run<32>(h); run<64>(h); run<128>(h); run<256>(h); }
  return 0;
}
