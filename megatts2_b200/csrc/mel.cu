// Mel front end: reflect-padded STFT (n_fft = win = 1024, hop 256) -> |.| -> banded slaney
// filterbank -> log(max(., clamp)).  Replaces extract_mel_spec (modules/tokenizer.py:107-125,
// speechbrain mel_spectogram -> torchaudio MelSpectrogram).
//
// One warp per frame: the 1024 real samples are packed as 512 complex points and transformed
// by three register-resident radix-8 Stockham passes (8*8*8) that exchange data through a
// per-warp padded shared buffer (conflict-free, __syncwarp only), then untangled to the 513
// one-sided bins.  A CTA owns 8 consecutive frames of one clip, stages their 2816-sample span
// once (each sample is needed by 4 frames) and writes an 80 x 8 output tile coalesced.
#include <math.h>
#include <mutex>

#include "common.cuh"

namespace mtts {

constexpr int MEL_NFFT = 1024, MEL_HOP = 256, MEL_FPB = 8, MEL_NZ = 512;
constexpr int MEL_ZPAD = MEL_NZ + MEL_NZ / 8;   // float2 slots incl. 1 pad per 8
constexpr int MEL_MAGP = 520;

__device__ float2 g_tw1024[1024];   // exp(-2*pi*i*k/1024), filled once per process in fp64

__global__ void mel_init_twiddle_kernel() {
  pdl_entry();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < 1024) {
    double s, c;
    sincospi(2.0 * (double)k / 1024.0, &s, &c);
    g_tw1024[k] = make_float2((float)c, (float)(-s));
  }
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void bf2(float2& a, float2& b) {
  const float2 t = a;
  a = make_float2(t.x + b.x, t.y + b.y);
  b = make_float2(t.x - b.x, t.y - b.y);
}
__device__ __forceinline__ void fft4_inplace(float2& a0, float2& a1, float2& a2, float2& a3) {
  bf2(a0, a2);
  bf2(a1, a3);
  a3 = make_float2(a3.y, -a3.x);   // * (-i)
  bf2(a0, a1);
  bf2(a2, a3);
}
// forward 8-point DFT; natural-order result is (v0,v4,v2,v6,v1,v5,v3,v7)
__device__ __forceinline__ void fft8_inplace(float2* v) {
  const float s = 0.70710678118654752440f;
  bf2(v[0], v[4]);
  bf2(v[1], v[5]);
  bf2(v[2], v[6]);
  bf2(v[3], v[7]);
  v[5] = make_float2((v[5].x + v[5].y) * s, (v[5].y - v[5].x) * s);
  v[6] = make_float2(v[6].y, -v[6].x);
  v[7] = make_float2((-v[7].x + v[7].y) * s, (-v[7].x - v[7].y) * s);
  fft4_inplace(v[0], v[1], v[2], v[3]);
  fft4_inplace(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ int zpad(int i) { return i + (i >> 3); }

__global__ void __launch_bounds__(MEL_FPB * 32, 3)
mel_kernel(const float* __restrict__ wav, int64_t wav_sb, int L_max, const int32_t* __restrict__ lens,
           const float* __restrict__ window,
           const float* __restrict__ fb_w, const int32_t* __restrict__ fb_off, const int32_t* __restrict__ fb_start,
           int n_mels, float clamp_min, float* __restrict__ out, int64_t out_sb, int64_t out_sm, int64_t out_sf,
           int vec_ok) {
  pdl_entry();
  constexpr int SPAN = MEL_NFFT + (MEL_FPB - 1) * MEL_HOP;   // 2816
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                                  // [SPAN]
  float* win = xs + SPAN;                            // [1024]
  float2* zb = reinterpret_cast<float2*>(win + MEL_NFFT);   // [FPB][ZPAD]
  float* mag = reinterpret_cast<float*>(zb + MEL_FPB * MEL_ZPAD);   // [FPB][MAGP]
  float* otile = mag + MEL_FPB * MEL_MAGP;           // [128][FPB]
  float* fbw_s = otile + 128 * MEL_FPB;              // [<= 1536] banded filterbank taps

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * MEL_FPB;
  // ragged batches (bulk extraction): clip b holds lens[b] <= L_max valid samples; its reflection and frame
  // count follow ITS length.  CTA-uniform exit for the frames past a short clip's end.
  const int L = lens ? min(lens[b], L_max) : L_max;
  const int F = 1 + L / MEL_HOP;
  if (f0 >= F || L <= MEL_NFFT / 2) return;
  const float* wv = wav + (int64_t)b * wav_sb;
  const int g0 = f0 * MEL_HOP - MEL_NFFT / 2;
  // stage the CTA's sample span.  Interior tiles (no reflection, 16-byte aligned) use batched float4 loads so
  // that all of a thread's global requests are in flight together; edge tiles take the scalar reflect path.
  const bool interior = vec_ok && g0 >= 0 && (g0 + SPAN) <= L;
  if (interior) {
    constexpr int NV = SPAN / 4;                               // 704 float4
    constexpr int PER = (NV + MEL_FPB * 32 - 1) / (MEL_FPB * 32);
    const float4* src = reinterpret_cast<const float4*>(wv + g0);
    float4 v[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = tid + q * MEL_FPB * 32;
      if (i < NV) v[q] = __ldg(src + i);
    }
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = tid + q * MEL_FPB * 32;
      if (i < NV) *reinterpret_cast<float4*>(xs + 4 * i) = v[q];
    }
  } else {
    for (int i = tid; i < SPAN; i += MEL_FPB * 32) {
      int g = g0 + i;
      if (g < 0) g = -g;
      if (g >= L) g = 2 * (L - 1) - g;
      xs[i] = (g >= 0 && g < L) ? __ldg(wv + g) : 0.f;
    }
  }
  {
    const float4* wsrc = reinterpret_cast<const float4*>(window);   // 1024 floats, 16-byte aligned (checked on host)
    *reinterpret_cast<float4*>(win + 4 * tid) = __ldg(wsrc + tid);  // 256 threads x float4 = 1024
  }
  const int n_taps = min(fb_off[n_mels], 1536);      // taps beyond the staged window fall back to global loads
  for (int i = tid; i < n_taps; i += MEL_FPB * 32) fbw_s[i] = __ldg(fb_w + i);
  __syncthreads();

  const int f = f0 + w;
  if (f < F) {   // warp-uniform
    float2* z = zb + w * MEL_ZPAD;
    float* mg = mag + w * MEL_MAGP;
    const float* xf = xs + w * MEL_HOP;
    float2 v[2][8];
    // ---- pass 0 (Ns = 1): no twiddles; inputs straight from the windowed samples
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int n = 2 * (j + 64 * r);   // even; xf and win are 8-byte aligned
        const float2 xv = *reinterpret_cast<const float2*>(xf + n);
        const float2 wv2 = *reinterpret_cast<const float2*>(win + n);
        v[h][r] = make_float2(xv.x * wv2.x, xv.y * wv2.y);
      }
      fft8_inplace(v[h]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
      const int idx = j * 8;
      z[zpad(idx + 0)] = v[h][0]; z[zpad(idx + 1)] = v[h][4]; z[zpad(idx + 2)] = v[h][2]; z[zpad(idx + 3)] = v[h][6];
      z[zpad(idx + 4)] = v[h][1]; z[zpad(idx + 5)] = v[h][5]; z[zpad(idx + 6)] = v[h][3]; z[zpad(idx + 7)] = v[h][7];
    }
    // ---- pass 1 (Ns = 8) and pass 2 (Ns = 64).  Twiddles w^(r*k): k = j & (Ns-1).  For Ns = 8 both halves of a lane
    // (j = lane, lane + 32) share k; for Ns = 64 the second half's twiddle is the first half's times exp(-2 pi i r/16),
    // a compile-time constant - so a lane fetches 7 table entries per pass (not 14), and fetches them BEFORE the
    // warp barrier that precedes the pass so their latency overlaps the exchange through shared memory.
#pragma unroll
    for (int pass = 1; pass < 3; ++pass) {
      const int Ns = (pass == 1) ? 8 : 64;
      const int tws = (pass == 1) ? 16 : 2;   // g_tw1024 index step: 2 * 512 / (Ns * 8)
      float2 tw[8];
      {
        const int k0 = lane & (Ns - 1);
#pragma unroll
        for (int r = 1; r < 8; ++r) tw[r] = g_tw1024[r * k0 * tws];
      }
      __syncwarp();   // the previous pass's stores are visible
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float2 a = z[zpad(j + 64 * r)];
          if (r > 0) {
            float2 t = tw[r];
            if (pass == 2 && h == 1) {
              // exp(-2 pi i r / 16), r = 1..7
              constexpr float C16[8] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.f,
                                        -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
              constexpr float S16[8] = {0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f, -1.f,
                                        -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
              t = cmul(t, make_float2(C16[r], S16[r]));
            }
            a = cmul(a, t);
          }
          v[h][r] = a;
        }
        fft8_inplace(v[h]);
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        const int idx = (j / Ns) * Ns * 8 + (j & (Ns - 1));
        z[zpad(idx + 0 * Ns)] = v[h][0]; z[zpad(idx + 1 * Ns)] = v[h][4]; z[zpad(idx + 2 * Ns)] = v[h][2];
        z[zpad(idx + 3 * Ns)] = v[h][6]; z[zpad(idx + 4 * Ns)] = v[h][1]; z[zpad(idx + 5 * Ns)] = v[h][5];
        z[zpad(idx + 6 * Ns)] = v[h][3]; z[zpad(idx + 7 * Ns)] = v[h][7];
      }
    }
    __syncwarp();
    // ---- untangle the packed real transform: bins k and 512-k from Z[k], Z[512-k]
    for (int k = lane; k <= 256; k += 32) {
      if (k == 0) {
        const float2 z0 = z[0];
        mg[0] = fabsf(z0.x + z0.y);
        mg[512] = fabsf(z0.x - z0.y);
      } else {
        const float2 a = z[zpad(k)];
        const float2 c = z[zpad(512 - k)];
        const float2 xe = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));    // (a + conj c)/2
        const float2 xo = make_float2(0.5f * (a.y + c.y), -0.5f * (a.x - c.x));   // (a - conj c)/(2i)
        const float2 t = cmul(g_tw1024[k], xo);
        const float pr = xe.x + t.x, pi = xe.y + t.y;
        const float qr = xe.x - t.x, qi = xe.y - t.y;
        mg[k] = sqrtf(pr * pr + pi * pi);
        mg[512 - k] = sqrtf(qr * qr + qi * qi);
      }
    }
    __syncwarp();
    for (int m = lane; m < n_mels; m += 32) {
      const int o0 = fb_off[m], o1 = fb_off[m + 1], s0 = fb_start[m];
      // four independent partial sums: the single-accumulator form was one dependent LDS -> FMA chain per tap
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float* mgp = mg + s0 - o0;
      int i = o0;
      if (o1 <= 1536) {
        for (; i + 4 <= o1; i += 4) {
          a0 = fmaf(mgp[i], fbw_s[i], a0);
          a1 = fmaf(mgp[i + 1], fbw_s[i + 1], a1);
          a2 = fmaf(mgp[i + 2], fbw_s[i + 2], a2);
          a3 = fmaf(mgp[i + 3], fbw_s[i + 3], a3);
        }
        for (; i < o1; ++i) a0 = fmaf(mgp[i], fbw_s[i], a0);
      } else {
        for (; i < o1; ++i) a0 = fmaf(mgp[i], i < 1536 ? fbw_s[i] : __ldg(fb_w + i), a0);
      }
      const float acc = (a0 + a1) + (a2 + a3);
      otile[m * MEL_FPB + w] = logf(fmaxf(acc, clamp_min));
    }
  }
  __syncthreads();
  const int nf = min(MEL_FPB, F - f0);
  float* ob = out + (int64_t)b * out_sb;
  if (out_sf <= out_sm) {   // frame index is the fast output dim
    for (int i = tid; i < n_mels * MEL_FPB; i += MEL_FPB * 32) {
      const int m = i / MEL_FPB, fl = i - m * MEL_FPB;
      if (fl < nf) ob[(int64_t)m * out_sm + (int64_t)(f0 + fl) * out_sf] = otile[i];
    }
  } else {
    for (int i = tid; i < n_mels * MEL_FPB; i += MEL_FPB * 32) {
      const int fl = i / n_mels, m = i - fl * n_mels;
      if (fl < nf) ob[(int64_t)m * out_sm + (int64_t)(f0 + fl) * out_sf] = otile[m * MEL_FPB + fl];
    }
  }
}

static std::mutex g_mel_mu;
static bool g_mel_ready[64] = {false};

int mel_spectrogram(const float* wav, int64_t wav_sb, int B, int L, const int32_t* lens, const float* window,
                    const float* fb_w, const int32_t* fb_off, const int32_t* fb_start, int n_mels, float clamp_min,
                    float* out, int64_t out_sb, int64_t out_sm, int64_t out_sf, cudaStream_t st) {
  MTTS_REQUIRE(wav && window && fb_w && fb_off && fb_start && out, "null pointer");
  MTTS_REQUIRE(L > MEL_NFFT / 2, "reflect padding needs L > n_fft/2 (torch.stft center=True)");
  MTTS_REQUIRE(n_mels > 0 && n_mels <= 128, "n_mels out of range");
  MTTS_REQUIRE(B >= 0 && B <= 65535 * 32, "bad batch");
  MTTS_REQUIRE((((uintptr_t)window) & 15) == 0, "window table must be 16-byte aligned");
  const int vec_ok = ((((uintptr_t)wav) & 15) == 0) && (wav_sb % 4 == 0);
  if (B == 0) return 0;
  const int F = 1 + L / MEL_HOP;
  const size_t smem = sizeof(float) * ((MEL_NFFT + (MEL_FPB - 1) * MEL_HOP) + MEL_NFFT + 2 * MEL_FPB * MEL_ZPAD +
                                       MEL_FPB * MEL_MAGP + 128 * MEL_FPB + 1536);
  {
    std::lock_guard<std::mutex> lk(g_mel_mu);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return fail(MTTS_ERR_UNSUPPORTED, "%s: device ordinal %lld out of range", "mel", dev);
    if (!g_mel_ready[dev]) {
      cudaError_t e = cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(MTTS_ERR_CUDA, "%s: cudaFuncSetAttribute failed: %lld", "mel", (long long)e);
      launch_k(mel_init_twiddle_kernel, 4, 256, 0, st);
      MTTS_CHECK_LAUNCH();
      cudaStreamSynchronize(st);   // one-time table init only; never on the steady-state path
      g_mel_ready[dev] = true;
    }
  }
  // grid.y is limited to 65535: fold larger batches over several launches
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = (B - b0 < 65535) ? (B - b0) : 65535;
    dim3 grid((unsigned)cdiv64(F, MEL_FPB), (unsigned)nb);
    launch_k(mel_kernel, grid, MEL_FPB * 32, smem, st, wav + (int64_t)b0 * wav_sb, wav_sb, L, lens ? lens + b0 : nullptr, window, fb_w,
                                               fb_off, fb_start, n_mels, clamp_min, out + (int64_t)b0 * out_sb, out_sb, out_sm,
                                               out_sf, vec_ok);
    MTTS_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace mtts
