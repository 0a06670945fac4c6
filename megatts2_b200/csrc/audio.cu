// Host-side audio front / back ends of Megatts.forward moved onto the device:
//   resample to 16 kHz      librosa.load(wav, sr=16000)        models/megatts2.py:335   (prepare_ds.py:113)
//   peak normalise          librosa.util.normalize(y)          models/megatts2.py:336   (prepare_ds.py:124)
//   waveform -> PCM         torchaudio.save('test.wav', ...)   models/megatts2.py:375
// The resampler is the band-limited polyphase FIR of torchaudio.functional.resample (the reference pins torchaudio; librosa
// and its soxr backend are un-vendored and unpinned - SURVEY.md 8c - so the kernel follows the published torchaudio
// definition, with the filter table built on the host in fp64 exactly as torchaudio builds it):
//   y[i * up + p] = sum_{k < taps} xpad[i * down + k] * h[p][k],   xpad = x shifted by `width` zeros on the left
// HBM-bound elementwise / FIR work; nothing here belongs on the tensor cores.
#include <float.h>

#include "kernels.h"

namespace mtts {

// CTA = RS_FR input strides (frames) of one clip: the frames' input span is staged once in shared memory; thread (p, f)
// walks phase p of frame f.  h is (up, taps) row-major and stays L2-resident.
constexpr int RS_FR = 32;
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ x, int64_t x_sb, const int32_t* __restrict__ lens_in, int L_in,
                const float* __restrict__ h, int up, int down, int width, int taps, float* __restrict__ y, int64_t y_sb,
                int L_out, const int32_t* __restrict__ lens_out) {
  pdl_entry();
  extern __shared__ float xs[];                      // RS_FR * down + taps
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * RS_FR;
  const int Lb = lens_in ? min(lens_in[b], L_in) : L_in;
  const int Lo = lens_out ? min(lens_out[b], L_out) : L_out;
  const int span = RS_FR * down + taps;
  const int64_t g0 = (int64_t)i0 * down - width;     // first input sample of the span
  const float* xb = x + (int64_t)b * x_sb;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int64_t g = g0 + i;
    xs[i] = (g >= 0 && g < Lb) ? xb[g] : 0.f;
  }
  __syncthreads();
  float* yb = y + (int64_t)b * y_sb;
  for (int o = threadIdx.x; o < RS_FR * up; o += blockDim.x) {
    const int f = o / up, p = o - f * up;
    const int64_t n = (int64_t)(i0 + f) * up + p;
    if (n >= L_out) continue;
    float acc = 0.f;
    if (n < Lo) {
      const float* hp = h + (int64_t)p * taps;
      const float* xp = xs + f * down;
      // four partial sums: the taps are independent FMAs (the order is fixed, so results are run-to-run identical)
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int k = 0;
      for (; k + 4 <= taps; k += 4) {
        a0 = fmaf(xp[k], __ldg(hp + k), a0);
        a1 = fmaf(xp[k + 1], __ldg(hp + k + 1), a1);
        a2 = fmaf(xp[k + 2], __ldg(hp + k + 2), a2);
        a3 = fmaf(xp[k + 3], __ldg(hp + k + 3), a3);
      }
      for (; k < taps; ++k) a0 = fmaf(xp[k], __ldg(hp + k), a0);
      acc = (a0 + a1) + (a2 + a3);
    }
    yb[n] = acc;                                     // samples past the clip's own output length are zero
  }
}

int resample(const float* x, int64_t x_sb, int B, int L_in, const int32_t* lens_in, const float* h, int up, int down,
             int width, int taps, float* y, int64_t y_sb, int L_out, const int32_t* lens_out, cudaStream_t st) {
  MTTS_REQUIRE(x && h && y && up > 0 && down > 0 && width >= 0 && taps > 0, "bad arguments");
  MTTS_REQUIRE(B >= 0 && B <= 65535 && L_in >= 0 && L_out >= 0, "bad sizes");
  if (B == 0 || L_out == 0) return 0;
  const size_t smem = sizeof(float) * ((size_t)RS_FR * down + taps);
  MTTS_REQUIRE(smem <= 200 * 1024, "ratio too large for the staged span (down * 32 + taps floats)");
  static std::atomic<uint64_t> configured{0};
  const int dev = cur_device();
  if (smem > 48 * 1024 && !(configured.load(std::memory_order_relaxed) & (1ull << dev))) {
    cudaError_t e = cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail(MTTS_ERR_CUDA, "%s: cudaFuncSetAttribute failed: %lld", "resample", (long long)e);
    configured.fetch_or(1ull << dev, std::memory_order_relaxed);
  }
  const int64_t frames = cdiv64(L_out, up);
  dim3 grid((unsigned)cdiv64(frames, RS_FR), (unsigned)B);
  launch_k(resample_kernel, grid, 256, smem, st, x, x_sb, lens_in, L_in, h, up, down, width, taps, y, y_sb, L_out, lens_out);
  MTTS_CHECK_LAUNCH();
  return 0;
}

// ---- peak normalisation: x / max|x| per clip (librosa.util.normalize, norm = inf); clips whose peak is below the
// smallest normal float are left unchanged, like librosa's threshold = tiny
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, int64_t x_sb, int L, const int32_t* __restrict__ lens, uint32_t* __restrict__ peak) {
  pdl_entry();
  const int b = blockIdx.y;
  const int Lb = lens ? min(lens[b], L) : L;
  const float* xb = x + (int64_t)b * x_sb;
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Lb; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(xb[i]));
  m = warp_max(m);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sm[w]);
    atomicMax(peak + b, __float_as_uint(m));         // non-negative floats order like their bit patterns
  }
}
__global__ void __launch_bounds__(256)
peak_scale_kernel(float* __restrict__ x, int64_t x_sb, int L, const int32_t* __restrict__ lens, const uint32_t* __restrict__ peak) {
  pdl_entry();
  const int b = blockIdx.y;
  const int Lb = lens ? min(lens[b], L) : L;
  const float pk = __uint_as_float(peak[b]);
  if (!(pk >= FLT_MIN)) return;                      // librosa: length < tiny -> divide by 1
  float* xb = x + (int64_t)b * x_sb;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Lb; i += (int64_t)gridDim.x * blockDim.x)
    xb[i] = __fdiv_rn(xb[i], pk);
}
__global__ void zero_u32_kernel(uint32_t* p, int n) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0u;
}

int peak_normalize(float* x, int64_t x_sb, int B, int L, const int32_t* lens, void* scratch, cudaStream_t st) {
  MTTS_REQUIRE(x && scratch && B >= 0 && B <= 65535 && L >= 0, "bad arguments");
  if (B == 0 || L == 0) return 0;
  uint32_t* peak = reinterpret_cast<uint32_t*>(scratch);          // B words
  launch_k(zero_u32_kernel, (unsigned)cdiv64(B, 256), 256, 0, st, peak, B);
  MTTS_CHECK_LAUNCH();
  const unsigned gx = (unsigned)(cdiv64(L, 256 * 8) < 64 ? cdiv64(L, 256 * 8) : 64);
  dim3 grid(gx ? gx : 1, (unsigned)B);
  launch_k(absmax_kernel, grid, 256, 0, st, (const float*)x, x_sb, L, lens, peak);
  MTTS_CHECK_LAUNCH();
  launch_k(peak_scale_kernel, grid, 256, 0, st, x, x_sb, L, lens, (const uint32_t*)peak);
  MTTS_CHECK_LAUNCH();
  return 0;
}

// ---- fp32 samples -> 16-bit PCM (the optional integer encoding of the wav writer): round-to-nearest-even of x * 32768,
// saturated to [-32768, 32767]
__global__ void pcm16_kernel(const float* __restrict__ x, int64_t n, int16_t* __restrict__ out) {
  pdl_entry();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i] * 32768.0f;
  v = fminf(fmaxf(v, -32768.0f), 32767.0f);
  out[i] = (int16_t)__float2int_rn(v);
}
int pcm16(const float* x, int64_t n, int16_t* out, cudaStream_t st) {
  MTTS_REQUIRE(x && out && n >= 0, "bad arguments");
  if (n == 0) return 0;
  launch_k(pcm16_kernel, (unsigned)cdiv64(n, 256), 256, 0, st, x, n, out);
  MTTS_CHECK_LAUNCH();
  return 0;
}

}  // namespace mtts

using namespace mtts;
extern "C" {
int mtts_resample_f32(const float* x, int64_t x_sb, int32_t B, int32_t L_in, const int32_t* lens_in, const float* h,
                      int32_t up, int32_t down, int32_t width, int32_t taps, float* y, int64_t y_sb, int32_t L_out,
                      const int32_t* lens_out, void* stream) {
  return resample(x, x_sb, B, L_in, lens_in, h, up, down, width, taps, y, y_sb, L_out, lens_out, (cudaStream_t)stream);
}
int mtts_peak_normalize_f32(float* x, int64_t x_sb, int32_t B, int32_t L, const int32_t* lens, void* scratch, void* stream) {
  return peak_normalize(x, x_sb, B, L, lens, scratch, (cudaStream_t)stream);
}
int mtts_pcm16_f32(const float* x, int64_t n, int16_t* out, void* stream) { return pcm16(x, n, out, (cudaStream_t)stream); }
}
