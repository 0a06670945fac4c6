/* Plain-C restatement of the integer-valued step of the path: the VQ nearest-code search
 * EuclideanCodebook.quantize (reference modules/quantization/core_vq.py:175-183):
 *     dist = -( sum(x^2) - 2 x.e_k + sum(e_k^2) ),  idx = first argmax_k dist
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): built by __graft_entry__.build() into
 * oracle/_build/libvq_oracle.so and used by tests as a second, torch-free checker.
 * fp32 accumulation in index order, like a naive reading of the reference expression. */
#include <stdint.h>

void vq_argmin_oracle(const float* x, const float* embed, int64_t n, int d, int k, int64_t* idx) {
  for (int64_t i = 0; i < n; ++i) {
    const float* xi = x + i * d;
    float xx = 0.f;
    for (int c = 0; c < d; ++c) xx += xi[c] * xi[c];
    float best = 0.f;
    int64_t bk = -1;
    for (int j = 0; j < k; ++j) {
      const float* e = embed + (int64_t)j * d;
      float dot = 0.f, ee = 0.f;
      for (int c = 0; c < d; ++c) { dot += xi[c] * e[c]; ee += e[c] * e[c]; }
      const float dist = -((xx - 2.0f * dot) + ee);
      if (bk < 0 || dist > best) { best = dist; bk = j; }   /* strict > keeps the first maximum */
    }
    idx[i] = bk;
  }
}
