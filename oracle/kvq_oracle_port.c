/* CPU port of the KVQuant decode hot path  --  TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * Plain C restatement (OpenMP over tokens) of the kernel semantics, each block citing the reference:
 *   K scores : deployment/kvquant/quant_cuda_kernel.cu:3075-3207 (dense, RoPE) + 487-520 (outlier SpMV)
 *   softmax  : deployment/transformers/.../modeling_llama.py:1959-1977 (scale by 1/sqrt(128), fp32 softmax)
 *   V output : quant_cuda_kernel.cu:3238-3419 (dense) + 449-469 (outlier SpMV)
 * cos/sin are hoisted to once per (token, pair) -- the reference kernel evaluates them per (head, channel,
 * token); hoisting only favours this baseline.  Checked against oracle/kvq_oracle.py in tests/test_oracle_port.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define D 128

static inline uint32_t code_of(const uint32_t* w, int64_t stride, int bits, int c) {
  if (bits == 4) return (w[(c >> 3) * stride] >> ((c & 7) * 4)) & 0xFu;
  if (bits == 2) return (w[(c >> 4) * stride] >> ((c & 15) * 2)) & 0x3u;
  {
    const int g = (c >> 5) * 3, l = c & 31;
    if (l < 10) return (w[g * stride] >> (3 * l)) & 7u;
    if (l == 10) return ((w[g * stride] >> 30) | (w[(g + 1) * stride] << 2)) & 7u;
    if (l < 21) return (w[(g + 1) * stride] >> ((3 * l) & 31)) & 7u;
    if (l == 21) return ((w[(g + 1) * stride] >> 31) | (w[(g + 2) * stride] << 1)) & 7u;
    return (w[(g + 2) * stride] >> ((3 * l) & 31)) & 7u;
  }
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline is to use every host core it may run on */
void kvq_port_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int kvq_port_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* scores[h*L + t] (unscaled), float accumulation as in the kernels */
void kvq_port_k_scores(int bits, const float* q, const int32_t* kcache, const float* klut, const float* kout,
                       const int32_t* kidx, int n_out, int H, int64_t Lmax, int64_t L, float theta, int pos_offset,
                       float* scores) {
  const int N = 1 << bits, W = D * bits / 32;
  float th[D / 2];
  for (int j = 0; j < D / 2; ++j) th[j] = powf(theta, (-2 * (float)j) / (float)D);
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < L; ++t) {
    float cs[D / 2], sn[D / 2];
    const int pos = (int)t + pos_offset;
    for (int j = 0; j < D / 2; ++j) { cs[j] = cosf(th[j] * pos); sn[j] = sinf(th[j] * pos); }
    for (int h = 0; h < H; ++h) {
      const uint32_t* w = (const uint32_t*)kcache + (int64_t)h * W * Lmax + t;
      const float* qh = q + h * D;
      const float* lh = klut + (int64_t)h * D * N;
      float res = 0.f;
      for (int c = 0; c < D; ++c) {
        const float kv = lh[c * N + code_of(w, Lmax, bits, c)];
        const int j = c & 63;
        const float sg = c < 64 ? 1.f : -1.f;
        res += kv * cs[j] * qh[c];
        res += sg * kv * sn[j] * qh[(c + 64) & 127];
      }
      scores[(int64_t)h * L + t] = res;
    }
    if (kout) {
      for (int i = 0; i < n_out; ++i) {
        const float v = kout[t * n_out + i];
        const int col = kidx[t * n_out + i];
        const int h = col / D, c = col % D, j = c & 63;
        const float sg = c < 64 ? 1.f : -1.f;
        scores[(int64_t)h * L + t] += v * cs[j] * q[col] + sg * v * sn[j] * q[h * D + ((c + 64) & 127)];
      }
    }
  }
}

/* out[h*128+c] = sum_t p[h*L+t] * V(h,c,t) */
void kvq_port_v_out(int bits, const float* p, const int32_t* vcache, const float* vlut, const float* vout,
                    const int32_t* vidx, int n_out, int H, int64_t Lmax, int64_t L, float* out) {
  const int N = 1 << bits, W = D * bits / 32;
  const int hidden = H * D;
  int nthr = kvq_port_threads();
  float* part = (float*)calloc((size_t)nthr * hidden, sizeof(float));
#pragma omp parallel
  {
#ifdef _OPENMP
    float* acc = part + (size_t)omp_get_thread_num() * hidden;
#else
    float* acc = part;
#endif
#pragma omp for schedule(static)
    for (int64_t t = 0; t < L; ++t) {
      const float* lt = vlut + t * N;
      for (int h = 0; h < H; ++h) {
        const uint32_t* w = (const uint32_t*)vcache + (int64_t)h * W * Lmax + t;
        const float pw = p[(int64_t)h * L + t];
        float* a = acc + h * D;
        for (int c = 0; c < D; ++c) a[c] += lt[code_of(w, Lmax, bits, c)] * pw;
      }
      if (vout) {
        for (int i = 0; i < n_out; ++i) {
          const int row = vidx[t * n_out + i];
          acc[row] += vout[t * n_out + i] * p[(int64_t)(row / D) * L + t];
        }
      }
    }
  }
  for (int j = 0; j < hidden; ++j) {
    float s = 0.f;
    for (int k = 0; k < nthr; ++k) s += part[(size_t)k * hidden + j];
    out[j] = s;
  }
  free(part);
}

/* whole decode-step attention of one layer (no sinks): scores -> /sqrt(128) -> softmax -> V.  scratch: H*L floats */
void kvq_port_attend(int bits, const float* q, const int32_t* kcache, const float* klut, const float* kout,
                     const int32_t* kidx, const int32_t* vcache, const float* vlut, const float* vout,
                     const int32_t* vidx, int n_out, int H, int64_t Lmax, int64_t L, float theta, int pos_offset,
                     float* out, float* scratch) {
  kvq_port_k_scores(bits, q, kcache, klut, kout, kidx, n_out, H, Lmax, L, theta, pos_offset, scratch);
  const float scale = 1.0f / sqrtf((float)D);
#pragma omp parallel for schedule(static)
  for (int h = 0; h < H; ++h) {
    float* s = scratch + (int64_t)h * L;
    float m = -INFINITY;
    for (int64_t t = 0; t < L; ++t) { s[t] *= scale; if (s[t] > m) m = s[t]; }
    double sum = 0.0;
    for (int64_t t = 0; t < L; ++t) { s[t] = expf(s[t] - m); sum += s[t]; }
    const float inv = (float)(1.0 / sum);
    for (int64_t t = 0; t < L; ++t) s[t] *= inv;
  }
  kvq_port_v_out(bits, scratch, vcache, vlut, vout, vidx, n_out, H, Lmax, L, out);
}
