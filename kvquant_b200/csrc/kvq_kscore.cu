// kvquant_b200 -- Q.K^T decode matvec over the packed pre-RoPE key cache, RoPE and the fixed-width outlier
// stream fused into one launch.
//
// Replaces (reference deployment/kvquant/quant_cuda_kernel.cu):
//   VecQuant{4,3,2}MatMulKernelNUQPerChannelTransposedRopeMHABatchedFusedOpt   3040-3209, 3692-4115, 4747-4996
//   SPMV_ATOMIC_ROPE_BALANCED                                                   472-521
//
//   S[h,t] = sum_c (LUT[h,c,code(h,c,t)] (+) outlier(h,c,t)) * (cos(th_j p) q[h,c] + s_c sin(th_j p) q[h,(c+64)%128])
//   j = c % 64, p = t + pos_offset.
//
// Design:
//   * The reference evaluates powf+cosf+sinf per (head, channel, token): 4096 sincos per token per layer make it
//     ALU-bound.  cos/sin depend on (j, p) only, so they come from a table rope[j][p] built ONCE with the
//     reference's own expressions (bit-identical values); a thread loads the 16 pairs it needs for its token and
//     reuses them across all heads of its CTA (G heads) -> 64 table reads per token per CTA.
//   * thread = token (coalesced 128-byte warp loads straight from the sequence-fastest cache rows), per-channel
//     premultiplied tables T[h][c][code] = (LUT*q[h,c], s_c*LUT*q[h,c^64]) in shared memory: 16 (8, 4) entries per
//     channel are 16 distinct consecutive 8-byte slots -> conflict-free multicast for any code pattern.
//   * per element: 1 code extract, 1 LDS.64, 2 FFMA.
#include "kvq_common.cuh"

namespace kvq {

template <int BITS> struct KCfg {
  static constexpr int N = 1 << BITS;
  static constexpr int W = Layout<BITS>::kWords;
  static constexpr int G = 128 / N;            // heads per CTA: 8 / 16 / 32 -> table = G*128*N*8 B = 128 KiB
  static constexpr int kThreads = 512;
  static constexpr int TT = kThreads;          // tokens per tile (thread = token)
  static constexpr int NR = (BITS == 2) ? 2 : 4;  // packed words a thread needs per (head, 16-pair chunk)
};

// packed-word rows (within the head) that hold channels [16a,16a+16) and their +64 partners
template <int BITS> __host__ __device__ constexpr int chunk_row(int a, int i) {
  if constexpr (BITS == 4) { return (i < 2) ? (2 * a + i) : (8 + 2 * a + (i - 2)); }
  else if constexpr (BITS == 2) { return (i == 0) ? a : (4 + a); }
  else { return (i < 2) ? (3 * (a >> 1) + (a & 1) + i) : (3 * ((a >> 1) + 2) + (a & 1) + (i - 2)); }
}

struct KParams {
  const float* q;            // [H,128]
  const uint32_t* cache;     // [H*W, Lmax]
  float* out;                // [H, out_stride]
  const float* lut;          // [H*128, N]
  const float* outliers;     // [>=L, n_out] or null (consumed by k_outlier_kernel, not by the dense kernel)
  const int32_t* outlier_idx;
  const float2* rope;        // [64, rope_npos]
  float* gmax;               // [H] or null (fused mode: running max of scaled scores)
  int64_t Lmax, L, out_stride, rope_npos;
  int H, n_out, pos_offset, tiles_per_cta;
  float scale;               // applied before the store (fused mode: 1/sqrt(128)); 1 for legacy
  int accumulate;            // 1: out = (out + S)*scale, 0: out = S*scale
};

// compile-time loop (immediate LDS offsets and PRMT selectors need constant expressions)
template <int K> struct IC { static constexpr int v = K; };
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (B < E) { f(IC<B>{}); static_for<B + 1, E>(f); }
}
// 8-byte shared load at [addr + IMM] (addr is a 32-bit shared-window address)
template <int IMM>
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm("ld.shared.v2.f32 {%0,%1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "n"(IMM));
  return v;
}
// packed fp32 FMA (sm_100 FFMA2): acc.xy += a.xy * b.xy in one issue slot
__device__ __forceinline__ void ffma2(float2& acc, const float2 a, const float2 b) {
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%0,%1};"
      " fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0,%1}, rc; }"
      : "+f"(acc.x), "+f"(acc.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

// One (head, 16-pair chunk A): 32 table lookups + 32 packed FMAs.  `base` = shared address of the head's table
// (256-byte aligned, so a code*8 byte can be PRMT-ed / OR-ed into its low byte); channel offsets are immediates.
//   4-bit / 2-bit: the codes of a word are pre-masked into byte lanes (4 resp. 8 logic ops per word), then ONE
//                  PRMT per element builds the address           -> PRMT + LDS.64 + FFMA2 per element;
//   3-bit:         one funnel shift + one LOP3 (and-or) per element (straddling codes come for free from the
//                  funnel shift)                                  -> SHF + LOP3 + LDS.64 + FFMA2 per element.
template <int BITS, int A>
__device__ __forceinline__ float k_chunk(const uint32_t* __restrict__ wn, const uint32_t base, const float2* __restrict__ cs) {
  constexpr int N = 1 << BITS;
  float2 acc[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  if constexpr (BITS == 4) {
    // wn[0..1]: channels 16A..16A+15, wn[2..3]: +64
    static_for<0, 4>([&](auto iw) {
      constexpr int wi = decltype(iw)::v;
      constexpr int c0 = 16 * A + (wi & 1) * 8 + (wi >> 1) * kHalf;
      const uint32_t e = (wn[wi] << 3) & 0x78787878u, o = (wn[wi] >> 1) & 0x78787878u;
      static_for<0, 8>([&](auto ik) {
        constexpr int k = decltype(ik)::v;
        const uint32_t addr = __byte_perm((k & 1) ? o : e, base, 0x7650 | (k >> 1));
        ffma2(acc[(wi >> 1) * 2 + (k & 1)], cs[(wi & 1) * 8 + k], lds_f2<(c0 + k) * N * 8>(addr));
      });
    });
  } else if constexpr (BITS == 2) {
    // wn[0]: channels 16A..16A+15, wn[1]: +64
    static_for<0, 2>([&](auto iw) {
      constexpr int wi = decltype(iw)::v;
      constexpr int c0 = 16 * A + wi * kHalf;
      const uint32_t m0 = (wn[wi] << 3) & 0x18181818u, m1 = (wn[wi] << 1) & 0x18181818u;
      const uint32_t m2 = (wn[wi] >> 1) & 0x18181818u, m3 = (wn[wi] >> 3) & 0x18181818u;
      static_for<0, 16>([&](auto ik) {
        constexpr int k = decltype(ik)::v;
        const uint32_t src = (k & 3) == 0 ? m0 : ((k & 3) == 1 ? m1 : ((k & 3) == 2 ? m2 : m3));
        const uint32_t addr = __byte_perm(src, base, 0x7650 | (k >> 2));
        ffma2(acc[wi * 2 + (k & 1)], cs[k], lds_f2<(c0 + k) * N * 8>(addr));
      });
    });
  } else {
    // wn[0..1]: the two words holding locs [16*(A&1), +16) of group A>>1; wn[2..3]: same for group (A>>1)+2
    static_for<0, 2>([&](auto ih) {
      constexpr int hf = decltype(ih)::v;
      const uint32_t w0 = wn[2 * hf], w1 = wn[2 * hf + 1];
      static_for<0, 16>([&](auto ik) {
        constexpr int k = decltype(ik)::v;
        constexpr int l = 16 * (A & 1) + k;               // position within the 32-channel group
        constexpr int bitpos = 3 * l - 32 * ((A & 1) ? 1 : 0);  // bit offset relative to w0 (may be negative / >= 32)
        constexpr int c = 16 * A + k + hf * kHalf;
        uint32_t x;
        if constexpr (bitpos < 0) {
          // (A&1)==1 and the code starts in the previous word: cannot happen for l >= 16 (3*16 = 48 >= 32)
          x = 0;
        } else if constexpr (bitpos + 3 <= 32) {
          x = (bitpos >= 3) ? (w0 >> (bitpos - 3)) : (w0 << (3 - bitpos));
        } else if constexpr (bitpos < 32) {
          x = __funnelshift_r(w0, w1, bitpos - 3);       // straddles w0 / w1
        } else {
          x = (bitpos - 32 >= 3) ? (w1 >> (bitpos - 32 - 3)) : (w1 << (3 - (bitpos - 32)));
        }
        const uint32_t addr = (x & 0x38u) | base;
        ffma2(acc[hf * 2 + (k & 1)], cs[k], lds_f2<c * N * 8>(addr));
      });
    });
  }
  return (acc[0].x + acc[0].y) + (acc[1].x + acc[1].y) + (acc[2].x + acc[2].y) + (acc[3].x + acc[3].y);
}

template <int BITS>
__global__ void __launch_bounds__(KCfg<BITS>::kThreads, 1) k_scores_kernel(const KParams p) {
  using C = KCfg<BITS>;
  constexpr int N = C::N, W = C::W, G = C::G, TT = C::TT, NR = C::NR;
  extern __shared__ unsigned char smem_raw[];
  // table base must be 256-byte aligned (the low address byte carries code*8)
  unsigned char* smem = smem_raw + ((256u - (smem_u32(smem_raw) & 255u)) & 255u);
  float2* s_tab = reinterpret_cast<float2*>(smem);                 // [G][128][N]
  float* s_q = reinterpret_cast<float*>(s_tab + G * kHeadDim * N);  // [G][128]
  float* s_part = s_q + G * kHeadDim;                              // [G][TT]

  const int tid = threadIdx.x;
  const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
  const int h0 = blockIdx.y * G;
  const int nh = min(G, p.H - h0);

  // ---- premultiplied tables -------------------------------------------------------------------------------
  for (int i = tid; i < nh * kHeadDim; i += C::kThreads) s_q[i] = p.q[(int64_t)h0 * kHeadDim + i];
  __syncthreads();
  for (int i = tid; i < nh * kHeadDim * N; i += C::kThreads) {
    const int hc = i / N;             // hl*128 + c
    const int c = hc & (kHeadDim - 1);
    const float l = p.lut[((int64_t)h0 * kHeadDim) * N + i];
    const float qa = s_q[hc];
    const float qb = s_q[hc ^ kHalf];  // (c+64)%128 within the same head
    s_tab[i] = make_float2(l * qa, (c < kHalf) ? (l * qb) : -(l * qb));
  }
  for (int i = tid; i < G * TT; i += C::kThreads) s_part[i] = 0.f;
  __syncthreads();
  const uint32_t tab0 = smem_u32(s_tab);

  const int64_t tile0 = (int64_t)blockIdx.x * p.tiles_per_cta;
  for (int ti = 0; ti < p.tiles_per_cta; ++ti) {
    const int64_t tbase = (tile0 + ti) * TT;
    if (tbase >= p.L) break;
    const int64_t t = tbase + tid;
    const bool live = t < p.L;

    // ---- dense part ---------------------------------------------------------------------------------------
    if (live) {
      const uint32_t* col = p.cache + (int64_t)h0 * W * p.Lmax + t;
      const float2* rp = p.rope + (t + p.pos_offset);
      static_for<0, 4>([&](auto ia) {
        constexpr int a = decltype(ia)::v;
        float2 cs[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cs[i] = ld_keep_f2(rp + (int64_t)(16 * a + i) * p.rope_npos, pol_keep);
        uint32_t wn[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) wn[i] = ld_stream_u32(col + (int64_t)chunk_row<BITS>(a, i) * p.Lmax, pol_stream);
        for (int hl = 0; hl < nh; ++hl) {
          uint32_t w[NR];
#pragma unroll
          for (int i = 0; i < NR; ++i) w[i] = wn[i];
          if (hl + 1 < nh) {
            const uint32_t* nxt = col + (int64_t)(hl + 1) * W * p.Lmax;
#pragma unroll
            for (int i = 0; i < NR; ++i) wn[i] = ld_stream_u32(nxt + (int64_t)chunk_row<BITS>(a, i) * p.Lmax, pol_stream);
          }
          const float part = k_chunk<BITS, a>(w, tab0 + (uint32_t)hl * (kHeadDim * N * 8), cs);
          atomicAdd(&s_part[hl * TT + tid], part);
        }
      });
    }
    __syncthreads();

    // ---- write back -----------------------------------------------------------------------------------------
    for (int hl = 0; hl < nh; ++hl) {
      float s = s_part[hl * TT + tid];
      s_part[hl * TT + tid] = 0.f;  // own column: ready for the next tile
      if (live) {
        float* o = p.out + (int64_t)(h0 + hl) * p.out_stride + t;
        if (p.accumulate) s += *o;   // legacy: mul += S;  sparse: the outlier pre-pass already deposited its part
        s *= p.scale;
        *o = s;
      }
      if (p.gmax != nullptr) {
        const float m = warp_max(live ? s : -INFINITY);
        if ((tid & 31) == 0 && m > -INFINITY) atomic_max_float(p.gmax + h0 + hl, m);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------
// Outlier pre-pass (replaces SPMV_ATOMIC_ROPE_BALANCED, quant_cuda_kernel.cu:472-521): thread = token, walks the
// token's n_out (value, channel) pairs, RoPE evaluated with the reference's own expressions (theta from a 64-entry
// powf table, cosf/sinf of theta*pos) -- 42 sincos per token instead of the dense path's former 4096.  The row is
// sorted by channel, so contributions of one head are consecutive: they are summed in a register and written once
// per head, without atomics (the thread owns column t of the score matrix in this launch).
// store_all = 1: every head's entry is written (value or 0)  -> initialises the fused score buffer;
// store_all = 0: only heads with outliers are touched, out += contribution (legacy accumulate semantics).
// ------------------------------------------------------------------------------------------------------------
constexpr int kOutThreads = 128;

__global__ void __launch_bounds__(kOutThreads) k_outlier_kernel(
    const float* __restrict__ q, const float* __restrict__ outliers, const int32_t* __restrict__ outlier_idx,
    float* __restrict__ out, int64_t out_stride, int64_t L, int H, int n_out, float rope_theta, int pos_offset,
    int store_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_q = reinterpret_cast<float*>(smem_raw);            // [H*128]
  float* s_theta = s_q + H * kHeadDim;                         // [64]
  float* s_val = s_theta + kHalf;                              // [128][n_out+1]
  int32_t* s_idx = reinterpret_cast<int32_t*>(s_val + kOutThreads * (n_out + 1));
  const int tid = threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.x * kOutThreads;
  const int ntok = (int)min((int64_t)kOutThreads, L - t0);
  for (int i = tid; i < H * kHeadDim; i += kOutThreads) s_q[i] = q[i];
  if (tid < kHalf) {
    const int headdim = kHeadDim;
    s_theta[tid] = powf(rope_theta, (-2 * __int2float_rd(tid % (headdim / 2)) / headdim));  // DK.cu:504
  }
  const int total = ntok * n_out;
  const int stride = n_out + 1;
  for (int e = tid; e < total; e += kOutThreads) {  // rows [t0, t0+ntok) are contiguous: coalesced
    const int r = e / n_out, k = e - r * n_out;
    s_val[r * stride + k] = outliers[t0 * n_out + e];
    s_idx[r * stride + k] = outlier_idx[t0 * n_out + e];
  }
  __syncthreads();
  if (tid >= ntok) return;
  const int64_t t = t0 + tid;
  const int pos = (int)t + pos_offset;
  float* ocol = out + t;
  int next_h = 0;       // store_all: next head that still has to be written
  int cur_h = -1;
  float acc = 0.f;
  for (int k = 0; k < n_out; ++k) {
    const float v = s_val[tid * stride + k];
    if (v == 0.f) continue;  // pads / non-outliers contribute exactly 0 in the reference too
    const int col = s_idx[tid * stride + k];
    const int h = col >> 7, c = col & (kHeadDim - 1);
    if (h != cur_h) {
      if (cur_h >= 0) {
        if (store_all) {
          for (; next_h < cur_h; ++next_h) ocol[(int64_t)next_h * out_stride] = 0.f;
          ocol[(int64_t)cur_h * out_stride] = acc;
          next_h = cur_h + 1;
        } else {
          ocol[(int64_t)cur_h * out_stride] += acc;
        }
      }
      cur_h = h;
      acc = 0.f;
    }
    const float theta = s_theta[c & (kHalf - 1)];
    const float sign = (c < kHalf) ? 1.f : -1.f;
    const float cs = cosf(theta * pos);
    const float sn = sinf(theta * pos);
    float dot = v * cs * s_q[col];
    dot += sign * v * sn * s_q[col ^ kHalf];
    acc += dot;
  }
  if (cur_h >= 0) {
    if (store_all) {
      for (; next_h < cur_h; ++next_h) ocol[(int64_t)next_h * out_stride] = 0.f;
      ocol[(int64_t)cur_h * out_stride] = acc;
      next_h = cur_h + 1;
    } else {
      ocol[(int64_t)cur_h * out_stride] += acc;
    }
  }
  if (store_all)
    for (; next_h < H; ++next_h) ocol[(int64_t)next_h * out_stride] = 0.f;
}

static int launch_k_outliers(const KParams& p, float rope_theta, int store_all, cudaStream_t st) {
  const size_t smem = (size_t)p.H * kHeadDim * 4 + kHalf * 4 + (size_t)kOutThreads * (p.n_out + 1) * 8;
  if (smem > 200 * 1024) return KVQ_E_UNSUPPORTED;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(k_outlier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr = smem;
  }
  const unsigned grid = (unsigned)((p.L + kOutThreads - 1) / kOutThreads);
  k_outlier_kernel<<<grid, kOutThreads, smem, st>>>(p.q, p.outliers, p.outlier_idx, p.out, p.out_stride, p.L, p.H,
                                                    p.n_out, rope_theta, p.pos_offset, store_all);
  KVQ_LAUNCH_CHECK();
  return 0;
}

// rope table: reference expressions quant_cuda_kernel.cu:3081 (theta) and 3123-3126 (cos/sin), once per (j, p)
__global__ void rope_table_kernel(float2* __restrict__ out, float rope_theta, int64_t n_pos) {
  const int j = blockIdx.y;
  const int headdim = kHeadDim;
  const int headdim2 = headdim / 2;
  const float theta = powf(rope_theta, (-2 * __int2float_rd(j % headdim2) / __int2float_rd(headdim)));
  for (int64_t pos64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pos64 < n_pos; pos64 += (int64_t)gridDim.x * blockDim.x) {
    const int pos = (int)pos64;
    const float c = cosf(theta * pos);
    const float s = sinf(theta * pos);
    out[(int64_t)j * n_pos + pos64] = make_float2(c, s);
  }
}

template <int BITS>
static int launch_k_scores(const KParams& p, cudaStream_t st) {
  using C = KCfg<BITS>;
  const size_t smem = 256 + (size_t)C::G * kHeadDim * C::N * sizeof(float2) + (size_t)C::G * kHeadDim * 4 + (size_t)C::G * C::TT * 4;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k_scores_kernel<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  const int n_groups = (p.H + C::G - 1) / C::G;
  const int64_t n_tiles = (p.L + C::TT - 1) / C::TT;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t max_splits = sms / n_groups > 0 ? sms / n_groups : 1;
  KParams q = p;
  q.tiles_per_cta = (int)((n_tiles + max_splits - 1) / max_splits);
  const int64_t splits = (n_tiles + q.tiles_per_cta - 1) / q.tiles_per_cta;
  const dim3 grid((unsigned)splits, (unsigned)n_groups);
  k_scores_kernel<BITS><<<grid, C::kThreads, smem, st>>>(q);
  KVQ_LAUNCH_CHECK();
  return 0;
}

int k_scores_dispatch(int bits, const KParams& p, cudaStream_t st) {
  switch (bits) {
    case 4: return launch_k_scores<4>(p, st);
    case 3: return launch_k_scores<3>(p, st);
    case 2: return launch_k_scores<2>(p, st);
    default: return KVQ_E_BITS;
  }
}

int k_scores_fused(int bits, const float* q, const int32_t* cache, float* scores, int64_t score_stride,
                   const float* lut, const float* outliers, const int32_t* outlier_idx, int n_out, int H,
                   int64_t Lmax, int64_t L, const float* rope, int64_t rope_npos, float theta, int pos_offset,
                   float* gmax, float scale, cudaStream_t st) {
  KParams p{};
  p.q = q; p.cache = reinterpret_cast<const uint32_t*>(cache); p.out = scores; p.lut = lut;
  p.outliers = outliers; p.outlier_idx = outlier_idx; p.rope = reinterpret_cast<const float2*>(rope);
  p.gmax = gmax; p.Lmax = Lmax; p.L = L; p.out_stride = score_stride; p.rope_npos = rope_npos;
  p.H = H; p.n_out = n_out; p.pos_offset = pos_offset; p.scale = scale; p.accumulate = 0;
  if (outliers != nullptr) {
    const int rc = launch_k_outliers(p, theta, /*store_all=*/1, st);  // initialises the score buffer
    if (rc != 0) return rc;
    p.accumulate = 1;
  }
  return k_scores_dispatch(bits, p, st);
}

}  // namespace kvq

using namespace kvq;

extern "C" {

int kvq_rope_table_build(float* rope_cos_sin, float theta, int64_t n_pos, void* stream) {
  if (!rope_cos_sin) return KVQ_E_NULL;
  if (n_pos <= 0 || n_pos > (int64_t)1 << 30) return KVQ_E_SHAPE;
  const dim3 grid((unsigned)((n_pos + 255) / 256 > 4096 ? 4096 : (n_pos + 255) / 256), kHalf);
  rope_table_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float2*>(rope_cos_sin), theta, n_pos);
  KVQ_LAUNCH_CHECK();
  return 0;
}

int kvq_k_matvec(int bits, const float* q, const int32_t* cache, float* mul, const float* lut, int B, int H,
                 int64_t Lmax, int64_t L, const float* outliers, const int32_t* outlier_idx, int n_out,
                 const float* rope_cos_sin, int64_t rope_npos, float theta, int pos_offset, void* stream) {
  if (!q || !cache || !mul || !lut || !rope_cos_sin) return KVQ_E_NULL;
  if (B <= 0 || H <= 0 || L < 0 || L > Lmax || pos_offset < 0) return KVQ_E_SHAPE;
  if ((outliers == nullptr) != (outlier_idx == nullptr)) return KVQ_E_NULL;
  if (outliers && (B != 1 || n_out <= 0)) return KVQ_E_SHAPE;  // reference: sparse part is batch-1 only (DK.cu:3605)
  if (rope_npos < L + pos_offset) return KVQ_E_SHAPE;
  if (L == 0) return 0;
  for (int b = 0; b < B; ++b) {
    KParams p{};
    p.q = q + (int64_t)b * H * kHeadDim;
    p.cache = reinterpret_cast<const uint32_t*>(cache);
    p.out = mul + (int64_t)b * H * L;
    p.lut = lut;
    p.outliers = outliers;
    p.outlier_idx = outlier_idx;
    p.rope = reinterpret_cast<const float2*>(rope_cos_sin);
    p.gmax = nullptr;
    p.Lmax = Lmax; p.L = L; p.out_stride = L; p.rope_npos = rope_npos;
    p.H = H; p.n_out = n_out; p.pos_offset = pos_offset;
    p.scale = 1.f; p.accumulate = 1;
    if (outliers != nullptr) {
      const int rc0 = launch_k_outliers(p, theta, /*store_all=*/0, static_cast<cudaStream_t>(stream));
      if (rc0 != 0) return rc0;
    }
    const int rc = k_scores_dispatch(bits, p, static_cast<cudaStream_t>(stream));
    if (rc != 0) return rc;
  }
  return 0;
}

}  // extern "C"
