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
  const float* outliers;     // [>=L, n_out] or null
  const int32_t* outlier_idx;
  const float2* rope;        // [64, rope_npos]
  float* gmax;               // [H] or null (fused mode: running max of scaled scores)
  int64_t Lmax, L, out_stride, rope_npos;
  int H, n_out, pos_offset, tiles_per_cta;
  float scale;               // applied to S before store (fused mode); 1 for legacy
  int accumulate;            // 1: out += S (legacy), 0: out = S*scale
};

template <int BITS>
__global__ void __launch_bounds__(KCfg<BITS>::kThreads, 1) k_scores_kernel(const KParams p) {
  using C = KCfg<BITS>;
  constexpr int N = C::N, W = C::W, G = C::G, TT = C::TT, NR = C::NR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_tab = reinterpret_cast<float2*>(smem_raw);           // [G][128][N]
  float* s_q = reinterpret_cast<float*>(s_tab + G * kHeadDim * N);  // [G][128]
  float* s_part = s_q + G * kHeadDim;                            // [G][TT]

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

  const int64_t tile0 = (int64_t)blockIdx.x * p.tiles_per_cta;
  for (int ti = 0; ti < p.tiles_per_cta; ++ti) {
    const int64_t tbase = (tile0 + ti) * TT;
    if (tbase >= p.L) break;
    const int64_t t = tbase + tid;
    const bool live = t < p.L;
    const int ntok = (int)min((int64_t)TT, p.L - tbase);

    // ---- outlier stream of this tile: rows [tbase, tbase+ntok) are contiguous in memory --------------------
    if (p.outliers != nullptr) {
      const int total = ntok * p.n_out;
      const float* ov = p.outliers + tbase * p.n_out;
      const int32_t* oi = p.outlier_idx + tbase * p.n_out;
      for (int e = tid; e < total; e += C::kThreads) {
        const float val = ov[e];
        const int idx = oi[e];
        const int hl = (idx >> 7) - h0;
        if (val != 0.f && hl >= 0 && hl < nh) {
          const int tl = e / p.n_out;
          const int c = idx & (kHeadDim - 1);
          const float2 cs = ld_keep_f2(p.rope + (int64_t)(c & (kHalf - 1)) * p.rope_npos + (tbase + tl + p.pos_offset), pol_keep);
          const float qa = s_q[hl * kHeadDim + c];
          const float qb = s_q[hl * kHeadDim + (c ^ kHalf)];
          const float sg = (c < kHalf) ? 1.f : -1.f;
          atomicAdd(&s_part[hl * TT + tl], val * (cs.x * qa + sg * cs.y * qb));
        }
      }
    }

    // ---- dense part ---------------------------------------------------------------------------------------
    if (live) {
      const uint32_t* col = p.cache + (int64_t)h0 * W * p.Lmax + t;
      const float2* rp = p.rope + (t + p.pos_offset);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float2 cs[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cs[i] = ld_keep_f2(rp + (int64_t)(16 * a + i) * p.rope_npos, pol_keep);
        uint32_t wn[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) wn[i] = ld_stream_u32(col + (int64_t)chunk_row<BITS>(a, i) * p.Lmax, pol_stream);
        for (int hl = 0; hl < nh; ++hl) {
          uint32_t w[W];
#pragma unroll
          for (int i = 0; i < NR; ++i) w[chunk_row<BITS>(a, i)] = wn[i];
          if (hl + 1 < nh) {
            const uint32_t* nxt = col + (int64_t)(hl + 1) * W * p.Lmax;
#pragma unroll
            for (int i = 0; i < NR; ++i) wn[i] = ld_stream_u32(nxt + (int64_t)chunk_row<BITS>(a, i) * p.Lmax, pol_stream);
          }
          const float2* tb = s_tab + hl * (kHeadDim * N);
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = 16 * a + i;
            const float2 e0 = tb[c * N + code_of<BITS>(w, c)];
            const float2 e1 = tb[(c + kHalf) * N + code_of<BITS>(w, c + kHalf)];
            acc0 = fmaf(cs[i].x, e0.x, acc0);
            acc0 = fmaf(cs[i].y, e0.y, acc0);
            acc1 = fmaf(cs[i].x, e1.x, acc1);
            acc1 = fmaf(cs[i].y, e1.y, acc1);
          }
          atomicAdd(&s_part[hl * TT + tid], acc0 + acc1);
        }
      }
    }
    __syncthreads();

    // ---- write back -----------------------------------------------------------------------------------------
    for (int hl = 0; hl < nh; ++hl) {
      float s = s_part[hl * TT + tid];
      s_part[hl * TT + tid] = 0.f;  // own column: ready for the next tile
      if (live) {
        float* o = p.out + (int64_t)(h0 + hl) * p.out_stride + t;
        if (p.accumulate) *o = *o + s;
        else { s *= p.scale; *o = s; }
      }
      if (p.gmax != nullptr) {
        const float m = warp_max(live ? s : -INFINITY);
        if ((tid & 31) == 0 && m > -INFINITY) atomic_max_float(p.gmax + h0 + hl, m);
      }
    }
    __syncthreads();
  }
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
  const size_t smem = (size_t)C::G * kHeadDim * C::N * sizeof(float2) + (size_t)C::G * kHeadDim * 4 + (size_t)C::G * C::TT * 4;
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
                   int64_t Lmax, int64_t L, const float* rope, int64_t rope_npos, int pos_offset, float* gmax,
                   float scale, cudaStream_t st) {
  KParams p{};
  p.q = q; p.cache = reinterpret_cast<const uint32_t*>(cache); p.out = scores; p.lut = lut;
  p.outliers = outliers; p.outlier_idx = outlier_idx; p.rope = reinterpret_cast<const float2*>(rope);
  p.gmax = gmax; p.Lmax = Lmax; p.L = L; p.out_stride = score_stride; p.rope_npos = rope_npos;
  p.H = H; p.n_out = n_out; p.pos_offset = pos_offset; p.scale = scale; p.accumulate = 0;
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
                 const float* rope_cos_sin, int64_t rope_npos, int pos_offset, void* stream) {
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
    const int rc = k_scores_dispatch(bits, p, static_cast<cudaStream_t>(stream));
    if (rc != 0) return rc;
  }
  return 0;
}

}  // extern "C"
