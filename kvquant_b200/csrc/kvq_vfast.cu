// kvquant_b200 -- score.V kernel of the fused attend, warp-specialised form (default; kvq_vnative.cu is the fallback).
//
// Replaces (reference deployment/kvquant/quant_cuda_kernel.cu):
//   VecQuant{4,3,2}MatMulKernelNUQPerChannelTransposedMHABatchedFusedOpt   3211-3433 (+3/2-bit)
//   SPMV_ATOMIC_BALANCED                                                    436-470
// and the softmax between the two matvecs (modeling_llama.py:1958-1977).
//
//     O[h,c] = sum_t w[h,t] (cent[code(h,c,t)] sf_t + off_t  (+) outlier)  =  sum_t (w sf_t) cent[code] + sum_t w off_t + outliers
//
// Same arithmetic and data movement as v_native_kernel (TMA slab ring, thread = packed word row, two codes per
// lookup in a private pair table).  What changed is who does what (ncu on v_native: ~50 % of its shared-memory /
// L1 wavefronts were the outlier reductions, `red.global.add.f32` per entry):
//   * warps 0-15  dense lookups only.
//   * warps 16-17 outliers: each owns a PRIVATE fp32 accumulator row in shared memory and walks the tile's
//     (value, index) stream 32 entries at a time with plain load / add / store -- race-free by construction (one
//     warp per row; equal indices inside a 32-entry chunk are merged with match.any first), no atomics, no global
//     clear / fence / read-back protocol, deterministic.  The rows live in the 128-byte holes of the pair table
//     (its rows are 256 bytes apart because PRMT builds `index * 256`, but only 128 bytes of a row are table) and in
//     whatever shared memory is left.
//   * warps 18-19 softmax weights of the NEXT tile (half of the heads each): exp(s - max), w sf_t, denominators and the
//     sum_t w off_t terms in registers (lane = token), and the TMA issue.
// HALF = true (the fp16 mode of kvq_attend): the pair table holds half2 {cent[lo], cent[hi]} (one 4-byte lookup = one
// wavefront per 64 elements instead of two) and the weights are fp16 (scaled by 2^8 against underflow); products are
// exact and accumulate in fp32 (FHFMA).  HALF = false: fp32 table and weights, bit-compatible with v_native.
#include "kvq_common.cuh"
#include <cuda_fp16.h>

namespace kvq {

constexpr int kVFCompute = 512;                    // compute threads (16 warps)
constexpr int kVFThreads = 640;                    // + warps 16..19
constexpr int kVFT = 32;                           // tokens per stage
constexpr int kVFMaxStages = 3;
constexpr int kVFMaxAcc = 2;
constexpr int kVFPre = 32;                         // outlier steps (token, 32-entry part) a warp prefetches per tile
constexpr float kVFHalfScale = 256.f;
constexpr uint32_t kVFSmemBudget = 227u * 1024u;

struct VFParams {
  const float* score;        // [H, score_stride] scaled scores
  const float* gmax;         // [H]
  const float* v_cent;       // [N]
  const float* v_aff;        // [Lmax][2]
  float* out_o;              // [n_cta][H][128]
  float* out_l;              // [n_cta][H]
  const float* outliers;
  const int32_t* outlier_idx;
  int64_t Lmax, L, score_stride;
  const int64_t* len_dev;
  int64_t len_add;
  int H, n_out, tiles_per_cta, n_stages, box_rows;
  int n_acc;                               // private outlier accumulators (= outlier warps)
  uint32_t acc_off[kVFMaxAcc];             // byte offset of accumulator a (from the aligned smem base)
  uint32_t acc_stride[kVFMaxAcc];          // bytes between consecutive 32-float rows (128 packed, 256 in table holes)
  uint32_t off_tab, off_w, off_ws, off_red, off_bar;
  uint32_t n_out_magic;                    // (unused)
};

template <int BITS> struct VFCfg {
  static constexpr int N = 1 << BITS;
  static constexpr int W = Layout<BITS>::kWords;
  static constexpr int IDXBITS = (BITS == 4) ? 8 : (BITS == 3 ? 6 : 4);
  static constexpr int TABN = 1 << IDXBITS;
  static constexpr int NP = (BITS == 3) ? 6 : 4;
};

__device__ __forceinline__ float2 vf_lds_f2(uint32_t addr) {
  float2 v;
  asm("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t vf_lds_u32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void vf_ffma2(float2& acc, const float2 a, const float2 b) {
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%0,%1};"
      " fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0,%1}, rc; }"
      : "+f"(acc.x), "+f"(acc.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}
// acc.x += e.lo * w.{lo|hi}, acc.y += e.hi * w.{lo|hi}   (e = half2 table entry, w = two tokens' fp16 weights)
template <int WH> __device__ __forceinline__ void vf_fh2(float2& acc, uint32_t e, uint32_t w) {
  if constexpr (WH == 0)
    asm("{ .reg .b16 el, eh, wl, wh; mov.b32 {el,eh}, %2; mov.b32 {wl,wh}, %3;"
        " fma.rn.f32.f16 %0, el, wl, %0; fma.rn.f32.f16 %1, eh, wl, %1; }" : "+f"(acc.x), "+f"(acc.y) : "r"(e), "r"(w));
  else
    asm("{ .reg .b16 el, eh, wl, wh; mov.b32 {el,eh}, %2; mov.b32 {wl,wh}, %3;"
        " fma.rn.f32.f16 %0, el, wh, %0; fma.rn.f32.f16 %1, eh, wh, %1; }" : "+f"(acc.x), "+f"(acc.y) : "r"(e), "r"(w));
}

// one lookup: idxoff = pair index * 256 (already positioned), tab = lane's column of the table
template <bool HALF, int WH>
__device__ __forceinline__ void vf_look(float2& acc, uint32_t addr, float wf, uint32_t wh) {
  if constexpr (HALF) vf_fh2<WH>(acc, vf_lds_u32(addr), wh);
  else vf_ffma2(acc, make_float2(wf, wf), vf_lds_f2(addr));
}

// 32 tokens of one unit (same unit geometry as vn_tile_unit, kvq_vnative.cu)
template <int BITS, int SUB, bool HALF>
__device__ __forceinline__ void vf_tile_unit(const unsigned char* stage, uint32_t row_off, uint32_t swz,
                                             uint32_t row_off2, uint32_t swz2, int part, uint32_t tab,
                                             const unsigned char* __restrict__ wsrow, float2* __restrict__ acc) {
#pragma unroll 2
  for (int q = 0; q < 8; ++q) {
    const uint4 wa = *reinterpret_cast<const uint4*>(stage + row_off + ((q ^ swz) << 4));
    uint4 wb = make_uint4(0, 0, 0, 0);
    if constexpr (BITS == 3 && SUB < 2) wb = *reinterpret_cast<const uint4*>(stage + row_off2 + ((q ^ swz2) << 4));
    float wsv[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t wsh[2] = {0u, 0u};
    if constexpr (HALF) {
      const uint2 h2 = *reinterpret_cast<const uint2*>(wsrow + 8 * q);      // 4 tokens x fp16
      wsh[0] = h2.x; wsh[1] = h2.y;
    } else {
      const float4 ws4 = *reinterpret_cast<const float4*>(wsrow + 16 * q);
      wsv[0] = ws4.x; wsv[1] = ws4.y; wsv[2] = ws4.z; wsv[3] = ws4.w;
    }
    const uint32_t wav[4] = {wa.x, wa.y, wa.z, wa.w};
    const uint32_t wbv[4] = {wb.x, wb.y, wb.z, wb.w};
    static_assert(true, "");
#pragma unroll
    for (int tt = 0; tt < 4; ++tt) {
      const float wf = wsv[tt];
      const uint32_t wh = wsh[tt >> 1];
      auto look = [&](float2& a, uint32_t addr) {
        if (tt & 1) vf_look<HALF, 1>(a, addr, wf, wh); else vf_look<HALF, 0>(a, addr, wf, wh);
      };
      if constexpr (BITS == 4) {
        const uint32_t w = wav[tt];
#pragma unroll
        for (int b = 0; b < 4; ++b) look(acc[b], __byte_perm(w, 0u, 0x4404 | (b << 4)) + tab);
      } else if constexpr (BITS == 2) {
        const uint32_t w = wav[tt] >> (16 * part);
#pragma unroll
        for (int b = 0; b < 4; ++b)
          look(acc[b], (b < 2 ? ((w << (8 - 4 * b)) & 0xF00u) : ((w >> (4 * b - 8)) & 0xF00u)) + tab);
      } else {
        const uint32_t w = wav[tt];
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          const int s = SUB + 6 * b;
          const uint32_t x = (s >= 8) ? (w >> (s - 8)) : (w << (8 - s));
          look(acc[b], (x & 0x3F00u) + tab);
        }
        if constexpr (SUB < 2) {
          const uint32_t c = ((w >> (30 + SUB)) | (wbv[tt] << (2 - SUB))) & 0x7u;
          look(acc[5], (c << 8) + tab);
        }
      }
    }
  }
}

template <int BITS, bool HALF, int HW>
__global__ void __launch_bounds__(kVFThreads, 1) v_fast_kernel(const __grid_constant__ CUtensorMap tmap, const VFParams p) {
  using C = VFCfg<BITS>;
  constexpr int N = C::N, W = C::W, NP = C::NP, TABN = C::TABN;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int rows = p.H * W;
  const uint32_t stage_bytes = (uint32_t)rows * (kVFT * 4);
  float* s_w = reinterpret_cast<float*>(smem + p.off_w);           // [2][32][H+1]  w = exp(s - max), token-major, padded:
                                                                    // written by lane = token, gathered by lane = entry (head varies)
  unsigned char* s_ws = smem + p.off_ws;                            // [2][H][32]  w * sf_t (f32, or f16 * 2^8)
  float* s_red = reinterpret_cast<float*>(smem + p.off_red);        // [2][H]      denominators, offset terms
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  constexpr uint32_t kWsRow = HALF ? kVFT * 2 : kVFT * 4;           // bytes per head in s_ws

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = p.n_stages;
  const int hidden = p.H * kHeadDim;
  const int ldw = p.H + 1;
  const int n_w = ldw * kVFT;
  const bool has_out = p.outliers != nullptr && p.n_acc > 0;

  // ---- pair table: row i (256 bytes apart) -> {cent[i & (N-1)], cent[i >> BITS]}; fp32: 16 half-warp-private float2
  //      slots (an LDS.64 is served one half-warp at a time), fp16: 32 lane-private half2 slots; bytes 128..255 of a
  //      row are not table (outlier accumulators may live there)
  for (int i = tid; i < TABN * 32; i += kVFThreads) {
    const int idx = i >> 5, sl = i & 31;
    const float lo = p.v_cent[idx & (N - 1)], hi = p.v_cent[(idx >> BITS) & (N - 1)];
    if constexpr (HALF) {
      const __half2 e = __floats2half2_rn(lo, hi);
      *reinterpret_cast<uint32_t*>(smem + p.off_tab + (uint32_t)idx * 256u + sl * 4) = *reinterpret_cast<const uint32_t*>(&e);
    } else if (sl < 16) {
      *reinterpret_cast<float2*>(smem + p.off_tab + (uint32_t)idx * 256u + sl * 8) = make_float2(lo, hi);
    }
  }
  const uint32_t tab = smem_u32(smem + p.off_tab) + (HALF ? lane * 4 : (lane & 15) * 8);
  // outlier accumulators
  for (int a = 0; a < p.n_acc; ++a)
    for (int i = tid; i < hidden; i += kVFThreads)
      *reinterpret_cast<float*>(smem + p.acc_off[a] + (uint32_t)(i >> 5) * p.acc_stride[a] + (i & 31) * 4) = 0.f;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&s_bar[s], 1);
    mbar_fence_init();
    prefetch_tensormap(&tmap);
  }

  int64_t L_eff = p.L;
  if (p.len_dev != nullptr) { const int64_t l = *p.len_dev + p.len_add; L_eff = l < 0 ? 0 : (l < p.L ? l : p.L); }
  const int64_t n_tiles_total = (L_eff + kVFT - 1) / kVFT;
  const int64_t tiles_per_cta = p.len_dev != nullptr ? (n_tiles_total + gridDim.x - 1) / gridDim.x : p.tiles_per_cta;
  const int64_t tile0 = (int64_t)blockIdx.x * tiles_per_cta;
  const int ntiles = (int)max((int64_t)0, min(tiles_per_cta, n_tiles_total - tile0));
  const int nbox = rows / p.box_rows;

  // ---- roles (warp-uniform; each role keeps its own state and loop, all meet at the same named-barrier count) --------
  constexpr int w_first = 18;                                        // weights warps 18, 19 (HW heads each at most)
  auto cta_sync = [] { asm volatile("bar.sync 0, %0;" ::"n"(kVFThreads) : "memory"); };
  __syncthreads();

  if (warp < 16) {
    // ================= dense lookups: thread = packed word row ========================================================
    int u_row[2], u_head[2], u_ch0[2], u_part[2];
    bool u_on[2];
    int sub = 0;
    if constexpr (BITS == 3) {
      sub = warp % 3;
      const int tri = warp / 3;
      const int ngroups = p.H * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int gi = tri * 32 + lane + i * 160;
        u_on[i] = (warp < 15) && gi < ngroups;
        u_row[i] = 3 * gi + sub;
        u_head[i] = gi >> 2;
        u_ch0[i] = (gi & 3) * 32 + (sub == 0 ? 0 : (sub == 1 ? 11 : 22));
        u_part[i] = 0;
      }
    } else {
      const int nunits = p.H * 16;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int u = tid + i * kVFCompute;
        u_on[i] = u < nunits;
        if constexpr (BITS == 4) { u_row[i] = u; u_part[i] = 0; u_head[i] = u >> 4; u_ch0[i] = (u & 15) * 8; }
        else { u_row[i] = u >> 1; u_part[i] = u & 1; u_head[i] = u >> 4; u_ch0[i] = ((u >> 1) & 7) * 16 + (u & 1) * 8; }
      }
    }
    uint32_t r_off[2], r_swz[2], r_off2[2], r_swz2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = u_on[i] ? u_row[i] : 0;
      r_off[i] = (uint32_t)r * 128u;
      r_swz[i] = (uint32_t)(r & 7);
      const int r2 = (r + 1 < rows) ? r + 1 : r;
      r_off2[i] = (uint32_t)r2 * 128u;
      r_swz2[i] = (uint32_t)(r2 & 7);
    }
    float2 acc[2][NP];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < NP; ++k) acc[i][k] = make_float2(0.f, 0.f);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < ntiles; ++it) {
      cta_sync();
      mbar_wait(&s_bar[s], ph);
      const unsigned char* stage = smem + (size_t)s * stage_bytes;
      const unsigned char* wsbuf = s_ws + (size_t)(it & 1) * p.H * kWsRow;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (u_on[i]) {
          const unsigned char* wsrow = wsbuf + (size_t)u_head[i] * kWsRow;
          if constexpr (BITS == 3) {
            if (sub == 0) vf_tile_unit<3, 0, HALF>(stage, r_off[i], r_swz[i], r_off2[i], r_swz2[i], 0, tab, wsrow, acc[i]);
            else if (sub == 1) vf_tile_unit<3, 1, HALF>(stage, r_off[i], r_swz[i], r_off2[i], r_swz2[i], 0, tab, wsrow, acc[i]);
            else vf_tile_unit<3, 2, HALF>(stage, r_off[i], r_swz[i], r_off2[i], r_swz2[i], 0, tab, wsrow, acc[i]);
          } else {
            vf_tile_unit<BITS, 0, HALF>(stage, r_off[i], r_swz[i], 0, 0, u_part[i], tab, wsrow, acc[i]);
          }
        }
      }
      if (++s == S) { s = 0; ph ^= 1u; }
    }
    cta_sync();   // every tile consumed, outlier accumulators final
    cta_sync();   // per-head scalars published
    for (int i = tid; i < p.H; i += kVFCompute) p.out_l[(int64_t)blockIdx.x * p.H + i] = s_red[i];
    float* obase = p.out_o + (int64_t)blockIdx.x * hidden;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (u_on[i]) {
        const float hoff = s_red[p.H + u_head[i]];
        const int nch = (BITS == 3) ? (sub == 2 ? 10 : 11) : 8;
#pragma unroll
        for (int k = 0; k < 2 * NP; ++k) {
          if (k < nch) {
            const int j = u_head[i] * kHeadDim + u_ch0[i] + k;
            float v = (k & 1) ? acc[i][k >> 1].y : acc[i][k >> 1].x;
            if constexpr (HALF) v *= (1.f / kVFHalfScale);
            v += hoff;
            for (int a = 0; a < p.n_acc; ++a)
              v += *reinterpret_cast<const float*>(smem + p.acc_off[a] + (uint32_t)(j >> 5) * p.acc_stride[a] + (j & 31) * 4);
            obase[j] = v;
          }
        }
      }
    }
  } else if (warp >= w_first) {
    // ================= softmax weights of the next tile (lane = token) + TMA issue ======================================
    const bool is_issuer = (warp == 19 && lane == 0);
    auto issue_tile = [&](int it) {
      const int s = it % S;
      const int64_t t0 = (tile0 + it) * kVFT;
      mbar_expect_tx(&s_bar[s], stage_bytes);
      unsigned char* dst = smem + (size_t)s * stage_bytes;
      for (int b = 0; b < nbox; ++b) tma_load_2d(dst + (size_t)b * p.box_rows * 128, &tmap, &s_bar[s], (int)t0, b * p.box_rows);
    };
    const int hsplit = (p.H + 1) >> 1;                                 // warp 19: heads [0, hsplit), warp 18: the rest
    const int wh0 = (warp == 19) ? 0 : hsplit;
    const int wh1 = (warp == 19) ? hsplit : p.H;
    float lacc[HW], oacc[HW];
#pragma unroll
    for (int i = 0; i < HW; ++i) { lacc[i] = 0.f; oacc[i] = 0.f; }
    // the per-head maxima sit in shared memory for the duration of the loop (s_red is free until the epilogue)
    for (int h = wh0 + lane; h < wh1; h += 32) s_red[h] = p.gmax[h];
    __syncwarp();
    auto make_weights = [&](int it) {     // weights of tile `it` -> buffer it & 1
      const int64_t t = (tile0 + it) * kVFT + lane;
      const bool ok = t < L_eff;
      float2 aff = make_float2(0.f, 0.f);
      if (ok) aff = *reinterpret_cast<const float2*>(p.v_aff + 2 * t);
      float* wb = s_w + (it & 1) * n_w;
      unsigned char* wsb = s_ws + (size_t)(it & 1) * p.H * kWsRow;
      const float* sp = p.score + t;
      constexpr int HB = (HW > 16) ? 8 : 16;   // heads per batch: all their score loads in flight before any store
#pragma unroll
      for (int b = 0; b < HW; b += HB) {
        float sc[HB];
#pragma unroll
        for (int i = 0; i < HB; ++i) {
          const int h = wh0 + b + i;
          sc[i] = (ok && h < wh1) ? __ldcg(sp + (int64_t)h * p.score_stride) : -INFINITY;
        }
#pragma unroll
        for (int i = 0; i < HB; ++i) {
          const int h = wh0 + b + i;
          if (h < wh1) {
            const float w = (sc[i] == -INFINITY) ? 0.f : __expf(sc[i] - s_red[h]);
            lacc[b + i] += w;
            oacc[b + i] = fmaf(w, aff.y, oacc[b + i]);
            wb[lane * ldw + h] = w;
            if constexpr (HALF) reinterpret_cast<__half*>(wsb + (size_t)h * kWsRow)[lane] = __float2half_rn(fminf(w * aff.x * kVFHalfScale, 65504.f));
            else reinterpret_cast<float*>(wsb + (size_t)h * kWsRow)[lane] = w * aff.x;
          }
        }
      }
    };
    if (ntiles > 0) {
      if (is_issuer)
        for (int it = 0; it < S - 1 && it < ntiles; ++it) issue_tile(it);
      make_weights(0);
    }
    for (int it = 0; it < ntiles; ++it) {
      cta_sync();
      if (is_issuer && it + S - 1 < ntiles) issue_tile(it + S - 1);
      if (it + 1 < ntiles) make_weights(it + 1);
    }
    cta_sync();
#pragma unroll
    for (int i = 0; i < HW; ++i) {
      const int h = wh0 + i;
      if (h < wh1) {     // warp-uniform
        const float a = warp_sum(lacc[i]), b = warp_sum(oacc[i]);
        if (lane == 0) { s_red[h] = a; s_red[p.H + h] = b; }
      }
    }
    cta_sync();
  } else {
    // ================= outliers: private accumulator row, plain read-modify-write =======================================
    const int oa = warp - 16;
    const bool active = has_out && oa < p.n_acc;
    const uint32_t my_acc = active ? smem_u32(smem + p.acc_off[oa]) : 0u;
    const uint32_t my_stride = active ? p.acc_stride[oa] : 0u;
    // Token-aligned walk: warp `oa` owns tokens oa, oa + n_acc, ... of the tile; one step = up to 32 entries of ONE
    // token's row.  The indices of a row are distinct, so the 32 lanes of a step touch 32 different accumulator
    // words: plain load / add / store, no atomics and nothing to merge.  Steps of one warp hit the same private row in
    // program order (same-address accesses of a warp are served in issue order; __syncwarp orders them formally).
    const int steps_per_tok = (p.n_out + 31) >> 5;
    const bool need_tail = p.n_out > 64 || (kVFT + p.n_acc - 1) / p.n_acc > kVFPre / 2;
    float opre_v[kVFPre];
    int opre_i[kVFPre];
    auto load_outliers = [&](int it) {
      const int64_t t0 = (tile0 + it) * kVFT;
#pragma unroll
      for (int k = 0; k < kVFPre; ++k) {
        const int tok = oa + (k / 2) * p.n_acc, part = k & 1;        // two steps per token cover n_out <= 64
        const int e = part * 32 + lane;
        const bool in = active && tok < kVFT && (t0 + tok) < L_eff && e < p.n_out && part < steps_per_tok;
        const int64_t off = (t0 + tok) * p.n_out + e;
        opre_v[k] = in ? __ldcs(p.outliers + off) : 0.f;
        opre_i[k] = in ? __ldcs(p.outlier_idx + off) : 0;
      }
    };
    // a lane with nothing to add (row tail, token past the end, zero value) targets a private dummy word with x = 0,
    // so that a step is straight-line code for the whole warp: no divergence, no reconvergence barriers
    const uint32_t dummy = smem_u32(s_red) + (uint32_t)(2 * p.H + (warp - 16) * 32 + lane) * 4u;
    auto scatter = [&](float v, int idx, int tok, const float* wbuf) {   // slow path (rows / tiles past the prefetch window)
      float x = 0.f;
      uint32_t a = dummy;
      if (v != 0.f) {
        x = v * wbuf[tok * ldw + (idx >> 7)];
        a = my_acc + (uint32_t)(idx >> 5) * my_stride + (uint32_t)(idx & 31) * 4u;
      }
      float cur;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cur) : "r"(a) : "memory");
      cur += x;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(cur) : "memory");
      __syncwarp();
    };
    if (active && ntiles > 0) load_outliers(0);
    for (int it = 0; it < ntiles; ++it) {
      cta_sync();
      if (active) {
        // this tile's entries were fetched at the end of the previous iteration (in flight across the barrier)
        const float* wbuf = s_w + (it & 1) * n_w;
        // per group of 8 steps -- phase 1 (independent per step, pipelines freely): x = value * w[head, token] and the
        // accumulator address, written over the prefetched (value, index) registers; phase 2: one plain
        // read-modify-write per step, in program order (the steps of this warp may hit the same word)
#pragma unroll
        for (int k0 = 0; k0 < kVFPre; k0 += 8) {
#pragma unroll
          for (int k = k0; k < k0 + 8; ++k) {
            const int tok = oa + (k / 2) * p.n_acc;
            const float v = opre_v[k];
            const int idx = opre_i[k];
            const bool on = (tok < kVFT) && (v != 0.f);
            opre_v[k] = on ? v * wbuf[tok * ldw + (idx >> 7)] : 0.f;
            opre_i[k] = on ? (int)(my_acc + (uint32_t)(idx >> 5) * my_stride + (uint32_t)(idx & 31) * 4u) : (int)dummy;
          }
#pragma unroll
          for (int k = k0; k < k0 + 8; ++k) {
            float cur;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cur) : "r"(opre_i[k]) : "memory");
            cur += opre_v[k];
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(opre_i[k]), "f"(cur) : "memory");
            __syncwarp();
          }
        }
        if (need_tail) {
          // tokens past the prefetch window (a single outlier warp) and rows wider than 64 entries: straight from
          // global memory
          const int64_t t0 = (tile0 + it) * kVFT;
          int slot = 0;
          for (int tok = oa; tok < kVFT && t0 + tok < L_eff; tok += p.n_acc, ++slot) {
            const bool pre = slot < kVFPre / 2;
            for (int e = (pre ? 64 : 0) + lane; e - lane < p.n_out; e += 32) {
              const bool in = e < p.n_out;
              const int64_t off = (t0 + tok) * p.n_out + e;
              scatter(in ? p.outliers[off] : 0.f, in ? p.outlier_idx[off] : 0, tok, wbuf);
            }
          }
        }
        if (it + 1 < ntiles) load_outliers(it + 1);
      }
    }
    cta_sync();
    cta_sync();
  }
}

int num_sms_cached();

template <int BITS, bool HALF, int HW>
static int launch_vf(VFParams p, const int32_t* cache, int* n_cta_out, cudaStream_t st) {
  using C = VFCfg<BITS>;
  const int rows = p.H * C::W;
  const uint32_t hidden = (uint32_t)p.H * kHeadDim;
  const uint32_t stage_bytes = (uint32_t)rows * (kVFT * 4);
  const uint32_t tab_span = (uint32_t)C::TABN * 256u;
  const uint32_t ws_bytes = 2u * p.H * (HALF ? kVFT * 2 : kVFT * 4);
  const uint32_t w_bytes = ((2u * (p.H + 1) * kVFT * 4) + 15u) & ~15u;
  const uint32_t red_bytes = (2u * p.H + 64u) * 4;     // per-head scalars + one dummy word per outlier-warp lane
  const int max_acc = kVFMaxAcc;
  const bool want_out = p.outliers != nullptr;
  int S = kVFMaxStages;
  for (; S >= 2; --S) {
    uint32_t used = stage_bytes * S;                    // stages first (1024-aligned)
    p.off_tab = used; used += tab_span;
    p.off_w = used; used += w_bytes;
    p.off_ws = used; used += ws_bytes;
    p.off_red = used; used += red_bytes;
    used = (used + 7u) & ~7u;
    p.off_bar = used; used += 8u * kVFMaxStages;
    used = (used + 127u) & ~127u;
    if (used + 1024u > kVFSmemBudget) continue;
    int n_acc = 0;
    if (want_out) {
      // rows of 32 floats: first in the 128-byte holes of the table (row stride 256), then packed in what is left
      const uint32_t rows_needed = hidden / 32;
      uint32_t holes = (uint32_t)C::TABN;
      uint32_t hole_next = 0;
      while (n_acc < max_acc && holes - hole_next >= rows_needed) {
        p.acc_off[n_acc] = p.off_tab + hole_next * 256u + 128u;
        p.acc_stride[n_acc] = 256u;
        hole_next += rows_needed;
        ++n_acc;
      }
      while (n_acc < max_acc && used + hidden * 4u + 1024u <= kVFSmemBudget) {
        p.acc_off[n_acc] = used;
        p.acc_stride[n_acc] = 128u;
        used += hidden * 4u;
        ++n_acc;
      }
      if (n_acc == 0) continue;      // try fewer stages
      // two accumulators at two stages beat three stages with one (the outlier walk is latency-bound)
      if (n_acc < 2 && S > 2) continue;
    }
    p.n_acc = n_acc;
    p.n_stages = S;
    const size_t total = (size_t)used + 1024u;
    static PerDeviceOnce attr_once;
    bool& attr_done = attr_once.cur();
    if (!attr_done) {
      cudaError_t e = cudaFuncSetAttribute(v_fast_kernel<BITS, HALF, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVFSmemBudget);
      if (e != cudaSuccess) return (int)e;
      attr_done = true;
    }
    CUtensorMap tmap;
    int nb = (rows + 255) / 256;
    while (rows % nb != 0 || (rows / nb) % 8 != 0) ++nb;
    p.box_rows = rows / nb;
    int rc = make_cache_tensor_map(&tmap, cache, (uint64_t)rows, (uint64_t)p.Lmax, kVFT, (uint32_t)p.box_rows, 128);
    if (rc != 0) return rc;
    const int64_t n_tiles = (p.L + kVFT - 1) / kVFT;
    const int sms = num_sms_cached();
    p.tiles_per_cta = (int)((n_tiles + sms - 1) / sms);
    const int n_cta = (int)((n_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta);
    p.n_out_magic = p.n_out > 0 ? (uint32_t)(((uint64_t)1 << 32) / (uint32_t)p.n_out + 1u) : 0u;
    v_fast_kernel<BITS, HALF, HW><<<n_cta, kVFThreads, total, st>>>(tmap, p);
    KVQ_LAUNCH_CHECK();
    *n_cta_out = n_cta;
    return 0;
  }
  return KVQ_E_UNSUPPORTED;
}

int v_fast_dispatch(int bits, int half_mode, const float* score, int64_t score_stride, const float* gmax,
                    const int32_t* cache, const float* v_cent, const float* v_aff, const float* outliers,
                    const int32_t* outlier_idx, int n_out, int H, int64_t Lmax, int64_t L, float* out_o, float* out_l,
                    int* n_cta, const int64_t* len_dev, int64_t len_add, cudaStream_t st) {
  if (H > 64 || n_out > 2048 / kVFT) return KVQ_E_UNSUPPORTED;
  VFParams p{};
  p.len_dev = len_dev; p.len_add = len_add;
  p.score = score; p.gmax = gmax; p.v_cent = v_cent; p.v_aff = v_aff; p.out_o = out_o; p.out_l = out_l;
  p.outliers = outliers; p.outlier_idx = outlier_idx; p.Lmax = Lmax; p.L = L; p.score_stride = score_stride;
  p.H = H; p.n_out = n_out;
  const bool wide = H > 32;      // heads per weights warp: 16 (H <= 32) or 32 (H <= 64)
  switch (bits * 2 + (half_mode ? 1 : 0)) {
    case 8: return wide ? launch_vf<4, false, 32>(p, cache, n_cta, st) : launch_vf<4, false, 16>(p, cache, n_cta, st);
    case 9: return wide ? launch_vf<4, true, 32>(p, cache, n_cta, st) : launch_vf<4, true, 16>(p, cache, n_cta, st);
    case 6: return wide ? launch_vf<3, false, 32>(p, cache, n_cta, st) : launch_vf<3, false, 16>(p, cache, n_cta, st);
    case 7: return wide ? launch_vf<3, true, 32>(p, cache, n_cta, st) : launch_vf<3, true, 16>(p, cache, n_cta, st);
    case 4: return wide ? launch_vf<2, false, 32>(p, cache, n_cta, st) : launch_vf<2, false, 16>(p, cache, n_cta, st);
    case 5: return wide ? launch_vf<2, true, 32>(p, cache, n_cta, st) : launch_vf<2, true, 16>(p, cache, n_cta, st);
    default: return KVQ_E_BITS;
  }
}

}  // namespace kvq
