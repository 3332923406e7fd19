// kvquant_b200 -- the uncapped "orig" dense-and-sparse path (4-bit only in the reference):
//   * append: pack one token, threshold-crossing elements get the zero-point code 7 and are emitted compacted in
//     channel order as (index, value - zeropoint)     (reference quant_cuda_kernel.cu:846-931 and 1078-1163);
//     the reference needs two launches + a blocking D2H read of the count in between (745-747); here one launch
//     emits indices, values and the count, and the host glue (kvquant_b200/quant_cuda.py) grows the CSR arrays.
//   * SpMV: nnz-balanced CSR (K, rows = tokens, RoPE per nonzero; 523-614) and CSC (V, cols = tokens; 616-689),
//     10 nonzeros per thread, `start` gives each thread its first row/col.
#include "kvq_common.cuh"

namespace kvq {

constexpr int kOrigThreads = 1024;

// ISV=false: per-channel thresholds/zeropoint arrays, per-channel LUT [hidden,16]
// ISV=true : scalar thresholds/zeropoint, per-token LUT row lut[slot*16 ...]
template <bool ISV>
__global__ void __launch_bounds__(kOrigThreads, 1) orig_append_kernel(
    uint32_t* __restrict__ cache, const float* __restrict__ lut, const float* __restrict__ newvec,
    const float* __restrict__ zp_arr, const float* __restrict__ lo_arr, const float* __restrict__ hi_arr,
    float zp_s, float lo_s, float hi_s, int32_t* __restrict__ out_idx, float* __restrict__ out_val,
    int32_t* __restrict__ out_count, int hidden, int64_t Lmax, int64_t slot) {
  extern __shared__ unsigned char s_code[];  // [hidden]
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (hidden + kOrigThreads - 1) / kOrigThreads;  // consecutive channels per thread
  const int j0 = tid * per;
  int cnt = 0;
  for (int e = 0; e < per; ++e) {
    const int j = j0 + e;
    if (j >= hidden) break;
    const float x = newvec[j];
    const float lo = ISV ? lo_s : lo_arr[j];
    const float hi = ISV ? hi_s : hi_arr[j];
    uint32_t code;
    if (x < lo || x > hi) { code = 7; ++cnt; }   // zero-point (quant_cuda_kernel.cu:905-909)
    else code = nearest_code<4>(lut + (ISV ? slot * 16 : (int64_t)j * 16), x);
    s_code[j] = (unsigned char)code;
  }
  // exclusive prefix sum of cnt over the block (channel order == thread order)
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    s_warp[lane] = wi - w;  // exclusive
    if (lane == 31) s_total = wi;
  }
  __syncthreads();
  int pos = s_warp[warp] + incl - cnt;
  for (int e = 0; e < per; ++e) {
    const int j = j0 + e;
    if (j >= hidden) break;
    const float x = newvec[j];
    const float lo = ISV ? lo_s : lo_arr[j];
    const float hi = ISV ? hi_s : hi_arr[j];
    if (x < lo || x > hi) {
      out_idx[pos] = j;
      out_val[pos] = x - (ISV ? zp_s : zp_arr[j]);
      ++pos;
    }
  }
  if (tid == 0) *out_count = s_total;
  // pack (4-bit), words ADDED to the cache like the reference's atomicAdd
  for (int R = tid; R < hidden / 8; R += kOrigThreads) {
    uint32_t w = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) w |= (uint32_t)s_code[R * 8 + q] << (4 * q);
    uint32_t* p = cache + (int64_t)R * Lmax + slot;
    *p = *p + w;
  }
}

// SPMV_ATOMIC_CSR_ROPE_BALANCED semantics (quant_cuda_kernel.cu:523-614)
__global__ void csr_k_spmv_kernel(const int* __restrict__ rows, const int* __restrict__ cols,
                                  const int* __restrict__ startrows, const float* __restrict__ mat,
                                  const float* __restrict__ vec, float* __restrict__ mul, int num_rows,
                                  int64_t seqlen, int num_threads, int nnz, float rope_theta, int pos_offset) {
  const int headdim = kHeadDim;
  const int per = (nnz + num_threads - 1) / num_threads;
  const int th = blockIdx.x * blockDim.x + threadIdx.x;
  if (th >= num_threads) return;
  int row = startrows[th];
  int nextrow = -1;
  if (row != -1) {
    nextrow = rows[row + 1];
    while (nextrow == th * per) {  // do not start on an empty row
      row += 1;
      if (row < num_rows) nextrow = rows[row + 1];
      else break;
    }
  }
  if (th * per >= nnz || row == -1) return;
  const int end = min(nnz, (th + 1) * per);
  for (int i = th * per; i < end; ++i) {
    const int col = cols[i];
    const float v = mat[i];
    const int head = col / headdim, ch = col % headdim;
    const float theta = powf(rope_theta, (-2 * __int2float_rd(ch % (headdim / 2)) / headdim));
    const float sign = (ch < headdim / 2) ? 1.f : -1.f;
    const float c = cosf(theta * (row + pos_offset));
    const float s = sinf(theta * (row + pos_offset));
    const int col2 = ((ch + headdim / 2) % headdim) + head * headdim;
    float dot = v * c * vec[col];
    dot += sign * v * s * vec[col2];
    atomicAdd(&mul[(int64_t)head * seqlen + row], dot);
    while (i + 1 == nextrow) {  // row finished (skip empty rows)
      row += 1;
      if (row < num_rows) nextrow = rows[row + 1];
      else { nextrow = -1; break; }
    }
  }
}

// SPMV_ATOMIC_CSC_BALANCED semantics (quant_cuda_kernel.cu:616-689)
__global__ void csc_v_spmv_kernel(const int* __restrict__ rows, const int* __restrict__ cols,
                                  const int* __restrict__ startcols, const float* __restrict__ mat,
                                  const float* __restrict__ vec, float* __restrict__ mul, int num_cols,
                                  int64_t seqlen, int num_threads, int nnz) {
  const int headdim = kHeadDim;
  const int per = (nnz + num_threads - 1) / num_threads;
  const int th = blockIdx.x * blockDim.x + threadIdx.x;
  if (th >= num_threads) return;
  int col = startcols[th];
  int nextcol = -1;
  if (col != -1) {
    nextcol = cols[col + 1];
    while (nextcol == th * per) {
      col += 1;
      if (col < num_cols) nextcol = cols[col + 1];
      else break;
    }
  }
  if (th * per >= nnz || col == -1) return;
  const int end = min(nnz, (th + 1) * per);
  for (int i = th * per; i < end; ++i) {
    const int row = rows[i];
    const int head = row / headdim;
    atomicAdd(&mul[row], mat[i] * vec[(int64_t)head * seqlen + col]);
    while (i + 1 == nextcol) {
      col += 1;
      if (col < num_cols) nextcol = cols[col + 1];
      else { nextcol = -1; break; }
    }
  }
}

}  // namespace kvq

using namespace kvq;

extern "C" {

int kvq_k_spmv_csr(const int32_t* rows, const int32_t* cols, const int32_t* start_rows, const float* vals,
                   const float* q, float* mul, int H, int64_t L, int num_rows, int num_threads, int nnz, float theta,
                   int pos_offset, void* stream) {
  if (!rows || !cols || !start_rows || !vals || !q || !mul) return KVQ_E_NULL;
  if (H <= 0 || L < 0 || num_threads < 0 || nnz < 0) return KVQ_E_SHAPE;
  if (num_threads == 0 || nnz == 0) return 0;
  csr_k_spmv_kernel<<<(num_threads + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      rows, cols, start_rows, vals, q, mul, num_rows, L, num_threads, nnz, theta, pos_offset);
  KVQ_LAUNCH_CHECK();
  return 0;
}

int kvq_v_spmv_csc(const int32_t* rows, const int32_t* cols, const int32_t* start_cols, const float* vals,
                   const float* score, float* mul, int H, int64_t L, int num_cols, int num_threads, int nnz,
                   void* stream) {
  if (!rows || !cols || !start_cols || !vals || !score || !mul) return KVQ_E_NULL;
  if (H <= 0 || L < 0 || num_threads < 0 || nnz < 0) return KVQ_E_SHAPE;
  if (num_threads == 0 || nnz == 0) return 0;
  csc_v_spmv_kernel<<<(num_threads + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      rows, cols, start_cols, vals, score, mul, num_cols, L, num_threads, nnz);
  KVQ_LAUNCH_CHECK();
  return 0;
}

int kvq_append_k_orig(int32_t* cache, const float* lut, const float* newvec, const float* zeropoint,
                      const float* thr_lower, const float* thr_upper, int32_t* out_cols, float* out_vals,
                      int32_t* out_count, int H, int64_t Lmax, int64_t slot, void* stream) {
  if (!cache || !lut || !newvec || !zeropoint || !thr_lower || !thr_upper || !out_cols || !out_vals || !out_count) return KVQ_E_NULL;
  if (H <= 0 || Lmax <= 0 || slot < 0 || slot >= Lmax) return KVQ_E_SHAPE;
  const int hidden = H * kHeadDim;
  orig_append_kernel<false><<<1, kOrigThreads, hidden, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint32_t*>(cache), lut, newvec, zeropoint, thr_lower, thr_upper, 0.f, 0.f, 0.f, out_cols,
      out_vals, out_count, hidden, Lmax, slot);
  KVQ_LAUNCH_CHECK();
  return 0;
}

int kvq_append_v_orig(int32_t* cache, const float* lut_tok, const float* newvec, float zeropoint, float thr_lower,
                      float thr_upper, int32_t* out_rows, float* out_vals, int32_t* out_count, int H, int64_t Lmax,
                      int64_t slot, void* stream) {
  if (!cache || !lut_tok || !newvec || !out_rows || !out_vals || !out_count) return KVQ_E_NULL;
  if (H <= 0 || Lmax <= 0 || slot < 0 || slot >= Lmax) return KVQ_E_SHAPE;
  const int hidden = H * kHeadDim;
  orig_append_kernel<true><<<1, kOrigThreads, hidden, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint32_t*>(cache), lut_tok, newvec, nullptr, nullptr, nullptr, zeropoint, thr_lower,
      thr_upper, out_rows, out_vals, out_count, hidden, Lmax, slot);
  KVQ_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
