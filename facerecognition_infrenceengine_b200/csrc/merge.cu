// k-way merge of best-first partial lists -> final top-k, threshold decision, row offset.
// Used (a) after every scan to fold the per-CTA partial lists, (b) as frg_merge_topk, the tail of
// the row-sharded multi-GPU match.  One warp per query.
//
// Order: better(a, b) = a.score > b.score || (a.score == b.score && a.row < b.row), i.e. the
// reference's strict '>' scan over gallery order (infrenceServer.py:538-542) extended to k slots.
#include "frg_internal.cuh"

namespace frg {

template <typename RowT>
__device__ __forceinline__ bool better(float sa, RowT ra, float sb, RowT rb) {
  return sa > sb || (sa == sb && ra < rb);
}

template <typename RowT> struct RowLimits;
template <> struct RowLimits<int32_t> { static __device__ __forceinline__ int32_t none() { return 0x7fffffff; } };
template <> struct RowLimits<int64_t> { static __device__ __forceinline__ int64_t none() { return 0x7fffffffffffffffLL; } };

// Internal score convention: larger is better.  Euclidean lists arrive either as distances
// (external, frg_merge_topk) or as -d^2 (internal partials, finalize_euclid): both are mapped to
// "larger is better" on load and mapped back on store.
template <typename RowT, int KMAX>
__global__ void __launch_bounds__(128)
merge_kernel(const float* __restrict__ scores, const RowT* __restrict__ rows, int parts, int nq,
             int k_in, int k_out, int metric, float threshold, int64_t row_offset, int internal_euclid,
             const int* __restrict__ q_index, const int* __restrict__ n_active,
             int64_t* __restrict__ out_rows, float* __restrict__ out_scores,
             uint8_t* __restrict__ out_accept) {
  const int lane = threadIdx.x & 31;
  // `slot` addresses the partial lists; with an index list (flagged queries) it differs from the
  // query the result belongs to
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slot >= nq || (n_active && slot >= *n_active)) return;
  const int q = q_index ? q_index[slot] : slot;
  const bool euclid = metric == FRG_METRIC_EUCLIDEAN;
  const float sentinel = euclid ? -INFINITY : kNoScore;

  float sc[KMAX];
  RowT ix[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { sc[j] = sentinel; ix[j] = RowLimits<RowT>::none(); }

  // candidates are read four at a time so that the global loads of one round are independent
  const int total = parts * k_in;
  for (int c0 = lane; c0 < total; c0 += 32 * 4) {
    RowT r[4];
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u * 32;
      r[u] = RowLimits<RowT>::none();
      s[u] = sentinel;
      if (c < total) {
        const int part = c / k_in, j = c - part * k_in;
        const size_t off = (size_t(part) * nq + slot) * k_in + j;
        r[u] = rows[off];
        s[u] = scores[off];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r[u] < 0 || r[u] == RowLimits<RowT>::none()) continue;
      float v = s[u];
      if (euclid && !internal_euclid) v = -v;        // external lists carry distances
      if (!(v > sentinel)) continue;                 // also drops NaN
      if (better<RowT>(v, r[u], sc[KMAX - 1], ix[KMAX - 1])) {
        sc[KMAX - 1] = v; ix[KMAX - 1] = r[u];
#pragma unroll
        for (int t = KMAX - 1; t > 0; --t) {
          if (better<RowT>(sc[t], ix[t], sc[t - 1], ix[t - 1])) {
            float ts = sc[t]; sc[t] = sc[t - 1]; sc[t - 1] = ts;
            RowT tr = ix[t]; ix[t] = ix[t - 1]; ix[t - 1] = tr;
          }
        }
      }
    }
  }

  // k_out rounds of warp arg-best over the lane heads; the winner pops its head
  for (int j = 0; j < k_out; ++j) {
    float bs = sc[0];
    RowT br = ix[0];
    int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const RowT orow = __shfl_xor_sync(0xffffffffu, br, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      // total order (score, row, lane) so that every lane converges on the same winner
      if (better<RowT>(os, orow, bs, br) || (os == bs && orow == br && ol < bl)) { bs = os; br = orow; bl = ol; }
    }
    if (lane == bl) {
#pragma unroll
      for (int t = 0; t < KMAX - 1; ++t) { sc[t] = sc[t + 1]; ix[t] = ix[t + 1]; }
      sc[KMAX - 1] = sentinel; ix[KMAX - 1] = RowLimits<RowT>::none();
    }
    if (lane == 0) {
      const bool filled = br != RowLimits<RowT>::none();
      float s_out;
      if (!filled) s_out = euclid ? INFINITY : kNoScore;
      else if (euclid) s_out = internal_euclid ? __fsqrt_rn(fmaxf(-bs, 0.f)) : -bs;
      else s_out = bs;
      out_rows[size_t(q) * k_out + j] = filled ? int64_t(br) + row_offset : int64_t(kNoRow);
      out_scores[size_t(q) * k_out + j] = s_out;
      if (j == 0 && out_accept) {
        // fp32 score >= fp32(threshold): infrenceServer.py:545 / peopleCount.py:876 under NumPy >= 2
        out_accept[q] = filled && (euclid ? (s_out <= threshold) : (s_out >= threshold)) ? 1 : 0;
      }
    }
  }
}

template <typename RowT>
static int launch_merge(const float* scores, const RowT* rows, int parts, int nq, int k_in, int k_out,
                        int metric, float threshold, int64_t row_offset, bool internal_euclid,
                        const int* q_index, const int* n_active, int64_t* out_rows, float* out_scores, uint8_t* out_accept, cudaStream_t st) {
  if (nq <= 0) return FRG_OK;
  if (k_out < 1 || k_out > FRG_MAX_K || k_in < 1) { set_error("merge: k out of range"); return FRG_ERR_INVALID; }
  const int grid = (nq + 3) / 4;
  const int ie = internal_euclid ? 1 : 0;
#define FRG_MERGE(K)                                                                                            \
  do {                                                                                                          \
    cudaFuncSetAttribute(merge_kernel<RowT, K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);           \
    merge_kernel<RowT, K><<<grid, 128, 0, st>>>(scores, rows, parts, nq, k_in, k_out, metric, threshold,        \
                                                row_offset, ie, q_index, n_active, out_rows, out_scores,       \
                                                out_accept);                                                    \
  } while (0)
  if (k_out == 1) FRG_MERGE(1);
  else if (k_out <= 4) FRG_MERGE(4);
  else if (k_out <= 8) FRG_MERGE(8);
  else FRG_MERGE(16);
#undef FRG_MERGE
  note_launch(nullptr);
  FRG_CUDA(cudaGetLastError());
  return FRG_OK;
}

int launch_merge_i64(const float* scores, const int64_t* rows, int parts, int nq, int k_in, int k_out,
                     int metric, float threshold, int64_t row_offset, bool finalize_euclid,
                     int64_t* out_rows, float* out_scores, uint8_t* out_accept, cudaStream_t st) {
  return launch_merge<int64_t>(scores, rows, parts, nq, k_in, k_out, metric, threshold, row_offset,
                               finalize_euclid, nullptr, nullptr, out_rows, out_scores, out_accept, st);
}

int launch_merge_i32(const float* scores, const int32_t* rows, int parts, int nq, int k_in, int k_out,
                     int metric, float threshold, int64_t row_offset, bool finalize_euclid,
                     int64_t* out_rows, float* out_scores, uint8_t* out_accept, cudaStream_t st) {
  return launch_merge<int32_t>(scores, rows, parts, nq, k_in, k_out, metric, threshold, row_offset,
                               finalize_euclid, nullptr, nullptr, out_rows, out_scores, out_accept, st);
}

int launch_merge_flagged(const float* scores, const int32_t* rows, int parts, int nq, int k, float threshold,
                         int64_t row_offset, const int* q_index, const int* n_active, int64_t* out_rows,
                         float* out_scores, uint8_t* out_accept, cudaStream_t st) {
  return launch_merge<int32_t>(scores, rows, parts, nq, k, k, FRG_METRIC_COSINE, threshold, row_offset, false,
                               q_index, n_active, out_rows, out_scores, out_accept, st);
}

}  // namespace frg
