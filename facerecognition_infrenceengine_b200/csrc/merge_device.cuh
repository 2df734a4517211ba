// Device-side k-way merge shared by merge.cu (stand-alone kernels) and scan_f32.cu (the fused
// flagged-query fallback).  Not part of the ABI.
#pragma once

#include "frg_internal.cuh"

namespace frg {

template <typename RowT>
__device__ __forceinline__ bool better(float sa, RowT ra, float sb, RowT rb) {
  return sa > sb || (sa == sb && ra < rb);
}

template <typename RowT> struct RowLimits;
template <> struct RowLimits<int32_t> { static __device__ __forceinline__ int32_t none() { return 0x7fffffff; } };
template <> struct RowLimits<int64_t> { static __device__ __forceinline__ int64_t none() { return 0x7fffffffffffffffLL; } };

// where the partial lists of a merge come from: plain arrays [part][slot] ...
template <typename RowT>
struct ArrayLists {
  const float* scores;
  const RowT* rows;
  int64_t score_stride, row_stride;          // elements between consecutive parts
  __device__ __forceinline__ void load(int part, size_t in_part, RowT* r, float* s) const {
    *r = rows[size_t(part) * row_stride + in_part];
    *s = scores[size_t(part) * score_stride + in_part];
  }
};

// Internal score convention: larger is better.  Euclidean lists arrive either as distances
// (external, frg_merge_topk) or as -d^2 (internal partials, finalize_euclid): both are mapped to
// "larger is better" on load and mapped back on store.
// One warp folds the `parts` best-first lists of one query.  `slot` addresses the partial lists; with
// an index list (flagged queries) it differs from `q`, the query the result belongs to.
// Lists: ArrayLists, or exchange.cu's packet reader (entries that arrive over NVLink while the warp waits).
template <typename RowT, int KMAX, typename Lists>
__device__ __forceinline__ void merge_lists(const Lists& lists, int parts, int k_in, int k_out, int metric,
                                            float threshold, int64_t row_offset, int internal_euclid, int slot,
                                            int q, int64_t* __restrict__ out_rows,
                                            float* __restrict__ out_scores, uint8_t* __restrict__ out_accept) {
  const int lane = threadIdx.x & 31;
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
        lists.load(part, size_t(slot) * k_in + j, &r[u], &s[u]);
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
    // total order (score, row, lane): every lane agrees on the winner (warp_argbest, frg_internal.cuh)
    const int bl = warp_argbest<RowT>(sc[0], ix[0], RowLimits<RowT>::none(), lane);
    const float bs = __shfl_sync(0xffffffffu, sc[0], bl);
    const RowT br = __shfl_sync(0xffffffffu, ix[0], bl);
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

template <typename RowT, int KMAX>
__device__ __forceinline__ void merge_one(const float* __restrict__ scores, const RowT* __restrict__ rows,
                                          int parts, int nq, int k_in, int k_out, int metric, float threshold,
                                          int64_t row_offset, int internal_euclid, int slot, int q,
                                          int64_t score_stride, int64_t row_stride,
                                          int64_t* __restrict__ out_rows, float* __restrict__ out_scores,
                                          uint8_t* __restrict__ out_accept) {
  (void)nq;
  const ArrayLists<RowT> lists{scores, rows, score_stride, row_stride};
  merge_lists<RowT, KMAX>(lists, parts, k_in, k_out, metric, threshold, row_offset, internal_euclid, slot, q,
                          out_rows, out_scores, out_accept);
}

}  // namespace frg
