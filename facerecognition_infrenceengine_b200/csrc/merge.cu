// k-way merge of best-first partial lists -> final top-k, threshold decision, row offset.
// Used (a) after every scan to fold the per-CTA partial lists, (b) as frg_merge_topk, the tail of
// the row-sharded multi-GPU match.  One warp per query.
//
// Order: better(a, b) = a.score > b.score || (a.score == b.score && a.row < b.row), i.e. the
// reference's strict '>' scan over gallery order (infrenceServer.py:538-542) extended to k slots.
#include "merge_device.cuh"

namespace frg {

template <typename RowT, int KMAX>
__global__ void __launch_bounds__(128)
merge_kernel(const float* __restrict__ scores, const RowT* __restrict__ rows, int parts, int nq,
             int k_in, int k_out, int metric, float threshold, int64_t row_offset, int internal_euclid,
             int64_t score_stride, int64_t row_stride,
             int64_t* __restrict__ out_rows, float* __restrict__ out_scores,
             uint8_t* __restrict__ out_accept) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  merge_one<RowT, KMAX>(scores, rows, parts, nq, k_in, k_out, metric, threshold, row_offset, internal_euclid, q, q,
                        score_stride, row_stride, out_rows, out_scores, out_accept);
}

template <typename RowT>
static int launch_merge(const float* scores, const RowT* rows, int parts, int nq, int k_in, int k_out,
                        int metric, float threshold, int64_t row_offset, bool internal_euclid,
                        int64_t score_stride, int64_t row_stride,
                        int64_t* out_rows, float* out_scores, uint8_t* out_accept, cudaStream_t st) {
  if (nq <= 0) return FRG_OK;
  if (score_stride <= 0) score_stride = int64_t(nq) * k_in;
  if (row_stride <= 0) row_stride = int64_t(nq) * k_in;
  if (k_out < 1 || k_out > FRG_MAX_K || k_in < 1) { set_error("merge: k out of range"); return FRG_ERR_INVALID; }
  const int grid = (nq + 3) / 4;
  const int ie = internal_euclid ? 1 : 0;
#define FRG_MERGE(K)                                                                                            \
  do {                                                                                                          \
    func_attr_once(merge_kernel<RowT, K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);           \
    merge_kernel<RowT, K><<<grid, 128, 0, st>>>(scores, rows, parts, nq, k_in, k_out, metric, threshold,        \
                                                row_offset, ie, score_stride, row_stride, out_rows, out_scores,  \
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
                               finalize_euclid, 0, 0, out_rows, out_scores, out_accept, st);
}

int launch_merge_i32(const float* scores, const int32_t* rows, int parts, int nq, int k_in, int k_out,
                     int metric, float threshold, int64_t row_offset, bool finalize_euclid,
                     int64_t* out_rows, float* out_scores, uint8_t* out_accept, cudaStream_t st) {
  return launch_merge<int32_t>(scores, rows, parts, nq, k_in, k_out, metric, threshold, row_offset,
                               finalize_euclid, 0, 0, out_rows, out_scores, out_accept, st);
}

int launch_merge_i64_strided(const float* scores, int64_t score_stride, const int64_t* rows, int64_t row_stride,
                             int parts, int nq, int k, int metric, float threshold, int64_t* out_rows,
                             float* out_scores, uint8_t* out_accept, cudaStream_t st) {
  return launch_merge<int64_t>(scores, rows, parts, nq, k, k, metric, threshold, 0, false, score_stride, row_stride,
                               out_rows, out_scores, out_accept, st);
}

}  // namespace frg
