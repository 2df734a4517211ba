// "First row whose score reaches the threshold" - the scan rule of two neighbours of the matching
// path:
//   * enrol-time duplicate check: first stored template, in cursor (= gallery) order, with
//     cos > 0.4                                              (trainingServer.py:170-200)
//   * unknown-person clustering: clusters are visited in creation order, the running best is
//     updated, and the first one with dot >= 0.65 is taken  (peopleCount.py:446-452) - every earlier
//     cluster scored < 0.65, so "first running maximum >= thr" == "first row >= thr".
// Exact fp32 arithmetic (same element->lane mapping and summation order as scan_f32).  One query per
// blockIdx.y; a warp owns whole rows in increasing order and the grid shares the best row found so
// far through a 64-bit atomicMin of (row << 32 | score bits), which also lets warps stop early.
#include "frg_internal.cuh"

namespace frg {

template <int NJ>
__global__ void __launch_bounds__(256)
first_match_kernel(const float* __restrict__ master, const int32_t* __restrict__ tags, int64_t rows,
                   const float* __restrict__ qn, int32_t tenant, float threshold, int strict,
                   unsigned long long* __restrict__ best) {
  constexpr int DIM = NJ * 128;
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.y;
  float4 qv[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) qv[j] = __ldg(reinterpret_cast<const float4*>(qn + size_t(q) * DIM) + j * 32 + lane);
  const int64_t gw = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t tw = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = gw; r < rows; r += tw) {
    // a lower row has already qualified: nothing this warp can still find matters
    if ((*reinterpret_cast<volatile unsigned long long*>(best + q) >> 32) < (unsigned long long)r) break;
    const int32_t tag = __ldg(tags + r);
    if (tag < 0 || (tenant >= 0 && tag != tenant)) continue;
    const float4* g = reinterpret_cast<const float4*>(master + r * DIM) + lane;
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float4 x = __ldg(g + j * 32);
      a = fmaf(x.x, qv[j].x, a); a = fmaf(x.y, qv[j].y, a); a = fmaf(x.z, qv[j].z, a); a = fmaf(x.w, qv[j].w, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    const bool pass = strict ? (a > threshold) : (a >= threshold);     // false for NaN
    if (pass) {
      if (lane == 0) atomicMin(best + q, ((unsigned long long)r << 32) | __float_as_uint(a));
      break;                                                           // later rows of this warp are higher
    }
  }
}

__global__ void first_match_init_kernel(unsigned long long* best, int nq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) best[i] = ~0ull;
}

__global__ void first_match_unpack_kernel(const unsigned long long* __restrict__ best, int nq, int64_t row_offset,
                                          int64_t* __restrict__ out_rows, float* __restrict__ out_scores) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const unsigned long long b = best[i];
  const bool found = b != ~0ull;
  out_rows[i] = found ? int64_t(b >> 32) + row_offset : int64_t(kNoRow);
  out_scores[i] = found ? __uint_as_float(uint32_t(b & 0xffffffffu)) : kNoScore;
}

int launch_first_match(const float* master, const int32_t* tags, int64_t rows, int dim, const float* qn, int nq,
                       int32_t tenant, float threshold, bool strict, int64_t row_offset, unsigned long long* scratch,
                       int sm_count, int64_t* out_rows, float* out_scores, cudaStream_t st) {
  if (nq <= 0) return FRG_OK;
  if (rows > 0xffffffffLL) { set_error("first_match: more than 2^32 rows in one shard"); return FRG_ERR_UNSUPPORTED; }
  first_match_init_kernel<<<(nq + 127) / 128, 128, 0, st>>>(scratch, nq);
  note_launch(nullptr);
  if (rows > 0) {
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = int64_t(sm_count) * 4;
    if (blocks > cap) blocks = cap;
    const dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(nq));
    const int s = strict ? 1 : 0;
    switch (dim) {
      case 128: first_match_kernel<1><<<grid, 256, 0, st>>>(master, tags, rows, qn, tenant, threshold, s, scratch); break;
      case 256: first_match_kernel<2><<<grid, 256, 0, st>>>(master, tags, rows, qn, tenant, threshold, s, scratch); break;
      case 512: first_match_kernel<4><<<grid, 256, 0, st>>>(master, tags, rows, qn, tenant, threshold, s, scratch); break;
      case 1024: first_match_kernel<8><<<grid, 256, 0, st>>>(master, tags, rows, qn, tenant, threshold, s, scratch); break;
      default: set_error("first_match: dim %d not built (128, 256, 512, 1024)", dim); return FRG_ERR_UNSUPPORTED;
    }
    note_launch(nullptr);
  }
  first_match_unpack_kernel<<<(nq + 127) / 128, 128, 0, st>>>(scratch, nq, row_offset, out_rows, out_scores);
  note_launch("first_match");
  FRG_CUDA(cudaGetLastError());
  return FRG_OK;
}

}  // namespace frg
