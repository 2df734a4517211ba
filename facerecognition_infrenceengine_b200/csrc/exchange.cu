// Multi-GPU tail as ONE kernel over NVLink peer memory: instead of an NCCL
// all-gather of every rank's local top-k followed by a merge launch, each rank
//   1. PUSHES its packed [rows | scores] block into slot `rank` of every rank's exchange buffer
//      (plain 8-byte stores to peer-mapped pointers: NVLink P2P writes, full NVSwitch bandwidth to all peers),
//   2. publishes the call's epoch in every rank's flag word (release at system scope, after a
//      last-CTA-done ticket so that all of its pushes are ordered before the flag),
//   3. waits until all ranks' flags carry the epoch (acquire at system scope), and
//   4. merges the `world` lists of each query from its OWN buffer (local HBM reads).
// The exchange is latency-sized (F*k*12 bytes per rank: 61 KB at F = 1024, k = 5), so what this saves is
// launches and NCCL's proxy/handshake latency, not bytes.  Buffers are double-buffered by epoch parity:
// a rank can be at most one call ahead of a peer (call e+1 cannot finish before the peer has published
// e+1, which it does only after its call e has completed in stream order), so slot parity e&1 is never
// overwritten while a slower peer still merges call e.
//
// Exchange buffer of every rank (identical layout, allocated symmetrically by the host side):
//   [0, 256)            uint32 flags[world]   flags[r] = last epoch rank r has fully pushed here
//   [256, 512)          uint32 ticket         local: CTAs of the running call that finished pushing
//   [512 + (parity*world + r) * block_cap ...)  rank r's block: int64 rows[nq*k], float scores[nq*k]
#include "merge_device.cuh"

namespace frg {

constexpr int kExchangeHeader = 512;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int KMAX>
__global__ void __launch_bounds__(128)
exchange_merge_kernel(unsigned char* const* __restrict__ peer_bufs, int rank, int world, int64_t block_cap,
                      uint32_t epoch, const int64_t* __restrict__ local_rows,
                      const float* __restrict__ local_scores, int nq, int k, int metric, float threshold,
                      int64_t* __restrict__ out_rows, float* __restrict__ out_scores,
                      uint8_t* __restrict__ out_accept) {
  const int parity = int(epoch & 1u);
  const size_t slot_off = size_t(kExchangeHeader) + (size_t(parity) * world + rank) * size_t(block_cap);
  const int64_t nslots = int64_t(nq) * k;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = int64_t(gridDim.x) * blockDim.x;

  // 1. push: rows (8 B each), then scores two at a time (nq*k is even: the host side checks)
  const uint2* src_r = reinterpret_cast<const uint2*>(local_rows);
  const uint2* src_s = reinterpret_cast<const uint2*>(local_scores);
  const int64_t n_r = nslots, n_s = nslots / 2;
  for (int i = 0; i < world; ++i) {
    const int peer = (rank + 1 + i) % world;              // start with the neighbour: spreads the first stores
    uint2* dst_r = reinterpret_cast<uint2*>(peer_bufs[peer] + slot_off);
    uint2* dst_s = dst_r + n_r;
    for (int64_t j = tid; j < n_r; j += nthreads) dst_r[j] = src_r[j];
    for (int64_t j = tid; j < n_s; j += nthreads) dst_s[j] = src_s[j];
  }

  // 2. all pushes of this rank ordered before its flags: fence per thread, ticket per CTA, last CTA publishes
  __threadfence_system();
  __syncthreads();
  unsigned char* mine = peer_bufs[rank];
  uint32_t* ticket = reinterpret_cast<uint32_t*>(mine + 256);
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0;                                        // for the next call (stream-ordered after this one)
      __threadfence_system();
      for (int i = 0; i < world; ++i) {
        const int peer = (rank + 1 + i) % world;
        st_release_sys(reinterpret_cast<uint32_t*>(peer_bufs[peer]) + rank, epoch);
      }
    }
    // 3. wait for every rank's block of this call (bounded: a dead peer must not hang the GPU for ever)
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(mine);
    const long long t0 = clock64();
    for (int r = 0; r < world; ++r) {
      // ">= epoch", wrap-safe: a faster peer may already have published the NEXT call's epoch here
      while (int32_t(ld_acquire_sys(flags + r) - epoch) < 0) {
        __nanosleep(200);
        if (clock64() - t0 > 60000000000ll) __trap();     // ~30 s at 2 GHz
      }
    }
  }
  __syncthreads();

  // 4. merge the `world` best-first lists of each query from the local buffer: one warp per query
  const unsigned char* base = mine + size_t(kExchangeHeader) + size_t(parity) * world * size_t(block_cap);
  const int64_t* rows = reinterpret_cast<const int64_t*>(base);
  const float* scores = reinterpret_cast<const float*>(base + size_t(nslots) * 8);
  const int warps_per_cta = blockDim.x >> 5;
  for (int q = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); q < nq; q += gridDim.x * warps_per_cta)
    merge_one<int64_t, KMAX>(scores, rows, world, nq, k, k, metric, threshold, 0, 0, q, q, block_cap / 4,
                             block_cap / 8, out_rows, out_scores, out_accept);
}

int launch_exchange_merge(unsigned char* const* peer_bufs, int rank, int world, int64_t block_cap, uint32_t epoch,
                          const int64_t* local_rows, const float* local_scores, int nq, int k, int metric,
                          float threshold, int sm_count, int64_t* out_rows, float* out_scores,
                          uint8_t* out_accept, cudaStream_t st) {
  if (nq <= 0) return FRG_OK;
  // every CTA spins on the flags: the whole grid must be resident at once
  int grid = (nq + 3) / 4;
  if (grid > sm_count) grid = sm_count;
  if (grid < 1) grid = 1;
#define FRG_XM(K)                                                                                              \
  exchange_merge_kernel<K><<<grid, 128, 0, st>>>(peer_bufs, rank, world, block_cap, epoch, local_rows,         \
                                                 local_scores, nq, k, metric, threshold, out_rows, out_scores, \
                                                 out_accept)
  if (k == 1) FRG_XM(1);
  else if (k <= 4) FRG_XM(4);
  else if (k <= 8) FRG_XM(8);
  else FRG_XM(16);
#undef FRG_XM
  note_launch(nullptr);
  FRG_CUDA(cudaGetLastError());
  return FRG_OK;
}

}  // namespace frg
