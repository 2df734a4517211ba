// Multi-GPU tail over NVLink peer memory, no collective-library call on the data path.
//
// Every rank's exchange buffer (peer-mapped into all ranks, e.g. a torch symmetric-memory allocation) holds,
// per call parity and source rank, the source's local top-k block as PACKETS: 8-byte words {payload, epoch},
// stored with single 8-byte writes, so a packet is valid exactly when its epoch word matches - the data is
// its own flag (the low-latency protocol of collective libraries): no fence, no separate flag, no grid-wide
// rendezvous.  Three packet planes per block: row low word, row high word, score bits.
//
//   * select_rescore_kernel (tc_match.cu) pushes each query's final local top-k to every rank THE MOMENT it
//     is final - the exchange overlaps the rest of the select kernel and the fallback launch;
//   * exchange_merge_kernel then pushes what select could not (queries redone by the exact fallback; every
//     query for variants without a select stage) and merges: one warp per query polls the `world` lists of
//     that query in its OWN buffer until their packets carry the call's epoch, and folds them.
//
// Double buffering by epoch parity: a rank can be at most one call ahead of a peer (it cannot finish call
// e+1 before the peer has pushed e+1, which the peer does only after its call e completed in stream order),
// so the packets of parity e&1 are never overwritten while a slower peer still reads call e.
#include "merge_device.cuh"

namespace frg {

__device__ __forceinline__ uint2 ld_packet(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}

// the packet planes this rank RECEIVES from `src` for the call's parity
struct PacketLists {
  const unsigned char* mine;       // this rank's exchange buffer
  int world;
  int64_t block_cap, nslots;
  uint32_t epoch;
  __device__ __forceinline__ void load(int part, size_t in_part, int64_t* r, float* s) const {
    const uint2* b = reinterpret_cast<const uint2*>(
        mine + kExchangeHeader + (size_t(epoch & 1u) * world + part) * size_t(block_cap));
    uint2 lo, hi, sc;
    const long long t0 = clock64();
    // bounded: a dead peer must not hang the GPU for ever (~30 s at 2 GHz, then the kernel traps)
    while ((lo = ld_packet(b + in_part)).y != epoch) { __nanosleep(100); if (clock64() - t0 > 60000000000ll) __trap(); }
    while ((hi = ld_packet(b + nslots + in_part)).y != epoch) { __nanosleep(100); if (clock64() - t0 > 60000000000ll) __trap(); }
    while ((sc = ld_packet(b + 2 * nslots + in_part)).y != epoch) { __nanosleep(100); if (clock64() - t0 > 60000000000ll) __trap(); }
    *r = int64_t((uint64_t(hi.x) << 32) | uint64_t(lo.x));
    *s = __uint_as_float(sc.x);
  }
};

// push_list / n_push: the queries select did not push (device-resident list, may be null = none);
// push_all: nothing was pushed yet (variants without a select stage, or results computed elsewhere)
template <int KMAX>
__global__ void __launch_bounds__(128)
exchange_merge_kernel(const XPush x, const int64_t* __restrict__ local_rows,
                      const float* __restrict__ local_scores, const int* __restrict__ push_list,
                      const int* __restrict__ n_push, int push_all, int nq, int k, int metric, float threshold,
                      int64_t* __restrict__ out_rows, float* __restrict__ out_scores,
                      uint8_t* __restrict__ out_accept) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  pdl_wait();             // the local match's results and its list of late queries

  // 1. late pushes: one warp per query, lanes over (peer, slot)
  const int np = push_all ? nq : (n_push ? *n_push : 0);
  for (int i = warp; i < np; i += nwarps) {
    const int q = push_all ? i : push_list[i];
    for (int c = lane; c < x.world * k; c += 32) {
      const int peer = c / k, j = c - peer * k;
      const int64_t slot = int64_t(q) * k + j;
      xpush_slot(x, (x.rank + 1 + peer) % x.world, slot, local_rows[slot], local_scores[slot]);
    }
  }

  // 2. merge: one warp per query polls and folds the `world` lists of that query
  const PacketLists lists{x.peer_bufs[x.rank], x.world, x.block_cap, x.nslots, x.epoch};
  for (int q = warp; q < nq; q += nwarps)
    merge_lists<int64_t, KMAX>(lists, x.world, k, k, metric, threshold, 0, 0, q, q, out_rows, out_scores, out_accept);
}

int launch_exchange_merge(const XPush& x, const int64_t* local_rows, const float* local_scores,
                          const int* push_list, const int* n_push, bool push_all, int nq, int k, int metric,
                          float threshold, int sm_count, int64_t* out_rows, float* out_scores,
                          uint8_t* out_accept, cudaStream_t st) {
  if (nq <= 0) return FRG_OK;
  int grid = (nq + 3) / 4;
  if (grid > 4 * sm_count) grid = 4 * sm_count;
  if (grid < 1) grid = 1;
  const int pa = push_all ? 1 : 0;
#define FRG_XM(K)                                                                                                 \
  FRG_CUDA(launch_kernel(exchange_merge_kernel<K>, dim3(grid), dim3(128), 0, st, true, x, local_rows, local_scores,  \
                         push_list, n_push, pa, nq, k, metric, threshold, out_rows, out_scores, out_accept))
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
