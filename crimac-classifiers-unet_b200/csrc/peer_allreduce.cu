// Gradient all-reduce over NVLink 5 / NVSwitch peer memory, as ONE kernel per bucket that the train step launches on a
// communication stream WHILE the rest of backward is still running.  The reference has no
// multi-GPU path; the semantics are DistributedDataParallel's: every replica ends up with the SUM over replicas of the
// flat fp32 gradient arena (the 1/world factor is folded into the SGD kernel), bit-identical on all replicas.
//
// Memory model: every rank's gradient arena is a symmetric allocation (torch.distributed._symmetric_memory: cuMemMap'ed
// on all GPUs of the node), so rank r holds device pointers to ALL arenas (`peers`), optionally the NVSwitch multicast
// address of the same allocation (`mc`), and a small symmetric signal pad per rank.
//
// Protocol of one bucket [off, off + count) (count in floats, multiple of 4), epoch e = number of earlier launches + 1:
//   1. ready:  rank r stores e into pad[p][bucket][0][r] of every peer p  (its gradients of this bucket are final: they were
//              written by earlier kernels of the same stream / graph).
//   2. every CTA waits until pad[r][bucket][0][0..world) >= e.
//   3. rank r owns the r-th slice of the bucket.  For each 16-byte vector of its slice it sums the `world` copies in
//      FIXED rank order 0..world-1 (one local + world-1 NVLink loads) and stores the sum into all `world` arenas (two-shot
//      all-reduce in one pass: reduce-scatter + all-gather).  With a multicast address the same is two instructions:
//      multimem.ld_reduce.add.v4.f32 (the switch adds) + multimem.st.
//   4. done:   the last CTA of rank r (atomic ticket) stores e into pad[p][bucket][1][r] of every peer.
//   5. every CTA waits until pad[r][bucket][1][0..world) >= e before it exits: all slices have landed in this rank's arena
//      (the SGD kernel is stream-ordered behind this kernel) and every peer has finished READING this rank's arena (the next
//      step may overwrite it).
// The epoch lives in device memory (state[bucket]) and is advanced by the kernel itself, so a captured CUDA graph can be
// replayed.  No CTA ever waits for another CTA of its own grid (only for remote flags), so the kernel cannot deadlock on
// partial residency; it uses 128 threads x <= 64 registers and no shared memory, and therefore fits on an SM next to
// a resident backward-data / weight-gradient CTA: the transfer overlaps the tensor-core work instead of displacing it.
#include "host_util.h"
#include "../../include/crimac_b200.h"

namespace {

constexpr int AR_THREADS = 128;
constexpr int AR_MAX_WORLD = 8;    // replicas per node (one process per GPU of an 8-GPU NVSwitch box)
constexpr int AR_PAD_SLOTS = 16;   // flag slots per (bucket, phase) row of the signal pad

struct ArParams {
  float* peers[AR_MAX_WORLD];      // arena base of every rank (peers[rank] = the local one)
  uint32_t* pads[AR_MAX_WORLD];    // signal pad base of every rank
  float* mc;                       // multicast address of the arena, or nullptr
  uint32_t* state;                 // local: [bucket][0] = epoch of the last finished launch, [bucket][1] = CTA ticket
  int rank, world, bucket;
  long off4, count4;               // bucket extent in float4 units
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float4* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float4* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// pad layout (uint32): [bucket][phase 0|1][source rank], AR_PAD_SLOTS slots per row
__device__ __forceinline__ uint32_t* pad_slot(uint32_t* pad, int bucket, int phase, int src) {
  return pad + (bucket * 2 + phase) * AR_PAD_SLOTS + src;
}

__device__ __forceinline__ void wait_all(uint32_t* pad, int bucket, int phase, int world, uint32_t epoch) {
  // lanes 0..world-1 of warp 0 each watch one source rank; bounded so that a lost peer traps instead of hanging the GPU
  if (threadIdx.x < world) {
    const uint32_t* slot = pad_slot(pad, bucket, phase, threadIdx.x);
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(slot) - epoch) < 0) {
      __nanosleep(200);
      if (clock64() - t0 > 20000000000LL) {  // ~10 s
        printf("crimac_b200: peer all-reduce timed out waiting for rank %d (bucket %d phase %d epoch %u)\n",
               static_cast<int>(threadIdx.x), bucket, phase, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
}

template <bool MC>
__global__ void __launch_bounds__(AR_THREADS, 8) peer_allreduce_kernel(const __grid_constant__ ArParams p) {
  uint32_t* my_pad = p.pads[p.rank];
  const uint32_t epoch = p.state[p.bucket * 2] + 1;  // every CTA reads it before the last CTA advances it (see the end)
  // 1. ready flags (CTA 0).  The gradients were produced by earlier kernels in stream order: globally visible already.
  if (blockIdx.x == 0 && threadIdx.x < p.world) st_release_sys(pad_slot(p.pads[threadIdx.x], p.bucket, 0, p.rank), epoch);
  // 2. wait for every rank's bucket
  wait_all(my_pad, p.bucket, 0, p.world, epoch);
  // 3. reduce my slice, broadcast the sums
  const long per = (p.count4 + p.world - 1) / p.world;
  const long lo = p.off4 + min(p.count4, per * p.rank), hi = p.off4 + min(p.count4, per * (p.rank + 1));
  const long stride = static_cast<long>(gridDim.x) * AR_THREADS;
  if (MC) {
    float4* mc = reinterpret_cast<float4*>(p.mc);
    constexpr int U = 4;
    for (long i0 = lo + blockIdx.x * static_cast<long>(AR_THREADS) + threadIdx.x; i0 < hi; i0 += U * stride) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i0 + u * stride < hi) v[u] = mc_ld_reduce(mc + i0 + u * stride);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i0 + u * stride < hi) mc_st(mc + i0 + u * stride, v[u]);
    }
  } else {
    // NVLink latency (~2 us) bounds a thread with one load in flight: each thread keeps U vectors x G peers = 8 loads
    // (128 bytes) in flight - 128 CTAs x 128 threads = 2 MB per rank - and walks the peers in FIXED rank order (the owner
    // of a slice is the only one that sums it, so every replica receives the same bits)
    constexpr int U = 4, G = 2;
    for (long i0 = lo + blockIdx.x * static_cast<long>(AR_THREADS) + threadIdx.x; i0 < hi; i0 += U * stride) {
      float4 acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r0 = 0; r0 < AR_MAX_WORLD; r0 += G) {
        if (r0 < p.world) {
          float4 v[G][U];
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int u = 0; u < U; ++u)
              v[g][u] = (r0 + g < p.world && i0 + u * stride < hi)
                            ? ld_peer(reinterpret_cast<const float4*>(p.peers[r0 + g]) + i0 + u * stride)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int u = 0; u < U; ++u) {
              acc[u].x += v[g][u].x;
              acc[u].y += v[g][u].y;
              acc[u].z += v[g][u].z;
              acc[u].w += v[g][u].w;
            }
        }
      }
#pragma unroll
      for (int r = 0; r < AR_MAX_WORLD; ++r)
        if (r < p.world) {
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (i0 + u * stride < hi) st_peer(reinterpret_cast<float4*>(p.peers[r]) + i0 + u * stride, acc[u]);
        }
    }
  }
  // 4. done flags: my stores must be visible system-wide before any peer sees the flag
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t ticket = atomicAdd(&p.state[p.bucket * 2 + 1], 1u);
    if (ticket == gridDim.x - 1) {
      p.state[p.bucket * 2 + 1] = 0;
      __threadfence_system();
      for (int r = 0; r < p.world; ++r) st_release_sys(pad_slot(p.pads[r], p.bucket, 1, p.rank), epoch);
    }
  }
  // 5. wait until every rank is done (all slices have landed here; nobody still reads my arena)
  wait_all(my_pad, p.bucket, 1, p.world, epoch);
  // advance the epoch once per launch: the LAST CTA to get here (second ticket round on a separate counter would need
  // another word; instead CTA 0 does it - every CTA has read `epoch` at its very start, and no CTA of the NEXT launch can
  // start before this grid has completed in stream order)
  if (blockIdx.x == 0 && threadIdx.x == 0) p.state[p.bucket * 2] = epoch;
}

}  // namespace

// Invariant used above: a CTA that started after CTA 0 had advanced the epoch would read epoch + 1.  It cannot happen:
// CTA 0 advances the epoch only after step 5, i.e. after this rank's own "done" flag, which the ticket of step 4 raises
// only once ALL CTAs of this grid have passed step 4 - every CTA has read `epoch` long before.

extern "C" int crimac_peer_allreduce(float* const* peer_arenas, void* const* peer_pads, float* multicast_arena,
                                     void* local_state, int rank, int world, int bucket, size_t offset, size_t count,
                                     int ctas, void* stream) {
  CRIMAC_REQUIRE(peer_arenas && peer_pads && local_state, "NULL argument");
  CRIMAC_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "rank / world (<= 8 replicas per node)");
  CRIMAC_REQUIRE(bucket >= 0 && bucket < CRIMAC_AR_MAX_BUCKETS, "bucket index");
  CRIMAC_REQUIRE(offset % 4 == 0 && count % 4 == 0, "bucket offset and size must be multiples of 4 floats");
  if (count == 0) return 0;
  ArParams p{};
  for (int r = 0; r < world; ++r) {
    CRIMAC_REQUIRE(peer_arenas[r] != nullptr && peer_pads[r] != nullptr, "NULL peer pointer");
    CRIMAC_REQUIRE((reinterpret_cast<uintptr_t>(peer_arenas[r]) & 15) == 0, "arenas must be 16-byte aligned");
    p.peers[r] = peer_arenas[r];
    p.pads[r] = static_cast<uint32_t*>(peer_pads[r]);
  }
  p.mc = multicast_arena;
  p.state = static_cast<uint32_t*>(local_state);
  p.rank = rank;
  p.world = world;
  p.bucket = bucket;
  p.off4 = static_cast<long>(offset / 4);
  p.count4 = static_cast<long>(count / 4);
  if (ctas <= 0) ctas = 128;
  const long per = (p.count4 + world - 1) / world;
  const long want = (per + AR_THREADS - 1) / AR_THREADS;
  if (ctas > want) ctas = static_cast<int>(want < 1 ? 1 : want);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps("peer_allreduce", 0, static_cast<double>(count) * 4.0 * 2.0 * (world - 1) / world, st);
  if (multicast_arena != nullptr)
    peer_allreduce_kernel<true><<<ctas, AR_THREADS, 0, st>>>(p);
  else
    peer_allreduce_kernel<false><<<ctas, AR_THREADS, 0, st>>>(p);
  CRIMAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}
