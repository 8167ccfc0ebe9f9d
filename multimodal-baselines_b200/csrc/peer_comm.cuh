// One-shot all-reduce over NVLink peer memory (one process per GPU, buffers shared by CUDA IPC).
//
// The only exchange step of the SIF path is the sum of the per-rank d x d Grams (360 KB at
// d = 300): far below the bandwidth regime, so the cost of a library
// all-reduce is its launch + protocol latency.  Here the exchange is part of the kernel that
// finishes the Gram: every rank publishes its local sum in its own exchange buffer, raises a
// flag in every peer's buffer, waits for the peers' flags, and then reads all ranks' buffers
// directly over NVLink, adding them in RANK ORDER -- so every rank computes bit-identical
// results (the replicated component solve relies on that) with no second pass.
//
// Exchange buffer of one rank (cudaMalloc'ed by mmb_comm_alloc, opened by the peers through
// cudaIpcOpenMemHandle):
//   [0, kCommSlotBytes)                 slot 0   data of even epochs
//   [kCommSlotBytes, 2 kCommSlotBytes)  slot 1   data of odd epochs
//   then  uint64 flags[2][kCommMaxRanks]         flags[e & 1][r] = e once rank r's data of epoch e is complete
//         uint32 counter                         CTAs of the local kernel that have published
// Double buffering is enough: a rank can only start epoch e + 2 after it has seen every peer's
// flag of epoch e + 1, which a peer raises only after its epoch-e kernel (including its reads of
// this rank's slot) has completed in stream order.
#pragma once

#include "common.cuh"

namespace mmb {

constexpr int kCommMaxRanks = 8;
// One slot holds the largest vector exchanged in one call: the d x d Gram (360 KB at d = 300), the
// N < d start block, and the flat head-parameter gradient of the data-parallel MMB step (843,400
// floats = 3.4 MB for MMB2).
constexpr size_t kCommSlotBytes = 4 * 1024 * 1024;
constexpr unsigned long long kCommAbortBit = 1ull << 63;   // flag value = epoch | abort: the rank gave up
constexpr size_t kCommFlagsOffset = 2 * kCommSlotBytes;
constexpr size_t kCommCounterOffset = kCommFlagsOffset + 2 * kCommMaxRanks * sizeof(unsigned long long);
constexpr size_t kCommBytes = kCommCounterOffset + 256;

// Largest co-resident grid of `kernel` (threads per CTA, dynamic smem), capped at `want` CTAs.
int coop_grid(const void* kernel, int threads, size_t smem, int64_t want);

struct PeerComm {
  void* buf[kCommMaxRanks];   // buf[rank] is this rank's own buffer
  int rank, world;
  unsigned long long epoch;   // > 0, same sequence on every rank
};

__device__ __forceinline__ char* comm_slot(const PeerComm& c, int r) {
  return (char*)c.buf[r] + (size_t)(c.epoch & 1ull) * kCommSlotBytes;
}

// Called by every CTA after its threads have written their part of the local data into
// comm_slot(c, c.rank) (and passed a __syncthreads()).  The last CTA to arrive raises this
// rank's flag in every rank's buffer.
__device__ __forceinline__ void comm_publish(const PeerComm& c) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned* counter = (unsigned*)((char*)c.buf[c.rank] + kCommCounterOffset);
    const unsigned prev = atomicAdd(counter, 1u);
    if (prev == gridDim.x * gridDim.y - 1) {
      *counter = 0u;   // every CTA has arrived; ready for the next launch
      __threadfence_system();
      for (int p = 0; p < c.world; ++p) {
        volatile unsigned long long* flag =
            (volatile unsigned long long*)((char*)c.buf[p] + kCommFlagsOffset) + (c.epoch & 1ull) * kCommMaxRanks + c.rank;
        *flag = c.epoch;
      }
      __threadfence_system();
    }
  }
}

// Every CTA: wait until all ranks' data of this epoch is complete.  Returns false after
// ~4 s without progress (a peer died) or as soon as a peer signals an abort (mmb_comm_abort): the
// caller sets MMB_STATUS_COMM_TIMEOUT and returns.
//
// The waiting CTAs depend on the LAST-arriving local CTA raising the flags, so the whole grid must be
// co-resident: every kernel that calls comm_wait is launched with cudaLaunchCooperativeKernel on a
// grid no larger than the occupancy API allows (coop_grid() in peer_comm.cu) and strides over its
// elements -- a concurrent kernel on another stream can delay the launch, never deadlock it.
__device__ __forceinline__ bool comm_wait(const PeerComm& c) {
  __shared__ int ok_s;
  if (threadIdx.x == 0) ok_s = 1;
  __syncthreads();
  if (threadIdx.x < c.world) {
    volatile unsigned long long* flag =
        (volatile unsigned long long*)((char*)c.buf[c.rank] + kCommFlagsOffset) + (c.epoch & 1ull) * kCommMaxRanks + threadIdx.x;
    const long long t0 = clock64();
    for (;;) {
      const unsigned long long f = *flag;
      if (f == c.epoch) break;
      if (f == (c.epoch | kCommAbortBit)) { ok_s = 0; break; }     // the peer failed before the exchange
      if (clock64() - t0 > 8000000000ll) { ok_s = 0; break; }
      __nanosleep(64);
    }
    __threadfence_system();
  }
  __syncthreads();
  return ok_s != 0;
}

}  // namespace mmb
