// Exchange-buffer management (CUDA IPC) and the stand-alone peer all-reduce; see peer_comm.cuh.
#include "peer_comm.cuh"

namespace mmb {

int gram_tc_allreduce(const float* X, int64_t N, int d, float* G, void* ws, size_t ws_bytes, const PeerComm& comm,
                      int* status, cudaStream_t st);
bool gram_tc_supported(int64_t N, int d);

template <typename T>
__global__ void __launch_bounds__(256)
    peer_allreduce_kernel(T* __restrict__ x, int n, const PeerComm comm, int* __restrict__ status) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  T* mine = (T*)comm_slot(comm, comm.rank);
  if (idx < n) mine[idx] = x[idx];
  __syncthreads();
  comm_publish(comm);
  if (!comm_wait(comm)) {
    if (threadIdx.x == 0) atomicOr(status, MMB_STATUS_COMM_TIMEOUT);
    return;
  }
  if (idx < n) {
    T s = (T)0;
    for (int r = 0; r < comm.world; ++r) s += *((const volatile T*)comm_slot(comm, r) + idx);
    x[idx] = s;
  }
}

static int make_comm(PeerComm* c, int rank, int world, void* const* bufs, uint64_t epoch) {
  MMB_REQUIRE(world >= 1 && world <= kCommMaxRanks && rank >= 0 && rank < world, "need 0 <= rank < world <= 8");
  MMB_REQUIRE(bufs && epoch > 0, "bufs must be given and epoch must be > 0");
  for (int r = 0; r < kCommMaxRanks; ++r) c->buf[r] = r < world ? bufs[r] : nullptr;
  for (int r = 0; r < world; ++r) MMB_REQUIRE(bufs[r], "null exchange buffer");
  c->rank = rank;
  c->world = world;
  c->epoch = epoch;
  return MMB_OK;
}

}  // namespace mmb

using namespace mmb;

extern "C" size_t mmb_comm_bytes(void) { return kCommBytes; }

extern "C" int mmb_comm_alloc(void** buf) {
  MMB_REQUIRE(buf, "null pointer");
  MMB_CUDA(cudaMalloc(buf, kCommBytes));
  MMB_CUDA(cudaMemset(*buf, 0, kCommBytes));
  MMB_CUDA(cudaDeviceSynchronize());
  return MMB_OK;
}

extern "C" int mmb_comm_free(void* buf) {
  if (buf) MMB_CUDA(cudaFree(buf));
  return MMB_OK;
}

extern "C" int mmb_comm_export(void* buf, void* handle64) {
  MMB_REQUIRE(buf && handle64, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t h;
  MMB_CUDA(cudaIpcGetMemHandle(&h, buf));
  memcpy(handle64, &h, sizeof(h));
  return MMB_OK;
}

extern "C" int mmb_comm_open(const void* handle64, void** peer_buf) {
  MMB_REQUIRE(handle64 && peer_buf, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  MMB_CUDA(cudaIpcOpenMemHandle(peer_buf, h, cudaIpcMemLazyEnablePeerAccess));
  return MMB_OK;
}

extern "C" int mmb_comm_close(void* peer_buf) {
  if (peer_buf) MMB_CUDA(cudaIpcCloseMemHandle(peer_buf));
  return MMB_OK;
}

extern "C" int mmb_allreduce_peer(void* x, int64_t n, int is_f64, int rank, int world, void* const* bufs,
                                  uint64_t epoch, int* status, mmb_stream_t stream) {
  MMB_REQUIRE(x && status && n > 0, "null pointer / empty vector");
  MMB_REQUIRE((size_t)n * (is_f64 ? 8 : 4) <= kCommSlotBytes, "vector larger than the exchange slot (512 KiB)");
  PeerComm c;
  int rc = make_comm(&c, rank, world, bufs, epoch);
  if (rc) return rc;
  const int grid = (int)ceil_div(n, 256);
  if (is_f64)
    peer_allreduce_kernel<double><<<grid, 256, 0, as_stream(stream)>>>((double*)x, (int)n, c, status);
  else
    peer_allreduce_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((float*)x, (int)n, c, status);
  MMB_LAUNCH_CHECK("peer_allreduce");
  return MMB_OK;
}

extern "C" int mmb_gram_allreduce_peer(const float* X, int64_t N, int d, float* G, void* ws, size_t ws_bytes,
                                       int mode, int rank, int world, void* const* bufs, uint64_t epoch,
                                       int* status, mmb_stream_t stream) {
  MMB_REQUIRE(G && status && d > 0 && N >= 0, "bad argument");
  MMB_REQUIRE((size_t)d * d * 4 <= kCommSlotBytes, "d x d Gram larger than the exchange slot");
  PeerComm c;
  int rc = make_comm(&c, rank, world, bufs, epoch);
  if (rc) return rc;
  const bool tc = (mode == MMB_GRAM_TF32X3 || (mode == MMB_GRAM_AUTO && N >= 4096)) && N > 0 && gram_tc_supported(N, d);
  if (tc) {
    MMB_REQUIRE(X && ws, "null pointer");
    return gram_tc_allreduce(X, N, d, G, ws, ws_bytes, c, status, as_stream(stream));
  }
  if (N > 0) {
    rc = mmb_gram(X, N, d, G, ws, ws_bytes, mode == MMB_GRAM_AUTO ? MMB_GRAM_FP32 : mode, stream);
    if (rc) return rc;
  } else {
    MMB_CUDA(cudaMemsetAsync(G, 0, (size_t)d * d * sizeof(float), as_stream(stream)));
  }
  return mmb_allreduce_peer(G, (int64_t)d * d, 0, rank, world, bufs, epoch, status, stream);
}
