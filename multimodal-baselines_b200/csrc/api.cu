// libmmb_b200.so: library plumbing (errors, device info, pinned memory), the Gram dispatcher
// and the composed SIF pipelines (device-resident and host-buffer variants).
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace mmb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return MMB_E_CUDA;
}

static unsigned long long g_launches = 0;
static char g_last_kernel[8][160];   // last kernel name recorded per stage tag (see note_kernel)

void count_launch(const char*) { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

// Kernels chosen by a run-time dispatch record their (template-expanded) name here so that the
// caller can say WHICH instantiation ran: tag 0 = embed.
void note_kernel(int tag, const char* name) {
  if (tag < 0 || tag >= 8) return;
  strncpy(g_last_kernel[tag], name, sizeof(g_last_kernel[tag]) - 1);
}

// Run-time switches (mmb_set_option); -1 = unset -> the environment variable / built-in default decides.
static int g_opt_embed_hot = -1, g_opt_embed_prescale = -1, g_opt_embed_warm = -1;
static int g_opt_overlap_sms = -1, g_opt_overlap_chunks = -1, g_opt_overlap_grid = -1;
static int env_int(const char* name, int dflt) { return getenv(name) ? atoi(getenv(name)) : dflt; }
// Gram beside the embed (mmb_sif_embed_gram): SMs given to the Gram (0 = no overlap, one stage after the other),
// chunks per call, embed CTAs per SM.
static int option_overlap_sms() {
  static const int env = env_int("MMB_OVERLAP_SMS", 0);   // measured slower than one stage after the other: off
  return g_opt_overlap_sms >= 0 ? g_opt_overlap_sms : env;
}
static int option_overlap_chunks() {
  static const int env = env_int("MMB_OVERLAP_CHUNKS", 10);
  const int v = g_opt_overlap_chunks > 0 ? g_opt_overlap_chunks : env;
  return v < 2 ? 2 : (v > 64 ? 64 : v);
}
static int option_overlap_grid() {
  static const int env = env_int("MMB_OVERLAP_GRID", 32);
  const int v = g_opt_overlap_grid > 0 ? g_opt_overlap_grid : env;
  return v < 4 ? 4 : (v > 256 ? 256 : v);
}
int option_embed_warm() {
  if (g_opt_embed_warm >= 0) return g_opt_embed_warm;
  static const int env = getenv("MMB_EMBED_WARM") ? atoi(getenv("MMB_EMBED_WARM")) : 0;
  return env < 0 ? 0 : (env > 512 ? 512 : env);
}
int option_embed_hot() {
  if (g_opt_embed_hot >= 0) return g_opt_embed_hot;
  static const int env = getenv("MMB_EMBED_HOT") ? atoi(getenv("MMB_EMBED_HOT")) : 0;
  return env;
}
int option_embed_prescale() {
  if (g_opt_embed_prescale >= 0) return g_opt_embed_prescale;
  static const int env = getenv("MMB_EMBED_PRESCALE") ? atoi(getenv("MMB_EMBED_PRESCALE")) : 1;
  return env;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
      cached = n;
      cached_dev = dev;
    }
  }
  return cached;
}

// implemented in gram_fp32.cu / gram_tc.cu
size_t gram_fp32_workspace_bytes(int64_t N, int d);
int gram_fp32(const float* X, int64_t N, int d, float* G, void* ws, size_t ws_bytes, cudaStream_t st);
size_t gram_tc_workspace_bytes(int64_t N, int d);
int gram_tc(const float* X, int64_t N, int d, float* G, void* ws, size_t ws_bytes, cudaStream_t st, int max_cta = 0);
bool gram_tc_supported(int64_t N, int d);

static int resolve_gram_mode(int64_t N, int d, int mode) {
  // small splits (MOSI / POM sizes) cannot fill a tensor-core pipeline: exact FP32 path
  if (mode == MMB_GRAM_AUTO) return (N >= 4096 && gram_tc_supported(N, d)) ? MMB_GRAM_TF32X3 : MMB_GRAM_FP32;
  return mode;
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

__global__ void accumulate_kernel(float* __restrict__ acc, const float* __restrict__ x, int n, int first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acc[i] = first ? x[i] : acc[i] + x[i];
}

__global__ void f32_to_f64_kernel(const float* __restrict__ in, double* __restrict__ out, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (double)__ldcs(in + i);
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_version(void) { return 100; }

extern "C" const char* mmb_last_error(void) { return g_err; }

extern "C" unsigned long long mmb_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" const char* mmb_last_kernel(int tag) { return (tag >= 0 && tag < 8) ? g_last_kernel[tag] : ""; }

extern "C" int mmb_set_option(const char* name, int value) {
  MMB_REQUIRE(name, "null pointer");
  if (!strcmp(name, "embed_hot")) g_opt_embed_hot = value;
  else if (!strcmp(name, "embed_prescale")) g_opt_embed_prescale = value;
  else if (!strcmp(name, "embed_warm")) g_opt_embed_warm = value < 0 ? 0 : (value > 512 ? 512 : value);
  else if (!strcmp(name, "overlap_sms")) g_opt_overlap_sms = value;
  else if (!strcmp(name, "overlap_chunks")) g_opt_overlap_chunks = value;
  else if (!strcmp(name, "overlap_grid")) g_opt_overlap_grid = value;
  else {
    set_error("mmb_set_option: unknown option '%s'", name);
    return MMB_E_INVALID;
  }
  return MMB_OK;
}

extern "C" int mmb_device_info(int* sm, int* cc_major, int* cc_minor) {
  int dev = 0;
  MMB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  MMB_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm) *sm = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return MMB_OK;
}

extern "C" int mmb_host_alloc(void** ptr, size_t bytes) {
  MMB_REQUIRE(ptr, "null pointer");
  MMB_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return MMB_OK;
}

extern "C" int mmb_host_alloc_wc(void** ptr, size_t bytes) {
  MMB_REQUIRE(ptr, "null pointer");
  MMB_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocWriteCombined));
  return MMB_OK;
}

extern "C" int mmb_host_free(void* ptr) {
  if (ptr) MMB_CUDA(cudaFreeHost(ptr));
  return MMB_OK;
}

extern "C" size_t mmb_gram_workspace_bytes(int64_t N, int d, int mode) {
  size_t a = gram_fp32_workspace_bytes(N, d);
  if (mode != MMB_GRAM_FP32 && gram_tc_supported(N, d)) {
    size_t b = gram_tc_workspace_bytes(N, d);
    if (b > a) a = b;
  }
  return a;
}

extern "C" int mmb_gram(const float* X, int64_t N, int d, float* G, void* ws, size_t ws_bytes, int mode,
                        mmb_stream_t stream) {
  MMB_REQUIRE(X && G && ws, "null pointer");
  MMB_REQUIRE(d > 0 && d % 4 == 0, "d must be a positive multiple of 4");
  MMB_REQUIRE(N >= 0, "negative N");
  MMB_REQUIRE((uintptr_t)X % 16 == 0, "X must be 16-byte aligned");
  mode = resolve_gram_mode(N, d, mode);
  if (mode == MMB_GRAM_FP32) return gram_fp32(X, N, d, G, ws, ws_bytes, as_stream(stream));
  if (mode == MMB_GRAM_TF32X3) {
    if (!gram_tc_supported(N, d)) {
      set_error("mmb_gram: the tcgen05 path needs d == 300 and N >= 1 on an sm_100 device");
      return MMB_E_UNSUPPORTED;
    }
    return gram_tc(X, N, d, G, ws, ws_bytes, as_stream(stream));
  }
  set_error("mmb_gram: unknown mode %d", mode);
  return MMB_E_INVALID;
}

// ---- composed pipeline, everything resident on the device ---------------------------------
namespace mmb {   // implemented in sif_embed.cu
int sif_prescale(const float* table, int64_t V, int d, const float* vocab_w, void* ws, cudaStream_t st);
int sif_embed_prescaled(const float* table, int64_t V, int d, const float* vocab_w, const void* ws, const int64_t* x,
                        int64_t N, int64_t L, float* emb, int* status, cudaStream_t st, int grid_mult = 8);
}  // namespace mmb


// ---- embed + Gram of a block, the Gram of chunk c beside the embed of chunk c + 1 (experiment, OFF by default) --
// Idea: the embed is bound by the L2->SM fabric, the tcgen05 Gram by the tensor pipe and shared memory of the SMs
// it runs on, so on DISJOINT SMs the two should barely compete.  The block is cut into chunks; chunk c's Gram runs
// on `overlap_sms` SMs (one persistent CTA each, launched first from a higher-priority stream so that it gets its
// SMs at the chunk boundary -- a Gram CTA's registers and shared memory leave no room for an embed CTA beside it,
// so the partition is the block scheduler's own) while chunk c + 1 is embedded on the others; only the last
// chunk's Gram (whole GPU) is exposed.  Chunk Grams are added in chunk order: deterministic.
// Measured (tools/overlap_probe.py, profiles/r02_overlap_probe.jsonl, 10 M utterances): 31.4 ms one stage after
// the other, 33.6 ms at best overlapped (40 SMs, 5 chunks) -- the embed slows down in proportion to the SMs it
// loses (its per-SM L1 pipe is at 75 % too), so a partition of the SMs cannot win.  overlap_sms = 0 keeps
// mmb_sif_embed_ws + mmb_gram; the entry point stays for the probe.
namespace {
struct SideStream {
  cudaStream_t st = nullptr;
};
SideStream g_side[64];

int side_stream(cudaStream_t* out) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  MMB_CUDA(cudaGetDevice(&dev));
  MMB_REQUIRE(dev >= 0 && dev < 64, "device ordinal out of range");
  if (!g_side[dev].st) {
    int lo = 0, hi = 0;
    MMB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = numerically lowest = greatest priority
    MMB_CUDA(cudaStreamCreateWithPriority(&g_side[dev].st, cudaStreamNonBlocking, hi));
  }
  *out = g_side[dev].st;
  return MMB_OK;
}
struct EventPair {
  cudaEvent_t fork = nullptr, join = nullptr;
  int init() {
    MMB_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    MMB_CUDA(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
    return MMB_OK;
  }
  ~EventPair() {
    if (fork) cudaEventDestroy(fork);
    if (join) cudaEventDestroy(join);
  }
};
}  // namespace

static bool embed_gram_overlaps(int64_t V, int d, int64_t N, int64_t L, int gram_mode) {
  if (option_overlap_sms() <= 0 || option_overlap_sms() >= sm_count()) return false;
  if (resolve_gram_mode(N, d, gram_mode) != MMB_GRAM_TF32X3 || !gram_tc_supported(N, d)) return false;
  if (mmb_sif_embed_workspace_bytes(V, d, N, L) == 0) return false;
  return N >= (int64_t)option_overlap_chunks() * 32768;     // chunks long enough to amortise their boundaries
}

extern "C" size_t mmb_sif_embed_gram_workspace_bytes(int64_t V, int d, int64_t N, int64_t L, int gram_mode) {
  return align_up(mmb_sif_embed_workspace_bytes(V, d, N, L)) + align_up(mmb_gram_workspace_bytes(N > 0 ? N : 1, d, gram_mode)) +
         align_up((size_t)d * d * sizeof(float));
}

extern "C" int mmb_sif_embed_gram(const float* table, int64_t V, int d, const float* vocab_w, const int64_t* x,
                                  int64_t N, int64_t L, float* emb, int* status, float* G, void* ws, size_t ws_bytes,
                                  int gram_mode, mmb_stream_t stream) {
  MMB_REQUIRE(table && vocab_w && x && emb && status && G && ws, "null pointer");
  MMB_REQUIRE(N > 0, "empty block");
  MMB_REQUIRE(ws_bytes >= mmb_sif_embed_gram_workspace_bytes(V, d, N, L, gram_mode), "workspace too small");
  MMB_REQUIRE((uintptr_t)ws % 256 == 0, "ws must be 256-byte aligned");
  const size_t scaled_bytes = mmb_sif_embed_workspace_bytes(V, d, N, L);
  const size_t gram_bytes = align_up(mmb_gram_workspace_bytes(N, d, gram_mode));
  char* ws_scaled = (char*)ws;
  char* ws_gram = ws_scaled + align_up(scaled_bytes);
  float* gchunk = (float*)(ws_gram + gram_bytes);
  cudaStream_t st = as_stream(stream);
  if (!embed_gram_overlaps(V, d, N, L, gram_mode)) {
    int rc = mmb_sif_embed_ws(table, V, d, vocab_w, x, N, L, emb, status, ws_scaled, scaled_bytes, stream);
    if (rc) return rc;
    return mmb_gram(emb, N, d, G, ws_gram, gram_bytes, gram_mode, stream);
  }
  MMB_REQUIRE(((uintptr_t)table % 16 == 0) && ((uintptr_t)emb % 16 == 0), "table / emb must be 16-byte aligned");
  cudaStream_t side;
  int rc = side_stream(&side);
  if (rc) return rc;
  EventPair ev;
  if ((rc = ev.init())) return rc;
  const int nch = option_overlap_chunks(), gram_sms = option_overlap_sms(), grid_mult = option_overlap_grid();
  const int64_t chunk = ceil_div(ceil_div(N, (int64_t)nch), (int64_t)16) * 16;
  if ((rc = sif_prescale(table, V, d, vocab_w, ws_scaled, st))) return rc;
  for (int64_t r0 = 0, c = 0; r0 < N; r0 += chunk, ++c) {
    const int64_t rows = (N - r0) < chunk ? (N - r0) : chunk;
    const bool last = r0 + rows >= N;
    rc = sif_embed_prescaled(table, V, d, vocab_w, ws_scaled, x + r0 * L, rows, L, emb + r0 * d, status, st, grid_mult);
    if (rc) return rc;
    MMB_CUDA(cudaEventRecord(ev.fork, st));
    MMB_CUDA(cudaStreamWaitEvent(side, ev.fork, 0));
    rc = gram_tc(emb + r0 * d, rows, d, c == 0 ? G : gchunk, ws_gram, gram_bytes, side, last ? 0 : gram_sms);
    if (rc) return rc;
    if (c > 0) {
      accumulate_kernel<<<(d * d + 255) / 256, 256, 0, side>>>(G, gchunk, d * d, 0);
      MMB_LAUNCH_CHECK("accumulate");
    }
  }
  MMB_CUDA(cudaEventRecord(ev.join, side));
  MMB_CUDA(cudaStreamWaitEvent(st, ev.join, 0));
  return MMB_OK;
}

struct SifWs {
  size_t gram, pc, s0, G, pcv, total;
};
static SifWs sif_ws_layout(int64_t N, int d, int npc) {
  SifWs w;
  const int k = npc + 10;
  size_t off = 0;
  w.gram = off; off += align_up(mmb_gram_workspace_bytes(N, d, MMB_GRAM_AUTO));
  w.pc = off;   off += align_up(mmb_pc_workspace_bytes(d, k));
  w.s0 = off;   off += align_up((size_t)d * k * sizeof(double));
  w.G = off;    off += align_up((size_t)d * d * sizeof(float));
  w.pcv = off;  off += align_up((size_t)(npc > 0 ? npc : 1) * d * sizeof(float));
  w.total = off;
  return w;
}

extern "C" size_t mmb_sif_workspace_bytes(int64_t N, int d, int npc) {
  return sif_ws_layout(N, d, npc > 0 ? npc : 1).total;
}

// Gram -> components -> projection on an (N, d) embedding block already on the device.
static int pc_removal_device(float* emb, int64_t N, int d, int npc, const double* Omega, float* pc_out,
                             float* G_out, void* ws, size_t ws_bytes, int gram_mode, cudaStream_t st) {
  const int k = npc + 10;
  MMB_REQUIRE(k <= 32, "npc must be <= 22");
  SifWs L = sif_ws_layout(N, d, npc);
  MMB_REQUIRE(ws && ws_bytes >= L.total, "workspace too small");
  MMB_REQUIRE(Omega, "Omega (seeded start block) is required when npc > 0");
  char* base = (char*)ws;
  float* G = G_out ? G_out : (float*)(base + L.G);
  float* pc = pc_out ? pc_out : (float*)(base + L.pcv);
  int rc = mmb_gram(emb, N, d, G, base + L.gram, L.pc - L.gram, gram_mode, st);
  if (rc) return rc;
  const double* S0 = Omega;
  const int transposed = N < d;
  if (transposed) {
    rc = mmb_start_block_xt(emb, N, d, Omega, k, (double*)(base + L.s0), st);
    if (rc) return rc;
    S0 = (const double*)(base + L.s0);
  }
  rc = mmb_pc_from_gram(G, d, S0, k, npc, transposed, 7, pc, base + L.pc, L.s0 - L.pc, st);
  if (rc) return rc;
  return mmb_remove_pc(emb, N, d, pc, npc, emb, st);
}

extern "C" int mmb_sif_embedding(const float* table, int64_t V, int d, const float* vocab_w,
                                 const int64_t* x, int64_t N, int64_t L, int npc, const double* Omega,
                                 float* emb, float* pc, float* G, void* ws, size_t ws_bytes,
                                 int gram_mode, int* status, mmb_stream_t stream) {
  // scratch for the pre-scaled table (large batches) sits behind the Gram / solve workspace when the caller
  // provided mmb_sif_workspace_bytes(N, d, npc) + mmb_sif_embed_workspace_bytes(V, d, N, L) bytes
  const size_t base_bytes = align_up(mmb_sif_workspace_bytes(N > 0 ? N : 1, d, npc));
  const size_t scaled_bytes = mmb_sif_embed_workspace_bytes(V, d, N, L);
  int rc;
  if (scaled_bytes && ws && ws_bytes >= base_bytes + scaled_bytes)
    rc = mmb_sif_embed_ws(table, V, d, vocab_w, x, N, L, emb, status, (char*)ws + base_bytes, scaled_bytes, stream);
  else
    rc = mmb_sif_embed(table, V, d, vocab_w, x, N, L, emb, status, stream);
  if (rc || npc <= 0 || N == 0) return rc;
  return pc_removal_device(emb, N, d, npc, Omega, pc, G, ws, ws_bytes, gram_mode, as_stream(stream));
}

// ---- the same call with host buffers --------------------------------------------------------
namespace {
// Staging memory of the *_host entry points comes from a PRIVATE stream-ordered pool per device whose
// release threshold keeps freed blocks for the next call (repeated calls do not pay cudaMalloc again).
// The device's default pool -- and with it every other user of cudaMallocAsync in the process -- is left
// alone; mmb_host_pipeline_trim() hands the cached blocks back to the driver.
cudaMemPool_t g_host_pool[64] = {};

int host_pool(cudaMemPool_t* out) {
  int dev = 0;
  MMB_CUDA(cudaGetDevice(&dev));
  MMB_REQUIRE(dev >= 0 && dev < 64, "device ordinal out of range");
  if (!g_host_pool[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool;
    MMB_CUDA(cudaMemPoolCreate(&pool, &props));
    uint64_t keep = UINT64_MAX;
    MMB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    g_host_pool[dev] = pool;
  }
  *out = g_host_pool[dev];
  return MMB_OK;
}

struct DevBuf {
  void* p = nullptr;
  cudaStream_t st = nullptr;
  int alloc(size_t bytes, cudaStream_t s) {
    st = s;
    cudaMemPool_t pool;
    int rc = host_pool(&pool);
    if (rc) return rc;
    MMB_CUDA(cudaMallocFromPoolAsync(&p, bytes ? bytes : 16, pool, s));
    return MMB_OK;
  }
  ~DevBuf() {
    if (p) cudaFreeAsync(p, st);
  }
};
struct Streams {
  cudaStream_t in = nullptr, comp = nullptr, out = nullptr;
  std::vector<cudaEvent_t> ev;
  int init() {
    MMB_CUDA(cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking));
    MMB_CUDA(cudaStreamCreateWithFlags(&comp, cudaStreamNonBlocking));
    MMB_CUDA(cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking));
    return MMB_OK;
  }
  int event(cudaEvent_t* e) {
    MMB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    ev.push_back(*e);
    return MMB_OK;
  }
  void sync_all() {
    if (in) cudaStreamSynchronize(in);
    if (out) cudaStreamSynchronize(out);
    if (comp) cudaStreamSynchronize(comp);
  }
  ~Streams() {
    for (auto e : ev) cudaEventDestroy(e);
    if (in) cudaStreamDestroy(in);
    if (comp) cudaStreamDestroy(comp);
    if (out) cudaStreamDestroy(out);
  }
};
}  // namespace

// Exchange parameters of the multi-GPU variant (world == 1: no exchange).
struct HostComm {
  int rank = 0, world = 1;
  void* const* bufs = nullptr;
  uint64_t epoch = 0;
  int64_t n_global = 0;
};

namespace {
// Declared AFTER the device buffers, hence destroyed BEFORE them: on every return path -- the early
// error returns included -- copies still in flight on the H2D / D2H streams finish before the buffers
// they use go back to the pool.
struct SyncGuard {
  Streams& S;
  ~SyncGuard() { S.sync_all(); }
};
// Multi-GPU: a rank that fails before it has enqueued the exchange tells its peers (abort flag) instead
// of letting them spin until the 4 s timeout.  Destroyed before SyncGuard, so the abort is synchronised too.
struct AbortGuard {
  const HostComm& hc;
  cudaStream_t st;
  bool armed;
  ~AbortGuard() {
    if (armed) mmb_comm_abort(hc.rank, hc.world, hc.bufs, hc.epoch, (mmb_stream_t)st);
  }
};
}  // namespace

static int sif_embedding_host_impl(const float* table_dev, int64_t V, int d, const float* vocab_w_dev,
                                   const int64_t* x_host, int64_t N, int64_t L, int npc,
                                   const double* Omega_host, void* emb_host, int emb_f64,
                                   float* pc_host, int gram_mode, int64_t chunk_rows, const HostComm& hc) {
  MMB_REQUIRE(table_dev && vocab_w_dev && (x_host || N == 0) && (emb_host || N == 0), "null pointer");
  MMB_REQUIRE(N >= 0 && L >= 0 && d > 0 && d % 4 == 0, "bad size");
  MMB_REQUIRE(npc >= 0 && npc + 10 <= 32, "npc must be in [0, 22]");
  MMB_REQUIRE(npc == 0 || Omega_host, "Omega is required when npc > 0");
  const bool dist = hc.world > 1;
  const int64_t n_global = dist ? hc.n_global : N;
  MMB_REQUIRE(!dist || n_global >= d, "the multi-GPU host path needs N_global >= d (use sif_dist for tiny splits)");
  if (N == 0 && !(dist && npc > 0)) return MMB_OK;
  if (chunk_rows <= 0) chunk_rows = 1 << 18;
  if (chunk_rows > N) chunk_rows = N > 0 ? N : 1;
  const int64_t nchunks = ceil_div(N, chunk_rows);
  const int k = npc + 10;

  Streams S;
  int rc = S.init();
  if (rc) return rc;
  DevBuf ids[2], emb, ws, omega, status, pcv, f64[2], gchunk, scaled;
  SyncGuard sync_guard{S};
  AbortGuard abort_guard{hc, S.comp, dist && npc > 0};
  const bool chunked_gram = npc > 0 && nchunks > 1;
  const size_t ids_chunk = (size_t)chunk_rows * L * sizeof(int64_t);
  if ((rc = ids[0].alloc(ids_chunk, S.comp))) return rc;
  if ((rc = ids[1].alloc(ids_chunk, S.comp))) return rc;
  if ((rc = emb.alloc((size_t)(N > 0 ? N : 1) * d * sizeof(float), S.comp))) return rc;
  if ((rc = status.alloc(sizeof(int), S.comp))) return rc;
  MMB_CUDA(cudaMemsetAsync(status.p, 0, sizeof(int), S.comp));
  const size_t ws_bytes = mmb_sif_workspace_bytes(N > 0 ? N : 1, d, npc);
  if (npc > 0) {
    const int64_t orows = n_global >= d ? d : N;
    if ((rc = ws.alloc(ws_bytes, S.comp))) return rc;
    if ((rc = omega.alloc((size_t)orows * k * sizeof(double), S.comp))) return rc;
    if ((rc = pcv.alloc((size_t)npc * d * sizeof(float), S.comp))) return rc;
    if (chunked_gram && (rc = gchunk.alloc((size_t)d * d * sizeof(float), S.comp))) return rc;
    MMB_CUDA(cudaMemcpyAsync(omega.p, Omega_host, (size_t)orows * k * sizeof(double),
                             cudaMemcpyHostToDevice, S.comp));
  }
  if (emb_f64) {
    const size_t b = (size_t)chunk_rows * d * sizeof(double);
    if ((rc = f64[0].alloc(b, S.comp))) return rc;
    if ((rc = f64[1].alloc(b, S.comp))) return rc;
  }
  // large batches: weights folded into the table once (sif_embed.cu); decided on the WHOLE batch, not per chunk
  const size_t scaled_bytes = mmb_sif_embed_workspace_bytes(V, d, N, L);
  if (scaled_bytes) {
    if ((rc = scaled.alloc(scaled_bytes, S.comp))) return rc;
    if ((rc = sif_prescale(table_dev, V, d, vocab_w_dev, scaled.p, S.comp))) return rc;
  }
  cudaEvent_t allocs_done;
  if ((rc = S.event(&allocs_done))) return rc;
  MMB_CUDA(cudaEventRecord(allocs_done, S.comp));
  MMB_CUDA(cudaStreamWaitEvent(S.in, allocs_done, 0));
  MMB_CUDA(cudaStreamWaitEvent(S.out, allocs_done, 0));

  // phase 1: ids H2D (stream `in`) overlapped with the embed kernel (stream `comp`)
  cudaEvent_t in_ready[2], buf_free[2];
  for (int b = 0; b < 2; ++b) {
    if ((rc = S.event(&in_ready[b]))) return rc;
    if ((rc = S.event(&buf_free[b]))) return rc;
  }
  for (int64_t c = 0; c < nchunks; ++c) {
    const int b = (int)(c & 1);
    const int64_t r0 = c * chunk_rows;
    const int64_t rows = (N - r0) < chunk_rows ? (N - r0) : chunk_rows;
    if (c >= 2) MMB_CUDA(cudaStreamWaitEvent(S.in, buf_free[b], 0));
    MMB_CUDA(cudaMemcpyAsync(ids[b].p, x_host + r0 * L, (size_t)rows * L * sizeof(int64_t),
                             cudaMemcpyHostToDevice, S.in));
    MMB_CUDA(cudaEventRecord(in_ready[b], S.in));
    MMB_CUDA(cudaStreamWaitEvent(S.comp, in_ready[b], 0));
    if (scaled_bytes)
      rc = sif_embed_prescaled(table_dev, V, d, vocab_w_dev, scaled.p, (const int64_t*)ids[b].p, rows, L,
                               (float*)emb.p + r0 * d, (int*)status.p, S.comp);
    else
      rc = mmb_sif_embed(table_dev, V, d, vocab_w_dev, (const int64_t*)ids[b].p, rows, L,
                         (float*)emb.p + r0 * d, (int*)status.p, S.comp);
    if (rc) return rc;
    MMB_CUDA(cudaEventRecord(buf_free[b], S.comp));
    if (chunked_gram) {
      // The compute stream idles while the next chunk's ids cross PCIe: take the chunk's Gram now
      // and add it to the running sum (chunk order -> deterministic), so that only the component
      // solve separates the last H2D from the first D2H.
      SifWs Lw = sif_ws_layout(N, d, npc);
      char* base = (char*)ws.p;
      float* G = (float*)(base + Lw.G);
      rc = mmb_gram((const float*)emb.p + r0 * d, rows, d, (float*)gchunk.p, base + Lw.gram, Lw.pc - Lw.gram,
                    gram_mode, S.comp);
      if (rc) return rc;
      accumulate_kernel<<<(d * d + 255) / 256, 256, 0, S.comp>>>(G, (const float*)gchunk.p, d * d, c == 0);
      MMB_LAUNCH_CHECK("accumulate");
    }
  }
  // phase 2: components from the Gram of the whole block (summed per chunk above, or taken here)
  float* pc_dev = (float*)pcv.p;
  if (npc > 0) {
    SifWs Lw = sif_ws_layout(N > 0 ? N : 1, d, npc);
    char* base = (char*)ws.p;
    float* G = (float*)(base + Lw.G);
    if (!chunked_gram) {
      if (N > 0) {
        rc = mmb_gram((const float*)emb.p, N, d, G, base + Lw.gram, Lw.pc - Lw.gram, gram_mode, S.comp);
        if (rc) return rc;
      } else {
        MMB_CUDA(cudaMemsetAsync(G, 0, (size_t)d * d * sizeof(float), S.comp));
      }
    }
    if (dist) {   // sum of the ranks' Grams over NVLink peer memory, rank order -> identical bits
      rc = mmb_allreduce_peer(G, (int64_t)d * d, 0, hc.rank, hc.world, hc.bufs, hc.epoch, (int*)status.p,
                              (mmb_stream_t)S.comp);
      if (rc) return rc;
      abort_guard.armed = false;   // this rank has taken part in the exchange
    }
    const double* S0 = (const double*)omega.p;
    const int transposed = n_global < d;
    if (transposed) {
      rc = mmb_start_block_xt((const float*)emb.p, N, d, (const double*)omega.p, k,
                              (double*)(base + Lw.s0), S.comp);
      if (rc) return rc;
      S0 = (const double*)(base + Lw.s0);
    }
    rc = mmb_pc_from_gram(G, d, S0, k, npc, transposed, 7, pc_dev, base + Lw.pc, Lw.s0 - Lw.pc, S.comp);
    if (rc) return rc;
  }
  // phase 3: projection (stream `comp`) overlapped with the D2H of finished chunks (`out`)
  cudaEvent_t out_ready[2], out_free[2];
  for (int b = 0; b < 2; ++b) {
    if ((rc = S.event(&out_ready[b]))) return rc;
    if ((rc = S.event(&out_free[b]))) return rc;
  }
  for (int64_t c = 0; c < nchunks; ++c) {
    const int b = (int)(c & 1);
    const int64_t r0 = c * chunk_rows;
    const int64_t rows = (N - r0) < chunk_rows ? (N - r0) : chunk_rows;
    float* blk = (float*)emb.p + r0 * d;
    if (npc > 0) {
      rc = mmb_remove_pc(blk, rows, d, pc_dev, npc, blk, S.comp);
      if (rc) return rc;
    }
    if (emb_f64) {
      if (c >= 2) MMB_CUDA(cudaStreamWaitEvent(S.comp, out_free[b], 0));
      const int64_t n = rows * d;
      f32_to_f64_kernel<<<(int)(ceil_div(n, 1024) < 4096 ? ceil_div(n, 1024) : 4096), 256, 0, S.comp>>>(
          blk, (double*)f64[b].p, n);
      MMB_LAUNCH_CHECK("f32_to_f64");
    }
    MMB_CUDA(cudaEventRecord(out_ready[b], S.comp));
    MMB_CUDA(cudaStreamWaitEvent(S.out, out_ready[b], 0));
    if (emb_f64) {
      MMB_CUDA(cudaMemcpyAsync((double*)emb_host + r0 * d, f64[b].p, (size_t)rows * d * sizeof(double),
                               cudaMemcpyDeviceToHost, S.out));
      MMB_CUDA(cudaEventRecord(out_free[b], S.out));
    } else {
      MMB_CUDA(cudaMemcpyAsync((float*)emb_host + r0 * d, blk, (size_t)rows * d * sizeof(float),
                               cudaMemcpyDeviceToHost, S.out));
    }
  }
  int h_status = 0;
  if (pc_host && npc > 0)
    MMB_CUDA(cudaMemcpyAsync(pc_host, pc_dev, (size_t)npc * d * sizeof(float), cudaMemcpyDeviceToHost, S.comp));
  MMB_CUDA(cudaMemcpyAsync(&h_status, status.p, sizeof(int), cudaMemcpyDeviceToHost, S.comp));
  MMB_CUDA(cudaStreamSynchronize(S.comp));
  MMB_CUDA(cudaStreamSynchronize(S.out));
  MMB_CUDA(cudaStreamSynchronize(S.in));
  if (h_status & MMB_STATUS_COMM_TIMEOUT) {
    set_error("peer all-reduce timed out: a rank never raised its flag");
    return MMB_E_COMM;
  }
  if (h_status & MMB_STATUS_BAD_INDEX) {
    set_error("index out of bounds: a token id is outside [-%lld, %lld)", (long long)V, (long long)V);
    return MMB_E_INDEX;
  }
  return MMB_OK;
}

extern "C" int mmb_host_pipeline_trim(size_t keep_bytes) {
  int dev = 0;
  MMB_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && g_host_pool[dev]) MMB_CUDA(cudaMemPoolTrimTo(g_host_pool[dev], keep_bytes));
  return MMB_OK;
}

extern "C" int mmb_sif_embedding_host(const float* table_dev, int64_t V, int d, const float* vocab_w_dev,
                                      const int64_t* x_host, int64_t N, int64_t L, int npc,
                                      const double* Omega_host, void* emb_host, int emb_f64,
                                      float* pc_host, int gram_mode, int64_t chunk_rows) {
  return sif_embedding_host_impl(table_dev, V, d, vocab_w_dev, x_host, N, L, npc, Omega_host, emb_host, emb_f64,
                                 pc_host, gram_mode, chunk_rows, HostComm());
}

extern "C" int mmb_sif_embedding_host_peer(const float* table_dev, int64_t V, int d, const float* vocab_w_dev,
                                           const int64_t* x_host, int64_t N_local, int64_t L, int npc,
                                           const double* Omega_host, void* emb_host, int emb_f64,
                                           float* pc_host, int gram_mode, int64_t chunk_rows, int64_t N_global,
                                           int rank, int world, void* const* bufs, uint64_t epoch) {
  MMB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
  MMB_REQUIRE(world == 1 || (bufs && epoch > 0), "exchange buffers and epoch > 0 are required");
  HostComm hc;
  hc.rank = rank; hc.world = world; hc.bufs = bufs; hc.epoch = epoch; hc.n_global = N_global;
  return sif_embedding_host_impl(table_dev, V, d, vocab_w_dev, x_host, N_local, L, npc, Omega_host, emb_host,
                                 emb_f64, pc_host, gram_mode, chunk_rows, hc);
}
