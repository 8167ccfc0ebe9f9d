// SIF weighted average for very large batches: the most frequent rows on the tensor cores.
// Reference: sif_functions.py:28-56 (get_weighted_average) with seq2weight (8-15) folded in; the
// semantics are those of sif_embed.cu (divisor counts every non-zero weight, id 0 is an ordinary row,
// negative ids wrap with weight 0, ids outside [-V, V) raise the status bit).
//
// Why: the gather kernel is bound by the L1 load path and the L2 -> SM fabric, and under a Zipf
// vocabulary half of all row reads go to a few dozen rows (the pad row and the most frequent words:
// the top 64 rows carry 51 % of the bench workload's non-pad tokens and all of its padding).  Their
// contribution to a tile of 128 utterances is a small dense product,
//      hot[d][u] = sum_k T'[hot_k][d] * count[k][u]        (T' = weight-scaled table, sif_embed.cu)
// with integer counts -- exact in TF32 -- so it runs on tcgen05 (kind::tf32, T' split hi + lo: two passes,
// 2^-22 relative per term) out of shared memory, while the warps gather only the remaining (cold) rows:
//      setup   : a histogram of the first 4096 utterances' ids picks the 64 most frequent rows (device
//                side, deterministic); every CTA copies those rows of T' into shared memory in the
//                canonical MN-major UMMA layout (SWIZZLE_128B, 32-byte atoms -- the layout of gram_tc.cu)
//                and builds a 1024-slot hash of their ids
//      gather  : warp per utterance; a token whose row is in the hash adds 1 to count[slot][utterance]
//                (one shared-memory word, owned by that warp), every other token is gathered as in
//                sif_embed_prescaled_kernel; the undivided cold sum goes to emb
//      MMA     : one thread issues 48 tcgen05.mma (M = 128 rows of d, N = 128 utterances, K = 8 hot rows
//                each; three M tiles cover d = 0..319) into 384 TMEM columns
//      finish  : TMEM lane = d, column = utterance: each warp reads 32 x 32 blocks of its lane quadrant and
//                completes emb = (cold + hot) / count with fully coalesced 128-byte accesses
// Deterministic (fixed orders everywhere); selected by mmb_sif_embed_ws for d = 300 and N * L >= 64 V.
#include "common.cuh"

namespace mmb {

namespace hot {

constexpr int kK = 64;                         // hot rows
constexpr int kTileU = 128;                    // utterances per tile (MMA N)
constexpr int kD = 300, kD4 = 75;
constexpr int kBoxes = 10;                     // 32-column boxes of the hot table (320 >= 300)
constexpr int kBoxBytes = kK * 128;            // one box: kK rows x 128 B
constexpr int kABytes = kBoxes * kBoxBytes;    // 81920 per split
constexpr int kCBytes = (kTileU / 32) * kBoxBytes;   // 32768: count tile, kK rows x 128 utterances
constexpr int kHashSize = 1024;
constexpr int kWarps = 32;
constexpr int kThreads = kWarps * 32;
constexpr int kSmemBytes = 2 * kABytes + kCBytes + kHashSize * 4 + kTileU * 4 + 64 + 1024 /*align*/;
constexpr unsigned kEmpty = 0xffffffffu;
constexpr int kSampleUtt = 4096;      // x L tokens: enough to rank the head of the distribution, cheap in contended atomics

__device__ __forceinline__ unsigned hashf(unsigned row) { return (row * 2654435761u) >> 22; }   // 10 bits

// ---- setup: which rows are hot ------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    hot_hist_kernel(const int64_t* __restrict__ ids, int64_t n_tokens, int V, int* __restrict__ counters) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tokens; i += stride) {
    const int64_t id = __ldg(ids + i);
    if (id >= 0 && id < V) atomicAdd(counters + id, 1);
  }
}

// The kK most frequent rows of the sample (ties -> lowest id); one CTA.  Only rows seen at least once per 4096
// sampled tokens are candidates (at most 4096 of them; a rarer row gains nothing from the tensor path), so the
// kK arg-max rounds run over a short shared-memory list instead of the V counters.
constexpr int kMaxCand = 4096;
__global__ void __launch_bounds__(1024)
    hot_select_kernel(const int* __restrict__ counters, int V, int64_t n_sample, int* __restrict__ hot_ids,
                      int n_select = kK) {
  __shared__ long long cand[kMaxCand];          // (count << 32) | (0x7fffffff - id): max = highest count, lowest id
  __shared__ long long best_s[32];
  __shared__ int where_s[32];
  __shared__ int n_cand;
  if (threadIdx.x == 0) n_cand = 0;
  __syncthreads();
  const int thr = (int)(n_sample / kMaxCand) + 1;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const int c = __ldg(counters + v);
    if (c >= thr) {
      const int slot = atomicAdd(&n_cand, 1);   // count >= n_sample / 4096 + 1 for at most 4095 rows
      if (slot < kMaxCand) cand[slot] = ((long long)c << 32) | (unsigned)(0x7fffffff - v);
    }
  }
  __syncthreads();
  const int n = n_cand < kMaxCand ? n_cand : kMaxCand;
  for (int r = 0; r < n_select; ++r) {
    long long best = -1;
    int where = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (cand[i] > best) { best = cand[i]; where = i; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const long long nb = __shfl_xor_sync(0xffffffffu, best, o);
      const int nw = __shfl_xor_sync(0xffffffffu, where, o);
      if (nb > best) { best = nb; where = nw; }
    }
    if ((threadIdx.x & 31) == 0) {
      best_s[threadIdx.x >> 5] = best;
      where_s[threadIdx.x >> 5] = where;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      long long b = best_s[threadIdx.x];
      int w = where_s[threadIdx.x];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const long long nb = __shfl_xor_sync(0xffffffffu, b, o);
        const int nw = __shfl_xor_sync(0xffffffffu, w, o);
        if (nb > b) { b = nb; w = nw; }
      }
      if (threadIdx.x == 0) {
        int id = -1;
        if (b >= 0) {
          id = 0x7fffffff - (int)(unsigned)(b & 0xffffffffll);
          cand[w] = -1;                          // taken
        }
        hot_ids[r] = id;
      }
    }
    __syncthreads();
  }
}

// ---- tcgen05 plumbing (same conventions as gram_tc.cu) ------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// MN-major shared-memory descriptor, SWIZZLE_128B with 32-byte atoms (UMMA layout type 1), version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_at(uint64_t base, uint32_t saddr) {
  return base | (uint64_t)((saddr >> 4) & 0x3fff);
}
// c F32, a/b TF32, both MN-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// byte offset of element (k, col) of a kK-row MN-major tile: 32-column boxes kBoxBytes apart, a row of a box is
// 128 B whose 32-byte chunks are XOR-swizzled with (k & 3)
__device__ __forceinline__ uint32_t tile_off(int k, int col) {
  return (uint32_t)(col >> 5) * kBoxBytes + (uint32_t)k * 128u + ((((uint32_t)(col & 31) >> 3) ^ ((uint32_t)k & 3u)) << 5) +
         ((uint32_t)col & 7u) * 4u;
}

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(f32x2& acc, f32x2 a, f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

constexpr int kRowBits = 26;

__global__ void __launch_bounds__(kThreads, 1)
    sif_embed_hot_kernel(const float4* __restrict__ tp4, int V, const int* __restrict__ flags,
                         const int* __restrict__ hot_ids, const int64_t* __restrict__ ids, int64_t N, int64_t L,
                         float* __restrict__ emb, int* __restrict__ status) {
  if (__ldg(flags) & 1) return;        // a zero vocabulary weight: the general kernel behind this one runs
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + kABytes;
  uint8_t* ctile = smem + 2 * kABytes;
  unsigned* hash = (unsigned*)(ctile + kCBytes);
  int* cnt_s = (int*)(hash + kHashSize);
  uint64_t* bar = (uint64_t*)(cnt_s + kTileU);
  uint32_t* tmem_slot = (uint32_t*)(bar + 1);
  int* work = (int*)(tmem_slot + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_mma = smem_u32(bar);

  // ---- prologue: barrier, TMEM, hash of the hot ids, hot rows of T' as the MMA's A operand (hi / lo) ----
  if (threadIdx.x == 0) {
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kHashSize; i += kThreads) hash[i] = kEmpty;
  __syncthreads();
  if (threadIdx.x < kK) {
    const int id = __ldg(hot_ids + threadIdx.x);
    if (id >= 0) {
      unsigned h = hashf((unsigned)id);
      const unsigned entry = ((unsigned)id << 6) | (unsigned)threadIdx.x;
      while (atomicCAS(hash + h, kEmpty, entry) != kEmpty) h = (h + 1) & (kHashSize - 1);
    }
  }
  for (int i = threadIdx.x; i < kK * 80; i += kThreads) {      // 80 float4 per row: d = 0..319
    const int k = i / 80, c4 = i % 80;
    const int id = __ldg(hot_ids + k);
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (id >= 0 && c4 < kD4) x = __ldg(tp4 + (size_t)id * kD4 + c4);
    float4 l;
    l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);   // the MMA reads the top 19 bits of x itself
    l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
    l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
    l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
    const uint32_t off = tile_off(k, c4 * 4);
    *(float4*)(a_hi + off) = x;
    *(float4*)(a_lo + off) = l;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const char* lane_base = (const char*)(tp4 + lane);
  const bool tail = lane + 64 < kD4;                      // third float4 chunk of the row: lanes 0..10
  const int64_t ntiles = (N + kTileU - 1) / kTileU;
  bool bad = false;
  uint32_t parity = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t u0 = tile * kTileU;
    // (a) clear the count tile
    for (int i = threadIdx.x; i < kCBytes / 16; i += kThreads) ((float4*)ctile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x < kTileU) cnt_s[threadIdx.x] = 1;
    if (threadIdx.x == 0) *work = 0;
    __syncthreads();
    // (b) gather: warps take utterances of the tile off a shared counter
    for (;;) {
      int u = 0;
      if (lane == 0) u = atomicAdd(work, 1);
      u = __shfl_sync(0xffffffffu, u, 0);
      const int64_t i = u0 + u;
      if (u >= kTileU || i >= N) break;
      f32x2 acc[3][2];
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c][0] = acc[c][1] = 0ull;
      int cnt = 0;
      for (int64_t base = 0; base < L; base += 32) {
        int row = -1, slot = -1;
        bool counts = false;
        if (base + lane < L) {
          const int64_t id = __ldcs(ids + i * L + base + lane);
          const int64_t r = id < 0 ? id + V : id;
          if (r >= 0 && r < V) {
            row = (int)r;
            counts = id >= 0;
            if (counts) {                                  // negative ids (weight 0) stay on the gather path
              unsigned h = hashf((unsigned)row);
              for (;;) {
                const unsigned e = hash[h];
                if (e == kEmpty) break;
                if ((e >> 6) == (unsigned)row) { slot = (int)(e & 63u); break; }
                h = (h + 1) & (kHashSize - 1);
              }
            }
          } else {
            bad = true;
          }
        }
        const unsigned nn = __ballot_sync(0xffffffffu, counts);
        cnt += __popc(nn);
        // hot tokens: one shared-memory word per (slot, utterance), owned by this warp
        const unsigned hgrp = __match_any_sync(0xffffffffu, slot);
        if (slot >= 0 && lane == __ffs(hgrp) - 1) {
          float* cp = (float*)(ctile + tile_off(slot, u));
          *cp += (float)__popc(hgrp);
        }
        // cold tokens: as sif_embed_prescaled_kernel
        const int crow = slot >= 0 ? -1 : row;
        const unsigned grp = __match_any_sync(0xffffffffu, crow);
        const bool head = (crow >= 0) && (lane == __ffs(grp) - 1);
        const unsigned packed = (unsigned)crow | ((unsigned)__popc(grp & nn) << kRowBits);
        unsigned heads = __ballot_sync(0xffffffffu, head);
        while (heads) {
          const int j0 = __ffs(heads) - 1;
          heads &= heads - 1;
          const int j1 = heads ? __ffs(heads) - 1 : -1;
          if (heads) heads &= heads - 1;
          const unsigned p0 = __shfl_sync(0xffffffffu, packed, j0);
          const unsigned p1 = __shfl_sync(0xffffffffu, packed, j1 < 0 ? j0 : j1);
          const float4* r0 = (const float4*)(lane_base + (size_t)(p0 & ((1u << kRowBits) - 1u)) * (kD4 * 16));
          const float4* r1 = (const float4*)(lane_base + (size_t)(p1 & ((1u << kRowBits) - 1u)) * (kD4 * 16));
          float4 v0[3], v1[3];
          v0[0] = __ldg(r0); v0[1] = __ldg(r0 + 32);
          if (tail) v0[2] = __ldg(r0 + 64);
          if (j1 >= 0) {
            v1[0] = __ldg(r1); v1[1] = __ldg(r1 + 32);
            if (tail) v1[2] = __ldg(r1 + 64);
          }
          const float m0 = (float)(p0 >> kRowBits), m1 = (float)(p1 >> kRowBits);
          const f32x2 w0 = pack2(m0, m0), w1 = pack2(m1, m1);
          ffma2(acc[0][0], w0, pack2(v0[0].x, v0[0].y)); ffma2(acc[0][1], w0, pack2(v0[0].z, v0[0].w));
          ffma2(acc[1][0], w0, pack2(v0[1].x, v0[1].y)); ffma2(acc[1][1], w0, pack2(v0[1].z, v0[1].w));
          if (tail) { ffma2(acc[2][0], w0, pack2(v0[2].x, v0[2].y)); ffma2(acc[2][1], w0, pack2(v0[2].z, v0[2].w)); }
          if (j1 >= 0) {
            ffma2(acc[0][0], w1, pack2(v1[0].x, v1[0].y)); ffma2(acc[0][1], w1, pack2(v1[0].z, v1[0].w));
            ffma2(acc[1][0], w1, pack2(v1[1].x, v1[1].y)); ffma2(acc[1][1], w1, pack2(v1[1].z, v1[1].w));
            if (tail) { ffma2(acc[2][0], w1, pack2(v1[2].x, v1[2].y)); ffma2(acc[2][1], w1, pack2(v1[2].z, v1[2].w)); }
          }
        }
      }
      // the undivided cold sum; the finish pass adds the hot part and divides
      float4* out4 = (float4*)(emb + (size_t)i * kD) + lane;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c < 2 || tail) {
          float4 r;
          unpack2(acc[c][0], r.x, r.y);
          unpack2(acc[c][1], r.z, r.w);
          out4[32 * c] = r;
        }
      }
      if (lane == 0) cnt_s[u] = cnt;
    }
    // (c) counts written through the generic proxy -> visible to the tensor core
    fence_proxy_async();
    __syncthreads();
    // (d) hot[d][u] = sum_k T'[k][d] count[k][u]: three M tiles (d 0..127, 128..255, 192..319) x N = 128
    if (threadIdx.x == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc(128, kTileU);
      const uint64_t dbase = make_desc(0, kBoxBytes, 512);
      const uint32_t hi = smem_u32(a_hi), lo = smem_u32(a_lo), cb = smem_u32(ctile);
#pragma unroll 1
      for (int ks = 0; ks < kK / 8; ++ks) {
        const uint32_t koff = ks * 1024;                 // 8 K rows = two 4-row swizzle atoms of 512 B
#pragma unroll
        for (int p = 0; p < 2; ++p) {                    // small terms first
          const uint32_t ab = (p == 0 ? lo : hi) + koff;
          const uint64_t b0 = desc_at(dbase, cb + koff);
          const uint32_t accf = (ks == 0 && p == 0) ? 0u : 1u;
          umma_tf32(tmem + 0, desc_at(dbase, ab), b0, idesc, accf);
          umma_tf32(tmem + 128, desc_at(dbase, ab + 4 * kBoxBytes), b0, idesc, accf);
          umma_tf32(tmem + 256, desc_at(dbase, ab + 6 * kBoxBytes), b0, idesc, accf);
        }
      }
      umma_commit(bar_mma);
    }
    // (e) finish: emb = (cold + hot) / count.  TMEM lane = d within the M tile, column = utterance.
    mbar_wait(bar_mma, parity);
    parity ^= 1u;
    tc_fence_after();
    {
      const int q = warp & 3, r8 = warp >> 2;
      const int nd = q >= 2 ? 3 : 2;                     // M tile 2 holds d = 256..319 in lanes 64..127
      for (int b = r8; b < nd * 4; b += 8) {
        const int mt = b >> 2, ub = b & 3;
        const int d = (mt < 2 ? 128 * mt : 192) + 32 * q + lane;
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(128 * mt + 32 * ub), r);
        if (d < kD) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int u = 32 * ub + j;
            const int64_t i = u0 + u;
            if (i < N) {
              float* p = emb + (size_t)i * kD + d;
              __stcs(p, __fdiv_rn(*p + __uint_as_float(r[j]), (float)cnt_s[u]));
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, MMB_STATUS_BAD_INDEX);
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace hot

// sample histogram (V ints) + up to 512 selected row ids
size_t sif_embed_hot_extra_bytes(int64_t V) { return ((size_t)V * sizeof(int) + 255) / 256 * 256 + 2048; }

// The k (<= 512) most frequent rows among the first kSampleUtt utterances' ids -> ids_out (device, k ints, -1 padded).
int sif_select_frequent_rows(const int64_t* x, int64_t N, int64_t L, int64_t V, int k, void* ws_hot, int** ids_out,
                             cudaStream_t st) {
  using namespace hot;
  MMB_REQUIRE(k > 0 && k <= 512, "1..512 rows");
  int* counters = (int*)ws_hot;
  int* sel = (int*)((char*)ws_hot + ((size_t)V * sizeof(int) + 255) / 256 * 256);
  MMB_CUDA(cudaMemsetAsync(counters, 0, (size_t)V * sizeof(int), st));
  const int64_t n_sample = (N < kSampleUtt ? N : kSampleUtt) * L;
  hot_hist_kernel<<<sm_count() * 4, 256, 0, st>>>(x, n_sample, (int)V, counters);
  MMB_LAUNCH_CHECK("hot_hist");
  hot_select_kernel<<<1, 1024, 0, st>>>(counters, (int)V, n_sample, sel, k);
  MMB_LAUNCH_CHECK("hot_select");
  *ids_out = sel;
  return MMB_OK;
}

bool sif_embed_hot_eligible(int64_t V, int d, int64_t N, int64_t L) {
  if (option_embed_hot() == 0) return false;
  return d == hot::kD && V < ((int64_t)1 << hot::kRowBits) && N * L >= 64 * V && N >= 4096 && L >= 1;
}

// ws_hot: V ints (sample histogram) + kK ints (hot ids); tp / flags from sif_prescale.
int sif_embed_hot(const float* tp, const int* flags, int64_t V, const int64_t* x, int64_t N, int64_t L, float* emb,
                  int* status, void* ws_hot, cudaStream_t st) {
  using namespace hot;
  int* hot_ids = nullptr;
  int rc = sif_select_frequent_rows(x, N, L, V, kK, ws_hot, &hot_ids, st);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MMB_CUDA(cudaFuncSetAttribute(sif_embed_hot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int64_t ntiles = ceil_div(N, kTileU);
  const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  sif_embed_hot_kernel<<<grid, kThreads, kSmemBytes, st>>>((const float4*)tp, (int)V, flags, hot_ids, x, N, L, emb,
                                                           status);
  MMB_LAUNCH_CHECK("sif_embed_hot");
  note_kernel(0, "sif_embed_hot_kernel");
  return MMB_OK;
}

}  // namespace mmb
