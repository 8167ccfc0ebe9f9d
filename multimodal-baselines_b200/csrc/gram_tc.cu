// G = X^T X on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the 3xTF32 split
// (x = hi + lo, G ~= hi^T hi + hi^T lo + lo^T hi), FP32 accumulation in TMEM.
// Reference: sif_functions.py:58-67 (compute_pc); the Gram is all that sklearn's randomized
// SVD needs of X.  Specialised for d = 300.
//
// Data flow per CTA (persistent over a contiguous range of K = utterance rows):
//   TMA      : X rows [k0, k0+16) arrive as ten 32-column boxes (cols 300..319 zero-filled by
//              the TMA unit) in the canonical MN-major SWIZZLE_128B operand layout -- X is
//              row-major (N, 300), i.e. "MN-major" for both operands of X^T X, so no
//              transpose is ever needed: A = B = the same shared-memory tile.
//   split    : 4 warps turn the FP32 tile into hi = x & 0xffffe000 (in place) and
//              lo = x - hi (second buffer); element-wise, so swizzle-agnostic.
//   MMA      : one elected thread issues tcgen05.mma (M = 128) for the upper-triangular blocks
//                 tile A  rows   0..127 x cols   0..303   (N = 256 + 48)   TMEM cols   0..303
//                 tile B  rows 128..255 x cols 128..303   (N = 176)        TMEM cols 304..479
//              with three passes (lo*hi, hi*lo, hi*hi) per 8-row K step.
//              The remaining 44x44 block (rows/cols 256..299) does not fit in the 512 TMEM
//              columns next to A and B, so a few CTAs ("role C") run the same pipeline on
//              the four boxes covering cols 192..319 only (A rows 192..319, N = 48).
//   epilogue : TMEM -> registers (tcgen05.ld) -> per-CTA partial tile in global memory;
//              a second kernel sums the partials in CTA order (deterministic) and mirrors
//              the upper triangle.
#include <cuda.h>

#include "common.cuh"

namespace mmb {

namespace tc {

constexpr int kD = 300;
constexpr int kBK = 16;                         // X rows per pipeline stage
constexpr int kStages = 4;
constexpr int kBoxes = 10;                      // 32-column boxes (320 >= 300)
constexpr int kBoxBytes = kBK * 128;            // one box: kBK rows x 128 B
constexpr int kHiBytes = kBoxes * kBoxBytes;    // 20480
constexpr int kStageBytes = 2 * kHiBytes;       // hi + lo
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kThreads = 192;                   // warp0 TMA, warp1 MMA, warps 2-5 split/epilogue
constexpr int kColsAB = 480, kColsC = 48;
constexpr int kPartialStride = 128 * kColsAB;   // floats per CTA partial (role C uses a prefix)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, MN-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, [61,64) layout = 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6), a/b format
// TF32 = 2 at [7,10)/[10,13), a/b major MN = 1 at bit 15/16, N>>3 at [17,23), M>>4 at [24,29).
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

struct Params {
  int64_t N;
  int n_ab;           // CTAs [0, n_ab) run role AB, the rest role C
  int n_c;
  int passes;         // 3 = 3xTF32, 1 = plain TF32 (debug)
  uint32_t lbo, sbo;  // descriptor strides in bytes (runtime so a bring-up sweep can vary them)
  float* partial;
};

__global__ void __launch_bounds__(kThreads, 1)
    gram_tc_kernel(const __grid_constant__ CUtensorMap tmap, const Params prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * kStages + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_ready = bar_full + 8 * kStages,
                 bar_empty = bar_ready + 8 * kStages, bar_done = bar_empty + 8 * kStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const bool role_c = (int)blockIdx.x >= prm.n_ab;
  const int role_rank = role_c ? blockIdx.x - prm.n_ab : blockIdx.x;
  const int role_size = role_c ? prm.n_c : prm.n_ab;
  const int64_t kblocks = (prm.N + kBK - 1) / kBK;
  const int64_t per = (kblocks + role_size - 1) / role_size;
  const int64_t kb0 = role_rank * per;
  int64_t kb1 = kb0 + per;
  if (kb1 > kblocks) kb1 = kblocks;
  const int64_t nkb = kb1 > kb0 ? kb1 - kb0 : 0;
  const int box0 = role_c ? 6 : 0, nbox = role_c ? 4 : kBoxes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_ready + 8 * s, 4);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int64_t i = 0; i < nkb; ++i) {
        const int s = (int)(i % kStages);
        const uint32_t ph = (uint32_t)((i / kStages) & 1);
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        mbar_expect_tx(bar_full + 8 * s, (uint32_t)(nbox * kBoxBytes));
        const uint32_t dst = smem_base + s * kStageBytes;
        const int row = (int)((kb0 + i) * kBK);
        for (int b = 0; b < nbox; ++b)
          tma_load_2d(dst + (box0 + b) * kBoxBytes, &tmap, (box0 + b) * 32, row, bar_full + 8 * s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc256 = make_idesc(128, 256), idesc48 = make_idesc(128, 48),
                         idesc176 = make_idesc(128, 176);
      const uint32_t blk = prm.lbo;  // byte distance between 32-column blocks
      for (int64_t i = 0; i < nkb; ++i) {
        const int s = (int)(i % kStages);
        const uint32_t ph = (uint32_t)((i / kStages) & 1);
        mbar_wait(bar_ready + 8 * s, ph);
        tc_fence_after();
        const uint32_t hi = smem_base + s * kStageBytes, lo = hi + kHiBytes;
        for (int ks = 0; ks < kBK / 8; ++ks) {
          const uint32_t koff = ks * 1024;  // one 8-row K group = 8 x 128 B
          for (int p = 0; p < prm.passes; ++p) {
            // passes: small terms first; the last pass is hi*hi
            const uint32_t abase = (prm.passes == 3 && p == 0) ? lo : hi;
            const uint32_t bbase = (prm.passes == 3 && p == 1) ? lo : hi;
            const uint32_t acc = (i > 0 || ks > 0 || p > 0) ? 1u : 0u;
            if (!role_c) {
              const uint64_t a0 = make_desc(abase + koff, prm.lbo, prm.sbo);
              const uint64_t a4 = make_desc(abase + 4 * blk + koff, prm.lbo, prm.sbo);
              const uint64_t b0 = make_desc(bbase + koff, prm.lbo, prm.sbo);
              const uint64_t b8 = make_desc(bbase + 8 * blk + koff, prm.lbo, prm.sbo);
              const uint64_t b4 = make_desc(bbase + 4 * blk + koff, prm.lbo, prm.sbo);
              umma_tf32(tmem + 0, a0, b0, idesc256, acc);     // rows 0..127   x cols 0..255
              umma_tf32(tmem + 256, a0, b8, idesc48, acc);    // rows 0..127   x cols 256..303
              umma_tf32(tmem + 304, a4, b4, idesc176, acc);   // rows 128..255 x cols 128..303
            } else {
              const uint64_t a6 = make_desc(abase + 6 * blk + koff, prm.lbo, prm.sbo);
              const uint64_t b8 = make_desc(bbase + 8 * blk + koff, prm.lbo, prm.sbo);
              umma_tf32(tmem + 0, a6, b8, idesc48, acc);      // rows 192..319 x cols 256..303
            }
          }
        }
        umma_commit(bar_empty + 8 * s);   // frees the stage when these MMAs have read it
      }
      umma_commit(bar_done);              // accumulators complete
    }
  } else {
    // ===== split (hi/lo) workers, then epilogue =====
    const int t = threadIdx.x - 64;  // 0..127
    for (int64_t i = 0; i < nkb; ++i) {
      const int s = (int)(i % kStages);
      const uint32_t ph = (uint32_t)((i / kStages) & 1);
      mbar_wait(bar_full + 8 * s, ph);
      float4* hi = (float4*)(smem + s * kStageBytes + box0 * kBoxBytes);
      float4* lo = (float4*)(smem + s * kStageBytes + kHiBytes + box0 * kBoxBytes);
      const int n16 = nbox * kBoxBytes / 16;
      for (int e = t; e < n16; e += 128) {
        const float4 x = hi[e];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
        h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
        h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
        h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
        l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
        hi[e] = h;
        lo[e] = l;
      }
      fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready + 8 * s);
    }
    // epilogue: this warp may touch TMEM lanes [32*(warp%4), +32)
    if (nkb > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
    }
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* out = prm.partial + (size_t)blockIdx.x * kPartialStride + (size_t)row * (role_c ? kColsC : kColsAB);
    const int ncols = role_c ? kColsC : kColsAB;
    if (nkb > 0) {
      for (int c = 0; c < ncols; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c, r);
        const int w = (ncols - c) < 32 ? (ncols - c) : 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (j < w)
            *(float4*)(out + c + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                  __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    } else {
      for (int c = 0; c < ncols; c += 4) *(float4*)(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// G[i][j] = G[j][i] = sum over the CTAs of the owning role, in CTA order.
__global__ void __launch_bounds__(256)
    gram_tc_reduce_kernel(const float* __restrict__ partial, int n_ab, int n_c, float* __restrict__ G) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kD * kD) return;
  const int i = idx / kD, j = idx % kD;
  if (j < i) return;
  float s = 0.f;
  if (i < 128) {
    const float* p = partial + (size_t)i * kColsAB + j;
    for (int c = 0; c < n_ab; ++c) s += p[(size_t)c * kPartialStride];
  } else if (i < 256) {
    const float* p = partial + (size_t)(i - 128) * kColsAB + 304 + (j - 128);
    for (int c = 0; c < n_ab; ++c) s += p[(size_t)c * kPartialStride];
  } else {
    const float* p = partial + (size_t)n_ab * kPartialStride + (size_t)(i - 192) * kColsC + (j - 256);
    for (int c = 0; c < n_c; ++c) s += p[(size_t)c * kPartialStride];
  }
  G[(size_t)i * kD + j] = s;
  G[(size_t)j * kD + i] = s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static void plan(int64_t N, int* n_ab, int* n_c) {
  const int sms = sm_count();
  const int64_t kblocks = (N + kBK - 1) / kBK;
  int c = (int)(sms * 0.14 + 0.5);   // role C costs ~1/7 of role AB per K block (smem-bound estimate)
  if (c < 1) c = 1;
  int ab = sms - c;
  if (ab < 1) ab = 1;
  if (kblocks < ab) ab = (int)kblocks;
  if (kblocks < c) c = (int)kblocks;
  *n_ab = ab;
  *n_c = c;
}

}  // namespace tc

bool gram_tc_supported(int64_t N, int d) {
  if (d != tc::kD || N < 1) return false;
  static thread_local int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cached = (major == 10 && tc::encode_fn() != nullptr) ? 1 : 0;
  }
  return cached == 1;
}

size_t gram_tc_workspace_bytes(int64_t N, int d) {
  (void)d;
  int ab, c;
  tc::plan(N, &ab, &c);
  return (size_t)(ab + c) * tc::kPartialStride * sizeof(float);
}

int gram_tc(const float* X, int64_t N, int d, float* G, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace tc;
  MMB_REQUIRE(d == kD, "tcgen05 Gram is specialised for d == 300");
  MMB_REQUIRE(N < ((int64_t)1 << 31) - 64, "N too large for 32-bit TMA coordinates");
  int n_ab, n_c;
  plan(N, &n_ab, &n_c);
  MMB_REQUIRE(ws_bytes >= gram_tc_workspace_bytes(N, d), "workspace too small");
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("gram_tc: cuTensorMapEncodeTiled is not available");
    return MMB_E_UNSUPPORTED;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)kD, (cuuint64_t)N};
  const cuuint64_t gstride[1] = {(cuuint64_t)kD * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)kBK};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)X, gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gram_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MMB_E_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    MMB_CUDA(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  Params prm;
  prm.N = N;
  prm.n_ab = n_ab;
  prm.n_c = n_c;
  prm.passes = 3;
  prm.lbo = kBoxBytes;  // 32-column blocks are kBK*128 bytes apart
  prm.sbo = 1024;       // 8-row K groups are 1024 bytes apart
  if (const char* e = getenv("MMB_TC_PASSES")) prm.passes = atoi(e) == 1 ? 1 : 3;
  if (const char* e = getenv("MMB_TC_LBO")) prm.lbo = (uint32_t)atoi(e);
  if (const char* e = getenv("MMB_TC_SBO")) prm.sbo = (uint32_t)atoi(e);
  prm.partial = (float*)ws;
  gram_tc_kernel<<<n_ab + n_c, kThreads, kSmemBytes, st>>>(tmap, prm);
  MMB_LAUNCH_CHECK("gram_tc");
  gram_tc_reduce_kernel<<<(kD * kD + 255) / 256, 256, 0, st>>>((const float*)ws, n_ab, n_c, G);
  MMB_LAUNCH_CHECK("gram_tc_reduce");
  return MMB_OK;
}

}  // namespace mmb
