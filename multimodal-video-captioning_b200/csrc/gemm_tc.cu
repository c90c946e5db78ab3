// gemm_tc.cu -- bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (sm_100a).
//
//   C[M,N] = sum_k A[m,k] * B[n,k]  (+ beta*C) (+ bias[n])
//
// A [M,K] and B [N,K] are bf16, K-contiguous (the layout every nn.Linear /
// nn.LSTM weight of the reference already has: weight[out,in], and the layout
// of every activation matrix [rows, features]).  This is the tensor-core form
// of the LSTM gate contraction (features_captioning.py:84), the attention
// projections (temporal_attention.py:20-21), the vocabulary projection
// (features_captioning.py:87) and all their backward GEMMs.
//
// Kernel anatomy (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d tiles of A (128x64) and
//               B (BNx64) into a STAGES-deep shared-memory ring, 128B swizzle,
//               completion on mbarriers (expect_tx);
//   warp 1      allocates TMEM, then one lane issues tcgen05.mma
//               (cta_group::1, kind::f16, M=128, N=BN, K=16) four times per
//               stage, accumulating in TMEM; tcgen05.commit releases the stage
//               back to the producer and finally signals the epilogue;
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns at a time, add
//               bias / beta*C, store fp32 (and optionally bf16) to global.
// Out-of-range rows / K tails are zero-filled by TMA; stores are predicated.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace mvc {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;   // 64 bf16 = 128 bytes = one swizzle-128B row

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("mvc gemm_tc: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 B, 8-row
// swizzle atoms 1024 B apart (SBO), LBO unused (=1), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                             // version = 1
  d |= (uint64_t)2 << 61;                             // layout type = SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16: D=f32, A=B=bf16, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int M, int N,
                    int K, float beta, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ Cb, int64_t ldcb) {
  using S = TcSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + S::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BAR_OFF + 8 * (2 * STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int nkb = (K + TC_BK - 1) / TC_BK;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_expect_tx(full_bar(stage), S::STAGE_BYTES);
        const uint32_t a_dst = smem_base + stage * S::STAGE_BYTES;
        tma_load_2d(a_dst, &map_a, full_bar(stage), kb * TC_BK, m0);
        tma_load_2d(a_dst + S::A_BYTES, &map_b, full_bar(stage), kb * TC_BK, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * S::STAGE_BYTES;
        const uint64_t adesc = make_sw128_desc(a_addr);
        const uint64_t bdesc = make_sw128_desc(a_addr + S::A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));            // frees the smem stage once these MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);                 // accumulator complete
    }
    __syncwarp();
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int64_t gm = (int64_t)m0 + row;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      if (gm < M) {
        const int gn0 = n0 + c * 32;
        float* crow = C ? C + gm * ldc : nullptr;
        __nv_bfloat16* brow = Cb ? Cb + gm * ldcb : nullptr;
        const bool vec_ok = crow && (gn0 + 31 < N) && ((reinterpret_cast<uintptr_t>(crow + gn0) & 15u) == 0);
        if (vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = __uint_as_float(v[j]); o.y = __uint_as_float(v[j + 1]);
            o.z = __uint_as_float(v[j + 2]); o.w = __uint_as_float(v[j + 3]);
            if (bias) {
              o.x += bias[gn0 + j]; o.y += bias[gn0 + j + 1]; o.z += bias[gn0 + j + 2]; o.w += bias[gn0 + j + 3];
            }
            float4* dst = reinterpret_cast<float4*>(crow + gn0 + j);
            if (beta != 0.f) {
              const float4 old = *dst;
              o.x += beta * old.x; o.y += beta * old.y; o.z += beta * old.z; o.w += beta * old.w;
            }
            *dst = o;
            if (brow) {
              brow[gn0 + j] = __float2bfloat16(o.x); brow[gn0 + j + 1] = __float2bfloat16(o.y);
              brow[gn0 + j + 2] = __float2bfloat16(o.z); brow[gn0 + j + 3] = __float2bfloat16(o.w);
            }
          }
        } else {
#pragma unroll 1
          for (int j = 0; j < 32; ++j) {
            const int gn = gn0 + j;
            if (gn >= N) break;
            float o = __uint_as_float(v[j]);
            if (bias) o += bias[gn];
            if (crow) {
              if (beta != 0.f) o += beta * crow[gn];
              crow[gn] = o;
            }
            if (brow) brow[gn] = __float2bfloat16(o);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  int64_t rows, cols, ld;
  int box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (size_t)k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= (size_t)k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (size_t)k.ld * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    h ^= (size_t)k.box_rows + (h << 6) + (h >> 2);
    return h;
  }
};

// [rows, cols] bf16 matrix, row pitch ld elements, box = box_rows x 64, 128B swizzle.
static int get_tensor_map(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, rows, cols, ld, box_rows};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  MVC_CHECK(enc, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVC_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%lld cols=%lld ld=%lld", (int)r, ptr,
            (long long)rows, (long long)cols, (long long)ld);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

template <int BN, int STAGES>
static int launch_tc(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, float beta, float* C,
                     int64_t ldc, const float* bias, void* Cb, int64_t ldcb, cudaStream_t st) {
  CUtensorMap ma, mb;
  MVC_TRY(get_tensor_map(A, M, K, lda, TC_BM, &ma));
  MVC_TRY(get_tensor_map(B, N, K, ldb, BN, &mb));
  auto kern = gemm_bf16_tc_kernel<BN, STAGES>;
  constexpr int smem = TcSmem<BN, STAGES>::TOTAL;
  static bool configured = false;
  if (!configured) {
    MVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((unsigned)cdiv(N, BN), (unsigned)cdiv(M, TC_BM));
  ProfScope prof(PK_GEMM_TC, M, N, K, st);
  kern<<<grid, 192, smem, st>>>(ma, mb, M, N, K, beta, C, ldc, bias, (__nv_bfloat16*)Cb, ldcb);
  MVC_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvc

using namespace mvc;

extern "C" int mvc_gemm_bf16(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, float beta,
                             float* C, int64_t ldc, const float* bias, void* Cb, int64_t ldcb, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  MVC_CHECK(A && B && (C || Cb), "mvc_gemm_bf16: null operand");
  MVC_CHECK(K > 0, "mvc_gemm_bf16: K must be positive");
  MVC_CHECK(lda % 8 == 0 && ldb % 8 == 0, "mvc_gemm_bf16: lda (%lld) / ldb (%lld) must be multiples of 8",
            (long long)lda, (long long)ldb);
  MVC_CHECK((reinterpret_cast<uintptr_t>(A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(B) & 15u) == 0,
            "mvc_gemm_bf16: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t mt = cdiv(M, TC_BM);
  if (mt * cdiv(N, 128) >= kNumSMs) return launch_tc<128, 4>(M, N, K, A, lda, B, ldb, beta, C, ldc, bias, Cb, ldcb, st);
  if (mt * cdiv(N, 64) >= kNumSMs) return launch_tc<64, 6>(M, N, K, A, lda, B, ldb, beta, C, ldc, bias, Cb, ldcb, st);
  return launch_tc<32, 8>(M, N, K, A, lda, B, ldb, beta, C, ldc, bias, Cb, ldcb, st);
}
