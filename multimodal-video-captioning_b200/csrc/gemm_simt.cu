// gemm_simt.cu -- fp32 FFMA GEMM with arbitrary element strides.
//
// This is the arithmetic of the exact-greedy-ids path: every nn.Linear /
// nn.LSTM contraction of the reference (temporal_attention.py:20-23,
// features_captioning.py:84,87, reconstructor.py:73,155) evaluated in fp32
// with a fixed k-ascending reduction order per output element (no split-K, no
// atomics), so results are run-to-run deterministic.
//
// Tiling: BMxBN output tile per 256-thread CTA, BK = 16, register
// double-buffering of the global loads, each thread owns RMxRN groups of 4x4
// outputs placed BM/RM (BN/RN) apart so shared-memory reads are 16-byte and
// conflict-free.
#include "common.cuh"

namespace mvc {

constexpr int GEMM_BK = 16;

template <int BM, int BN, int RM, int RN>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int64_t a_rs, int64_t a_cs,
                const float* __restrict__ Bm, int64_t b_rs, int64_t b_cs, float beta, float* __restrict__ C,
                int64_t ldc, const float* __restrict__ bias) {
  constexpr int BK = GEMM_BK;
  constexpr int TM = 4 * RM, TN = 4 * RN;
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads per CTA");
  constexpr int LA = BM * BK / 256, LB = BN * BK / 256;
  constexpr int PADM = BM + 4, PADN = BN + 4;

  __shared__ __align__(16) float As[BK][PADM];
  __shared__ __align__(16) float Bs[BK][PADN];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const bool a_kfast = (a_cs == 1), b_kfast = (b_cs == 1);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[LA], rb[LB];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const int idx = tid + i * 256;
      const int r = a_kfast ? idx / BK : idx % BM;
      const int kk = a_kfast ? idx % BK : idx / BM;
      const int gm = m0 + r, gk = k0 + kk;
      ra[i] = (gm < M && gk < K) ? A[gm * a_rs + gk * a_cs] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const int idx = tid + i * 256;
      const int r = b_kfast ? idx / BK : idx % BN;
      const int kk = b_kfast ? idx % BK : idx / BN;
      const int gn = n0 + r, gk = k0 + kk;
      rb[i] = (gn < N && gk < K) ? Bm[gn * b_rs + gk * b_cs] : 0.f;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const int idx = tid + i * 256;
      const int r = a_kfast ? idx / BK : idx % BM;
      const int kk = a_kfast ? idx % BK : idx / BM;
      As[kk][r] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const int idx = tid + i * 256;
      const int r = b_kfast ? idx / BK : idx % BN;
      const int kk = b_kfast ? idx % BK : idx / BN;
      Bs[kk][r] = rb[i];
    }
  };

  load_tiles(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    store_tiles();
    __syncthreads();
    if (k0 + BK < K) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < RM; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&As[kk][g * (BM / RM) + ty * 4]);
        a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < RN; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[kk][g * (BN / RN) + tx * 4]);
        b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int gi = 0; gi < RM; ++gi)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gm = m0 + gi * (BM / RM) + ty * 4 + i;
      if (gm >= M) continue;
#pragma unroll
      for (int gj = 0; gj < RN; ++gj)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int gn = n0 + gj * (BN / RN) + tx * 4 + j;
          if (gn >= N) continue;
          float v = alpha * acc[gi * 4 + i][gj * 4 + j];
          if (bias) v += bias[gn];
          float* c = C + gm * ldc + gn;
          if (beta != 0.f) v += beta * (*c);
          *c = v;
        }
    }
}

}  // namespace mvc

extern "C" int mvc_gemm_f32(int M, int N, int K, float alpha, const float* A, int64_t a_rs, int64_t a_cs,
                            const float* B, int64_t b_rs, int64_t b_cs, float beta, float* C, int64_t ldc,
                            const float* bias, void* stream) {
  using namespace mvc;
  if (M <= 0 || N <= 0) return 0;
  MVC_CHECK(A && B && C, "mvc_gemm_f32: null operand");
  MVC_CHECK(K >= 0, "mvc_gemm_f32: negative K");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(PK_GEMM_F32, M, N, K, st);
  const int64_t big_tiles = cdiv(M, 128) * cdiv(N, 128);
  if (big_tiles >= 2 * kNumSMs) {
    dim3 grid((unsigned)cdiv(N, 128), (unsigned)cdiv(M, 128));
    gemm_f32_kernel<128, 128, 2, 2><<<grid, 256, 0, st>>>(M, N, K, alpha, A, a_rs, a_cs, B, b_rs, b_cs, beta, C, ldc, bias);
  } else {
    dim3 grid((unsigned)cdiv(N, 64), (unsigned)cdiv(M, 64));
    gemm_f32_kernel<64, 64, 1, 1><<<grid, 256, 0, st>>>(M, N, K, alpha, A, a_rs, a_cs, B, b_rs, b_cs, beta, C, ldc, bias);
  }
  MVC_LAUNCH_CHECK();
  return 0;
}
