// attention.cu -- fused additive (Bahdanau) soft attention, forward and backward.
//
// Reference arithmetic: TemporalAttention.forward, temporal_attention.py:19-33,
// with the two projections taken as inputs (wq = W.q, uk = U.k; U.k is loop
// invariant and hoisted by the callers, W.q is a small per-step GEMM).
//
// Forward (one launch per decoder step): grid = (B, FS).  Every CTA recomputes
// the T scores of its batch row (T*A tanh, ~11k), does the masked softmax in
// shared memory with warp-shuffle reductions, then produces its F/FS slice of
// the context sum with 16-byte vector loads of the keys (bf16x8 or fp32x4),
// T-rows split over thread groups and combined through shared memory.  Bound
// by the read of keys [B,T,F] + uk [B,T,A]: HBM on first touch, L2 afterwards
// (MSVD-shaped bf16 working set = 27 MB << 126 MB L2).
#include "common.cuh"

namespace mvc {

template <bool FAST>
__device__ __forceinline__ float tanh_sel(float x) {
  if constexpr (FAST) return tanh_fast(x);
  else return tanhf(x);
}

template <typename KT>
struct VecOf;
template <>
struct VecOf<float> {
  static constexpr int N = 4;
  using Raw = float4;
  __device__ static void unpack(const Raw& r, float* f) { f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w; }
};
template <>
struct VecOf<__nv_bfloat16> {
  static constexpr int N = 8;
  using Raw = uint4;
  __device__ static void unpack(const Raw& r, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
};

// dynamic smem layout (floats): sQ[A] (wq+bias), sW[A], sE[T], sRed[32], sPart[G*chunk]
template <typename KT, bool FAST, bool VEC>
__global__ void __launch_bounds__(256)
soft_attention_fwd_kernel(int T, int A, int F, int chunk, const float* __restrict__ wq, const float* __restrict__ uk,
                          const float* __restrict__ bias, const float* __restrict__ w, const KT* __restrict__ keys,
                          int keys_batch, int64_t k_sb, int64_t k_st, const uint8_t* __restrict__ mask, int64_t m_sb,
                          int64_t m_st, float* __restrict__ ctx_f32, int64_t ctx_ld, __nv_bfloat16* __restrict__ ctx_bf16,
                          int64_t ctxb_ld, float* __restrict__ alpha) {
  extern __shared__ __align__(16) float smem[];
  float* sQ = smem;
  float* sW = sQ + A;
  float* sE = sW + A;
  float* sRed = sE + T;
  float* sPart = sRed + 32;

  const int b = blockIdx.x, kb = b % keys_batch;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;

  for (int a = tid; a < A; a += blockDim.x) {
    sQ[a] = wq[(int64_t)b * A + a] + bias[a];
    sW[a] = w[a];
  }
  __syncthreads();

  // scores: one warp per frame, lanes over the bottleneck dim
  const float* ukb = uk + (int64_t)kb * T * A;
  for (int t = wid; t < T; t += nw) {
    float e = 0.f;
    const float* row = ukb + (int64_t)t * A;
    for (int a = lane; a < A; a += 32) e = fmaf(sW[a], tanh_sel<FAST>(sQ[a] + row[a]), e);
    e = warp_sum(e);
    if (lane == 0) {
      if (mask && !mask[b * m_sb + t * m_st]) e = -INFINITY;
      sE[t] = e;
    }
  }
  __syncthreads();

  // softmax over T
  float mx = -INFINITY;
  for (int t = tid; t < T; t += blockDim.x) mx = fmaxf(mx, sE[t]);
  mx = block_max(mx, sRed);
  float s = 0.f;
  for (int t = tid; t < T; t += blockDim.x) {
    const float p = FAST ? __expf(sE[t] - mx) : expf(sE[t] - mx);
    sE[t] = p;
    s += p;
  }
  s = block_sum(s, sRed);
  const float inv = 1.f / s;
  for (int t = tid; t < T; t += blockDim.x) {
    const float p = sE[t] * inv;
    sE[t] = p;
    if (blockIdx.y == 0) alpha[(int64_t)b * T + t] = p;
  }
  __syncthreads();

  // context slice [f0, f1)
  const int f0 = blockIdx.y * chunk;
  const int f1 = min(F, f0 + chunk);
  if (f0 >= f1) return;
  const KT* kbase = keys + (int64_t)kb * k_sb;
  constexpr int VN = VEC ? VecOf<KT>::N : 1;
  const int nvec = (f1 - f0 + VN - 1) / VN;          // host guarantees nvec <= blockDim
  int G = blockDim.x / nvec;
  if (G > T) G = T;
  if (G < 1) G = 1;
  const int g = tid / nvec, v = tid - g * nvec;
  float acc[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) acc[i] = 0.f;
  if (g < G) {
    const int f = f0 + v * VN;
    for (int t = g; t < T; t += G) {
      const float p = sE[t];
      const KT* src = kbase + (int64_t)t * k_st + f;
      if constexpr (VEC) {
        const typename VecOf<KT>::Raw raw = *reinterpret_cast<const typename VecOf<KT>::Raw*>(src);
        float x[VN];
        VecOf<KT>::unpack(raw, x);
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[i] = fmaf(p, x[i], acc[i]);
      } else {
        acc[0] = fmaf(p, ld_as_float(src), acc[0]);
      }
    }
    if (G > 1) {
#pragma unroll
      for (int i = 0; i < VN; ++i) sPart[(g * nvec + v) * VN + i] = acc[i];
    }
  }
  if (G > 1) {
    __syncthreads();
    if (g == 0) {
      for (int gg = 1; gg < G; ++gg)
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[i] += sPart[(gg * nvec + v) * VN + i];
    }
  }
  if (g == 0 && v < nvec) {
    const int f = f0 + v * VN;
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      if (f + i < f1) {
        if (ctx_f32) ctx_f32[b * ctx_ld + f + i] = acc[i];
        if (ctx_bf16) ctx_bf16[b * ctxb_ld + f + i] = __float2bfloat16(acc[i]);
      }
    }
  }
}

// Backward: one CTA per batch row.
// dynamic smem (floats): sD[F] (dctx), sAl[T], sDa[T] (dalpha -> de), sRed[32]
template <typename KT, bool FAST, bool VEC>
__global__ void __launch_bounds__(512)
soft_attention_bwd_kernel(int T, int A, int F, const float* __restrict__ wq, const float* __restrict__ uk,
                          const float* __restrict__ bias, const float* __restrict__ w, const KT* __restrict__ keys,
                          int64_t k_sb, int64_t k_st, const float* __restrict__ alpha, const float* __restrict__ dctx,
                          int64_t dctx_ld, float* __restrict__ dwq, float* __restrict__ duk,
                          float* __restrict__ dw_partial, float* __restrict__ dkeys, int64_t dk_sb, int64_t dk_st) {
  extern __shared__ __align__(16) float smem[];
  float* sD = smem;
  float* sAl = sD + ((F + 3) & ~3);
  float* sDa = sAl + T;
  float* sRed = sDa + T;

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  for (int f = tid; f < F; f += blockDim.x) sD[f] = dctx[b * dctx_ld + f];
  for (int t = tid; t < T; t += blockDim.x) sAl[t] = alpha[(int64_t)b * T + t];
  __syncthreads();

  // dalpha[t] = dctx . keys[t]   (warp per frame)
  const KT* kbase = keys + (int64_t)b * k_sb;
  constexpr int VN = VEC ? VecOf<KT>::N : 1;
  for (int t = wid; t < T; t += nw) {
    const KT* row = kbase + (int64_t)t * k_st;
    float d = 0.f;
    if constexpr (VEC) {
      for (int f = lane * VN; f < F; f += 32 * VN) {
        const typename VecOf<KT>::Raw raw = *reinterpret_cast<const typename VecOf<KT>::Raw*>(row + f);
        float x[VN];
        VecOf<KT>::unpack(raw, x);
#pragma unroll
        for (int i = 0; i < VN; ++i) d = fmaf(sD[f + i], x[i], d);
      }
    } else {
      for (int f = lane; f < F; f += 32) d = fmaf(sD[f], ld_as_float(row + f), d);
    }
    d = warp_sum(d);
    if (lane == 0) sDa[t] = d;
  }
  __syncthreads();

  // softmax backward: de = alpha * (dalpha - sum_t alpha*dalpha)
  float part = 0.f;
  for (int t = tid; t < T; t += blockDim.x) part += sAl[t] * sDa[t];
  const float dot = block_sum(part, sRed);
  for (int t = tid; t < T; t += blockDim.x) sDa[t] = sAl[t] * (sDa[t] - dot);
  __syncthreads();

  // through w . tanh(wq + uk + bias)
  const float* ukb = uk + (int64_t)b * T * A;
  float* dukb = duk ? duk + (int64_t)b * T * A : nullptr;
  for (int a = tid; a < A; a += blockDim.x) {
    const float q = wq[(int64_t)b * A + a] + bias[a];
    const float wa = w[a];
    float sq = 0.f, sw = 0.f;
    for (int t = 0; t < T; ++t) {
      const float th = tanh_sel<FAST>(q + ukb[(int64_t)t * A + a]);
      const float de = sDa[t];
      const float dpre = de * wa * (1.f - th * th);
      sq += dpre;
      sw = fmaf(de, th, sw);
      if (dukb) dukb[(int64_t)t * A + a] += dpre;
    }
    dwq[(int64_t)b * A + a] = sq;
    if (dw_partial) dw_partial[(int64_t)b * A + a] += sw;
  }

  // optional gradient w.r.t. the keys themselves (local reconstructor: keys = decoder hiddens)
  if (dkeys) {
    float* dkb = dkeys + (int64_t)b * dk_sb;
    for (int64_t i = tid; i < (int64_t)T * F; i += blockDim.x) {
      const int t = (int)(i / F), f = (int)(i - (int64_t)t * F);
      dkb[(int64_t)t * dk_st + f] += sAl[t] * sD[f];
    }
  }
}

template <typename T>
static bool aligned16(const T* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace mvc

using namespace mvc;

extern "C" int mvc_soft_attention_fwd(int B, int T, int A, int F, const float* wq, const float* uk, const float* bias,
                                      const float* w, const void* keys, int keys_bf16, int keys_batch, int64_t k_sb,
                                      int64_t k_st, const uint8_t* mask, int64_t m_sb, int64_t m_st, float* ctx_f32,
                                      int64_t ctx_ld, void* ctx_bf16, int64_t ctxb_ld, float* alpha, int fast_math,
                                      void* stream) {
  if (B == 0) return 0;
  MVC_CHECK(wq && uk && bias && w && keys && alpha, "mvc_soft_attention_fwd: null argument");
  MVC_CHECK(T > 0 && A > 0 && F > 0 && keys_batch > 0, "mvc_soft_attention_fwd: bad dims");
  const int VN = keys_bf16 ? 8 : 4;
  const bool vec = (F % VN == 0) && (k_sb % VN == 0) && (k_st % VN == 0) &&
                   (reinterpret_cast<uintptr_t>(keys) % 16 == 0);
  const int vn = vec ? VN : 1;
  // F-split: enough CTAs for ~2 waves, and at most 256 vectors per CTA
  int fs_min = (int)cdiv(F, 256 * vn);
  int fs = (int)cdiv(2 * kNumSMs, B);
  if (fs < fs_min) fs = fs_min;
  int fs_max = (int)cdiv(F, vn);
  if (fs > fs_max) fs = fs_max;
  int chunk = (int)cdiv(cdiv(F, fs), vn) * vn;
  fs = (int)cdiv(F, chunk);
  const int nvec = chunk / vn;
  int G = 256 / nvec;
  if (G > T) G = T;
  if (G < 1) G = 1;
  (void)G; (void)nvec;
  const size_t smem = sizeof(float) * (2 * (size_t)A + T + 32 + (size_t)256 * vn);
  MVC_CHECK(smem <= 200 * 1024, "mvc_soft_attention_fwd: A=%d T=%d needs %zu B of shared memory", A, T, smem);
  dim3 grid(B, fs);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(PK_ATTN_FWD, B, T, F, st);
#define LAUNCH_FWD(KT, FAST, VEC)                                                                              \
  do {                                                                                                         \
    auto kern = soft_attention_fwd_kernel<KT, FAST, VEC>;                                                      \
    if (smem > 48 * 1024) MVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, 256, smem, st>>>(T, A, F, chunk, wq, uk, bias, w, (const KT*)keys, keys_batch, k_sb, k_st, mask, \
                                  m_sb, m_st, ctx_f32, ctx_ld, (__nv_bfloat16*)ctx_bf16, ctxb_ld, alpha);     \
  } while (0)
  if (keys_bf16) {
    if (fast_math) { if (vec) LAUNCH_FWD(__nv_bfloat16, true, true); else LAUNCH_FWD(__nv_bfloat16, true, false); }
    else { if (vec) LAUNCH_FWD(__nv_bfloat16, false, true); else LAUNCH_FWD(__nv_bfloat16, false, false); }
  } else {
    if (fast_math) { if (vec) LAUNCH_FWD(float, true, true); else LAUNCH_FWD(float, true, false); }
    else { if (vec) LAUNCH_FWD(float, false, true); else LAUNCH_FWD(float, false, false); }
  }
#undef LAUNCH_FWD
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_soft_attention_bwd(int B, int T, int A, int F, const float* wq, const float* uk, const float* bias,
                                      const float* w, const void* keys, int keys_bf16, int64_t k_sb, int64_t k_st,
                                      const float* alpha, const float* dctx, int64_t dctx_ld, float* dwq, float* duk,
                                      float* dw_partial, float* dkeys, int64_t dk_sb, int64_t dk_st, int fast_math,
                                      void* stream) {
  if (B == 0) return 0;
  MVC_CHECK(wq && uk && bias && w && keys && alpha && dctx && dwq, "mvc_soft_attention_bwd: null argument");
  const int VN = keys_bf16 ? 8 : 4;
  const bool vec = (F % VN == 0) && (k_sb % VN == 0) && (k_st % VN == 0) &&
                   (reinterpret_cast<uintptr_t>(keys) % 16 == 0);
  const size_t smem = sizeof(float) * ((size_t)((F + 3) & ~3) + 2 * (size_t)T + 32);
  MVC_CHECK(smem <= 200 * 1024, "mvc_soft_attention_bwd: F=%d T=%d needs %zu B of shared memory", F, T, smem);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(PK_ATTN_BWD, B, T, F, st);
#define LAUNCH_BWD(KT, FAST, VEC)                                                                              \
  do {                                                                                                         \
    auto kern = soft_attention_bwd_kernel<KT, FAST, VEC>;                                                      \
    if (smem > 48 * 1024) MVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<B, 512, smem, st>>>(T, A, F, wq, uk, bias, w, (const KT*)keys, k_sb, k_st, alpha, dctx, dctx_ld, dwq, \
                               duk, dw_partial, dkeys, dk_sb, dk_st);                                          \
  } while (0)
  if (keys_bf16) {
    if (fast_math) { if (vec) LAUNCH_BWD(__nv_bfloat16, true, true); else LAUNCH_BWD(__nv_bfloat16, true, false); }
    else { if (vec) LAUNCH_BWD(__nv_bfloat16, false, true); else LAUNCH_BWD(__nv_bfloat16, false, false); }
  } else {
    if (fast_math) { if (vec) LAUNCH_BWD(float, true, true); else LAUNCH_BWD(float, true, false); }
    else { if (vec) LAUNCH_BWD(float, false, true); else LAUNCH_BWD(float, false, false); }
  }
#undef LAUNCH_BWD
  MVC_LAUNCH_CHECK();
  return 0;
}
