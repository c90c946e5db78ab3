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
#include <cstdlib>
#include <mutex>
#include <unordered_set>

#include "ptx.cuh"
#include "step.cuh"

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

// eight fp16 values (the projected keys of the decode loops)
__device__ __forceinline__ void unpack_f16x8(const uint4& r, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}

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


// ===================================================================== staged kernels (sm_100a fast path)
// One CTA per (row, F-chunk).  The keys chunk [T x chunk] is pulled into shared memory by the TMA
// engine (T 1-D bulk copies on one mbarrier) and the U.k rows are prefetched into registers; both are
// loop invariant, so under programmatic dependent launch they are issued BEFORE griddepcontrol.wait
// and overlap the kernel that is still producing this step's query.  Thread layout for the scores:
// warp <-> frame t (round robin), lane <-> bottleneck unit a = lane + 32k.
constexpr int ATT_MAXR = 8;      // frame rounds per warp held in registers (forward, 9 warps)
constexpr int ATT_MAXR_BWD = 4;  // backward, 16 warps

template <typename KT, bool FAST, int AV, int NT>
__global__ void __launch_bounds__(NT, 2)
attn_fwd_staged_kernel(const AttnFwdArgs a, int chunk) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int NW = NT / 32;
  constexpr int VN = VecOf<KT>::N;
  const int T = a.T, A = AV * 32, F = a.F;
  const int kb = blockIdx.x;                     // key block; rows kb, kb + keys_batch, ... (beams) share it
  const int nq = a.B / a.keys_batch;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // The gridDim.y CTAs of one key block form a thread-block cluster: CTA r stages F-chunk r of the keys and scores
  // the frames [r*Tc, (r+1)*Tc); the scores are exchanged through distributed shared memory, so no work is repeated
  // however finely the row is split.  Everything loop invariant (keys chunk, U.k slab) arrives by TMA bulk copies:
  // register prefetches of U.k queued ~300 LDGs per SM in front of every other load (3 us before the first score).
  const int cl = gridDim.y, rank = blockIdx.y;
  const int f0 = rank * chunk, f1 = min(F, f0 + chunk), ncols = f1 - f0;
  const int Tc = (T + cl - 1) / cl, t0 = min(T, rank * Tc), t1 = min(T, t0 + Tc);
  const int Tp = (T + 3) & ~3;
  const size_t stage_bytes = ((size_t)T * chunk * sizeof(KT) + 127) & ~size_t(127);
  KT* sK = reinterpret_cast<KT*>(smem_raw);
  float* sU = reinterpret_cast<float*>(smem_raw + stage_bytes);     // [Tc][A] U.k rows of this CTA's frames
  float* sQ = sU + (size_t)Tc * A;
  float* sW = sQ + A;
  float* sE = sW + A;                            // [2][Tp]: raw scores, double buffered over queries
  float* sP = sE + 2 * Tp;                       // [Tp]: soft-max weights of the current query
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sP + Tp);
  const uint32_t bar_u = smem_u32(mbar), bar_k = smem_u32(mbar + 1);
  unsigned long long* prof = a.prof ? a.prof + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
#define AT_STAMP(i) do { if (prof && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); prof[i] = t_; } } while (0)
  AT_STAMP(0);
  if (wid == 0) {
    const KT* kbase = reinterpret_cast<const KT*>(a.keys) + (int64_t)kb * a.k_sb + f0;
    const uint32_t row_bytes = (uint32_t)(ncols * sizeof(KT));
    const bool one_copy = (ncols == chunk && a.k_st == chunk);       // contiguous [T, F] block
    if (lane == 0) {
      mbar_init(bar_u, 1);
      mbar_init(bar_k, 1);
      fence_mbar_init();
      // U.k slab first (the scores need it first), then the keys chunk
      const uint32_t ub = (uint32_t)((t1 - t0) * A * sizeof(float));
      if (ub) {
        mbar_expect_tx(bar_u, ub);
        bulk_load_1d(smem_u32(sU), a.uk + ((int64_t)kb * T + t0) * A, ub, bar_u);
      }
      if (ncols > 0) {
        mbar_expect_tx(bar_k, row_bytes * (uint32_t)T);
        if (one_copy) bulk_load_1d(smem_u32(sK), kbase, row_bytes * (uint32_t)T, bar_k);
      }
    }
    __syncwarp();
    // per-frame copies of an F-chunk: issued by the 32 lanes in parallel (one cp.async.bulk costs its issuing thread
    // ~0.1 us, so a single thread would spend 3 us on T = 30 of them)
    if (ncols > 0 && !one_copy)
      for (int t = lane; t < T; t += 32)
        bulk_load_1d(smem_u32(sK + (size_t)t * chunk), kbase + (int64_t)t * a.k_st, row_bytes, bar_k);
  }
  for (int i = tid; i < A; i += NT) sW[i] = a.w[i];
  __syncthreads();                  // barriers initialised, sW visible
  if (cl > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // "I am running" (waited below)
  AT_STAMP(1);
  pdl_trigger();
  pdl_wait();                       // the query (wq) comes from the preceding kernel
  if (cl > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");     // every peer CTA has started
  AT_STAMP(2);
  const uint32_t sE_u32 = smem_u32(sE);
  for (int qi = 0; qi < nq; ++qi) {
    const int b = qi * a.keys_batch + kb;
    float* sEq = sE + (qi & 1) * Tp;
    if (qi) __syncthreads();        // previous query's readers of sQ / sP are done
    for (int i = tid; i < A; i += NT) sQ[i] = a.wq[(int64_t)b * A + i] + a.bias[i];
    __syncthreads();
    if (qi == 0 && t1 > t0) mbar_wait(bar_u, 0);
    AT_STAMP(3);
    float qv[AV], wv[AV];
#pragma unroll
    for (int k = 0; k < AV; ++k) { qv[k] = sQ[lane + 32 * k]; wv[k] = sW[lane + 32 * k]; }
    for (int t = t0 + wid; t < t1; t += NW) {
      const float* urow = sU + (size_t)(t - t0) * A + lane;
      float e = 0.f;
#pragma unroll
      for (int k = 0; k < AV; ++k) e = fmaf(wv[k], tanh_sel<FAST>(qv[k] + urow[32 * k]), e);
      e = warp_sum(e);
      if (lane == 0) {
        if (a.mask && !a.mask[b * a.m_sb + t * a.m_st]) e = -INFINITY;
        if (cl == 1) {
          sEq[t] = e;
        } else {
          const uint32_t off = sE_u32 + (uint32_t)(((qi & 1) * Tp + t) * 4);
          for (int pr = 0; pr < cl; ++pr) {
            uint32_t dst;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(off), "r"(pr));
            asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst), "f"(e) : "memory");
          }
        }
      }
    }
    if (cl > 1) {
      // all scores of this query have landed in every CTA of the cluster.  One barrier per query is enough: the
      // buffers alternate, and a peer cannot start query qi + 2 before this CTA has passed the barrier of qi + 1.
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
      __syncthreads();
    }
    AT_STAMP(4);
    // soft-max over T by warp 0 (T <= 72: at most three values per lane)
    if (wid == 0) {
      float v[3];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int t = lane + 32 * i;
        v[i] = t < T ? sEq[t] : -INFINITY;
        mx = fmaxf(mx, v[i]);
      }
      mx = warp_max(mx);
      float sm = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int t = lane + 32 * i;
        v[i] = t < T ? (FAST ? __expf(v[i] - mx) : expf(v[i] - mx)) : 0.f;
        sm += v[i];
      }
      sm = warp_sum(sm);
      const float inv = 1.f / sm;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int t = lane + 32 * i;
        if (t < T) {
          const float pv = v[i] * inv;
          sP[t] = pv;
          if (rank == 0) a.alpha[(int64_t)b * T + t] = pv;
        }
      }
    }
    __syncthreads();
    AT_STAMP(5);
    if (ncols <= 0) continue;
    if (qi == 0) mbar_wait(bar_k, 0);   // keys chunk has landed
    AT_STAMP(6);
    const int nvec = ncols / VN;
    for (int v = tid; v < nvec; v += NT) {
      float acc[VN];
#pragma unroll
      for (int i = 0; i < VN; ++i) acc[i] = 0.f;
#pragma unroll 6
      for (int t = 0; t < T; ++t) {
        const typename VecOf<KT>::Raw raw = *reinterpret_cast<const typename VecOf<KT>::Raw*>(sK + (size_t)t * chunk + v * VN);
        float x[VN];
        VecOf<KT>::unpack(raw, x);
        const float pw = sP[t];
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[i] = fmaf(pw, x[i], acc[i]);
      }
      const int f = f0 + v * VN;
      if (a.ctx_f32) {          // vectorised path: f, ctx_ld multiples of VN and 16-byte aligned base (checked on the host)
        float4* dst = reinterpret_cast<float4*>(a.ctx_f32 + b * a.ctx_ld + f);
#pragma unroll
        for (int i = 0; i < VN / 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
      }
      if (a.ctx_bf16) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.ctx_bf16) + b * a.ctxb_ld + f;
        if constexpr (VN == 8) {
          uint4 pk;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[0], acc[1]), h1 = __floats2bfloat162_rn(acc[2], acc[3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[4], acc[5]), h3 = __floats2bfloat162_rn(acc[6], acc[7]);
          pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(dst) = pk;
        } else {
          uint2 pk;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[0], acc[1]), h1 = __floats2bfloat162_rn(acc[2], acc[3]);
          pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(dst) = pk;
        }
      }
    }
    AT_STAMP(7);
  }   // queries
#undef AT_STAMP
}

// Streaming variant for grids of several waves (greedy decode at B = 512: 3.5 key blocks per SM): ONE persistent CTA
// per SM walks its key blocks and keeps the TMA engine busy across them.  The [T, F] key stage is refilled in two
// frame halves -- half 0 of the next block is requested as soon as the context sum has consumed half 0 of the current
// one, half 1 likewise -- and the U.k slab is double buffered, so the loads of block i + 1 overlap the second half of
// the context sum, the stores and the query / score / soft-max phases of block i + 1.  (A fresh CTA per block
// serialises load -> compute: 4.7 us per block of which 3.6 us is the block's share of HBM time.)
// Requires: one query per key block, bf16 keys with k_st == F (contiguous block), F / 8 <= NT.
template <bool FAST, int AV, int NT, bool F16 = false>
__global__ void __launch_bounds__(NT, 1)
attn_fwd_stream_kernel(const AttnFwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  using KT = __nv_bfloat16;
  constexpr int NW = NT / 32;
  constexpr int VN = 8;
  const int T = a.T, A = AV * 32, F = a.F;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int Tp = (T + 3) & ~3;
  const int Th = (T + 1) / 2;                                       // frames in half 0
  const size_t stage_bytes = ((size_t)T * F * sizeof(KT) + 127) & ~size_t(127);
  KT* sK = reinterpret_cast<KT*>(smem_raw);
  float* sU = reinterpret_cast<float*>(smem_raw + stage_bytes);     // [2][T][A]
  float* sQ = sU + 2 * (size_t)T * A;                               // [2][A] query + bias, double buffered
  float* sW = sQ + 2 * A;
  float* sE = sW + A;                                               // [Tp]
  float* sP = sE + Tp;                                              // [Tp]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sP + Tp);
  const uint32_t bar_k0 = smem_u32(mbar), bar_k1 = smem_u32(mbar + 1), bar_u0 = smem_u32(mbar + 2);
  const uint32_t half0_bytes = (uint32_t)((size_t)Th * F * sizeof(KT)), half1_bytes = (uint32_t)((size_t)(T - Th) * F * sizeof(KT));
  const uint32_t slab_bytes = (uint32_t)((size_t)T * A * sizeof(float));
  const int nblk = a.keys_batch;
  auto load_keys = [&](int kb, int half) {      // tid 0 only
    const KT* src = reinterpret_cast<const KT*>(a.keys) + (int64_t)kb * a.k_sb + (half ? (size_t)Th * F : 0);
    const uint32_t bytes = half ? half1_bytes : half0_bytes;
    const uint32_t bar = half ? bar_k1 : bar_k0;
    if (bytes) {
      mbar_expect_tx(bar, bytes);
      bulk_load_1d(smem_u32(sK + (half ? (size_t)Th * F : 0)), src, bytes, bar);
    } else {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    }
  };
  auto load_slab = [&](int kb, int buf) {       // tid 0 only
    mbar_expect_tx(bar_u0 + 8u * buf, slab_bytes);
    bulk_load_1d(smem_u32(sU + (size_t)buf * T * A), a.uk + (int64_t)kb * T * A, slab_bytes, bar_u0 + 8u * buf);
  };
  if (tid == 0) {
    mbar_init(bar_k0, 1);
    mbar_init(bar_k1, 1);
    mbar_init(bar_u0, 1);
    mbar_init(bar_u0 + 8u, 1);
    fence_mbar_init();
    if ((int)blockIdx.x < nblk) {
      load_slab(blockIdx.x, 0);
      load_keys(blockIdx.x, 0);
      load_keys(blockIdx.x, 1);
    }
  }
  for (int i = tid; i < A; i += NT) sW[i] = a.w[i];
  __syncthreads();
  pdl_trigger();
  pdl_wait();                       // the queries (wq) come from the preceding kernel
  float wv[AV];
#pragma unroll
  for (int k = 0; k < AV; ++k) wv[k] = sW[lane + 32 * k];
  if ((int)blockIdx.x < nblk)
    for (int i = tid; i < A; i += NT) sQ[i] = a.wq[(int64_t)blockIdx.x * A + i] + a.bias[i];
  int it = 0;
  for (int kb = blockIdx.x; kb < nblk; kb += gridDim.x, ++it) {
    const int b = kb;
    const int ub = it & 1;
    const int nxt = kb + gridDim.x;
    // U.k slab of the next block into the other buffer (its last readers finished before the previous iteration's
    // barriers)
    if (tid == 0 && nxt < nblk) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of that buffer, async-proxy writes next
      load_slab(nxt, ub ^ 1);
    }
    // the query of the NEXT block is fetched now (global round trip hidden behind this block's work) and parked
    // in the other sQ buffer at the end of the iteration
    float qnext[(AV * 32 + NT - 1) / NT];
#pragma unroll
    for (int r = 0; r < (AV * 32 + NT - 1) / NT; ++r) {
      const int i = tid + r * NT;
      qnext[r] = (nxt < nblk && i < A) ? a.wq[(int64_t)nxt * A + i] + a.bias[i] : 0.f;
    }
    __syncthreads();                // sQ[ub] (written at the end of the previous iteration / in the prologue) is visible
    mbar_wait(bar_u0 + 8u * ub, (uint32_t)(it >> 1) & 1u);
    float qv[AV];
#pragma unroll
    for (int k = 0; k < AV; ++k) qv[k] = sQ[ub * A + lane + 32 * k];
    const float* slab = sU + (size_t)ub * T * A;
    for (int t = wid; t < T; t += NW) {
      const float* urow = slab + (size_t)t * A + lane;
      float e = 0.f;
#pragma unroll
      for (int k = 0; k < AV; ++k) e = fmaf(wv[k], tanh_sel<FAST>(qv[k] + urow[32 * k]), e);
      e = warp_sum(e);
      if (lane == 0) {
        if (a.mask && !a.mask[b * a.m_sb + t * a.m_st]) e = -INFINITY;
        sE[t] = e;
      }
    }
    __syncthreads();
    if (wid == 0) {                 // soft-max over T (T <= 72: at most three values per lane)
      float v[3];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int t = lane + 32 * i;
        v[i] = t < T ? sE[t] : -INFINITY;
        mx = fmaxf(mx, v[i]);
      }
      mx = warp_max(mx);
      float sm = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int t = lane + 32 * i;
        v[i] = t < T ? (FAST ? __expf(v[i] - mx) : expf(v[i] - mx)) : 0.f;
        sm += v[i];
      }
      sm = warp_sum(sm);
      const float inv = 1.f / sm;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int t = lane + 32 * i;
        if (t < T) {
          const float pv = v[i] * inv;
          sP[t] = pv;
          a.alpha[(int64_t)b * T + t] = pv;
        }
      }
    }
    __syncthreads();
    // context sum, frame half by frame half; each half's stage is handed back to the TMA engine as soon as every
    // thread has read it
    const int nvec = F / VN;
    const int v = tid;
    float acc[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) acc[i] = 0.f;
    mbar_wait(bar_k0, (uint32_t)it & 1u);
    if (v < nvec) {
#pragma unroll 5
      for (int t = 0; t < Th; ++t) {
        const uint4 raw = *reinterpret_cast<const uint4*>(sK + (size_t)t * F + v * VN);
        float x[VN];
        if constexpr (F16) unpack_f16x8(raw, x);
        else VecOf<KT>::unpack(raw, x);
        const float pw = sP[t];
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[i] = fmaf(pw, x[i], acc[i]);
      }
    }
    __syncthreads();
    if (tid == 0 && nxt < nblk) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads above, async-proxy writes next
      load_keys(nxt, 0);
    }
    mbar_wait(bar_k1, (uint32_t)it & 1u);
    if (v < nvec) {
#pragma unroll 5
      for (int t = Th; t < T; ++t) {
        const uint4 raw = *reinterpret_cast<const uint4*>(sK + (size_t)t * F + v * VN);
        float x[VN];
        if constexpr (F16) unpack_f16x8(raw, x);
        else VecOf<KT>::unpack(raw, x);
        const float pw = sP[t];
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[i] = fmaf(pw, x[i], acc[i]);
      }
      const int f = v * VN;
      if (a.ctx_f32) {
        float4* dst = reinterpret_cast<float4*>(a.ctx_f32 + b * a.ctx_ld + f);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      if (a.ctx_bf16) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.ctx_bf16) + b * a.ctxb_ld + f;
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[0], acc[1]), h1 = __floats2bfloat162_rn(acc[2], acc[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[4], acc[5]), h3 = __floats2bfloat162_rn(acc[6], acc[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst) = pk;
      }
    }
#pragma unroll
    for (int r = 0; r < (AV * 32 + NT - 1) / NT; ++r) {
      const int i = tid + r * NT;
      if (i < A) sQ[(ub ^ 1) * A + i] = qnext[r];
    }
    __syncthreads();
    if (tid == 0 && nxt < nblk) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      load_keys(nxt, 1);
    }
  }
}

// Multi-query variant (beam search: the `nq` beams of one video share its staged key block).  Up to QB queries are
// processed per pass: the U.k row of a frame is read once for all their scores, every staged key vector once for all
// their context sums, and the soft-max of query q runs on warp q -- instead of nq sequential single-query passes,
// each with its own global round trip for wq and its own four block-wide synchronisations.
template <typename KT, bool FAST, int AV, int NT, int QB>
__global__ void __launch_bounds__(NT, 1)
attn_fwd_staged_mq_kernel(const AttnFwdArgs a, int chunk) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int NW = NT / 32;
  constexpr int VN = VecOf<KT>::N;
  static_assert(QB <= NW, "one soft-max warp per query");
  const int T = a.T, A = AV * 32, F = a.F;
  const int kb = blockIdx.x;
  const int nq = a.B / a.keys_batch;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cl = gridDim.y, rank = blockIdx.y;
  const int f0 = rank * chunk, f1 = min(F, f0 + chunk), ncols = f1 - f0;
  const int Tc = (T + cl - 1) / cl, t0 = min(T, rank * Tc), t1 = min(T, t0 + Tc);
  const int Tp = (T + 3) & ~3;
  const size_t stage_bytes = ((size_t)T * chunk * sizeof(KT) + 127) & ~size_t(127);
  KT* sK = reinterpret_cast<KT*>(smem_raw);
  float* sU = reinterpret_cast<float*>(smem_raw + stage_bytes);     // [Tc][A]
  float* sQ = sU + (size_t)Tc * A;                                  // [QB][A]
  float* sW = sQ + (size_t)QB * A;                                  // [A]
  float* sE = sW + A;                                               // [2][QB][Tp] raw scores, double buffered over passes
  float* sP = sE + 2 * QB * Tp;                                     // [QB][Tp] soft-max weights of the current pass
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sP + QB * Tp);
  const uint32_t bar_u = smem_u32(mbar), bar_k = smem_u32(mbar + 1);
  unsigned long long* prof = a.prof ? a.prof + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
#define AT_STAMP(i) do { if (prof && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); prof[i] = t_; } } while (0)
  AT_STAMP(0);
  if (wid == 0) {
    const KT* kbase = reinterpret_cast<const KT*>(a.keys) + (int64_t)kb * a.k_sb + f0;
    const uint32_t row_bytes = (uint32_t)(ncols * sizeof(KT));
    const bool one_copy = (ncols == chunk && a.k_st == chunk);
    if (lane == 0) {
      mbar_init(bar_u, 1);
      mbar_init(bar_k, 1);
      fence_mbar_init();
      const uint32_t ub = (uint32_t)((t1 - t0) * A * sizeof(float));
      if (ub) {
        mbar_expect_tx(bar_u, ub);
        bulk_load_1d(smem_u32(sU), a.uk + ((int64_t)kb * T + t0) * A, ub, bar_u);
      }
      if (ncols > 0) {
        mbar_expect_tx(bar_k, row_bytes * (uint32_t)T);
        if (one_copy) bulk_load_1d(smem_u32(sK), kbase, row_bytes * (uint32_t)T, bar_k);
      }
    }
    __syncwarp();
    if (ncols > 0 && !one_copy)
      for (int t = lane; t < T; t += 32)
        bulk_load_1d(smem_u32(sK + (size_t)t * chunk), kbase + (int64_t)t * a.k_st, row_bytes, bar_k);
  }
  for (int i = tid; i < A; i += NT) sW[i] = a.w[i];
  __syncthreads();
  if (cl > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  AT_STAMP(1);
  pdl_trigger();
  pdl_wait();
  if (cl > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  AT_STAMP(2);
  const uint32_t sE_u32 = smem_u32(sE);
  float wv[AV];
#pragma unroll
  for (int k = 0; k < AV; ++k) wv[k] = sW[lane + 32 * k];
  int pass = 0;
  for (int q0 = 0; q0 < nq; q0 += QB, ++pass) {
    const int nqp = min(QB, nq - q0);
    float* sEp = sE + (pass & 1) * QB * Tp;
    if (q0) __syncthreads();        // previous pass's readers of sQ / sP are done
    for (int i = tid; i < nqp * A; i += NT) {
      const int q = i / A, c = i - q * A;
      sQ[i] = a.wq[((int64_t)(q0 + q) * a.keys_batch + kb) * A + c] + a.bias[c];
    }
    __syncthreads();
    if (q0 == 0 && t1 > t0) mbar_wait(bar_u, 0);
    AT_STAMP(3);
    for (int t = t0 + wid; t < t1; t += NW) {
      const float* urow = sU + (size_t)(t - t0) * A + lane;
      float u[AV];
#pragma unroll
      for (int k = 0; k < AV; ++k) u[k] = urow[32 * k];
      // branch-free over the QB slots (slots >= nqp compute on stale shared memory and are never stored): a uniform
      // branch per query cut the loop into 9-instruction basic blocks, each exposing its own shared-memory latency
      float e[QB];
#pragma unroll
      for (int q = 0; q < QB; ++q) {
        e[q] = 0.f;
        const float* qrow = sQ + q * A + lane;
#pragma unroll
        for (int k = 0; k < AV; ++k) e[q] = fmaf(wv[k], tanh_sel<FAST>(qrow[32 * k] + u[k]), e[q]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < QB; ++q) e[q] += __shfl_xor_sync(0xffffffffu, e[q], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < QB; ++q) {
          if (q < nqp) {
            float ev = e[q];
            const int b = (q0 + q) * a.keys_batch + kb;
            if (a.mask && !a.mask[b * a.m_sb + t * a.m_st]) ev = -INFINITY;
            if (cl == 1) {
              sEp[q * Tp + t] = ev;
            } else {
              const uint32_t off = sE_u32 + (uint32_t)((((pass & 1) * QB + q) * Tp + t) * 4);
              for (int pr = 0; pr < cl; ++pr) {
                uint32_t dst;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(off), "r"(pr));
                asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst), "f"(ev) : "memory");
              }
            }
          }
        }
      }
    }
    if (cl > 1) {
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
      __syncthreads();
    }
    AT_STAMP(4);
    // soft-max of query q on warp q (T <= 72: at most three values per lane)
    if (wid < nqp) {
      const float* er = sEp + wid * Tp;
      const int b = (q0 + wid) * a.keys_batch + kb;
      float v[3];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int t = lane + 32 * i;
        v[i] = t < T ? er[t] : -INFINITY;
        mx = fmaxf(mx, v[i]);
      }
      mx = warp_max(mx);
      float sm = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int t = lane + 32 * i;
        v[i] = t < T ? (FAST ? __expf(v[i] - mx) : expf(v[i] - mx)) : 0.f;
        sm += v[i];
      }
      sm = warp_sum(sm);
      const float inv = 1.f / sm;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int t = lane + 32 * i;
        if (t < T) {
          const float pv = v[i] * inv;
          sP[wid * Tp + t] = pv;
          if (rank == 0) a.alpha[(int64_t)b * T + t] = pv;
        }
      }
    }
    __syncthreads();
    AT_STAMP(5);
    if (ncols <= 0) continue;
    if (q0 == 0) mbar_wait(bar_k, 0);
    AT_STAMP(6);
    const int nvec = ncols / VN;
    for (int v = tid; v < nvec; v += NT) {
      float acc[QB][VN];
#pragma unroll
      for (int q = 0; q < QB; ++q)
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[q][i] = 0.f;
#pragma unroll 6
      for (int t = 0; t < T; ++t) {
        const typename VecOf<KT>::Raw raw = *reinterpret_cast<const typename VecOf<KT>::Raw*>(sK + (size_t)t * chunk + v * VN);
        float x[VN];
        VecOf<KT>::unpack(raw, x);
        float pw[QB];
#pragma unroll
        for (int q = 0; q < QB; ++q) pw[q] = sP[q * Tp + t];
#pragma unroll
        for (int q = 0; q < QB; ++q)
#pragma unroll
          for (int i = 0; i < VN; ++i) acc[q][i] = fmaf(pw[q], x[i], acc[q][i]);
      }
      const int f = f0 + v * VN;
#pragma unroll
      for (int q = 0; q < QB; ++q) {
        if (q < nqp) {
          const int64_t b = (int64_t)(q0 + q) * a.keys_batch + kb;
          if (a.ctx_f32) {
            float4* dst = reinterpret_cast<float4*>(a.ctx_f32 + b * a.ctx_ld + f);
#pragma unroll
            for (int i = 0; i < VN / 4; ++i)
              dst[i] = make_float4(acc[q][4 * i], acc[q][4 * i + 1], acc[q][4 * i + 2], acc[q][4 * i + 3]);
          }
          if (a.ctx_bf16) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.ctx_bf16) + b * a.ctxb_ld + f;
            if constexpr (VN == 8) {
              uint4 pk;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[q][0], acc[q][1]), h1 = __floats2bfloat162_rn(acc[q][2], acc[q][3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[q][4], acc[q][5]), h3 = __floats2bfloat162_rn(acc[q][6], acc[q][7]);
              pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(dst) = pk;
            } else {
              uint2 pk;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[q][0], acc[q][1]), h1 = __floats2bfloat162_rn(acc[q][2], acc[q][3]);
              pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
              *reinterpret_cast<uint2*>(dst) = pk;
            }
          }
        }
      }
    }
    AT_STAMP(7);
  }   // passes
#undef AT_STAMP
}

// Backward, one CTA per batch row; keys row block staged by TMA bulk copies, U.k in registers.
template <typename KT, bool FAST, int AV, int NT>
__global__ void __launch_bounds__(NT)
attn_bwd_staged_kernel(const AttnBwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int NW = NT / 32;
  constexpr int VN = VecOf<KT>::N;
  const int T = a.T, A = AV * 32, F = a.F;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  size_t stage_bytes = ((size_t)T * F * sizeof(KT) + 127) & ~size_t(127);
  if (stage_bytes < sizeof(float) * NW * 2 * A) stage_bytes = sizeof(float) * NW * 2 * A;
  KT* sK = reinterpret_cast<KT*>(smem_raw);
  // dctx, stored so that lane l's element (m*32*VN + l*VN + i) sits at word (m*VN + i)*32 + l: conflict-free
  auto dpos = [](int f) { return ((f / (32 * VN)) * VN + (f % VN)) * 32 + (f % (32 * VN)) / VN; };
  float* sD = reinterpret_cast<float*>(smem_raw + stage_bytes);       // dctx [F rounded to 32*VN]
  float* sAl = sD + (F + 32 * VN - 1) / (32 * VN) * (32 * VN);        // alpha [T]
  float* sDa = sAl + ((T + 3) & ~3);                                  // dalpha -> de [T]
  float* sQ = sDa + ((T + 3) & ~3);                                   // wq + bias [A]
  float* sW = sQ + A;                                                 // w [A]
  float* sRed = sW + A;
  float* sAcc = reinterpret_cast<float*>(smem_raw);                   // [NW][2][A] per-warp partials; reuses the keys
                                                                      // stage once dalpha is done
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sRed + 32);
  const uint32_t bar = smem_u32(mbar);

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    const KT* kbase = reinterpret_cast<const KT*>(a.keys) + (int64_t)b * a.k_sb;
    const uint32_t row_bytes = (uint32_t)(F * sizeof(KT));
    mbar_expect_tx(bar, row_bytes * (uint32_t)T);
    if (a.k_st == F) {
      bulk_load_1d(smem_u32(sK), kbase, row_bytes * (uint32_t)T, bar);      // contiguous [T, F] block: one copy
    } else {
      for (int t = 0; t < T; ++t) bulk_load_1d(smem_u32(sK + (size_t)t * F), kbase + (int64_t)t * a.k_st, row_bytes, bar);
    }
  }
  float ur[ATT_MAXR_BWD][AV];
  const float* ukb = a.uk + (int64_t)b * T * A;
#pragma unroll
  for (int r = 0; r < ATT_MAXR_BWD; ++r) {
    const int t = wid + r * NW;
    if (t < T) {
#pragma unroll
      for (int k = 0; k < AV; ++k) ur[r][k] = __ldg(ukb + (int64_t)t * A + lane + 32 * k);
    }
  }
  for (int i = tid; i < A; i += NT) sW[i] = a.w[i];
  pdl_trigger();
  pdl_wait();
  for (int f = tid; f < F; f += NT) sD[dpos(f)] = a.dctx[b * a.dctx_ld + f];
  for (int t = tid; t < T; t += NT) sAl[t] = a.alpha[(int64_t)b * T + t];
  for (int i = tid; i < A; i += NT) sQ[i] = a.wq[(int64_t)b * A + i] + a.bias[i];
  __syncthreads();
  mbar_wait(bar, 0);
  // dalpha[t] = dctx . keys[t]
  for (int t = wid; t < T; t += NW) {
    const KT* row = sK + (size_t)t * F;
    float d = 0.f;
    for (int m = 0, f = lane * VN; f < F; f += 32 * VN, ++m) {
      const typename VecOf<KT>::Raw raw = *reinterpret_cast<const typename VecOf<KT>::Raw*>(row + f);
      float x[VN];
      VecOf<KT>::unpack(raw, x);
#pragma unroll
      for (int i = 0; i < VN; ++i) d = fmaf(sD[(m * VN + i) * 32 + lane], x[i], d);
    }
    d = warp_sum(d);
    if (lane == 0) sDa[t] = d;
  }
  __syncthreads();                  // all reads of the staged keys are done: the stage area is free (sAcc)
  float part = 0.f;
  for (int t = tid; t < T; t += NT) part += sAl[t] * sDa[t];
  const float dot = block_sum(part, sRed);
  for (int t = tid; t < T; t += NT) sDa[t] = sAl[t] * (sDa[t] - dot);      // de[t]
  __syncthreads();
  // dpre[t,a] = de[t] w[a] (1 - tanh^2);  warp <-> t, lane <-> a
  float sq[AV], sw[AV];
#pragma unroll
  for (int k = 0; k < AV; ++k) { sq[k] = 0.f; sw[k] = 0.f; }
  float* dukb = a.duk ? a.duk + (int64_t)b * T * A : nullptr;
  // duk accumulates over the decode steps: fetch all old values first (independent loads in flight together; a
  // load-add-store per element in program order exposed one global round trip per element, ~0.7 us x 32)
  float dold[ATT_MAXR_BWD][AV];
  if (dukb) {
#pragma unroll
    for (int r = 0; r < ATT_MAXR_BWD; ++r) {
      const int t = wid + r * NW;
      if (t < T) {
#pragma unroll
        for (int k = 0; k < AV; ++k) dold[r][k] = dukb[(int64_t)t * A + lane + 32 * k];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ATT_MAXR_BWD; ++r) {
    const int t = wid + r * NW;
    if (t < T) {
      const float de = sDa[t];
#pragma unroll
      for (int k = 0; k < AV; ++k) {
        const int i = lane + 32 * k;
        const float th = tanh_sel<FAST>(sQ[i] + ur[r][k]);
        const float dpre = de * sW[i] * (1.f - th * th);
        sq[k] += dpre;
        sw[k] = fmaf(de, th, sw[k]);
        if (dukb) dukb[(int64_t)t * A + i] = dold[r][k] + dpre;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < AV; ++k) {
    sAcc[((size_t)wid * 2 + 0) * A + lane + 32 * k] = sq[k];
    sAcc[((size_t)wid * 2 + 1) * A + lane + 32 * k] = sw[k];
  }
  __syncthreads();
  for (int i = tid; i < A; i += NT) {
    float q = 0.f, w2 = 0.f;
#pragma unroll 4
    for (int w = 0; w < NW; ++w) {          // fixed order: deterministic
      q += sAcc[((size_t)w * 2 + 0) * A + i];
      w2 += sAcc[((size_t)w * 2 + 1) * A + i];
    }
    a.dwq[(int64_t)b * A + i] = q;
    if (a.dwq_bf16) reinterpret_cast<__nv_bfloat16*>(a.dwq_bf16)[(int64_t)b * A + i] = __float2bfloat16(q);
    if (a.dw_partial) a.dw_partial[(int64_t)b * A + i] += w2;
  }
  if (a.dkeys) {
    // dkeys[t, f] += alpha[t] * dctx[f]: four independent read-modify-writes in flight per thread
    float* dkb = a.dkeys + (int64_t)b * a.dk_sb;
    const int total = T * F;
    for (int base = tid; base < total; base += 4 * NT) {
      float old[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * NT;
        if (i < total) { const int t = i / F, f = i - t * F; old[u] = dkb[(int64_t)t * a.dk_st + f]; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * NT;
        if (i < total) { const int t = i / F, f = i - t * F; dkb[(int64_t)t * a.dk_st + f] = old[u] + sAl[t] * sD[dpos(f)]; }
      }
    }
  }
}

template <typename T>
static bool aligned16(const T* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace mvc

using namespace mvc;

namespace mvc {

constexpr size_t kAttnMaxSmem = 220 * 1024;

static int ensure_big_smem(const void* kern) {
  static std::mutex mu;
  static std::unordered_set<const void*> done;
  std::lock_guard<std::mutex> lk(mu);
  if (done.count(kern)) return 0;
  MVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnMaxSmem));
  // all of the unified L1 / shared array as shared memory: several staged CTAs per SM when their stages are small
  MVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  done.insert(kern);
  return 0;
}

static int launch_ex(const void* kern, dim3 grid, dim3 block, size_t smem, bool pdl, void** args, cudaStream_t st,
                     int cluster_y = 1) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl && pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_y > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = (unsigned)cluster_y; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  MVC_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
  return 0;
}

template <typename KT, bool FAST>
static const void* pick_fwd_staged(int A) {
  switch (A) {
    case 32: return (const void*)attn_fwd_staged_kernel<KT, FAST, 1, 288>;
    case 64: return (const void*)attn_fwd_staged_kernel<KT, FAST, 2, 288>;
    case 128: return (const void*)attn_fwd_staged_kernel<KT, FAST, 4, 288>;
    case 256: return (const void*)attn_fwd_staged_kernel<KT, FAST, 8, 288>;
    default: return nullptr;
  }
}
// queries per pass of the multi-query kernel: 5 (the reference's beam width, features_captioning.py:131) or 8
template <bool FAST, bool F16 = false>
static const void* pick_fwd_stream(int A) {
  switch (A) {
    case 32: return (const void*)attn_fwd_stream_kernel<FAST, 1, 288, F16>;
    case 64: return (const void*)attn_fwd_stream_kernel<FAST, 2, 288, F16>;
    case 128: return (const void*)attn_fwd_stream_kernel<FAST, 4, 288, F16>;
    case 256: return (const void*)attn_fwd_stream_kernel<FAST, 8, 288, F16>;
    default: return nullptr;
  }
}
static size_t stream_smem(int T, int A, int F) {
  const size_t tp = (size_t)((T + 3) & ~3);
  return (((size_t)T * F * 2 + 127) & ~size_t(127)) + sizeof(float) * (2 * (size_t)T * A + 3 * (size_t)A + 2 * tp) + 64;
}
// would launch_attention_fwd take the streaming kernel for one query per key block of F 16-bit values?
bool attention_stream_eligible(int rows, int keys_batch, int T, int A, int F) {
  return rows == keys_batch && keys_batch > kNumSMs && F % 8 == 0 && F / 8 <= 288 && T <= ATT_MAXR * 9 &&
         (A == 32 || A == 64 || A == 128 || A == 256) && stream_smem(T, A, F) <= kAttnMaxSmem &&
         !getenv("MVC_B200_ATTN_NOSTREAM");
}

template <int QB>
static const void* pick_fwd_staged_mq(int A) {
  switch (A) {
    case 32: return (const void*)attn_fwd_staged_mq_kernel<__nv_bfloat16, true, 1, 288, QB>;
    case 64: return (const void*)attn_fwd_staged_mq_kernel<__nv_bfloat16, true, 2, 288, QB>;
    case 128: return (const void*)attn_fwd_staged_mq_kernel<__nv_bfloat16, true, 4, 288, QB>;
    case 256: return (const void*)attn_fwd_staged_mq_kernel<__nv_bfloat16, true, 8, 288, QB>;
    default: return nullptr;
  }
}

template <typename KT, bool FAST>
static const void* pick_bwd_staged(int A) {
  switch (A) {
    case 32: return (const void*)attn_bwd_staged_kernel<KT, FAST, 1, 512>;
    case 64: return (const void*)attn_bwd_staged_kernel<KT, FAST, 2, 512>;
    case 128: return (const void*)attn_bwd_staged_kernel<KT, FAST, 4, 512>;
    case 256: return (const void*)attn_bwd_staged_kernel<KT, FAST, 8, 512>;
    default: return nullptr;
  }
}

static unsigned long long* g_attn_prof = nullptr;
void set_attn_prof(unsigned long long* p) { g_attn_prof = p; }

int launch_attention_fwd(const AttnFwdArgs& a, bool pdl, cudaStream_t st) {
  const int B = a.B, T = a.T, A = a.A, F = a.F;
  if (B == 0) return 0;
  MVC_CHECK(a.wq && a.uk && a.bias && a.w && a.keys && a.alpha, "mvc_soft_attention_fwd: null argument");
  MVC_CHECK(T > 0 && A > 0 && F > 0 && a.keys_batch > 0, "mvc_soft_attention_fwd: bad dims");
  const int VN = a.keys_bf16 ? 8 : 4;
  const size_t es = a.keys_bf16 ? 2 : 4;
  const bool vec = (F % VN == 0) && (a.k_sb % VN == 0) && (a.k_st % VN == 0) &&
                   (reinterpret_cast<uintptr_t>(a.keys) % 16 == 0);
  // the staged kernel also stores the context with 16-byte vectors and pulls U.k by bulk copies
  const bool vec_out = (!a.ctx_f32 || (reinterpret_cast<uintptr_t>(a.ctx_f32) % 16 == 0 && a.ctx_ld % 4 == 0)) &&
                       (!a.ctx_bf16 || (reinterpret_cast<uintptr_t>(a.ctx_bf16) % 16 == 0 && a.ctxb_ld % 8 == 0)) &&
                       (reinterpret_cast<uintptr_t>(a.uk) % 16 == 0);
  ProfScope prof(PK_ATTN_FWD, B, T, F, st);
  // ---- staged fast path
  const void* kern = nullptr;
  if (vec && vec_out && T <= ATT_MAXR * 9 && B % a.keys_batch == 0) {
    if (a.keys_bf16) kern = a.fast_math ? pick_fwd_staged<__nv_bfloat16, true>(A) : pick_fwd_staged<__nv_bfloat16, false>(A);
    else kern = a.fast_math ? pick_fwd_staged<float, true>(A) : pick_fwd_staged<float, false>(A);
  }
  // several queries per key block (beam search), bf16 keys, fast math: the multi-query kernel
  const int nq = a.keys_batch > 0 ? B / a.keys_batch : 1;
  const int qb = nq <= 5 ? 5 : 8;
  const void* kern_mq = qb == 5 ? pick_fwd_staged_mq<5>(A) : pick_fwd_staged_mq<8>(A);
  const bool mq = kern && nq > 1 && a.keys_bf16 && a.fast_math && kern_mq;
  if (mq) kern = kern_mq;
  // several waves of key blocks, one query each (greedy decode): the streaming kernel, one persistent CTA per SM
  if (kern && !mq && nq == 1 && a.keys_bf16 && a.k_st == F && a.keys_batch > kNumSMs && F / 8 <= 288 &&
      !getenv("MVC_B200_ATTN_NOSTREAM")) {
    const void* ks = a.keys_f16 ? (a.fast_math ? pick_fwd_stream<true, true>(A) : pick_fwd_stream<false, true>(A))
                                : (a.fast_math ? pick_fwd_stream<true>(A) : pick_fwd_stream<false>(A));
    const size_t smem = stream_smem(T, A, F);
    if (ks && smem <= kAttnMaxSmem) {
      MVC_TRY(ensure_big_smem(ks));
      AttnFwdArgs args = a;
      void* params[] = {(void*)&args};
      MVC_TRY(launch_ex(ks, dim3(kNumSMs), dim3(288), smem, pdl, params, st, 1));
      MVC_LAUNCH_CHECK();
      return 0;
    }
  }
  MVC_CHECK(!a.keys_f16, "mvc_soft_attention_fwd: fp16 (projected) keys are served by the streaming kernel only");
  if (kern) {
    const size_t tp = (size_t)((T + 3) & ~3);
    const size_t tail0 = mq ? sizeof(float) * ((qb + 1) * (size_t)A + 3 * qb * tp) + 16
                            : sizeof(float) * (2 * (size_t)A + 3 * tp) + 16;
    // F-chunks per key block = CTAs per cluster (1, 2, 4 or 8; scores are shared through DSMEM, so splitting costs no
    // repeated work): the fewest whose staged keys fit in shared memory, more while the grid cannot fill the SMs.
    // An unsplit row whose [T, F] block is contiguous arrives by ONE bulk copy; an F-chunk needs T strided copies,
    // and each cp.async.bulk occupies the SM's TMA unit for ~0.1 us whatever its size (measured: 3.6 us for 30
    // copies of 2 KB), so grids that already fill the machine are never split further.
    int fs = 1;
    auto smem_for = [&](int fsplit, int* chunk_out) {
      int chunk = (int)cdiv(cdiv(F, fsplit), VN) * VN;
      *chunk_out = chunk;
      const size_t uslab = sizeof(float) * (size_t)cdiv(T, fsplit) * A;     // U.k rows of the CTA's frames
      return (((size_t)T * chunk * es + 127) & ~size_t(127)) + uslab + tail0;
    };
    int chunk = F;
    while (smem_for(fs, &chunk) > kAttnMaxSmem && fs < 8) fs *= 2;
    while ((int64_t)a.keys_batch * fs < kNumSMs && fs < 8 && chunk > 64 * VN) { fs *= 2; smem_for(fs, &chunk); }
    if (const char* e = getenv("MVC_B200_ATTN_FS")) { fs = atoi(e); }   // tuning aid
    smem_for(fs, &chunk);
    fs = (int)cdiv(F, chunk);       // CTAs per cluster actually needed at this chunk width
    const size_t smem = (((size_t)T * chunk * es + 127) & ~size_t(127)) + sizeof(float) * (size_t)cdiv(T, fs) * A + tail0;
    if (smem <= kAttnMaxSmem && fs <= 8) {
      MVC_TRY(ensure_big_smem(kern));
      AttnFwdArgs args = a;
      args.prof = g_attn_prof;
      void* params[] = {(void*)&args, (void*)&chunk};
      MVC_TRY(launch_ex(kern, dim3(a.keys_batch, fs), dim3(288), smem, pdl, params, st, fs));
      MVC_LAUNCH_CHECK();
      return 0;
    }
  }
  // ---- generic path (any A / T / alignment): direct global loads
  const int vn = vec ? VN : 1;
  int fs_min = (int)cdiv(F, 256 * vn);
  int fs = (int)cdiv(2 * kNumSMs, B);
  if (fs < fs_min) fs = fs_min;
  int fs_max = (int)cdiv(F, vn);
  if (fs > fs_max) fs = fs_max;
  int chunk = (int)cdiv(cdiv(F, fs), vn) * vn;
  fs = (int)cdiv(F, chunk);
  const size_t smem = sizeof(float) * (2 * (size_t)A + T + 32 + (size_t)256 * vn);
  MVC_CHECK(smem <= 200 * 1024, "mvc_soft_attention_fwd: A=%d T=%d needs %zu B of shared memory", A, T, smem);
  dim3 grid(B, fs);
#define LAUNCH_FWD(KT, FAST, VEC)                                                                              \
  do {                                                                                                         \
    auto k2 = soft_attention_fwd_kernel<KT, FAST, VEC>;                                                        \
    if (smem > 48 * 1024) MVC_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k2<<<grid, 256, smem, st>>>(T, A, F, chunk, a.wq, a.uk, a.bias, a.w, (const KT*)a.keys, a.keys_batch, a.k_sb, a.k_st, \
                                a.mask, a.m_sb, a.m_st, a.ctx_f32, a.ctx_ld, (__nv_bfloat16*)a.ctx_bf16, a.ctxb_ld, \
                                a.alpha);                                                                      \
  } while (0)
  if (a.keys_bf16) {
    if (a.fast_math) { if (vec) LAUNCH_FWD(__nv_bfloat16, true, true); else LAUNCH_FWD(__nv_bfloat16, true, false); }
    else { if (vec) LAUNCH_FWD(__nv_bfloat16, false, true); else LAUNCH_FWD(__nv_bfloat16, false, false); }
  } else {
    if (a.fast_math) { if (vec) LAUNCH_FWD(float, true, true); else LAUNCH_FWD(float, true, false); }
    else { if (vec) LAUNCH_FWD(float, false, true); else LAUNCH_FWD(float, false, false); }
  }
#undef LAUNCH_FWD
  MVC_LAUNCH_CHECK();
  return 0;
}

__global__ void cast_rows_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(src[i]);
}

int launch_attention_bwd(const AttnBwdArgs& a, bool pdl, cudaStream_t st) {
  const int B = a.B, T = a.T, A = a.A, F = a.F;
  if (B == 0) return 0;
  MVC_CHECK(a.wq && a.uk && a.bias && a.w && a.keys && a.alpha && a.dctx && a.dwq, "mvc_soft_attention_bwd: null argument");
  const int VN = a.keys_bf16 ? 8 : 4;
  const size_t es = a.keys_bf16 ? 2 : 4;
  const bool vec = (F % VN == 0) && (a.k_sb % VN == 0) && (a.k_st % VN == 0) &&
                   (reinterpret_cast<uintptr_t>(a.keys) % 16 == 0);
  ProfScope prof(PK_ATTN_BWD, B, T, F, st);
  const void* kern = nullptr;
  if (vec && T <= ATT_MAXR_BWD * 16) {
    if (a.keys_bf16) kern = a.fast_math ? pick_bwd_staged<__nv_bfloat16, true>(A) : pick_bwd_staged<__nv_bfloat16, false>(A);
    else kern = a.fast_math ? pick_bwd_staged<float, true>(A) : pick_bwd_staged<float, false>(A);
  }
  if (kern) {
    size_t stage = ((size_t)T * F * es + 127) & ~size_t(127);
    if (stage < sizeof(float) * 16 * 2 * (size_t)A) stage = sizeof(float) * 16 * 2 * (size_t)A;
    const size_t smem = stage + sizeof(float) * ((size_t)cdiv(F, 32 * VN) * 32 * VN + 2 * (size_t)((T + 3) & ~3) +
                                                 2 * (size_t)A + 32) + 16;
    if (smem <= kAttnMaxSmem) {
      MVC_TRY(ensure_big_smem(kern));
      AttnBwdArgs args = a;
      void* params[] = {(void*)&args};
      MVC_TRY(launch_ex(kern, dim3(B), dim3(512), smem, pdl, params, st));
      MVC_LAUNCH_CHECK();
      return 0;
    }
  }
  const size_t smem = sizeof(float) * ((size_t)((F + 3) & ~3) + 2 * (size_t)T + 32);
  MVC_CHECK(smem <= 200 * 1024, "mvc_soft_attention_bwd: F=%d T=%d needs %zu B of shared memory", F, T, smem);
#define LAUNCH_BWD(KT, FAST, VEC)                                                                              \
  do {                                                                                                         \
    auto k2 = soft_attention_bwd_kernel<KT, FAST, VEC>;                                                        \
    if (smem > 48 * 1024) MVC_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k2<<<B, 512, smem, st>>>(T, A, F, a.wq, a.uk, a.bias, a.w, (const KT*)a.keys, a.k_sb, a.k_st, a.alpha, a.dctx, \
                             a.dctx_ld, a.dwq, a.duk, a.dw_partial, a.dkeys, a.dk_sb, a.dk_st);                \
  } while (0)
  if (a.keys_bf16) {
    if (a.fast_math) { if (vec) LAUNCH_BWD(__nv_bfloat16, true, true); else LAUNCH_BWD(__nv_bfloat16, true, false); }
    else { if (vec) LAUNCH_BWD(__nv_bfloat16, false, true); else LAUNCH_BWD(__nv_bfloat16, false, false); }
  } else {
    if (a.fast_math) { if (vec) LAUNCH_BWD(float, true, true); else LAUNCH_BWD(float, true, false); }
    else { if (vec) LAUNCH_BWD(float, false, true); else LAUNCH_BWD(float, false, false); }
  }
#undef LAUNCH_BWD
  MVC_LAUNCH_CHECK();
  if (a.dwq_bf16) {
    const int64_t n = (int64_t)B * A;
    cast_rows_bf16_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(a.dwq, (__nv_bfloat16*)a.dwq_bf16, n);
    MVC_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace mvc

extern "C" int mvc_soft_attention_fwd(int B, int T, int A, int F, const float* wq, const float* uk, const float* bias,
                                      const float* w, const void* keys, int keys_bf16, int keys_batch, int64_t k_sb,
                                      int64_t k_st, const uint8_t* mask, int64_t m_sb, int64_t m_st, float* ctx_f32,
                                      int64_t ctx_ld, void* ctx_bf16, int64_t ctxb_ld, float* alpha, int fast_math,
                                      void* stream) {
  AttnFwdArgs a{};
  a.B = B; a.T = T; a.A = A; a.F = F; a.wq = wq; a.uk = uk; a.bias = bias; a.w = w;
  a.keys = keys; a.keys_bf16 = keys_bf16; a.keys_batch = keys_batch; a.k_sb = k_sb; a.k_st = k_st;
  a.mask = mask; a.m_sb = m_sb; a.m_st = m_st; a.ctx_f32 = ctx_f32; a.ctx_ld = ctx_ld; a.ctx_bf16 = ctx_bf16;
  a.ctxb_ld = ctxb_ld; a.alpha = alpha; a.fast_math = fast_math;
  return launch_attention_fwd(a, false, (cudaStream_t)stream);
}

extern "C" int mvc_soft_attention_bwd(int B, int T, int A, int F, const float* wq, const float* uk, const float* bias,
                                      const float* w, const void* keys, int keys_bf16, int64_t k_sb, int64_t k_st,
                                      const float* alpha, const float* dctx, int64_t dctx_ld, float* dwq, float* duk,
                                      float* dw_partial, float* dkeys, int64_t dk_sb, int64_t dk_st, int fast_math,
                                      void* stream) {
  AttnBwdArgs a{};
  a.B = B; a.T = T; a.A = A; a.F = F; a.wq = wq; a.uk = uk; a.bias = bias; a.w = w;
  a.keys = keys; a.keys_bf16 = keys_bf16; a.k_sb = k_sb; a.k_st = k_st; a.alpha = alpha; a.dctx = dctx; a.dctx_ld = dctx_ld;
  a.dwq = dwq; a.dwq_bf16 = nullptr; a.duk = duk; a.dw_partial = dw_partial; a.dkeys = dkeys; a.dk_sb = dk_sb; a.dk_st = dk_st;
  a.fast_math = fast_math;
  return launch_attention_bwd(a, false, (cudaStream_t)stream);
}

extern "C" int mvc_debug_set_attn_prof(unsigned long long* dev_buf) {
  mvc::set_attn_prof(dev_buf);
  return 0;
}
