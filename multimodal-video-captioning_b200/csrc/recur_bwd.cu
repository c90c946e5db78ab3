// recur_bwd.cu -- BPTT through the SA-LSTM decoder recurrence as ONE persistent, cluster-cooperative kernel
// (the mirror image of recur_fwd.cu; autograd of features_captioning.py:77-119).
//
//   grid = 32 clusters x 4 CTAs (H = 512).  CTA (cluster c, rank r) keeps for the whole kernel
//     shared memory : wcat^T[output columns [n0, n0+96) of d[ctx;h], gate-K-slice r (4H/4 columns)]   (UMMA B operand)
//                     attention.W^T[units [r*H/4, +H/4), :]                                           (dh += dwq . W)
//     tensor memory : the keys of batch row b = 4c + r  (416 columns)  +  the 96-column accumulator
//     registers     : U.k rows of row b, the running d(U.k) of row b, d(w) partials, dc of "its" (row, unit) pairs
//   per step s = S-1 .. 0:
//     P2  d[ctx;h]_s tile = dG_s[128, K-slice] . wcat^T slice   (tcgen05, A = bf16 gate gradients streamed by TMA);
//         partial parked in smem, cluster barrier, rank r sums rows [32r,32r+32) through DSMEM -> dxh (fp32)
//     grid barrier
//     P3  attention backward of row b: dalpha_t = dctx . key_t out of TMEM, softmax Jacobian, dpre = de w (1 - tanh^2)
//         (tanh recomputed from the saved query + register-resident U.k), dwq_s, d(U.k) += , d(w) +=
//     P4  cluster exchange of the four rows' dwq (DSMEM), dh_s = dxh[:, F:] + dwq . W  (mma.sync, W^T slice resident)
//     P1' LSTM cell backward of step s-1 for the (row, unit) pairs this thread owns: dG_{s-1} (fp32 + bf16)
//     grid barrier
//   d(U.k) and d(w) leave the registers once, after the last step.  All sums have a fixed order: deterministic.
#include <cuda.h>

#include <mutex>

#include "ptx.cuh"
#include "step.cuh"
#include "recur.cuh"

namespace mvc {

constexpr int RB_THREADS = 320;      // warps 0-7 compute, warp 8 TMA producer, warp 9 MMA issuer
constexpr int RB_CS = 4;
constexpr int RB_BN = 96;            // UMMA N (>= output columns per cluster, multiple of 16)
constexpr int RB_STAGES = 3;
constexpr int RB_STAGE_BYTES = 128 * 64 * 2;
constexpr int RB_MAXKB = 8;          // gate K blocks per rank (4H / 4 / 64): H <= 512
constexpr int RB_KB_BYTES = RB_BN * 64 * 2;                  // 12288
constexpr int RB_B_BYTES = RB_MAXKB * RB_KB_BYTES;           // 98304
constexpr int RB_RING_BYTES = RB_STAGES * RB_STAGE_BYTES;    // 49152
constexpr int RB_PS = 88;            // partial tile row pitch (floats); 84 live columns
constexpr int RB_KEY_COLS = 416;     // TMEM columns [0, 416) keys, [416, 512) accumulator
constexpr int RB_ACC_COL = 416;
constexpr int RB_R = 6;              // frame rounds per warp (T <= 48)

__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cl_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cl_map(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 cl_ld4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void cl_st_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ uint64_t sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16b(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void gbar(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v < target && clock64() - t0 > 4000000000LL) {
        printf("mvc recur_bwd: grid barrier timed out (block %d, %u of %u)\n", blockIdx.x, v, target);
        __trap();
      }
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

template <int AV, int WPL>
__global__ void __launch_bounds__(RB_THREADS, 1)
recur_bwd_kernel(const __grid_constant__ CUtensorMap map_dg, const __grid_constant__ CUtensorMap map_wt,
                 const __grid_constant__ RecurBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int B = p.B, T = p.T, F = p.F, H = p.H, K = p.K, S = p.S;
  constexpr int A = AV * 32;
  constexpr int AP = A + 8;                      // pitch (bf16) of W^T slice rows and of the dwq exchange rows
  const int UPR = H / RB_CS;                     // hidden units per rank (dh, cell backward)
  const int ncol = ((K + 31) / 32 + 3) & ~3;     // output columns of d[ctx;h] per cluster (84 for K = 2688)

  uint8_t* ring = smem + RB_B_BYTES;                              // A stages | partial tile | per-warp dwq partials
  __nv_bfloat16* sWT = reinterpret_cast<__nv_bfloat16*>(ring + RB_RING_BYTES);      // [UPR][AP]  W^T slice
  __nv_bfloat16* sDq = sWT + (size_t)UPR * AP;                    // [4][AP]   dwq of the cluster's rows (bf16)
  float* sDa = reinterpret_cast<float*>(sDq + RB_CS * AP);        // [8][64]   per-warp dalpha partials
  float* sDe = sDa + 8 * 64;                                      // [64]      de
  float* sAl = sDe + 64;                                          // [64]      alpha of this step
  float* sWv = sAl + 64;                                          // [A]
  float* sBias = sWv + A;                                         // [A]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + A);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (RB_STAGES + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * RB_STAGES);
  const uint32_t w_bar = bar0 + 8u * (2 * RB_STAGES + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * RB_STAGES + 2);
  float* sAcc = reinterpret_cast<float*>(ring);                   // [8][A] (ring idle during P3)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cl_rank();
  const int cl = blockIdx.x / RB_CS;
  const int brow = blockIdx.x;
  const bool has_row = brow < B;
  const int n0 = cl * ncol;                      // first d[ctx;h] column of this cluster
  const int nkb_all = (4 * H) / 64;
  const int kb0 = nkb_all * rank / RB_CS, kb1 = nkb_all * (rank + 1) / RB_CS;
  const int nkb = kb1 - kb0;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring_base = smem_base + RB_B_BYTES;

  // ---------------------------------------------------------------- setup
  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dg) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wt) : "memory");
    for (int s = 0; s < RB_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t tmem_acc = tmem_lane + RB_ACC_COL;

  if (warp == 8 && lane == 0) {
    mbar_expect_tx(w_bar, (uint32_t)nkb * RB_KB_BYTES);
    for (int i = 0; i < nkb; ++i) tma_load_2d(smem_base + i * RB_KB_BYTES, &map_wt, w_bar, (kb0 + i) * 64, n0);
  }
  float ur[RB_R][AV], dur[RB_R][AV], dwr[AV];
  if (warp < 8) {
    const int vec_per_row = A / 8;
    for (int i = tid; i < UPR * vec_per_row; i += 256) {
      const int u = i / vec_per_row, k8 = i - u * vec_per_row;
      const uint4 v = *reinterpret_cast<const uint4*>(p.attWT + (size_t)(rank * UPR + u) * A + k8 * 8);
      *reinterpret_cast<uint4*>(sWT + (size_t)u * AP + k8 * 8) = v;
    }
    for (int i = tid; i < A; i += 256) { sWv[i] = p.att_w[i]; sBias[i] = p.att_b[i]; }
    if (tid < 64) { sDe[tid] = 0.f; sAl[tid] = 0.f; }
    const float* ukb = p.uk + (size_t)(has_row ? brow : 0) * T * A;
#pragma unroll
    for (int r = 0; r < RB_R; ++r) {
      const int t = warp + r * 8;
#pragma unroll
      for (int k = 0; k < AV; ++k) {
        ur[r][k] = (t < T) ? __ldg(ukb + (size_t)t * A + lane + 32 * k) : 0.f;
        dur[r][k] = 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < AV; ++k) dwr[k] = 0.f;
  }
  if (warp < 4) {
    // keys of my batch row -> TMEM columns [0, T*WPL): lane L keeps features [2*WPL*L, +2*WPL) of every frame
    const int L = tid;
    const uint32_t* krow = reinterpret_cast<const uint32_t*>(p.feats + (size_t)(has_row ? brow : 0) * T * F);
    const int fw = F / 2;
#pragma unroll
    for (int c = 0; c < RB_KEY_COLS / 32; ++c) {
      if (c * 32 < T * WPL) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int w = c * 32 + j;
          const int t = w / WPL, k = w % WPL;
          const int word = L * WPL + k;
          v[j] = (has_row && t < T && word < fw) ? __ldg(krow + (size_t)t * fw + word) : 0u;
        }
        tmem_st32(tmem_lane + (uint32_t)(c * 32), v);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  // (row, unit) pairs owned by this thread for dh / cell backward: the C fragment layout of mma.sync m16n8k16 --
  // warp w holds n-tiles {2w, 2w+1} (8 units each) of the rank's UPR units, lanes 0..15 hold rows 0..3
  const int r4 = lane >> 2, kq = (lane & 3) * 2;
  const bool live = warp < 8 && r4 < RB_CS && (2 * warp) * 8 < UPR;
  const int prow = cl * RB_CS + (r4 & 3);        // batch row of my pairs
  const bool prow_ok = live && prow < B;
  float dcr[2][2] = {{0.f, 0.f}, {0.f, 0.f}};   // dc of my 4 pairs, carried across steps
  float dhc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};   // dh carried into the cell backward

  unsigned gb = 0;
  uint32_t ring_it = 0;
  const unsigned nctas = gridDim.x;
  long long* prof = (p.prof && blockIdx.x == 0 && tid == 0) ? p.prof : nullptr;
#define RB_STAMP(i) do { if (prof) prof[step_i * 10 + (i)] = clock64(); } while (0)

  // LSTM cell backward of step `s` for my pairs (elementwise; lstm_cell_bwd_kernel restated per pair)
  auto cell_backward = [&](int s) {
    if (!prow_ok) return;
    const size_t grow = (size_t)s * B + prow;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int nt = 2 * warp + i;
      if (nt * 8 >= UPR) continue;
      const int j = rank * UPR + nt * 8 + kq;    // units j, j+1
      const float* a = p.act + grow * (size_t)(4 * H);
      const int cb = (j >> 4) * 64 + (j & 15);
      const float2 ig = *reinterpret_cast<const float2*>(a + cb), fg = *reinterpret_cast<const float2*>(a + cb + 16);
      const float2 gg = *reinterpret_cast<const float2*>(a + cb + 32), og = *reinterpret_cast<const float2*>(a + cb + 48);
      const float2 cn = *reinterpret_cast<const float2*>(p.c + ((size_t)(s + 1) * B + prow) * H + j);
      const float2 cp = *reinterpret_cast<const float2*>(p.c + grow * H + j);
      float2 dhe = make_float2(0.f, 0.f);
      if (p.dh_ext) dhe = *reinterpret_cast<const float2*>(p.dh_ext + grow * H + j);
      const float igv[2] = {ig.x, ig.y}, fgv[2] = {fg.x, fg.y}, ggv[2] = {gg.x, gg.y}, ogv[2] = {og.x, og.y};
      const float cnv[2] = {cn.x, cn.y}, cpv[2] = {cp.x, cp.y}, dhv[2] = {dhe.x + dhc[i][0], dhe.y + dhc[i][1]};
      float di[2], df[2], dg[2], dO[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float tc = tanh_ex2(cnv[e]);
        const float dct = dcr[i][e] + dhv[e] * ogv[e] * (1.f - tc * tc);
        di[e] = dct * ggv[e] * igv[e] * (1.f - igv[e]);
        df[e] = dct * cpv[e] * fgv[e] * (1.f - fgv[e]);
        dg[e] = dct * igv[e] * (1.f - ggv[e] * ggv[e]);
        dO[e] = dhv[e] * tc * ogv[e] * (1.f - ogv[e]);
        dcr[i][e] = dct * fgv[e];
      }
      float* d = p.dG + grow * (size_t)(4 * H) + cb;
      *reinterpret_cast<float2*>(d) = make_float2(di[0], di[1]);
      *reinterpret_cast<float2*>(d + 16) = make_float2(df[0], df[1]);
      *reinterpret_cast<float2*>(d + 32) = make_float2(dg[0], dg[1]);
      *reinterpret_cast<float2*>(d + 48) = make_float2(dO[0], dO[1]);
      __nv_bfloat16* db = p.dG_b + grow * (size_t)(4 * H) + cb;
      *reinterpret_cast<__nv_bfloat162*>(db) = __floats2bfloat162_rn(di[0], di[1]);
      *reinterpret_cast<__nv_bfloat162*>(db + 16) = __floats2bfloat162_rn(df[0], df[1]);
      *reinterpret_cast<__nv_bfloat162*>(db + 32) = __floats2bfloat162_rn(dg[0], dg[1]);
      *reinterpret_cast<__nv_bfloat162*>(db + 48) = __floats2bfloat162_rn(dO[0], dO[1]);
    }
  };

  cell_backward(S - 1);                          // dh carried into the last step is zero
  gbar(p.sync, (++gb) * nctas);

  for (int s = S - 1; s >= 0; --s) {
    const int step_i = S - 1 - s;
    RB_STAMP(0);
    // ================================================================= P2: d[ctx;h]_s = dG_s . wcat
    const int rl = (tid & 255) >> 3, q8 = tid & 7;     // reducer: 8 warps, thread = (row, every 8th float4)
    const int row = rank * 32 + rl;
    if (warp == 8) {
      if (lane == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");
        uint32_t it = ring_it;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int stage = (int)(it % RB_STAGES);
          const uint32_t par = (it / RB_STAGES) & 1u;
          mbar_wait(empty_bar(stage), par ^ 1u);
          mbar_expect_tx(full_bar(stage), RB_STAGE_BYTES);
          tma_load_2d(ring_base + stage * RB_STAGE_BYTES, &map_dg, full_bar(stage), (kb0 + i) * 64, s * B);
        }
      }
      __syncwarp();
    } else if (warp == 9) {
      if (lane == 0) {
        constexpr uint32_t idesc = idesc_bf16b(128, RB_BN);
        if (step_i == 0) mbar_wait(w_bar, 0);
        uint32_t it = ring_it;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int stage = (int)(it % RB_STAGES);
          const uint32_t par = (it / RB_STAGES) & 1u;
          mbar_wait(full_bar(stage), par);
          tc_fence_after();
          const uint64_t adesc = sw128(ring_base + stage * RB_STAGE_BYTES);
          const uint64_t bdesc = sw128(smem_base + i * RB_KB_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + RB_ACC_COL, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (i | k) ? 1u : 0u);
          umma_commit(empty_bar(stage));
        }
        umma_commit(tmem_full_bar);
      }
      __syncwarp();
    } else if (warp < 4) {
      mbar_wait(tmem_full_bar, (uint32_t)(step_i & 1));
      tc_fence_after();
      const int pr = warp * 32 + lane;
      float* part = reinterpret_cast<float*>(ring);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_acc + (uint32_t)(c * 32), v);
        float4* dst = reinterpret_cast<float4*>(part + (size_t)pr * RB_PS + c * 32);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (c * 32 + j < RB_PS)
            dst[j >> 2] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                      __uint_as_float(v[j + 3]));
      }
      tc_fence_before();
    }
    ring_it += (uint32_t)nkb;
    RB_STAMP(1);
    cl_arrive();
    cl_wait();
    RB_STAMP(2);
    if (warp < 8 && row < B) {
      const uint32_t pbase = ring_base + (uint32_t)(row * RB_PS) * 4u;
      float* out = p.dxh + (size_t)row * K + n0;
      for (int v4 = q8; v4 * 4 < ncol; v4 += 8) {
        if (n0 + v4 * 4 >= K) break;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int sr = 0; sr < RB_CS; ++sr) {
          const float4 t4 = cl_ld4(cl_map(pbase + (uint32_t)(v4 * 16), (uint32_t)sr));
          acc.x += t4.x; acc.y += t4.y; acc.z += t4.z; acc.w += t4.w;
        }
        *reinterpret_cast<float4*>(out + v4 * 4) = acc;
      }
    }
    if (warp < 8) asm volatile("fence.proxy.async;" ::: "memory");
    RB_STAMP(3);
    gbar(p.sync, (++gb) * nctas);                 // dxh of every row / column is complete
    RB_STAMP(4);

    // ================================================================= P3: attention backward of my row
    if (warp < 8) {
      const int half = warp >> 2, L = tid & 127;
      if (has_row) {
        // dctx slice of lane L: features [2*WPL*L, +2*WPL)
        float dcx[2 * WPL];
        const float* dcrow = p.dxh + (size_t)brow * K;
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
          const int f = (L * WPL + k) * 2;
          float2 v2 = make_float2(0.f, 0.f);
          if (f < F) v2 = __ldcg(reinterpret_cast<const float2*>(dcrow + f));
          dcx[2 * k] = v2.x; dcx[2 * k + 1] = v2.y;
        }
        if (lane < 2) {                             // zero this warp's dalpha slots; fetch alpha
#pragma unroll
          for (int i = 0; i < 32; ++i) sDa[warp * 64 + lane * 32 + i] = 0.f;
        }
        if (warp == 0) {
          const float* al = p.alpha + ((size_t)s * B + brow) * T;
          if (lane < T) sAl[lane] = al[lane];
          if (lane + 32 < T) sAl[lane + 32] = al[lane + 32];
        }
        __syncwarp();
        // dalpha_t = dctx . key_t : per-lane partial over its feature slice, warp-reduced per frame
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < RB_KEY_COLS / 32; ++c) {
          if ((c & 1) == half && c * 32 < T * WPL) {
            uint32_t v[32];
            tmem_ld32(tmem_lane + (uint32_t)(c * 32), v);
            const int t_lo = (c * 32) / WPL, t_hi = (c * 32 + 31) / WPL;     // compile-time after unrolling
#pragma unroll
            for (int tt = 0; tt < 32 / WPL + 2; ++tt) {
              const int t = t_lo + tt;
              if (t <= t_hi && t < 64) {
                float pt = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int w = c * 32 + j;
                  if (w / WPL == t) {
                    const int k = w % WPL;
                    const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[j]));
                    pt = fmaf(dcx[2 * k], x.x, pt);
                    pt = fmaf(dcx[2 * k + 1], x.y, pt);
                  }
                }
                pt = warp_sum(pt);
                if (lane == 0) sDa[warp * 64 + t] += pt;
              }
            }
          }
        }
        tc_fence_before();
      }
      cbar();
      if (has_row && warp == 0) {
        // softmax Jacobian: de_t = alpha_t (dalpha_t - sum alpha dalpha)
        float da0 = 0.f, da1 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { da0 += sDa[w * 64 + lane]; da1 += sDa[w * 64 + lane + 32]; }
        const float a0 = sAl[lane], a1 = sAl[lane + 32];            // zero beyond T
        const float dot = warp_sum(a0 * da0 + a1 * da1);
        sDe[lane] = a0 * (da0 - dot);
        sDe[lane + 32] = a1 * (da1 - dot);
      }
      cbar();
      if (has_row) {
        const float* wqr = p.wq + ((size_t)s * B + brow) * A;
        float qb[AV], wv[AV], sq[AV];
#pragma unroll
        for (int k = 0; k < AV; ++k) { qb[k] = wqr[lane + 32 * k] + sBias[lane + 32 * k]; wv[k] = sWv[lane + 32 * k]; sq[k] = 0.f; }
#pragma unroll
        for (int r = 0; r < RB_R; ++r) {
          const int t = warp + r * 8;
          if (t < T) {
            const float de = sDe[t];
#pragma unroll
            for (int k = 0; k < AV; ++k) {
              const float th = tanh_fast(qb[k] + ur[r][k]);
              const float dpre = de * wv[k] * (1.f - th * th);
              sq[k] += dpre;
              dwr[k] = fmaf(de, th, dwr[k]);
              dur[r][k] += dpre;
            }
          }
        }
#pragma unroll
        for (int k = 0; k < AV; ++k) sAcc[warp * A + lane + 32 * k] = sq[k];
      }
      cbar();
      // dwq of my row: fixed-order sum over the 8 warps; fp32 + bf16 to global, bf16 into every rank's exchange row
      if (tid < A) {
        float qsum = 0.f;
        if (has_row) {
#pragma unroll
          for (int w = 0; w < 8; ++w) qsum += sAcc[w * A + tid];
          p.dwq[((size_t)s * B + brow) * A + tid] = qsum;
          p.dwq_b[((size_t)s * B + brow) * A + tid] = __float2bfloat16(qsum);
        }
        const float other = __shfl_down_sync(0xffffffffu, qsum, 1);
        if ((tid & 1) == 0) {
          __nv_bfloat162 pk = __floats2bfloat162_rn(qsum, other);
          const uint32_t bits = *reinterpret_cast<uint32_t*>(&pk);
          const uint32_t local = smem_u32(sDq + (size_t)rank * AP + tid);
#pragma unroll
          for (int dst = 0; dst < RB_CS; ++dst) cl_st_u32(cl_map(local, (uint32_t)dst), bits);
        }
      }
    }
    RB_STAMP(5);
    cl_arrive();
    cl_wait();                                      // the four rows' dwq are in every rank's sDq
    RB_STAMP(6);
    if (s > 0) {
      // ============================================================= P4: dh_s = dxh[:, F:] + dwq . W ; P1': cell backward of s-1
      if (warp < 8) {                               // mma.sync is warp-collective: every lane takes part
        const bool arow_live = r4 < RB_CS;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int nt = 2 * warp + i;
          if (nt * 8 < UPR) {                        // warp-uniform
            const __nv_bfloat16* arow = sDq + (size_t)(r4 & 3) * AP + kq;
            const __nv_bfloat16* brow2 = sWT + (size_t)(nt * 8 + r4) * AP + kq;
            float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
            for (int k0 = 0; k0 < A; k0 += 16) {
              const uint32_t a0 = arow_live ? *reinterpret_cast<const uint32_t*>(arow + k0) : 0u;
              const uint32_t a2 = arow_live ? *reinterpret_cast<const uint32_t*>(arow + k0 + 8) : 0u;
              const uint32_t b0 = *reinterpret_cast<const uint32_t*>(brow2 + k0);
              const uint32_t b1 = *reinterpret_cast<const uint32_t*>(brow2 + k0 + 8);
              asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                           : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                           : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
            }
            float2 dx = make_float2(0.f, 0.f);
            if (prow_ok) dx = __ldcg(reinterpret_cast<const float2*>(p.dxh + (size_t)prow * K + F + rank * UPR + nt * 8 + kq));
            dhc[i][0] = c0 + dx.x;
            dhc[i][1] = c1 + dx.y;
          }
        }
      }
      cell_backward(s - 1);
    }
    RB_STAMP(7);
    gbar(p.sync, (++gb) * nctas);                 // dG_{s-1} complete; dxh / exchange rows free again
    RB_STAMP(8);
  }
#undef RB_STAMP

  // ---------------------------------------------------------------- d(U.k), d(w) leave the registers
  if (warp < 8 && has_row) {
    float* dukb = p.duk + (size_t)brow * T * A;
#pragma unroll
    for (int r = 0; r < RB_R; ++r) {
      const int t = warp + r * 8;
      if (t < T) {
#pragma unroll
        for (int k = 0; k < AV; ++k) dukb[(size_t)t * A + lane + 32 * k] = dur[r][k];
      }
    }
#pragma unroll
    for (int k = 0; k < AV; ++k) sAcc[warp * A + lane + 32 * k] = dwr[k];
  }
  __syncthreads();
  if (tid < A && has_row) {
    float wsum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) wsum += sAcc[w * A + tid];
    p.dwpart[(size_t)brow * A + tid] = wsum;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 encode_fn3() {
  static EncodeTiledFn3 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn3>(q);
  });
  return fn;
}
static int make_map3(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
  EncodeTiledFn3 enc = encode_fn3();
  MVC_CHECK(enc, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVC_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

static size_t recur_bwd_smem(int H, int A) {
  const size_t upr = (size_t)H / RB_CS;
  return 1024 + RB_B_BYTES + RB_RING_BYTES + (upr + RB_CS) * (A + 8) * 2 + sizeof(float) * (8 * 64 + 64 + 64 + 2 * (size_t)A) +
         8 * (2 * RB_STAGES + 2) + 16;
}
static inline int key_wpl_b(int F) { return (F + 255) / 256; }
static const void* recur_bwd_kernel_for(int A, int F) {
  if (A != 256) return nullptr;
  switch (key_wpl_b(F)) {
    case 9: return (const void*)recur_bwd_kernel<8, 9>;
    case 8: return (const void*)recur_bwd_kernel<8, 8>;
    case 1: return (const void*)recur_bwd_kernel<8, 1>;
    default: return nullptr;
  }
}

bool recur_bwd_supported(int B, int T, int F, int H, int A) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("MVC_B200_PERSISTENT");
    disabled = (e && e[0] == '0') ? 1 : 0;
  }
  if (disabled) return false;
  if (!(B >= 1 && B <= 128 && T >= 1 && T <= 8 * RB_R && F % 8 == 0 && H == 512)) return false;   // 32 clusters
  if (!recur_bwd_kernel_for(A, F)) return false;
  const int K = F + H;
  const int ncol = ((K + 31) / 32 + 3) & ~3;
  if (ncol > RB_PS || ncol > RB_BN) return false;
  if (T * key_wpl_b(F) > RB_KEY_COLS) return false;
  if ((4 * H) / 64 / RB_CS > RB_MAXKB) return false;
  if (recur_bwd_smem(H, A) > 227 * 1024) return false;
  static std::mutex mu;
  static int max_clusters[3] = {-1, -1, -1};
  const int ki = key_wpl_b(F) == 9 ? 0 : (key_wpl_b(F) == 8 ? 1 : 2);
  std::lock_guard<std::mutex> lk(mu);
  if (max_clusters[ki] < 0) {
    const void* kern = recur_bwd_kernel_for(A, F);
    const size_t smem = 227 * 1024;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      max_clusters[ki] = 0;
    } else {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(RB_CS);
      cfg.blockDim = dim3(RB_THREADS);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = RB_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
      max_clusters[ki] = n;
    }
  }
  return max_clusters[ki] >= 32;
}

static long long* g_recur_bwd_prof = nullptr;

int recur_bwd_launch(const RecurBwdParams& p, const void* wcatT, cudaStream_t st) {
  const void* kern = recur_bwd_kernel_for(p.A, p.F);
  MVC_CHECK(kern && recur_bwd_supported(p.B, p.T, p.F, p.H, p.A), "persistent backward recurrence: unsupported dims");
  CUtensorMap mg, mw;
  MVC_TRY(make_map3(p.dG_b, (int64_t)p.S * p.B, 4 * (int64_t)p.H, 4 * (int64_t)p.H, 128, &mg));
  MVC_TRY(make_map3(wcatT, p.K, 4 * (int64_t)p.H, 4 * (int64_t)p.H, RB_BN, &mw));
  MVC_CUDA(cudaMemsetAsync(p.sync, 0, sizeof(unsigned), st));
  const size_t smem = recur_bwd_smem(p.H, p.A);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(32 * RB_CS);
  cfg.blockDim = dim3(RB_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = RB_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RecurBwdParams pp = p;
  pp.prof = g_recur_bwd_prof;
  void* args[] = {(void*)&mg, (void*)&mw, (void*)&pp};
  ProfScope prof(PK_STEP_FUSED, p.B, -p.S, p.K, st);
  MVC_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
  MVC_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvc

namespace mvc { void r2_set_bwd_prof(long long* p); }
extern "C" int mvc_debug_set_recur_bwd_prof(long long* dev_buf) {
  mvc::g_recur_bwd_prof = dev_buf;
  mvc::r2_set_bwd_prof(dev_buf);
  return 0;
}
