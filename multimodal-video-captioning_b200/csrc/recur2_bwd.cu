// recur2_bwd.cu -- BPTT through the SA-LSTM recurrence as ONE persistent kernel, mirror image of recur2_fwd.cu
// (autograd of features_captioning.py:77-119; design notes: recur2.cuh).
//
//   grid = 32 clusters x 4 CTAs.  CTA (cluster c, rank r):
//     row owner of batch row b = 4c + r : LSTM cell backward of the whole row (dc in registers), attention backward
//         with dalpha_t = dG_s[b,:] . P[b,t,:] out of TMEM (P[b] resident, as in the forward), softmax Jacobian,
//         dpre = de w (1 - tanh^2) (tanh recomputed from the saved query + register-resident U.k), dwq_s,
//         d(U.k) and d(w) accumulated in registers over all steps;
//     query-projection backward         : attention.W^T units [128r, 128r+128) in shared memory:
//         dh_att[4 rows, my units] = dwq . W (mma.sync), scattered to the row owners over DSMEM;
//     recurrent GEMM                    : W_hh^T rows [16c, 16c+16) x gate-K-slice r resident in shared memory;
//         ghb[128, 16] partial = dG_s[:, slice] . W_hh on tcgen05 (A = bf16 gate gradients by TMA); each K-slice partial
//         goes straight from TMEM to its own plane of global `ghb` [4][128][H], summed by the row owner in rank order.
//   step s = S-1 .. 0:   compute warps 0-7                         |  GEMM warps 8-11
//     wait X (ghb of step s+1) and dh_att(s+1) (mbarrier)          |  wait Y >= rows*(S-s)   (dG_s of every row)
//     dh_{s+1} = dh_ext + ghb + dh_att; cell backward -> dG_s; Y++ |  TMA dG_s k-slices -> tcgen05.mma -> TMEM
//     dalpha (TMEM), de, dpre, dwq_s -> cluster (DSMEM, mbarrier)  |  -> ghb plane r; X++
//     dh_att slices -> owners (DSMEM, mbarrier)                    |
//   All sums have a fixed order: deterministic.
#include <cuda_fp16.h>

#include <mutex>

#include "recur2.cuh"
#include "step.cuh"

namespace mvc {
using namespace r2;

constexpr int B2_NKB = 4 * R2_H / 64 / R2_CS;        // 8 gate k-blocks per rank
constexpr int B2_BN = 16;                            // h columns per cluster
constexpr int B2_KB_BYTES = B2_BN * 128;             // 2048
constexpr int B2_B_BYTES = B2_NKB * B2_KB_BYTES;     // 16384
constexpr int B2_STAGE_BYTES = 128 * 128;
constexpr int B2_STAGES = 6;
constexpr int B2_RING_BYTES = B2_STAGES * B2_STAGE_BYTES;
constexpr int B2_AP = R2_A + 8;                      // bf16 pitch of W^T slice rows / dwq exchange rows
constexpr int B2_UPR = R2_H / R2_CS;                 // 128 hidden units per rank

__global__ void __launch_bounds__(R2_THREADS, 1)
recur2_bwd_kernel(const __grid_constant__ CUtensorMap map_dg, const __grid_constant__ CUtensorMap map_wt,
                  const __grid_constant__ Recur2BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (not by integer round-trip): the pointer keeps its shared address space, so every
  // access below compiles to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int H = R2_H, A = R2_A, AV = R2_AV, AP = B2_AP, UPR = B2_UPR;
  const int B = p.B, T = p.T, S = p.S;

  uint8_t* ring = smem + B2_B_BYTES;
  __nv_bfloat16* sWT = reinterpret_cast<__nv_bfloat16*>(ring + B2_RING_BYTES);            // [UPR][AP]  W^T slice
  __nv_bfloat16* sDq = sWT + (size_t)UPR * AP;                                            // [4][AP] dwq of the cluster's rows
  float* sDh = reinterpret_cast<float*>(sDq + R2_CS * AP);                                // [H] dh_att of my row
  float* sDg = sDh + H;                                                                   // [4H] dG_s of my row
  float* sAcc = sDg + 4 * H;                                                              // [8][A] per-warp dwq partials
  float* sDa = sAcc + 8 * A;                                                              // [8][64] per-warp dalpha partials
  float* sDe = sDa + 8 * 64;                                                              // [64]
  float* sAl = sDe + 64;                                                                  // [64]
  float* sWv = sAl + 64;                                                                  // [A]
  float* sBias = sWv + A;                                                                 // [A]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + A);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (B2_STAGES + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * B2_STAGES);
  const uint32_t w_bar = tmem_full_bar + 8u;
  const uint32_t dq_full = w_bar + 8u;          // 32 arrivals (8 warps x 4 owners): the four rows' dwq_s are in my exchange buffer
  const uint32_t dh_full = dq_full + 8u;        // 32 arrivals (8 warps x 4 ranks): my row's dh_att is complete in sDh
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * B2_STAGES + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_rank();
  const int cl = blockIdx.x / R2_CS;
  const int brow = blockIdx.x;
  const bool has_row = brow < B;
  const int n0 = cl * B2_BN;
  const int kb0 = rank * B2_NKB;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring_base = smem_base + B2_B_BYTES;
  const unsigned nctas = gridDim.x;
  const unsigned nlive = (unsigned)(B < (int)gridDim.x ? B : (int)gridDim.x);
  unsigned* cntY = p.sync;                        // Y[r] at cntY + 32 r: arrivals of the row owners of cluster rank r
  unsigned* cntX = p.sync + 128;                  // X[r] at cntX + 32 r: arrivals of the GEMM groups of cluster rank r
  // live rows per cluster rank (row b is owned by CTA b: rank b % 4), CTAs per rank
  const unsigned nl0 = (unsigned)((nlive + 3) / 4), nl1 = (unsigned)((nlive + 2) / 4), nl2 = (unsigned)((nlive + 1) / 4),
                 nl3 = (unsigned)(nlive / 4), ncr = nctas / R2_CS;

  // ---------------------------------------------------------------- setup
  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dg) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wt) : "memory");
    for (int s = 0; s < B2_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(w_bar, 1);
    mbar_init(dq_full, R2_CS * 8);
    mbar_init(dh_full, R2_CS * 8);
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
  const uint32_t tmem_p = tmem_lane + R2_PCOL;

  if (warp == 8 && lane == 0) {
    mbar_expect_tx(w_bar, (uint32_t)B2_B_BYTES);
    for (int i = 0; i < B2_NKB; ++i) tma_load_2d(smem_base + i * B2_KB_BYTES, &map_wt, w_bar, (kb0 + i) * 64, n0);
  }
  float ur[R2_R][AV], dur[R2_R][AV], dwr[AV];
  if (warp < 8) {
    const int vec_per_row = A / 8;
    for (int i = tid; i < UPR * vec_per_row; i += 256) {
      const int u = i / vec_per_row, k8 = i - u * vec_per_row;
      const uint4 v = *reinterpret_cast<const uint4*>(p.attWT + (size_t)(rank * UPR + u) * A + k8 * 8);
      *reinterpret_cast<uint4*>(sWT + (size_t)u * AP + k8 * 8) = v;
    }
    for (int i = tid; i < A; i += 256) { sWv[i] = p.att_w[i]; sBias[i] = p.att_b[i]; }
    if (tid < 64) { sDe[tid] = 0.f; sAl[tid] = 0.f; }
    for (int i = tid; i < 8 * 64; i += 256) sDa[i] = 0.f;
    const float* ukb = p.uk + (size_t)(has_row ? brow : 0) * T * A;
#pragma unroll
    for (int r = 0; r < R2_R; ++r) {
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
    const uint4* prow = reinterpret_cast<const uint4*>(p.P + (size_t)(has_row ? brow : 0) * T * (4 * H)) + 2 * tid;
#pragma unroll
    for (int c = 0; c < R2_R * 2; ++c) {
      uint32_t v[32];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int t = c * 4 + f;
        uint4 a = make_uint4(0u, 0u, 0u, 0u), b2 = a;
        if (has_row && t < T) {
          a = __ldg(prow + (size_t)t * (4 * H / 8));
          b2 = __ldg(prow + (size_t)t * (4 * H / 8) + 1);
        }
        v[f * 8 + 0] = a.x; v[f * 8 + 1] = a.y; v[f * 8 + 2] = a.z; v[f * 8 + 3] = a.w;
        v[f * 8 + 4] = b2.x; v[f * 8 + 5] = b2.y; v[f * 8 + 6] = b2.z; v[f * 8 + 7] = b2.w;
      }
      tmem_st32(tmem_p + (uint32_t)(c * 32), v);
    }
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();

  long long* prof = (p.prof && blockIdx.x == 0) ? p.prof : nullptr;
#define B2_STAMP(i) do { if (prof && tid == 0) prof[j * 10 + (i)] = clock64(); } while (0)
#define B2_STAMP_G(i) do { if (prof && tid == 256) prof[j * 10 + (i)] = clock64(); } while (0)

  if (warp < 8) {
    // ================================================================= row owner: cell backward + attention backward
    float dcr[4] = {0.f, 0.f, 0.f, 0.f};         // dc of units 4*tid .. 4*tid+3 (threads < 128), carried across steps
    const int half = warp >> 2;
    const int L = tid & 127;                      // TMEM lane / gate-column group of this thread
    for (int j = 0; j < S; ++j) {
      const int s = S - 1 - j;
      const size_t grow = (size_t)s * B + brow;
      B2_STAMP(0);
      // prefetch what does not depend on the previous step
      float qb[AV];
      if (has_row) {
        const float* wqr = p.wq + grow * A;
#pragma unroll
        for (int k = 0; k < AV; ++k) qb[k] = __ldg(wqr + lane + 32 * k) + sBias[lane + 32 * k];
        if (warp == 1) {
          const float* al = p.alpha + grow * T;
          sAl[lane] = lane < T ? __ldg(al + lane) : 0.f;
          sAl[lane + 32] = lane + 32 < T ? __ldg(al + lane + 32) : 0.f;
        }
      }
      if (tid < 128) {
        float4 a4[4], cn4, cp4, dhe4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_row) {
          const float4* ar = reinterpret_cast<const float4*>(p.act + grow * (size_t)(4 * H)) + 4 * tid;
#pragma unroll
          for (int q = 0; q < 4; ++q) a4[q] = __ldcs(ar + q);
          cn4 = __ldcs(reinterpret_cast<const float4*>(p.c + ((size_t)(s + 1) * B + brow) * H) + tid);
          cp4 = __ldcs(reinterpret_cast<const float4*>(p.c + grow * H) + tid);
          if (p.dh_ext) dhe4 = __ldcs(reinterpret_cast<const float4*>(p.dh_ext + grow * H) + tid);
        }
        float dh[4] = {dhe4.x, dhe4.y, dhe4.z, dhe4.w};
        if (j > 0) {
          // dh_{s+1} += dG_{s+1} . W_hh (global, all clusters) + dwq_{s+1} . W (my cluster)
          if (tid == 0) poll_counters4(cntX, ncr * j, ncr * j, ncr * j, ncr * j);
          named_bar<2, 128>();
          float4 g4[R2_CS];
          if (has_row) {
#pragma unroll
            for (int r = 0; r < R2_CS; ++r)
              g4[r] = __ldcg(reinterpret_cast<const float4*>(p.ghb + ((size_t)r * 128 + brow) * H) + tid);
          }
          mbar_wait_cl(dh_full, (uint32_t)((j - 1) & 1));
          B2_STAMP(1);
          if (has_row) {
            const float4 d4 = *reinterpret_cast<const float4*>(sDh + 4 * tid);
            float4 t4 = g4[0];
#pragma unroll
            for (int r = 1; r < R2_CS; ++r) { t4.x += g4[r].x; t4.y += g4[r].y; t4.z += g4[r].z; t4.w += g4[r].w; }
            dh[0] += t4.x + d4.x; dh[1] += t4.y + d4.y; dh[2] += t4.z + d4.z; dh[3] += t4.w + d4.w;
          }
        }
        if (has_row) {
          const float cnv[4] = {cn4.x, cn4.y, cn4.z, cn4.w}, cpv[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
          float dg[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float ig = a4[u].x, fg = a4[u].y, gg = a4[u].z, og = a4[u].w;
            const float tc = tanh_ex2(cnv[u]);
            const float dct = dcr[u] + dh[u] * og * (1.f - tc * tc);
            dg[4 * u] = dct * gg * ig * (1.f - ig);
            dg[4 * u + 1] = dct * cpv[u] * fg * (1.f - fg);
            dg[4 * u + 2] = dct * ig * (1.f - gg * gg);
            dg[4 * u + 3] = dh[u] * tc * og * (1.f - og);
            dcr[u] = dct * fg;
          }
          // publish the bf16 row FIRST (A operand of the recurrent GEMM) and the fp32 copy for the attention backward
          float4* dsm = reinterpret_cast<float4*>(sDg) + tid;        // [q][lane group]: conflict-free 16-byte accesses
          uint32_t pk[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            dsm[q * 128] = make_float4(dg[4 * q], dg[4 * q + 1], dg[4 * q + 2], dg[4 * q + 3]);
            __nv_bfloat162 lo = __floats2bfloat162_rn(dg[4 * q], dg[4 * q + 1]), hi = __floats2bfloat162_rn(dg[4 * q + 2], dg[4 * q + 3]);
            pk[2 * q] = *reinterpret_cast<uint32_t*>(&lo);
            pk[2 * q + 1] = *reinterpret_cast<uint32_t*>(&hi);
          }
          uint4* d16 = reinterpret_cast<uint4*>(p.dG_b + grow * (size_t)(4 * H)) + 2 * tid;
          d16[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          d16[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          named_bar<2, 128>();
          if (tid == 0) signal_counter(cntY + R2_CNT_STRIDE * rank);     // dG_s of my row is in global memory
          float4* d32 = reinterpret_cast<float4*>(p.dG + grow * (size_t)(4 * H)) + 4 * tid;
#pragma unroll
          for (int q = 0; q < 4; ++q) d32[q] = make_float4(dg[4 * q], dg[4 * q + 1], dg[4 * q + 2], dg[4 * q + 3]);
        }
      }
      named_bar<1, 256>();                          // sDg / sAl visible to all eight warps
      B2_STAMP(2);
      // ---- dalpha_t = dG_s[b,:] . P[b,t,:]: per-lane partial over its 16 gate columns, 12 frames per thread
      // (even 4-frame chunks on warps 0-3, odd ones on warps 4-7), then a 16-value butterfly across the warp
      if (has_row) {
        float dg[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v4 = *(reinterpret_cast<const float4*>(sDg) + q * 128 + L);
          dg[4 * q] = v4.x; dg[4 * q + 1] = v4.y; dg[4 * q + 2] = v4.z; dg[4 * q + 3] = v4.w;
        }
        float pt[32];                               // 24 live: chunk slot cs = i / 4 (chunk 2 cs + half), frame i % 4
#pragma unroll
        for (int i = 0; i < 32; ++i) pt[i] = 0.f;
        tc_fence_after();
#pragma unroll
        for (int cs = 0; cs < R2_R; ++cs) {
          const int c = 2 * cs + half;
          if (c * 4 < T) {
            uint32_t v[32];
            tmem_ld32(tmem_p + (uint32_t)(c * 32), v);
#pragma unroll
            for (int f = 0; f < 4; ++f) {
              float a0 = 0.f, a1 = 0.f;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&v[f * 8 + k]));
                a0 = fmaf(dg[2 * k], x.x, a0);
                a1 = fmaf(dg[2 * k + 1], x.y, a1);
              }
              pt[cs * 4 + f] = a0 + a1;
            }
          }
        }
        tc_fence_before();
        // butterfly over the warp: after the five rounds lane l holds the warp total of value l
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) {
          const bool up = (lane & w) != 0;
#pragma unroll
          for (int i = 0; i < w; ++i) {
            const float keep = up ? pt[w + i] : pt[i], give = up ? pt[i] : pt[w + i];
            pt[i] = keep + __shfl_xor_sync(0xffffffffu, give, w);
          }
        }
        if (lane < 4 * R2_R) {
          const int t = (2 * (lane >> 2) + half) * 4 + (lane & 3);
          sDa[warp * 64 + t] = pt[0];
        }
      }
      named_bar<1, 256>();
      B2_STAMP(3);
      if (has_row) {
        // softmax Jacobian de_t = alpha_t (dalpha_t - sum alpha dalpha), redundantly in every warp (no barrier): lane l
        // keeps frames l and l + 32; frame t was covered by the warp group of its chunk parity ((t/4) & 1)
        float de2[2];
        {
          float da[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int t = lane + 32 * e;
            const int w0 = ((t >> 2) & 1) * 4;
            da[e] = sDa[(w0 + 0) * 64 + t] + sDa[(w0 + 1) * 64 + t] + sDa[(w0 + 2) * 64 + t] + sDa[(w0 + 3) * 64 + t];
          }
          const float a0 = sAl[lane], a1 = sAl[lane + 32];            // zero beyond T
          const float dot = warp_sum(a0 * da[0] + a1 * da[1]);
          de2[0] = a0 * (da[0] - dot);
          de2[1] = a1 * (da[1] - dot);
        }
        float wv[AV], sq[AV];
#pragma unroll
        for (int k = 0; k < AV; ++k) { wv[k] = sWv[lane + 32 * k]; sq[k] = 0.f; }
#pragma unroll
        for (int r = 0; r < R2_R; ++r) {
          const int t = warp + r * 8;               // warp-uniform; t < 48
          const float de = __shfl_sync(0xffffffffu, t < 32 ? de2[0] : de2[1], t & 31);     // zero beyond T
#pragma unroll
          for (int k = 0; k < AV; ++k) {
            const float th = tanh_fast(qb[k] + ur[r][k]);
            const float dpre = de * wv[k] * (1.f - th * th);
            sq[k] += dpre;
            dwr[k] = fmaf(de, th, dwr[k]);
            dur[r][k] += dpre;
          }
        }
#pragma unroll
        for (int k = 0; k < AV; ++k) sAcc[warp * A + lane + 32 * k] = sq[k];
      }
      named_bar<1, 256>();
      // dwq of my row: fixed-order sum over the 8 warps; bf16 into every rank's exchange row (critical path) first,
      // fp32 + bf16 copies for the weight-gradient GEMMs to global afterwards
      float qsum = 0.f;
      if (has_row) {
#pragma unroll
        for (int w = 0; w < 8; ++w) qsum += sAcc[w * A + tid];
      }
      if (s > 0) {
        const float other = __shfl_down_sync(0xffffffffu, qsum, 1);
        if ((tid & 1) == 0) {
          __nv_bfloat162 pk2 = __floats2bfloat162_rn(qsum, other);
          const uint32_t bits = *reinterpret_cast<uint32_t*>(&pk2);
          const uint32_t local = smem_u32(sDq + (size_t)rank * AP + tid);
#pragma unroll
          for (int d = 0; d < R2_CS; ++d) st_dsmem_u32(mapa(local, (uint32_t)d), bits);
        }
        __syncwarp();
        if (lane < R2_CS) mbar_arrive_remote(mapa(dq_full, (uint32_t)lane));
      }
      if (has_row) {
        p.dwq[grow * A + tid] = qsum;
        p.dwq_b[grow * A + tid] = __float2bfloat16(qsum);
      }
      B2_STAMP(4);
      if (s > 0) {
        mbar_wait_cl(dq_full, (uint32_t)(j & 1));   // the four rows' dwq_s are in sDq
        B2_STAMP(5);
        // dh_att[4 rows][my 128 units] = dwq . W : warp w owns n-tiles {2w, 2w+1} (8 units each); lanes 0..15 hold rows 0..3
        const int r4 = lane >> 2, kq = (lane & 3) * 2;
        const bool arow_live = r4 < R2_CS;
        // four independent accumulator chains (2 n-tiles x even / odd k-steps) keep the tensor pipe busy
        float acc4[2][2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc4[i][hf][e] = 0.f;
        const __nv_bfloat16* arow = sDq + (size_t)(r4 & 3) * AP + kq;
#pragma unroll
        for (int k0 = 0; k0 < A; k0 += 32) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int k = k0 + 16 * hf;
            const uint32_t a0 = arow_live ? *reinterpret_cast<const uint32_t*>(arow + k) : 0u;
            const uint32_t a2 = arow_live ? *reinterpret_cast<const uint32_t*>(arow + k + 8) : 0u;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const __nv_bfloat16* wrow = sWT + (size_t)((2 * warp + i) * 8 + r4) * AP + kq;
              const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wrow + k);
              const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wrow + k + 8);
              asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                           : "+f"(acc4[i][hf][0]), "+f"(acc4[i][hf][1]), "+f"(acc4[i][hf][2]), "+f"(acc4[i][hf][3])
                           : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int nt = 2 * warp + i;
          if (arow_live)
            st_dsmem_f32x2(mapa(smem_u32(sDh + rank * UPR + nt * 8 + kq), (uint32_t)r4), acc4[i][0][0] + acc4[i][1][0],
                           acc4[i][0][1] + acc4[i][1][1]);
        }
        __syncwarp();
        if (lane < R2_CS) mbar_arrive_remote(mapa(dh_full, (uint32_t)lane));
      }
      B2_STAMP(6);
    }
    // ---------------------------------------------------------------- d(U.k), d(w) leave the registers
    named_bar<1, 256>();
    if (has_row) {
      float* dukb = p.duk + (size_t)brow * T * A;
#pragma unroll
      for (int r = 0; r < R2_R; ++r) {
        const int t = warp + r * 8;
        if (t < T) {
#pragma unroll
          for (int k = 0; k < AV; ++k) dukb[(size_t)t * A + lane + 32 * k] = dur[r][k];
        }
      }
#pragma unroll
      for (int k = 0; k < AV; ++k) sAcc[warp * A + lane + 32 * k] = dwr[k];
    }
    named_bar<1, 256>();
    if (has_row) {
      float wsum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) wsum += sAcc[w * A + tid];
      p.dwpart[(size_t)brow * A + tid] = wsum;
    }
  } else {
    // ================================================================= recurrent GEMM group (warps 8-11)
    uint32_t it_p = 0, it_c = 0;
    for (int j = 0; j + 1 < S; ++j) {
      const int s = S - 1 - j;                      // ghb(s) = dG_s . W_hh, consumed by the cell backward of step s-1
      if (warp == 8) {
        if (lane == 0) {
          poll_counters4(cntY, nl0 * (j + 1), nl1 * (j + 1), nl2 * (j + 1), nl3 * (j + 1));   // dG_s of every row is published
          asm volatile("fence.proxy.async;" ::: "memory");
          for (int i = 0; i < B2_NKB; ++i, ++it_p) {
            const int stage = (int)(it_p % B2_STAGES);
            const uint32_t par = (it_p / B2_STAGES) & 1u;
            mbar_wait(empty_bar(stage), par ^ 1u);
            mbar_expect_tx(full_bar(stage), B2_STAGE_BYTES);
            tma_load_2d(ring_base + stage * B2_STAGE_BYTES, &map_dg, full_bar(stage), (kb0 + i) * 64, s * B);
          }
        }
        __syncwarp();
      } else if (warp == 9) {
        if (lane == 0) {
          constexpr uint32_t idesc = idesc_bf16(128, B2_BN);
          if (j == 0) mbar_wait(w_bar, 0);
          for (int i = 0; i < B2_NKB; ++i, ++it_c) {
            const int stage = (int)(it_c % B2_STAGES);
            const uint32_t par = (it_c / B2_STAGES) & 1u;
            mbar_wait(full_bar(stage), par);
            tc_fence_after();
            const uint64_t adesc = sw128_desc(ring_base + stage * B2_STAGE_BYTES);
            const uint64_t bdesc = sw128_desc(smem_base + i * B2_KB_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (i | k) ? 1u : 0u);
            umma_commit(empty_bar(stage));
          }
          umma_commit(tmem_full_bar);
        }
        __syncwarp();
      }
      B2_STAMP_G(7);
      mbar_wait(tmem_full_bar, (uint32_t)(j & 1));
      tc_fence_after();
      {
        // this K-slice's partial [128 x 16]: TMEM -> plane `rank` of ghb (row = TMEM lane)
        const int prow = (warp & 3) * 32 + lane;
        uint32_t v[32];
        tmem_ld32(tmem_lane, v);                    // 16 live accumulator columns
        if (prow < B) {
          float4* dst = reinterpret_cast<float4*>(p.ghb + ((size_t)rank * 128 + prow) * H + n0);
#pragma unroll
          for (int q = 0; q < B2_BN / 4; ++q)
            dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                 __uint_as_float(v[4 * q + 3]));
        }
      }
      tc_fence_before();
      B2_STAMP_G(8);
      named_bar<3, 128>();
      if (tid == 256) signal_counter(cntX + R2_CNT_STRIDE * rank);
      B2_STAMP_G(9);
    }
  }
#undef B2_STAMP
#undef B2_STAMP_G

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------ host
int r2_make_map(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out);

size_t recur2_bwd_smem() {
  return 1024 + B2_B_BYTES + B2_RING_BYTES + (size_t)(B2_UPR + R2_CS) * B2_AP * 2 +
         sizeof(float) * (R2_H + 4 * R2_H + 8 * R2_A + 8 * 64 + 64 + 64 + 2 * R2_A) + 8 * (2 * B2_STAGES + 4) + 16;
}
const void* recur2_bwd_kernel_ptr() { return (const void*)recur2_bwd_kernel; }

static long long* g_recur2_bwd_prof = nullptr;
void r2_set_bwd_prof(long long* p) { g_recur2_bwd_prof = p; }

int recur2_bwd_launch(const Recur2BwdParams& p, const void* whhT_um, cudaStream_t st, bool sync_cleared) {
  MVC_TRY(r2_apply_spin_limit());
  MVC_CHECK(recur2_supported(p.B, p.T, p.F, R2_H, R2_A), "persistent backward recurrence: unsupported dims");
  CUtensorMap mg, mw;
  MVC_TRY(r2_make_map(p.dG_b, (int64_t)p.S * p.B, 4 * (int64_t)R2_H, 4 * (int64_t)R2_H, 128, &mg));
  MVC_TRY(r2_make_map(whhT_um, R2_H, 4 * (int64_t)R2_H, 4 * (int64_t)R2_H, B2_BN, &mw));
  if (!sync_cleared) MVC_CUDA(cudaMemsetAsync(p.sync, 0, sizeof(unsigned) * 256, st));
  const size_t smem = recur2_bwd_smem();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(R2_H / 16) * R2_CS);
  cfg.blockDim = dim3(R2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = R2_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  Recur2BwdParams pp = p;
  pp.prof = g_recur2_bwd_prof;
  void* args[] = {(void*)&mg, (void*)&mw, (void*)&pp};
  ProfScope prof(PK_STEP_FUSED, p.B, -p.S, p.K, st);
  MVC_CUDA(cudaLaunchKernelExC(&cfg, (const void*)recur2_bwd_kernel, args));
  MVC_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvc
