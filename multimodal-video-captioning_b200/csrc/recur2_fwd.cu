// recur2_fwd.cu -- teacher-forced SA-LSTM time loop (features_captioning.py:77-119) as ONE persistent kernel with
// projected keys resident in tensor memory (design notes: recur2.cuh).
//
//   grid = 32 clusters x 4 CTAs = 128 CTAs, one per SM.  CTA (cluster c, rank r):
//     row owner of batch row b = 4c + r : attention + LSTM cell of that row, P[b] (T x 4H bf16) in TMEM columns
//                                         [64, 64 + 8T), U.k rows and the cell state c in registers;
//     query projection                  : attention.W units [64r, 64r+64) in shared memory, applied to the cluster's
//                                         four h rows (mma.sync), results scattered to the row owners over DSMEM;
//     recurrent GEMM                    : W_hh rows of units [16c, 16c+16) x K-slice r resident in shared memory as a
//                                         128B-swizzled UMMA operand; gh[128, 64] partial = h_s[:, slice] . W^T on
//                                         tcgen05 (A by TMA); each K-slice partial goes straight from TMEM to its own
//                                         plane of global `gh` [4][128][4H]; the row owner adds the four planes in
//                                         rank order (deterministic).
//   step s:   compute warps 0-7                          |  GEMM warps 8-11
//     wait h_s rows of the cluster (mbarrier)            |  wait Y >= rows*s   (h_s of every row published)
//     wq slice -> owners (DSMEM) -> mbarrier             |  TMA h_s k-slices -> tcgen05.mma -> TMEM -> gh plane r
//     scores, softmax, sum_t alpha_t P[b,t,:] (TMEM)     |  X += 1 per warp
//     wait X; gates = gx + sum_r gh[r] + P-sum; cell     |
//     h_{s+1}: global (bf16 + fp32), DSMEM to the four ranks, mbarrier arrive, Y += 1
#include <cuda_fp16.h>

#include <mutex>

#include "recur2.cuh"
#include "step.cuh"

namespace mvc {
using namespace r2;

constexpr int F2_NKB = R2_H / 64 / R2_CS;            // 2 k-blocks of h per rank
constexpr int F2_BN = 64;                            // gate columns per cluster (16 units x 4 gates)
constexpr int F2_KB_BYTES = F2_BN * 128;             // one resident W_hh k-block: 64 rows x 64 bf16
constexpr int F2_B_BYTES = F2_NKB * F2_KB_BYTES;     // 16384
constexpr int F2_STAGE_BYTES = 128 * 128;            // A stage: 128 rows x 64 bf16
constexpr int F2_STAGES = 2;
constexpr int F2_RING_BYTES = F2_STAGES * F2_STAGE_BYTES;
constexpr int F2_HP = R2_H + 8;                      // bf16 pitch of the resident attention.W slice / h rows
constexpr int F2_UPR = R2_A / R2_CS;                 // 64 query units per rank
constexpr int F2_GCP = 17;                           // pitch of the P-sum hand-over buffer (floats)

__global__ void __launch_bounds__(R2_THREADS, 1)
recur2_fwd_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ Recur2FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (not by integer round-trip): the pointer keeps its shared address space, so every
  // access below compiles to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int H = R2_H, A = R2_A, AV = R2_AV, HP = F2_HP, UPR = F2_UPR;
  const int B = p.B, T = p.T, K = p.K, S = p.S;

  uint8_t* ring = smem + F2_B_BYTES;
  __nv_bfloat16* sWatt = reinterpret_cast<__nv_bfloat16*>(ring + F2_RING_BYTES);          // [UPR][HP]
  __nv_bfloat16* sHb = sWatt + (size_t)UPR * HP;                                          // [4][HP] h rows of the cluster
  float* sQ = reinterpret_cast<float*>(sHb + (size_t)R2_CS * HP);                         // [A] wq of my row
  float* sWv = sQ + A;
  float* sBias = sWv + A;
  float* sE = sBias + A;                                                                  // [64] scores -> alpha
  float* sGc = sE + 64;                                                                   // [128][F2_GCP]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sGc + 128 * F2_GCP);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (F2_STAGES + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * F2_STAGES);
  const uint32_t w_bar = tmem_full_bar + 8u;
  const uint32_t hb_full = w_bar + 8u;          // 16 arrivals: 4 cell warps of each of the 4 row owners published h_s here
  const uint32_t q_full = hb_full + 8u;         // 32 arrivals: 8 warps of every rank delivered their slice of my row's query
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * F2_STAGES + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_rank();
  const int cl = blockIdx.x / R2_CS;
  const int brow = blockIdx.x;
  const bool has_row = brow < B;
  const int n0 = cl * F2_BN;
  const int kb0 = rank * F2_NKB;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring_base = smem_base + F2_B_BYTES;
  const unsigned nctas = gridDim.x;
  const unsigned nlive = (unsigned)(B < (int)gridDim.x ? B : (int)gridDim.x);
  unsigned* cntY = p.sync;                        // Y[r] at cntY + 32 r: arrivals of the row owners of cluster rank r
  unsigned* cntX = p.sync + 128;                  // X[r] at cntX + 32 r: arrivals of the GEMM groups of cluster rank r
  // live rows per cluster rank (row b is owned by CTA b: rank b % 4), CTAs per rank
  const unsigned nl0 = (unsigned)((nlive + 3) / 4), nl1 = (unsigned)((nlive + 2) / 4), nl2 = (unsigned)((nlive + 1) / 4),
                 nl3 = (unsigned)(nlive / 4), ncr = nctas / R2_CS;

  // ---------------------------------------------------------------- one-time setup
  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_h) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < F2_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(w_bar, 1);
    mbar_init(hb_full, R2_CS * 4);
    mbar_init(q_full, R2_CS * 8);
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

  if (warp == 8 && lane == 0) {                  // resident W_hh slice, once
    mbar_expect_tx(w_bar, (uint32_t)F2_B_BYTES);
    for (int i = 0; i < F2_NKB; ++i) tma_load_2d(smem_base + i * F2_KB_BYTES, &map_w, w_bar, (kb0 + i) * 64, n0);
  }
  float ur[R2_R][AV];                            // U.k rows of my batch row (loop invariant)
  if (warp < 8) {
    const int vec_per_row = H / 8;
    for (int i = tid; i < UPR * vec_per_row; i += 256) {
      const int u = i / vec_per_row, k8 = i - u * vec_per_row;
      const uint4 v = *reinterpret_cast<const uint4*>(p.attW + (size_t)(rank * UPR + u) * H + k8 * 8);
      *reinterpret_cast<uint4*>(sWatt + (size_t)u * HP + k8 * 8) = v;
    }
    for (int i = tid; i < A; i += 256) { sWv[i] = p.att_w[i]; sBias[i] = p.att_b[i]; sQ[i] = 0.f; }
    if (tid < 64) sE[tid] = 0.f;
    for (int i = tid; i < R2_CS * HP / 2; i += 256) reinterpret_cast<uint32_t*>(sHb)[i] = 0u;       // h_0 = 0
    const float* ukb = p.uk + (size_t)(has_row ? brow : 0) * T * A;
#pragma unroll
    for (int r = 0; r < R2_R; ++r) {
      const int t = warp + r * 8;
#pragma unroll
      for (int k = 0; k < AV; ++k) ur[r][k] = (t < T) ? __ldg(ukb + (size_t)t * A + lane + 32 * k) : 0.f;
    }
  }
  if (warp < 4) {
    // P[b] -> TMEM, resident for the whole kernel: lane L keeps, for frame t, the 8 words (gate columns
    // [16L, 16L+16) = i,f,g,o of units 4L..4L+3) at columns [64 + 8t, 64 + 8t + 8)
    const uint4* prow = reinterpret_cast<const uint4*>(p.P + (size_t)(has_row ? brow : 0) * T * (4 * H)) + 2 * tid;
#pragma unroll
    for (int c = 0; c < R2_R * 2; ++c) {         // 12 chunks of 32 columns = 4 frames each
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
  cluster_sync_all();                            // every CTA's mbarriers are initialised before any remote arrive

  long long* prof = (p.prof && blockIdx.x == 0) ? p.prof : nullptr;
#define F2_STAMP(i) do { if (prof && tid == 0) prof[s * 10 + (i)] = clock64(); } while (0)
#define F2_STAMP_G(i) do { if (prof && tid == 256) prof[s * 10 + (i)] = clock64(); } while (0)

  if (warp < 8) {
    // ================================================================= attention + cell (row owner)
    float cst[4] = {0.f, 0.f, 0.f, 0.f};         // cell state of units 4*tid .. 4*tid+3 (threads < 128)
    const int half = warp >> 2;
    for (int s = 0; s < S; ++s) {
      F2_STAMP(0);
      const size_t grow = (size_t)s * B + brow;
      float4 gx4[4];
      if (tid < 128 && has_row) {
        const float4* gxr = reinterpret_cast<const float4*>(p.gx + grow * (size_t)(4 * H)) + 4 * tid;
#pragma unroll
        for (int q = 0; q < 4; ++q) gx4[q] = __ldcs(gxr + q);
      }
      if (s > 0) {
        mbar_wait_cl(hb_full, (uint32_t)((s - 1) & 1));        // h_s of the cluster's four rows is in sHb
        F2_STAMP(1);
        // wq[4 rows][rank*UPR + 8*warp + 0..7] = h . W_slice^T (mma.sync m16n8k16, rows 4..15 of the A tile are zero)
        const int r4 = lane >> 2, kq = (lane & 3) * 2;
        const __nv_bfloat16* arow = sHb + (size_t)(r4 & 3) * HP + kq;
        const __nv_bfloat16* wrow = sWatt + (size_t)(warp * 8 + r4) * HP + kq;
        const bool live = r4 < R2_CS;
        float acc4[4][4];                           // four independent accumulator chains over k
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc4[i][e] = 0.f;
#pragma unroll
        for (int k0 = 0; k0 < H; k0 += 64) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k = k0 + 16 * i;
            const uint32_t a0 = live ? *reinterpret_cast<const uint32_t*>(arow + k) : 0u;
            const uint32_t a2 = live ? *reinterpret_cast<const uint32_t*>(arow + k + 8) : 0u;
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wrow + k);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wrow + k + 8);
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(acc4[i][0]), "+f"(acc4[i][1]), "+f"(acc4[i][2]), "+f"(acc4[i][3])
                         : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
          }
        }
        const float c0 = (acc4[0][0] + acc4[1][0]) + (acc4[2][0] + acc4[3][0]);
        const float c1 = (acc4[0][1] + acc4[1][1]) + (acc4[2][1] + acc4[3][1]);
        if (live) st_dsmem_f32x2(mapa(smem_u32(sQ + rank * UPR + warp * 8 + kq), (uint32_t)r4), c0, c1);
        __syncwarp();
        if (lane < R2_CS) mbar_arrive_remote(mapa(q_full, (uint32_t)lane));      // this warp's 8 units of all four rows
        mbar_wait_cl(q_full, (uint32_t)((s - 1) & 1));         // every rank's slice of my row's query has landed
      }
      F2_STAMP(2);
      if (has_row) {
        float* wq_out = p.wq_out + grow * A;
        for (int i = tid; i < A; i += 256) wq_out[i] = sQ[i];
        float qb[AV], wv[AV];
#pragma unroll
        for (int k = 0; k < AV; ++k) { qb[k] = sQ[lane + 32 * k] + sBias[lane + 32 * k]; wv[k] = sWv[lane + 32 * k]; }
        float er[8];                                 // per-lane partial scores of frames warp + 8r (r < 6)
#pragma unroll
        for (int r = 0; r < 8; ++r) er[r] = 0.f;
#pragma unroll
        for (int r = 0; r < R2_R; ++r) {
          float e0 = 0.f, e1 = 0.f;
#pragma unroll
          for (int k = 0; k < AV; k += 2) {
            e0 = fmaf(wv[k], tanh_fast(qb[k] + ur[r][k]), e0);
            e1 = fmaf(wv[k + 1], tanh_fast(qb[k + 1] + ur[r][k + 1]), e1);
          }
          er[r] = e0 + e1;
        }
        // 8-value butterfly: lane l ends with the warp total of value l >> 2 (9 shuffles instead of 30)
#pragma unroll
        for (int w = 4; w >= 1; w >>= 1) {
          const bool up = (lane & (4 * w)) != 0;
#pragma unroll
          for (int i = 0; i < w; ++i) {
            const float keep = up ? er[w + i] : er[i], give = up ? er[i] : er[w + i];
            er[i] = keep + __shfl_xor_sync(0xffffffffu, give, 4 * w);
          }
        }
        er[0] += __shfl_xor_sync(0xffffffffu, er[0], 2);
        er[0] += __shfl_xor_sync(0xffffffffu, er[0], 1);
        const int tr = warp + 8 * (lane >> 2);
        if ((lane & 3) == 0 && (lane >> 2) < R2_R && tr < T) sE[tr] = er[0];
      }
      named_bar<1, 256>();
      // softmax over T <= 48 frames, redundantly in every warp (no second barrier): lane l keeps alpha of frames l, l+32
      float al_lo = 0.f, al_hi = 0.f;
      if (has_row) {
        const float e0 = lane < T ? sE[lane] : -INFINITY, e1 = lane + 32 < T ? sE[lane + 32] : -INFINITY;
        const float mx = warp_max(fmaxf(e0, e1));
        const float p0 = lane < T ? __expf(e0 - mx) : 0.f, p1 = lane + 32 < T ? __expf(e1 - mx) : 0.f;
        const float inv = 1.f / warp_sum(p0 + p1);
        al_lo = p0 * inv; al_hi = p1 * inv;
        if (warp == 7) {
          float* al = p.alpha + grow * T;
          if (lane < T) al[lane] = al_lo;
          if (lane + 32 < T) al[lane + 32] = al_hi;
        }
      }
      F2_STAMP(3);
      // sum_t alpha_t P[b,t,:] straight out of TMEM: TMEM lane L owns gate columns [16L, 16L+16).  Warps w and w+4
      // share a lane quarter: even 4-frame chunks go to warps 0-3, odd ones to warps 4-7.
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      if (has_row) {
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < R2_R * 2; ++c) {
          if ((c & 1) == half && c * 4 < T) {
            uint32_t v[32];
            tmem_ld32(tmem_p + (uint32_t)(c * 32), v);
            float al[4];                          // alpha of frames 4c .. 4c+3 (zero beyond T)
#pragma unroll
            for (int f = 0; f < 4; ++f)
              al[f] = __shfl_sync(0xffffffffu, (c * 4 + f) < 32 ? al_lo : al_hi, (c * 4 + f) & 31);
#pragma unroll
            for (int f = 0; f < 4; ++f) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&v[f * 8 + k]));
                acc[2 * k] = fmaf(al[f], x.x, acc[2 * k]);
                acc[2 * k + 1] = fmaf(al[f], x.y, acc[2 * k + 1]);
              }
            }
          }
        }
        tc_fence_before();
        if (half == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) sGc[(tid - 128) * F2_GCP + i] = acc[i];
        }
      }
      named_bar<1, 256>();
      F2_STAMP(4);
      if (tid < 128) {
        // ============================================================ LSTM cell of my row: units 4*tid .. 4*tid+3
        if (s > 0) {
          if (tid == 0) poll_counters4(cntX, ncr * s, ncr * s, ncr * s, ncr * s);   // gh planes of h_s . W_hh^T are complete
          named_bar<2, 128>();
        }
        F2_STAMP(5);
        float pre[16], hn[4];
        uint32_t w0 = 0u, w1 = 0u;
        const size_t nrow = (size_t)(s + 1) * B + brow;
        if (has_row) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pre[i] = acc[i] + sGc[tid * F2_GCP + i];
#pragma unroll
          for (int q = 0; q < 4; ++q) { pre[4 * q] += gx4[q].x; pre[4 * q + 1] += gx4[q].y; pre[4 * q + 2] += gx4[q].z; pre[4 * q + 3] += gx4[q].w; }
          if (s > 0) {
            float4 g4[R2_CS][4];
#pragma unroll
            for (int r = 0; r < R2_CS; ++r) {
              const float4* ghr = reinterpret_cast<const float4*>(p.gh + ((size_t)r * 128 + brow) * (4 * H)) + 4 * tid;
#pragma unroll
              for (int q = 0; q < 4; ++q) g4[r][q] = __ldcg(ghr + q);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 t4 = g4[0][q];
#pragma unroll
              for (int r = 1; r < R2_CS; ++r) { t4.x += g4[r][q].x; t4.y += g4[r][q].y; t4.z += g4[r][q].z; t4.w += g4[r][q].w; }
              pre[4 * q] += t4.x; pre[4 * q + 1] += t4.y; pre[4 * q + 2] += t4.z; pre[4 * q + 3] += t4.w;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float ig = sigmoid_ex2(pre[4 * u]), fg = sigmoid_ex2(pre[4 * u + 1]), gg = tanh_ex2(pre[4 * u + 2]),
                        og = sigmoid_ex2(pre[4 * u + 3]);
            pre[4 * u] = ig; pre[4 * u + 1] = fg; pre[4 * u + 2] = gg; pre[4 * u + 3] = og;
            cst[u] = fg * cst[u] + ig * gg;
            hn[u] = og * tanh_ex2(cst[u]);
          }
          // publish h_{s+1} FIRST (the recurrent GEMM and the cluster's query projection wait for it) ...
          __nv_bfloat162 h01 = __floats2bfloat162_rn(hn[0], hn[1]), h23 = __floats2bfloat162_rn(hn[2], hn[3]);
          w0 = *reinterpret_cast<uint32_t*>(&h01); w1 = *reinterpret_cast<uint32_t*>(&h23);
          *reinterpret_cast<uint2*>(p.xh + nrow * K + p.F + 4 * tid) = make_uint2(w0, w1);
          const uint32_t hloc = smem_u32(sHb + (size_t)rank * HP + 4 * tid);
#pragma unroll
          for (int d = 0; d < R2_CS; ++d) st_dsmem_u32x2(mapa(hloc, (uint32_t)d), w0, w1);
        }
        __syncwarp();
        if (lane < R2_CS) mbar_arrive_remote(mapa(hb_full, (uint32_t)lane));
        named_bar<2, 128>();
        if (tid == 0 && has_row) signal_counter(cntY + R2_CNT_STRIDE * rank);     // h_{s+1} of my row is in xh slot s+1
        F2_STAMP(6);
        // ... then everything that is only saved for the backward pass
        if (has_row) {
          *reinterpret_cast<float4*>(p.c + nrow * H + 4 * tid) = make_float4(cst[0], cst[1], cst[2], cst[3]);
          if (p.out_hid) *reinterpret_cast<float4*>(p.out_hid + nrow * H + 4 * tid) = make_float4(hn[0], hn[1], hn[2], hn[3]);
          if (p.act) {
            float4* ar = reinterpret_cast<float4*>(p.act + grow * (size_t)(4 * H)) + 4 * tid;
#pragma unroll
            for (int q = 0; q < 4; ++q) ar[q] = make_float4(pre[4 * q], pre[4 * q + 1], pre[4 * q + 2], pre[4 * q + 3]);
          }
        }
      }
    }
  } else {
    // ================================================================= recurrent GEMM group (warps 8-11)
    uint32_t it_p = 0, it_c = 0;
    for (int s = 1; s < S; ++s) {
      if (warp == 8) {
        if (lane == 0) {
          poll_counters4(cntY, nl0 * s, nl1 * s, nl2 * s, nl3 * s);      // h_s of every row is in xh slot s
          asm volatile("fence.proxy.async;" ::: "memory");
          for (int i = 0; i < F2_NKB; ++i, ++it_p) {
            const int stage = (int)(it_p % F2_STAGES);
            const uint32_t par = (it_p / F2_STAGES) & 1u;
            mbar_wait(empty_bar(stage), par ^ 1u);
            mbar_expect_tx(full_bar(stage), F2_STAGE_BYTES);
            tma_load_2d(ring_base + stage * F2_STAGE_BYTES, &map_h, full_bar(stage), (kb0 + i) * 64, s * B);
          }
        }
        __syncwarp();
      } else if (warp == 9) {
        if (lane == 0) {
          constexpr uint32_t idesc = idesc_bf16(128, F2_BN);
          if (s == 1) mbar_wait(w_bar, 0);
          for (int i = 0; i < F2_NKB; ++i, ++it_c) {
            const int stage = (int)(it_c % F2_STAGES);
            const uint32_t par = (it_c / F2_STAGES) & 1u;
            mbar_wait(full_bar(stage), par);
            tc_fence_after();
            const uint64_t adesc = sw128_desc(ring_base + stage * F2_STAGE_BYTES);
            const uint64_t bdesc = sw128_desc(smem_base + i * F2_KB_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (i | k) ? 1u : 0u);
            umma_commit(empty_bar(stage));
          }
          umma_commit(tmem_full_bar);
        }
        __syncwarp();
      }
      F2_STAMP_G(7);
      // this K-slice's partial tile [128 x 64]: TMEM -> plane `rank` of gh (row = TMEM lane), no exchange
      mbar_wait(tmem_full_bar, (uint32_t)((s - 1) & 1));
      tc_fence_after();
      {
        const int prow = (warp & 3) * 32 + lane;
        float4* dst = reinterpret_cast<float4*>(p.gh + ((size_t)rank * 128 + prow) * (4 * H) + n0);
#pragma unroll
        for (int c = 0; c < F2_BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_lane + (uint32_t)(c * 32), v);
          if (prow < B) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              dst[c * 8 + (j >> 2)] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                  __uint_as_float(v[j + 3]));
          }
        }
      }
      tc_fence_before();
      F2_STAMP_G(8);
      named_bar<3, 128>();
      if (tid == 256) signal_counter(cntX + R2_CNT_STRIDE * rank);
      F2_STAMP_G(9);
    }
  }
#undef F2_STAMP
#undef F2_STAMP_G

  // ---------------------------------------------------------------- teardown (no CTA may exit while a peer can still
  // address its shared memory)
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFnR2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnR2 encode_fn_r2() {
  static EncodeTiledFnR2 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnR2>(q);
  });
  return fn;
}
int r2_make_map(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
  EncodeTiledFnR2 enc = encode_fn_r2();
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

static size_t recur2_fwd_smem() {
  return 1024 + F2_B_BYTES + F2_RING_BYTES + (size_t)(F2_UPR + R2_CS) * F2_HP * 2 +
         sizeof(float) * (3 * R2_A + 64 + 128 * F2_GCP) + 8 * (2 * F2_STAGES + 4) + 16;
}
size_t recur2_bwd_smem();
const void* recur2_bwd_kernel_ptr();

// Both persistent kernels need their 32 clusters of 4 CTAs co-resident (the counters and mbarriers spin): checked once
// per process with the occupancy calculator; MVC_B200_PERSISTENT=0 forces the launch chain.
bool recur2_supported(int B, int T, int F, int H, int A) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("MVC_B200_PERSISTENT");
    const char* e2 = getenv("MVC_B200_RECUR2");          // =0: keep the first-generation persistent kernels (A/B runs)
    disabled = ((e && e[0] == '0') || (e2 && e2[0] == '0')) ? 1 : 0;
  }
  if (disabled) return false;
  if (!(B >= 1 && B <= 128 && T >= 1 && T <= 8 * R2_R && F >= 8 && F % 8 == 0 && H == R2_H && A == R2_A)) return false;
  static std::mutex mu;
  static int ok = -1;
  std::lock_guard<std::mutex> lk(mu);
  if (ok < 0) {
    ok = 1;
    const void* kerns[2] = {(const void*)recur2_fwd_kernel, recur2_bwd_kernel_ptr()};
    const size_t smems[2] = {recur2_fwd_smem(), recur2_bwd_smem()};
    for (int i = 0; i < 2 && ok; ++i) {
      if (smems[i] > 227 * 1024 ||
          cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smems[i]) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        break;
      }
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(R2_CS);
      cfg.blockDim = dim3(R2_THREADS);
      cfg.dynamicSmemBytes = smems[i];
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = R2_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kerns[i], &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
      if (n < R2_H / 16) ok = 0;
    }
  }
  return ok == 1;
}

static long long* g_recur2_prof = nullptr;
void r2_set_fwd_prof(long long* p) { g_recur2_prof = p; }

int recur2_fwd_launch(const Recur2FwdParams& p, const void* whh_um, int64_t ldw, cudaStream_t st, bool sync_cleared) {
  MVC_TRY(r2_apply_spin_limit());
  MVC_CHECK(recur2_supported(p.B, p.T, p.F, R2_H, R2_A), "persistent recurrence: unsupported dims");
  CUtensorMap mh, mw;
  // A operand: the h halves of the xh slots, [(S+1)*B rows, H cols], row pitch K
  MVC_TRY(r2_make_map(p.xh + p.F, (int64_t)(p.S + 1) * p.B, R2_H, p.K, 128, &mh));
  MVC_TRY(r2_make_map(whh_um, (int64_t)4 * R2_H, R2_H, ldw, F2_BN, &mw));
  if (!sync_cleared) MVC_CUDA(cudaMemsetAsync(p.sync, 0, sizeof(unsigned) * 256, st));
  const size_t smem = recur2_fwd_smem();
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
  Recur2FwdParams pp = p;
  pp.prof = g_recur2_prof;
  void* args[] = {(void*)&mh, (void*)&mw, (void*)&pp};
  ProfScope prof(PK_STEP_FUSED, p.B, p.S, p.K, st);
  MVC_CUDA(cudaLaunchKernelExC(&cfg, (const void*)recur2_fwd_kernel, args));
  MVC_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ ctx rows (backward only)
// ctx[s,b,:] = sum_t alpha[s,b,t] feats[b,t,:] for all s: block = (chunk of 512 features, row b, chunk of 8 steps); the
// 8 x T attention weights of the block sit in shared memory (zero padded to a multiple of 4 frames) and are read as
// float4 (one LDS.128 per 16 FMAs); every thread owns 4 adjacent features (one 8-byte load per frame) and 8 steps.
// (The first version -- 2 features per thread, scalar broadcast loads of alpha, all steps per block -- was LDS bound:
// 62 us; this one is FMA bound.)
constexpr int CR_S = 8;
constexpr int CR_THREADS = 128;
__global__ void __launch_bounds__(CR_THREADS)
r2_ctx_rows_kernel(const __nv_bfloat16* __restrict__ feats, const float* __restrict__ alpha, int B, int T, int F, int S,
                   __nv_bfloat16* __restrict__ xh, int64_t ldx) {
  extern __shared__ float4 sAl4[];                // [CR_S][Tp / 4]
  float* sAl = reinterpret_cast<float*>(sAl4);
  const int b = blockIdx.y, s0 = blockIdx.z * CR_S;
  const int Tp = (T + 3) & ~3, T4 = Tp >> 2;
  for (int i = threadIdx.x; i < CR_S * Tp; i += CR_THREADS) {
    const int si = i / Tp, t = i - si * Tp;
    sAl[i] = (s0 + si < S && t < T) ? alpha[((size_t)(s0 + si) * B + b) * T + t] : 0.f;
  }
  __syncthreads();
  const int f = (blockIdx.x * CR_THREADS + threadIdx.x) * 4;
  if (f >= F) return;
  const uint2* kr = reinterpret_cast<const uint2*>(feats + (size_t)b * T * F + f);
  const size_t rs = (size_t)F >> 2;               // frame pitch in uint2 units
  float acc[CR_S][4];
#pragma unroll
  for (int i = 0; i < CR_S; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
  for (int t4 = 0; t4 < T4; ++t4) {
    float kf[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t4 * 4 + j;
      const uint2 k = t < T ? kr[(size_t)t * rs] : make_uint2(0u, 0u);
      const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&k.x));
      const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&k.y));
      kf[j][0] = lo.x; kf[j][1] = lo.y; kf[j][2] = hi.x; kf[j][3] = hi.y;
    }
#pragma unroll
    for (int i = 0; i < CR_S; ++i) {
      const float4 a = sAl4[i * T4 + t4];
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = fmaf(av[j], kf[j][e], acc[i][e]);
    }
  }
#pragma unroll
  for (int i = 0; i < CR_S; ++i)
    if (s0 + i < S) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(acc[i][0], acc[i][1]), hi = __floats2bfloat162_rn(acc[i][2], acc[i][3]);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(xh + ((size_t)(s0 + i) * B + b) * ldx + f) = o;
    }
}

// Tensor-core form of the same contraction (T <= 64): per row b it is the small GEMM ctx[S, F] = alpha_b[S, T] . keys_b[T, F].
// block = (256 features, row b, 32 steps), 4 warps x 64 features; alpha is split into two bf16 terms (hi + lo: the fp32
// weight to 2^-17) held as mma.sync A fragments in registers for the whole block, the key tile is staged once in shared
// memory ([t][f], 16-byte copies) and read as B fragments by ldmatrix.trans; fp32 accumulation, bf16 result.  The key
// tile is read once (not once per 8-step chunk) and the FMA pipe is left alone: 33 -> ~10 us, and it no longer slows the
// vocabulary projection it runs next to.
constexpr int CM_F = 256, CM_KP = CM_F + 8, CM_AP = 72;
__global__ void __launch_bounds__(128)
r2_ctx_rows_mma_kernel(const __nv_bfloat16* __restrict__ feats, const float* __restrict__ alpha, int B, int T, int F, int S,
                       __nv_bfloat16* __restrict__ xh, int64_t ldx) {
  __shared__ __align__(16) __nv_bfloat16 sK[64 * CM_KP];
  __shared__ __align__(16) __nv_bfloat16 sAh[32 * CM_AP];
  __shared__ __align__(16) __nv_bfloat16 sAl[32 * CM_AP];
  const int b = blockIdx.y, f0 = blockIdx.x * CM_F, s0 = blockIdx.z * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int Tk = (T + 15) & ~15;
  const int nf = F - f0 < CM_F ? F - f0 : CM_F;              // live features of this block (a multiple of 8)
  for (int i = tid; i < 32 * 64; i += 128) {
    const int m = i >> 6, t = i & 63;
    const float a = (s0 + m < S && t < T) ? alpha[((size_t)(s0 + m) * B + b) * T + t] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16(a);
    sAh[m * CM_AP + t] = hi;
    sAl[m * CM_AP + t] = __float2bfloat16(a - __bfloat162float(hi));
  }
  for (int i = tid; i < Tk * (CM_F / 8); i += 128) {
    const int t = i >> 5, c = i & 31;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (t < T && c * 8 < nf) v = *reinterpret_cast<const uint4*>(feats + ((size_t)b * T + t) * F + f0 + c * 8);
    *reinterpret_cast<uint4*>(sK + t * CM_KP + c * 8) = v;
  }
  __syncthreads();
  uint32_t ah[2][4][4], al[2][4][4];                         // [m tile][k step][fragment register]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int o = (mt * 16 + g) * CM_AP + ks * 16 + 2 * q;
      ah[mt][ks][0] = *reinterpret_cast<const uint32_t*>(sAh + o);
      ah[mt][ks][1] = *reinterpret_cast<const uint32_t*>(sAh + o + 8 * CM_AP);
      ah[mt][ks][2] = *reinterpret_cast<const uint32_t*>(sAh + o + 8);
      ah[mt][ks][3] = *reinterpret_cast<const uint32_t*>(sAh + o + 8 * CM_AP + 8);
      al[mt][ks][0] = *reinterpret_cast<const uint32_t*>(sAl + o);
      al[mt][ks][1] = *reinterpret_cast<const uint32_t*>(sAl + o + 8 * CM_AP);
      al[mt][ks][2] = *reinterpret_cast<const uint32_t*>(sAl + o + 8);
      al[mt][ks][3] = *reinterpret_cast<const uint32_t*>(sAl + o + 8 * CM_AP + 8);
    }
  const int mi = lane >> 3, mr = lane & 7;                   // ldmatrix: lanes 8i..8i+7 address the rows of matrix i
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    const int nb = warp * 64 + it * 16;                      // 16 features = two n tiles
    if (nb >= nf) break;
    float acc[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (ks * 16 < Tk) {
        // matrices: (k 0-7, n 0-7), (k 8-15, n 0-7), (k 0-7, n 8-15), (k 8-15, n 8-15), transposed on the way in
        const uint32_t addr = smem_u32(sK + (ks * 16 + (mi & 1) * 8 + mr) * CM_KP + nb + (mi >> 1) * 8);
        uint32_t bq[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(bq[0]), "=r"(bq[1]), "=r"(bq[2]), "=r"(bq[3]) : "r"(addr));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(acc[mt][nt][0]), "+f"(acc[mt][nt][1]), "+f"(acc[mt][nt][2]), "+f"(acc[mt][nt][3])
                         : "r"(ah[mt][ks][0]), "r"(ah[mt][ks][1]), "r"(ah[mt][ks][2]), "r"(ah[mt][ks][3]),
                           "r"(bq[2 * nt]), "r"(bq[2 * nt + 1]));
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(acc[mt][nt][0]), "+f"(acc[mt][nt][1]), "+f"(acc[mt][nt][2]), "+f"(acc[mt][nt][3])
                         : "r"(al[mt][ks][0]), "r"(al[mt][ks][1]), "r"(al[mt][ks][2]), "r"(al[mt][ks][3]),
                           "r"(bq[2 * nt]), "r"(bq[2 * nt + 1]));
          }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int fl = nb + nt * 8 + 2 * q;                  // feature inside the block
        if (fl < nf) {
          const int sa = s0 + mt * 16 + g, sb = sa + 8;
          if (sa < S)
            *reinterpret_cast<__nv_bfloat162*>(xh + ((size_t)sa * B + b) * ldx + f0 + fl) =
                __floats2bfloat162_rn(acc[mt][nt][0], acc[mt][nt][1]);
          if (sb < S)
            *reinterpret_cast<__nv_bfloat162*>(xh + ((size_t)sb * B + b) * ldx + f0 + fl) =
                __floats2bfloat162_rn(acc[mt][nt][2], acc[mt][nt][3]);
        }
      }
  }
}

int r2_ctx_rows(const void* feats_bf16, const float* alpha, int B, int T, int F, int S, void* xh_bf16, int64_t ldx,
                cudaStream_t st) {
  static int use_mma = -1;
  if (use_mma < 0) {
    const char* e = getenv("MVC_B200_CTX_MMA");
    use_mma = (e && e[0] == '0') ? 0 : 1;
  }
  if (use_mma && T <= 64 && F % 8 == 0 && ldx % 2 == 0 && (reinterpret_cast<uintptr_t>(feats_bf16) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(xh_bf16) & 3) == 0) {
    dim3 grid((unsigned)cdiv(F, CM_F), (unsigned)B, (unsigned)cdiv(S, 32));
    r2_ctx_rows_mma_kernel<<<grid, 128, 0, st>>>((const __nv_bfloat16*)feats_bf16, alpha, B, T, F, S,
                                                 (__nv_bfloat16*)xh_bf16, ldx);
    MVC_LAUNCH_CHECK();
    return 0;
  }
  MVC_CHECK(F % 4 == 0 && ldx % 4 == 0 && T <= 1024, "r2_ctx_rows: unsupported dims S=%d T=%d F=%d", S, T, F);
  dim3 grid((unsigned)cdiv(F, 4 * CR_THREADS), (unsigned)B, (unsigned)cdiv(S, CR_S));
  const int Tp = (T + 3) & ~3;
  r2_ctx_rows_kernel<<<grid, CR_THREADS, sizeof(float) * CR_S * Tp, st>>>((const __nv_bfloat16*)feats_bf16, alpha, B, T, F, S,
                                                                         (__nv_bfloat16*)xh_bf16, ldx);
  MVC_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvc
