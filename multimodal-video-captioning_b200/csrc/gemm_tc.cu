// gemm_tc.cu -- bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (sm_100a), with
// deterministic split-K and an optional fused LSTM-cell epilogue.
//
//   C[M,N] = sum_k A[m,k] * B[n,k]  (+ beta*C) (+ bias[n])
//
// A [M,K] and B [N,K] are bf16, K-contiguous (the layout every nn.Linear / nn.LSTM weight of
// the reference already has: weight[out,in], and the layout of every activation matrix
// [rows, features]).  This is the tensor-core form of the LSTM gate contraction
// (features_captioning.py:84), the attention projections (temporal_attention.py:20-21), the
// vocabulary projection (features_captioning.py:87) and all their backward GEMMs.
//
// Kernel anatomy (one 128 x BN output tile x one K-range per CTA, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d tiles of A (128x64) and B (BNx64) into a
//               STAGES-deep shared-memory ring, 128B swizzle, completion on mbarriers (expect_tx);
//   warp 1      allocates TMEM, then one lane issues tcgen05.mma (cta_group::1, kind::f16,
//               M=128, N=BN, K=16) four times per stage, accumulating in TMEM; tcgen05.commit
//               releases the stage back to the producer and finally signals the epilogue;
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns at a time.
// Out-of-range rows / K tails are zero-filled by TMA; stores are predicated.
//
// Split-K (the recurrence GEMMs have M = batch = one row tile, so N tiles alone cannot fill
// 148 SMs): blockIdx.z owns a contiguous range of 64-wide K blocks and the S splits of one output
// tile are launched as one thread-block CLUSTER (1,1,S).  Each CTA parks its fp32 partial tile in
// its own shared memory (the idle operand ring); after a cluster barrier CTA z reads rows
// [z*128/S, (z+1)*128/S) of all S partials through distributed shared memory, sums them in rank
// order (deterministic) and runs the epilogue for that row block -- no global scratch, no atomics.
// The reduction is warp-per-row: BN/4 lanes x float4 read one parked row of one partial per
// request (conflict-free, contiguous), and the result is stored as contiguous row segments
// (DSMEM moves ~17 B/clk/SM: a thread-per-row mapping with 4-way bank conflicts took 8.4 us
// per tile, this one 3 us).  Tile width and S come from a small cost model (tc_gemm()).
//
// Epilogues:
//   plain : + bias, + beta*C, fp32 store, optional bf16 copy;
//   cell  : BN = 128 and the weight rows are permuted so that one tile holds gates i,f,g,o of 32
//           hidden units (col = (j/16)*64 + gate*16 + j%16).  The tile is parked UNIT-major (the
//           four gates of a unit adjacent) whether K is split or not, and finished by a looped
//           warp-per-row pass with lane <-> hidden unit (cell_reduce_rows): hoisted input
//           projection / embedding-table row / bias / c_prev loads in flight together with the
//           DSMEM loads, ex2-based sigmoid / tanh, c and h (fp32 + bf16) stored as 128-byte row
//           segments -- the LSTM step never materialises pre-activations.  Multi-wave launches
//           (beam search) use a 3-stage ring so that two CTAs are resident per SM.
//
// The persistent 128x256 kernel further down serves the large GEMMs, with arg-max / top-k +
// log-sum-exp epilogues for the vocabulary projection of the decode loops (optionally carrying the
// next step's attention query projection as an auxiliary column block, TcAux).
//
// Programmatic dependent launch: when launched with the PDL attribute the producer prefetches
// the loop-invariant operand B (weights) before griddepcontrol.wait, so the weight traffic of
// step s+1 overlaps the tail of the kernel that produces its activations.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "ptx.cuh"
#include "tc_gemm.cuh"

namespace mvc {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;   // 64 bf16 = 128 bytes = one swizzle-128B row

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

// MN-major operand (the matrix is stored [K rows, MN contiguous], i.e. TRANSPOSED w.r.t. the K-major case), 128-byte
// swizzle, tile built from TMA boxes {64 MN elements, 64 K rows}: one box = 64 rows (k) of 128 B; 8 consecutive k rows
// form a 1024-byte swizzle atom (stride byte offset = 1024), boxes for MN = 64.. follow 8192 B later (leading byte
// offset = 8192).  One UMMA (K = 16) consumes two k-groups: the start address advances by 2048 B per step.
__device__ __forceinline__ uint64_t make_sw128_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;                   // leading byte offset: next 64-element block along MN
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: next group of 8 k rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_major(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
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
  static constexpr int TOK_OFF = BAR_OFF + (2 * STAGES + 1) * 8 + 16;        // int64 tokens of the CTA's reduce rows
  static constexpr int TOTAL = TOK_OFF + 128 * 8 + 1024;                     // + alignment slack
};



// ---- cluster / distributed shared memory
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// 16-bit copy of an output value: bf16, or fp16 with saturation (TcEpilogue::cb_f16)
__device__ __forceinline__ __nv_bfloat16 to_b16(float r, int f16) {
  if (f16) return __ushort_as_bfloat16(__half_as_ushort(__float2half_rn(fminf(fmaxf(r, -65504.f), 65504.f))));
  return __float2bfloat16(r);
}

// plain epilogue for NC (compile-time capacity; `n` live) consecutive columns of one row
template <int NC, typename Arr>
__device__ __forceinline__ void plain_store(const TcEpilogue& ep, int64_t gm, int gn0, int N, Arr& o, int n = NC) {
  float* crow = ep.C ? ep.C + gm * ep.ldc : nullptr;
  __nv_bfloat16* brow = ep.Cb ? ep.Cb + gm * ep.ldcb : nullptr;
  const bool vec_ok = crow && (gn0 + n - 1 < N) && ((reinterpret_cast<uintptr_t>(crow + gn0) & 15u) == 0);
  if (vec_ok) {
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
      if (j < n) {
        float4 r = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        if (ep.bias) {
          r.x += ep.bias[gn0 + j]; r.y += ep.bias[gn0 + j + 1]; r.z += ep.bias[gn0 + j + 2]; r.w += ep.bias[gn0 + j + 3];
        }
        float4* dst = reinterpret_cast<float4*>(crow + gn0 + j);
        if (ep.beta != 0.f) {
          const float4 old = *dst;
          r.x += ep.beta * old.x; r.y += ep.beta * old.y; r.z += ep.beta * old.z; r.w += ep.beta * old.w;
        }
        *dst = r;
        if (brow) {
          brow[gn0 + j] = to_b16(r.x, ep.cb_f16); brow[gn0 + j + 1] = to_b16(r.y, ep.cb_f16);
          brow[gn0 + j + 2] = to_b16(r.z, ep.cb_f16); brow[gn0 + j + 3] = to_b16(r.w, ep.cb_f16);
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int gn = gn0 + j;
      if (j < n && gn < N) {
        float r = o[j];
        if (ep.bias) r += ep.bias[gn];
        if (crow) {
          if (ep.beta != 0.f) r += ep.beta * crow[gn];
          crow[gn] = r;
        }
        if (brow) brow[gn] = to_b16(r, ep.cb_f16);
      }
    }
  }
}

// ---- fused LSTM cell epilogue (TC_MODE_CELL): the accumulator tile is parked UNIT-major in shared memory and finished
// by a warp-per-row pass (cell_reduce_rows below), split-K or not.
// Warp e of the 4 epilogue warps finishes rows z*R + e, e + 4, ... (R = 128 / S) of the tile, lane <-> hidden unit, so
// a warp touches 32 consecutive units of one row (c / h: one 128-byte segment; packed gate columns: two 64-byte
// segments per gate).
__device__ __forceinline__ void cell_finish_unit(const TcEpilogue& ep, int64_t gm, int tile, int u, const float (&g)[4], float cp) {
  const int H = ep.H;
  const int ug = tile * 32 + u;
  const float ig = sigmoid_ex2(g[0]), fg = sigmoid_ex2(g[1]), gg = tanh_ex2(g[2]), og = sigmoid_ex2(g[3]);
  const float cn = fg * cp + ig * gg;
  const float hn = og * tanh_ex2(cn);
  ep.c_out[gm * H + ug] = cn;
  if (ep.h32) ep.h32[gm * ep.h_ld + ug] = hn;
  if (ep.h32b) ep.h32b[gm * ep.h2_ld + ug] = hn;
  if (ep.hb) ep.hb[gm * ep.hb_ld + ug] = __float2bfloat16(hn);
  if (ep.act) {
    float* actr = ep.act + gm * (int64_t)(4 * H) + tile * 128;
    actr[gate_lcol(0, u)] = ig; actr[gate_lcol(1, u)] = fg; actr[gate_lcol(2, u)] = gg; actr[gate_lcol(3, u)] = og;
  }
}

// DSMEM reduction + cell for the warp's rows, S = cluster size (compile time): a real loop of four batches of 8 / S
// rows (straight-line code for all 16 rows was instruction-fetch bound: 2400 SASS instructions executed once, 80 %
// of the samples `stall_no_inst`).  Per batch the 8 float4 DSMEM loads (rows x partials) and the rows' global
// addends (input projection, embedding-table row, c_prev) are all in flight before anything consumes them.
template <int S>
__device__ __forceinline__ void cell_reduce_rows(const TcEpilogue& ep, uint32_t smem_base, int RS, int z, int ew, int lane,
                                                 int64_t m0, int M, int tile, const int64_t* s_tok) {
  constexpr int RB = 8 / S;                       // rows per batch
  constexpr int rows_per = TC_BM / S;
  const int H = ep.H;
  const int n0 = tile * 128;
  const int woff = ((lane >> 4) * 64 + (lane & 15) * 4);     // unit-major word offset inside the parked row
  int lc[4];
  float bias4[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    lc[c] = n0 + gate_lcol(c, lane);
    bias4[c] = ep.bias ? __ldg(ep.bias + lc[c]) : 0.f;
  }
#pragma unroll 1
  for (int b = 0; b < 4; ++b) {
    float4 t[8];
    float a[RB][4], e[RB][4], cp[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int rl = ew + 4 * (b * RB + r);
      const int row = z * rows_per + rl;
      const uint32_t addr = smem_base + (uint32_t)(row * RS + woff) * 4u;
#pragma unroll
      for (int s = 0; s < S; ++s) t[r * S + s] = ld_dsmem_v4(mapa_cluster(addr, (uint32_t)s));
      const int64_t gm = m0 + row;
      const bool ok = gm < M;
      const float* gxr = (ok && ep.gx) ? ep.gx + gm * ep.gx_ld : nullptr;
      const float* etr = (ok && ep.embtab) ? ep.embtab + s_tok[rl] * (int64_t)(4 * H) : nullptr;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        a[r][c] = gxr ? gxr[lc[c]] : 0.f;
        e[r][c] = etr ? etr[lc[c]] : 0.f;
      }
      cp[r] = (ok && ep.c_prev) ? ep.c_prev[gm * H + tile * 32 + lane] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int row = z * rows_per + ew + 4 * (b * RB + r);
      float g[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) g[c] = a[r][c] + e[r][c] + bias4[c];
#pragma unroll
      for (int s = 0; s < S; ++s) {               // rank order: deterministic
        g[0] += t[r * S + s].x; g[1] += t[r * S + s].y; g[2] += t[r * S + s].z; g[3] += t[r * S + s].w;
      }
      const int64_t gm = m0 + row;
      if (gm < M) cell_finish_unit(ep, gm, tile, lane, g, cp[r]);
    }
  }
}

template <int BN, int STAGES, int MODE>
__global__ void __launch_bounds__(192, (STAGES <= 3 ? 2 : 1))
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int M, int N,
                    int K, const __grid_constant__ TcEpilogue ep) {
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
  const int nkb_all = (K + TC_BK - 1) / TC_BK;
  const int splits = gridDim.z, z = blockIdx.z;
  const int kb0 = (int)((long long)nkb_all * z / splits), kb1 = (int)((long long)nkb_all * (z + 1) / splits);
  const int nkb = kb1 - kb0;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  unsigned long long* prof =
      ep.prof ? ep.prof + (size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
#define TC_STAMP(i, cond) do { if (prof && (cond)) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); prof[i] = t_; } } while (0)
  TC_STAMP(0, threadIdx.x == 0);
  int64_t* s_tok = reinterpret_cast<int64_t*>(smem + S::TOK_OFF);   // split-K cell epilogue: tokens of my reduce rows

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
  pdl_trigger();     // every CTA of this grid is resident by the time the dependent grid may start
  TC_STAMP(1, threadIdx.x == 0);

  if (warp == 0) {
    if (lane == 0) {
      // B (weights) does not depend on the preceding kernel: fill the ring's B halves before the
      // grid dependency resolves, then stream A.
      const int pre = ep.b_const ? (nkb < STAGES ? nkb : STAGES) : 0;
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(full_bar(kb), S::STAGE_BYTES);
        tma_load_2d(smem_base + kb * S::STAGE_BYTES + S::A_BYTES, &map_b, full_bar(kb), (kb0 + kb) * TC_BK, n0);
      }
      pdl_wait();
      TC_STAMP(2, true);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t a_dst = smem_base + stage * S::STAGE_BYTES;
        if (kb >= pre) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), S::STAGE_BYTES);
          tma_load_2d(a_dst + S::A_BYTES, &map_b, full_bar(stage), (kb0 + kb) * TC_BK, n0);
        }
        tma_load_2d(a_dst, &map_a, full_bar(stage), (kb0 + kb) * TC_BK, m0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      TC_STAMP(3, true);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && nkb > 0) {
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
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    pdl_wait();                                   // C / cell state may still be in use upstream
    if constexpr (MODE == TC_MODE_CELL) {
      // stage the tokens of the rows this CTA will finish (embedding-table gather in the cell epilogue)
      const int rows_per = TC_BM / splits;
      const int te = threadIdx.x - 64;
      const int64_t gmr = (int64_t)m0 + z * rows_per + te;
      if (te < rows_per) s_tok[te] = (ep.embtab && gmr < M) ? ep.tokens[gmr] : 0;
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    TC_STAMP(4, threadIdx.x == 64);
    if (splits == 1 && MODE == TC_MODE_PLAIN) {
      const int64_t gm = (int64_t)m0 + row;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(trow + (uint32_t)(c * 32), v);
        if (gm < M) {
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
          plain_store<32>(ep, gm, n0 + c * 32, N, o);
        }
      }
    } else {
      // split-K (and the unsplit cell epilogue): park this CTA's tile in its own shared memory (the operand ring is
      // idle now: every MMA has completed), row pitch BN + 4 words.  Plain: row-major.  Cell: UNIT-major, the four gates of a unit adjacent
      // (word (k*16 + u)*4 + g for packed column k*64 + g*16 + u), so that the reduction reads one float4 per unit.
      float* red = reinterpret_cast<float*>(smem);
      constexpr int RS = BN + 4;
      if constexpr (MODE == TC_MODE_CELL) {
#pragma unroll 1
        for (int k = 0; k < BN / 64; ++k) {
          uint32_t v0[32], v1[32];
          tmem_ld32(trow + (uint32_t)(k * 64), v0);          // gates i, f of units [16k, 16k + 16)
          tmem_ld32(trow + (uint32_t)(k * 64 + 32), v1);     // gates g, o
          float4* dst = reinterpret_cast<float4*>(red + (size_t)row * RS + k * 64);
#pragma unroll
          for (int u = 0; u < 16; ++u)
            dst[u] = make_float4(__uint_as_float(v0[u]), __uint_as_float(v0[16 + u]), __uint_as_float(v1[u]),
                                 __uint_as_float(v1[16 + u]));
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(trow + (uint32_t)(c * 32), v);
          float4* dst = reinterpret_cast<float4*>(red + (size_t)row * RS + c * 32);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            dst[j >> 2] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                      __uint_as_float(v[j + 3]));
        }
      }
    }
    if constexpr (MODE == TC_MODE_CELL) {
      if (splits == 1) {
        // unsplit cell epilogue = the same warp-per-row pass over the parked tile (S = 1: only this CTA's partial)
        asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps: whole tile parked
        cell_reduce_rows<1>(ep, smem_base, BN + 4, 0, warp - 2, lane, m0, M, blockIdx.x, s_tok);
      }
    }
    tc_fence_before();
  }
  if (splits > 1) {
    // Reduce across the cluster (the S CTAs of one output tile) through distributed shared memory: CTA z owns
    // rows [z*128/S, (z+1)*128/S) of the tile, reads that row block from all S partials in rank order
    // (deterministic), and runs the epilogue for it.
    cluster_arrive();
    cluster_wait();
    TC_STAMP(5, threadIdx.x == 64);
    if (warp >= 2) {
      // A warp reads whole rows: BN / 4 lanes x float4 per row (32 / (BN/4) rows per instruction), so every DSMEM
      // request is a conflict-free contiguous row segment of the peer's shared memory and every global store a
      // contiguous row segment of C.  All S partial loads of a row are issued before the adds.
      constexpr int RS = BN + 4;
      const int ew = warp - 2;
      const int rows_per = TC_BM / splits;
      if constexpr (MODE == TC_MODE_PLAIN) {
        constexpr int LPR = BN / 4, RPI = 32 / LPR;          // lanes per row, rows per warp instruction
        const int col = (lane % LPR) * 4;
#pragma unroll 2
        for (int rl = ew * RPI + lane / LPR; rl < rows_per; rl += 4 * RPI) {
          const int row = z * rows_per + rl;
          const uint32_t addr = smem_base + (uint32_t)(row * RS + col) * 4u;
          float4 t[8];
#pragma unroll
          for (int s = 0; s < 8; ++s)
            if (s < splits) t[s] = ld_dsmem_v4(mapa_cluster(addr, (uint32_t)s));
          float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int s = 0; s < 8; ++s)
            if (s < splits) { o[0] += t[s].x; o[1] += t[s].y; o[2] += t[s].z; o[3] += t[s].w; }
          const int64_t gm = (int64_t)m0 + row;
          if (gm < M) plain_store<4>(ep, gm, n0 + col, N, o, 4);
        }
      } else {
        // cell: lane <-> hidden unit of the tile (32 units), float4 = its four gates; g4r / cpr hold the addends
        if (splits == 2) cell_reduce_rows<2>(ep, smem_base, RS, z, ew, lane, m0, M, blockIdx.x, s_tok);
        else if (splits == 4) cell_reduce_rows<4>(ep, smem_base, RS, z, ew, lane, m0, M, blockIdx.x, s_tok);
        else cell_reduce_rows<8>(ep, smem_base, RS, z, ew, lane, m0, M, blockIdx.x, s_tok);
      }
    }
    TC_STAMP(6, threadIdx.x == 64);
    cluster_arrive();                                        // nobody may leave while its partial is being read
    cluster_wait();
  }
  __syncthreads();
  TC_STAMP(7, threadIdx.x == 64);
#undef TC_STAMP
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ persistent 128x256 kernel for the big GEMMs
// One CTA per SM loops over output tiles (tile = blockIdx.x + i*gridDim.x).  The fp32 accumulator is DOUBLE
// BUFFERED in TMEM (2 x 256 columns = all 512): while the four epilogue warps drain tile i, the MMA warp already
// accumulates tile i+1 and the producer keeps the TMA ring full across tile boundaries, so neither the prologue
// nor the epilogue of a tile is exposed.  The epilogue transposes each 32x32 chunk through shared memory so that
// every global store instruction writes 128 contiguous bytes of one output row.
constexpr int PG_BN = 256;
constexpr int PG_STAGES = 4;
constexpr int PG_A_BYTES = TC_BM * TC_BK * 2;            // 16 KB
constexpr int PG_B_BYTES = PG_BN * TC_BK * 2;            // 32 KB
constexpr int PG_STAGE_BYTES = PG_A_BYTES + PG_B_BYTES;  // 48 KB
constexpr int PG_EPI_OFF = PG_STAGES * PG_STAGE_BYTES;   // 4 warps x [32][36] floats staging
constexpr int PG_STG = 32 * 36;                           // floats per warp: pitch 36 = 16-byte aligned rows (float4 accesses)
constexpr int PG_EPI_BYTES = 4 * PG_STG * 4 + 4 * PG_BN * 4;    // per-warp transpose staging + per-warp bias tile
constexpr int PG_BAR_OFF = PG_EPI_OFF + PG_EPI_BYTES;
constexpr int PG_SMEM = PG_BAR_OFF + (2 * PG_STAGES + 4) * 8 + 16 + 1024;

// One 32-column chunk of a row into its sorted best-KK list (ties: the earlier, lower column stays in front).
template <int KK>
__device__ __forceinline__ void topk_insert_chunk(const float (&x)[32], float cmax, int gn0, float (&tv)[8], int (&tix)[8]) {
  if (cmax > tv[KK - 1]) {                         // some element of this chunk enters the list
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (x[j] > tv[KK - 1]) {
        tv[KK - 1] = x[j]; tix[KK - 1] = gn0 + j;
#pragma unroll
        for (int i = KK - 1; i > 0; --i) {
          if (tv[i] > tv[i - 1]) {                 // strict: an equal value stays behind the earlier (lower) column
            const float tf = tv[i]; tv[i] = tv[i - 1]; tv[i - 1] = tf;
            const int tn = tix[i]; tix[i] = tix[i - 1]; tix[i - 1] = tn;
          }
        }
      }
    }
  }
}

// MC = 1: the CTAs run as CLUSTER PAIRS over two vertically adjacent 128-row tiles of the same 256-column block.  Both
// need the same B tile, so each CTA fetches only HALF of it (128 of the 256 rows) with a TMA multicast that lands in both
// CTAs' shared memory: L2 reads per CTA and k-block drop from 48 KB to 32 KB (the 128x256-tile kernel is L2 -> SM
// bandwidth bound: 43 GB/s per SM x 148 = 6.3 TB/s at 25-33 % tensor-pipe active).  A stage may be refilled only when
// BOTH CTAs' MMAs have consumed it (the peer's multicast writes into my ring too): the MMA commit arrives on both CTAs'
// empty barriers (multicast commit, count 2).
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

template <int MODE, int A_MN = 0, int B_MN = 0, int MC = 0>
__global__ void __launch_bounds__(192, 1)
gemm_bf16_tc_persist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int M,
                            int N, int K, const __grid_constant__ TcEpilogue ep) {
  static_assert(!MC || (!A_MN && !B_MN), "the multicast pair variant takes K-major operands");
  extern __shared__ uint8_t smem_raw[];
  // aligned by offset, not by integer round-trip: the pointer keeps its shared address space (LDS / STS, not generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + PG_BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (PG_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * PG_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * PG_STAGES + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + PG_BAR_OFF + 8 * (2 * PG_STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (K + TC_BK - 1) / TC_BK;
  const int ntn = (N + PG_BN - 1) / PG_BN, ntm = (M + TC_BM - 1) / TC_BM;
  // work items: tiles, or (MC) pairs of vertically adjacent tiles shared by the two CTAs of a cluster
  int crank = 0;
  if constexpr (MC) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int ntiles = MC ? ntn * ((ntm + 1) / 2) : ntn * ntm;
  const int tile0 = MC ? (int)blockIdx.x / 2 : (int)blockIdx.x, tile_step = MC ? (int)gridDim.x / 2 : (int)gridDim.x;
  auto tile_m0 = [&](int tile) { return MC ? (2 * (tile / ntn) + crank) * TC_BM : (tile / ntn) * TC_BM; };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < PG_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), MC ? 2 : 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);               // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (MC) {                               // the peer's barriers are initialised before any multicast / commit
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      uint32_t it = 0;
      for (int tile = tile0; tile < ntiles; tile += tile_step) {
        const int m0 = tile_m0(tile), n0 = (tile % ntn) * PG_BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int stage = (int)(it % PG_STAGES);
          const uint32_t par = (it / PG_STAGES) & 1u;
          mbar_wait(empty_bar(stage), par ^ 1u);
          mbar_expect_tx(full_bar(stage), PG_STAGE_BYTES);
          const uint32_t a_dst = smem_base + stage * PG_STAGE_BYTES;
          if constexpr (A_MN) {                   // A stored [K, M]: two boxes of {64 m, 64 k}
#pragma unroll
            for (int i = 0; i < TC_BM / 64; ++i) tma_load_2d(a_dst + i * 8192, &map_a, full_bar(stage), m0 + i * 64, kb * TC_BK);
          } else {
            tma_load_2d(a_dst, &map_a, full_bar(stage), kb * TC_BK, m0);
          }
          if constexpr (B_MN) {                   // B stored [K, N]: four boxes of {64 n, 64 k}
#pragma unroll
            for (int i = 0; i < PG_BN / 64; ++i)
              tma_load_2d(a_dst + PG_A_BYTES + i * 8192, &map_b, full_bar(stage), n0 + i * 64, kb * TC_BK);
          } else if constexpr (MC) {              // my half of the shared B tile, delivered to both CTAs of the pair
            tma_load_2d_mc(a_dst + PG_A_BYTES + crank * (PG_B_BYTES / 2), &map_b, full_bar(stage), kb * TC_BK,
                           n0 + crank * (PG_BN / 2), (uint16_t)0x3);
          } else {
            tma_load_2d(a_dst + PG_A_BYTES, &map_b, full_bar(stage), kb * TC_BK, n0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_major(TC_BM, PG_BN, A_MN, B_MN);
      uint32_t it = 0;
      int ti = 0;
      for (int tile = tile0; tile < ntiles; tile += tile_step, ++ti) {
        const int buf = ti & 1;
        mbar_wait(tempty_bar(buf), (((uint32_t)ti >> 1) & 1u) ^ 1u);     // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * PG_BN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int stage = (int)(it % PG_STAGES);
          const uint32_t par = (it / PG_STAGES) & 1u;
          mbar_wait(full_bar(stage), par);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * PG_STAGE_BYTES;
          const uint64_t adesc = A_MN ? make_sw128_desc_mn(a_addr) : make_sw128_desc(a_addr);
          const uint64_t bdesc = B_MN ? make_sw128_desc_mn(a_addr + PG_A_BYTES) : make_sw128_desc(a_addr + PG_A_BYTES);
          // per K = 16 step: K-major operands advance 32 B inside the swizzle atom, MN-major ones two k-groups (2048 B)
          constexpr uint64_t a_step = A_MN ? (2048 >> 4) : 2, b_step = B_MN ? (2048 >> 4) : 2;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16(tacc, adesc + a_step * k, bdesc + b_step * k, idesc, (kb | k) ? 1u : 0u);
          if constexpr (MC) umma_commit_mc(empty_bar(stage), (uint16_t)0x3);
          else umma_commit(empty_bar(stage));
        }
        umma_commit(tfull_bar(buf));
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;                       // TMEM lane quarter of this warp
    float* stg = reinterpret_cast<float*>(smem + PG_EPI_OFF) + (warp - 2) * PG_STG;
    float* sbias = reinterpret_cast<float*>(smem + PG_EPI_OFF + 4 * PG_STG * 4) + (warp - 2) * PG_BN;   // bias of this tile
    pdl_wait();
    int ti = 0;
    for (int tile = tile0; tile < ntiles; tile += tile_step, ++ti) {
      const int buf = ti & 1;
      const int m0 = tile_m0(tile), n0 = (tile % ntn) * PG_BN;
      // this tile's bias values -> the warp's private smem copy (broadcast reads below instead of 256 LDGs per thread)
#pragma unroll
      const int nmain = (MODE != TC_MODE_PLAIN && ep.aux_C) ? ep.n_main : N;   // columns that carry a bias / reduction
      for (int i = lane; i < PG_BN; i += 32) sbias[i] = (ep.bias && n0 + i < nmain) ? __ldg(ep.bias + n0 + i) : 0.f;
      __syncwarp();
      mbar_wait(tfull_bar(buf), ((uint32_t)ti >> 1) & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * PG_BN);
      if (MODE != TC_MODE_PLAIN && ep.aux_C && n0 >= ep.aux_n0) {
        // auxiliary column block: plain fp32 store (32x32 chunks transposed through smem -> 128-byte row segments)
        const int64_t gm0 = (int64_t)m0 + q * 32;
        const int nrows = (int)((M - gm0) < 32 ? (M - gm0) : 32);
#pragma unroll 1
        for (int c = 0; c < PG_BN / 32; ++c) {
          const int gn0 = n0 + c * 32;
          if (gn0 >= N) break;
          uint32_t v[32];
          tmem_ld32(trow + (uint32_t)(c * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);
          __syncwarp();
          const int gn = gn0 + lane;
          if (gn < N && nrows > 0) {
            float* dst = ep.aux_C + gm0 * ep.aux_ld + (gn - ep.aux_n0);
#pragma unroll 4
            for (int rr = 0; rr < nrows; ++rr) {
              *dst = stg[rr * 33 + lane];
              dst += ep.aux_ld;
            }
          }
          __syncwarp();
        }
      } else if constexpr (MODE == TC_MODE_ARGMAX) {
        // K-E: row-wise arg-max of (acc + bias) over this tile's columns; the logits never leave the SM.
        // Thread = row; columns ascend, so a strict > keeps the lowest index on ties.
        float best = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll 1
        for (int c = 0; c < PG_BN / 32; ++c) {
          const int gn0 = n0 + c * 32;
          if (gn0 >= nmain) break;
          uint32_t v[32];
          tmem_ld32(trow + (uint32_t)(c * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gn = gn0 + j;
            if (gn < nmain) {
              const float x = __uint_as_float(v[j]) + sbias[c * 32 + j];
              if (x > best) { best = x; bi = gn; }
            }
          }
        }
        const int64_t gm = (int64_t)m0 + q * 32 + lane;
        if (gm < M) {
          ep.amax_val[(int64_t)(tile % ntn) * M + gm] = best;       // [tile][row]
          ep.amax_idx[(int64_t)(tile % ntn) * M + gm] = bi;
        }
      } else if constexpr (MODE == TC_MODE_TOPK) {
        // K-E (beam search): per row of this tile the 8 best (acc + bias) values with their columns (sorted, ties ->
        // lowest column) and the (max, sum exp) pair of an online log-sum-exp; the logits never leave the SM.
        float tv[8];
        int tix[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { tv[i] = -INFINITY; tix[i] = 0x7fffffff; }
        float mx = -INFINITY, sm = 0.f;
#pragma unroll 1
        for (int c = 0; c < PG_BN / 32; ++c) {
          const int gn0 = n0 + c * 32;
          if (gn0 >= nmain) break;
          uint32_t v[32];
          tmem_ld32(trow + (uint32_t)(c * 32), v);
          float x[32];
          float cmax = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gn = gn0 + j;
            x[j] = (gn < nmain) ? __uint_as_float(v[j]) + sbias[c * 32 + j] : -INFINITY;
            cmax = fmaxf(cmax, x[j]);
          }
          // online log-sum-exp, one rescale per 32-column chunk; the 32 exponentials are independent
          const float nmx = fmaxf(mx, cmax);
          float part = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) part += __expf(x[j] - nmx);
          sm = sm * __expf(mx - nmx) + part;
          mx = nmx;
          // sorted insertion into the row's best-KK list (KK = 5 covers the reference's beam width; the unused tail
          // slots stay -inf): 4 instead of 7 dependent compare-exchanges per candidate
          if (ep.topk_k <= 5) topk_insert_chunk<5>(x, cmax, gn0, tv, tix);
          else topk_insert_chunk<8>(x, cmax, gn0, tv, tix);
        }
        const int64_t gm = (int64_t)m0 + q * 32 + lane;
        if (gm < M) {
          const int64_t tn = tile % ntn;           // partials are [tile][k][row]: coalesced here and in the merge
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            ep.topk_val[(tn * 8 + i) * M + gm] = tv[i];
            ep.topk_idx[(tn * 8 + i) * M + gm] = tix[i];
          }
          ep.lse_max[tn * M + gm] = mx;
          ep.lse_sum[tn * M + gm] = sm;
        }
      } else {
      float lmx = -INFINITY, lsm = 0.f;           // K-D: running log-sum-exp of this row over the tile (optional)
#pragma unroll 1
      for (int c = 0; c < PG_BN / 32; ++c) {
        const int gn0 = n0 + c * 32;
        if (gn0 >= N) break;
        uint32_t v[32];
        tmem_ld32(trow + (uint32_t)(c * 32), v);
        if (ep.lse_max) {
          float cmax = -INFINITY;
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gn = gn0 + j;
            x[j] = (gn < N) ? __uint_as_float(v[j]) + sbias[c * 32 + j] : -INFINITY;
            cmax = fmaxf(cmax, x[j]);
          }
          const float nmx = fmaxf(lmx, cmax);
          float part = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) part += __expf(x[j] - nmx);
          lsm = lsm * __expf(lmx - nmx) + part;
          lmx = nmx;
        }
        // The warp's 32x32 chunk goes row-major through shared memory (thread = accumulator row writes 8 float4; pitch 36
        // keeps both sides conflict-free) and leaves as float4 per lane: 8 lanes cover the 32 columns of one row, so one
        // store instruction writes four 128-byte row segments -- 8 store instructions per chunk instead of 32, all
        // branches hoisted.  (The previous loop, one 4-byte store per lane and row with the null / beta checks inside, ran
        // at ~190 cycles per row on the single epilogue warp of each scheduler: 26 us per 128x256 tile, three times the
        // 9 us mainloop -- every big GEMM was epilogue bound, tools/gemm_probe.py.)
        {
          float4* srow = reinterpret_cast<float4*>(stg + lane * 36);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            srow[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                  __uint_as_float(v[4 * j + 3]));
        }
        __syncwarp();
        const int64_t gm0 = (int64_t)m0 + q * 32;
        const int cg = lane & 7, rsub = lane >> 3;
        const int gn = gn0 + cg * 4;
        const float4 b4 = *reinterpret_cast<const float4*>(sbias + c * 32 + cg * 4);
        float* Cp = ep.C;
        __nv_bfloat16* Cbp = ep.Cb;
        int64_t ldc = ep.ldc;
        const int64_t ldcb = ep.ldcb;
        const float beta = ep.beta;
        const int f16 = ep.cb_f16;
        int gnc = gn, lim = N;                      // column inside the fp32 destination / first column that is not mine
        if (ep.C2) {                                // two destinations split at column split_n (a multiple of 4)
          if (gn >= ep.split_n) { Cp = ep.C2; ldc = ep.ldc2; gnc = gn - ep.split_n; Cbp = nullptr; }   // 16-bit copy: first part only
          else lim = ep.split_n;
        }
        const bool c_vec = Cp && (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(Cp + gnc) & 15u) == 0 && gn + 3 < lim;
        const bool cb_vec = Cbp && (ldcb & 3) == 0 && (reinterpret_cast<uintptr_t>(Cbp) & 7u) == 0 && gn + 3 < N;
        float4 r4[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const float4 t4 = *reinterpret_cast<const float4*>(stg + (it * 4 + rsub) * 36 + cg * 4);
          r4[it] = make_float4(t4.x + b4.x, t4.y + b4.y, t4.z + b4.z, t4.w + b4.w);
        }
        if (gn < N) {
          if (Cp) {
            if (c_vec) {
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int64_t gm = gm0 + it * 4 + rsub;
                if (gm < M) {
                  float4* dst = reinterpret_cast<float4*>(Cp + gm * ldc + gnc);
                  if (beta != 0.f) {
                    const float4 o4 = *dst;
                    r4[it].x += beta * o4.x; r4[it].y += beta * o4.y; r4[it].z += beta * o4.z; r4[it].w += beta * o4.w;
                  }
                  *dst = r4[it];
                }
              }
            } else {
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int64_t gm = gm0 + it * 4 + rsub;
                if (gm < M) {
                  float* dst = Cp + gm * ldc + gnc;
                  float rr[4] = {r4[it].x, r4[it].y, r4[it].z, r4[it].w};
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (gn + e < lim) {
                      if (beta != 0.f) rr[e] += beta * dst[e];
                      dst[e] = rr[e];
                    }
                  r4[it] = make_float4(rr[0], rr[1], rr[2], rr[3]);
                }
              }
            }
          }
          if (Cbp) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int64_t gm = gm0 + it * 4 + rsub;
              if (gm < M) {
                __nv_bfloat16* dstb = Cbp + gm * ldcb + gn;
                if (cb_vec) {
                  const __nv_bfloat16 h0 = to_b16(r4[it].x, f16), h1 = to_b16(r4[it].y, f16), h2 = to_b16(r4[it].z, f16),
                                      h3 = to_b16(r4[it].w, f16);
                  uint2 pk;
                  pk.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                  pk.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
                  *reinterpret_cast<uint2*>(dstb) = pk;
                } else {
                  const float rr[4] = {r4[it].x, r4[it].y, r4[it].z, r4[it].w};
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (gn + e < N) dstb[e] = to_b16(rr[e], f16);
                }
              }
            }
          }
        }
        __syncwarp();
      }
      if (ep.lse_max) {
        const int64_t gm = (int64_t)m0 + q * 32 + lane;
        if (gm < M) {
          ep.lse_max[(int64_t)(tile % ntn) * M + gm] = lmx;
          ep.lse_sum[(int64_t)(tile % ntn) * M + gm] = lsm;
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(buf)) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) {                               // the peer's last commits / multicasts target my shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
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

static unsigned long long* g_gemm_prof = nullptr;
static int g_gemm_prof_m = 0, g_gemm_prof_n = 0, g_gemm_prof_k = 0;

template <int BN, int STAGES, int MODE>
static int launch_tc(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, TcEpilogue ep,
                     int splits, bool pdl, cudaStream_t st) {
  ep.prof = (g_gemm_prof && M == g_gemm_prof_m && N == g_gemm_prof_n && K == g_gemm_prof_k) ? g_gemm_prof : nullptr;
  CUtensorMap ma, mb;
  MVC_TRY(get_tensor_map(A, M, K, lda, TC_BM, &ma));
  MVC_TRY(get_tensor_map(B, N, K, ldb, BN, &mb));
  auto kern = gemm_bf16_tc_kernel<BN, STAGES, MODE>;
  constexpr int smem = TcSmem<BN, STAGES>::TOTAL;
  static bool configured = false;
  if (!configured) {
    MVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const dim3 grid((unsigned)cdiv(N, BN), (unsigned)cdiv(M, TC_BM), (unsigned)splits);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (splits > 1) {           // the S K-splits of one output tile form a thread-block cluster
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = (unsigned)splits;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  ProfScope prof(PK_GEMM_TC, M, N, K, st);
  MVC_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, M, N, K, ep));
  MVC_LAUNCH_CHECK();
  return 0;
}

// How many clusters of `cs` CTAs of this kernel can be resident at once (GPC packing makes this less than
// 148 / cs: e.g. 15 clusters of 8).  Cached per (kernel, cluster size).
template <int BN, int STAGES, int MODE>
static int max_active_clusters(int cs) {
  static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (cs <= 1) return kNumSMs;
  if (cache[cs] > 0) return cache[cs];
  auto kern = gemm_bf16_tc_kernel<BN, STAGES, MODE>;
  constexpr int smem = TcSmem<BN, STAGES>::TOTAL;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1, 1, (unsigned)cs);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)cs;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = kNumSMs / cs - 3;       // conservative guess
  }
  cache[cs] = n;
  return n;
}

// multicast-pair variant (plain epilogue, K-major operands): 74 clusters of 2 CTAs
static bool mc_enabled() {
  static int v = -1;
  if (v < 0) {
    // opt-in: measured on B200 at the C2 shapes the pair variant is NOT faster (P GEMM 94 vs 91 us, dW_c 49 vs 42 us):
    // L2 throughput sits at 17 % of peak in the 1-CTA kernel, so halving the L2 reads of B buys nothing
    const char* e = getenv("MVC_B200_GEMM_MULTICAST");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
static int launch_tc_persist_mc(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, TcEpilogue ep,
                                bool pdl, cudaStream_t st) {
  CUtensorMap ma, mb;
  MVC_TRY(get_tensor_map(A, M, K, lda, TC_BM, &ma));
  MVC_TRY(get_tensor_map(B, N, K, ldb, PG_BN / 2, &mb));           // each CTA of a pair fetches half of the B tile
  auto kern = gemm_bf16_tc_persist_kernel<TC_MODE_PLAIN, 0, 0, 1>;
  static bool configured = false;
  if (!configured) {
    MVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PG_SMEM));
    configured = true;
  }
  const int64_t pairs = cdiv(cdiv(M, TC_BM), 2) * cdiv(N, PG_BN);
  const int64_t max_clusters = kNumSMs / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * (pairs < max_clusters ? pairs : max_clusters)));
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = PG_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  ProfScope prof(PK_GEMM_TC, M, N, K, st);
  MVC_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, M, N, K, ep));
  MVC_LAUNCH_CHECK();
  return 0;
}

template <int MODE, int A_MN = 0, int B_MN = 0>
static int launch_tc_persist(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, TcEpilogue ep,
                             bool pdl, cudaStream_t st) {
  if constexpr (MODE == TC_MODE_PLAIN && !A_MN && !B_MN) {
    // two or more tile rows and enough tiles to fill the machine in pairs: share every B tile between two CTAs
    if (mc_enabled() && cdiv(M, TC_BM) >= 2 && cdiv(M, TC_BM) * cdiv(N, PG_BN) >= 64)
      return launch_tc_persist_mc(M, N, K, A, lda, B, ldb, ep, pdl, st);
  }
  CUtensorMap ma, mb;
  // K-major operand [rows = M|N, cols = K], box {64 k, rows}; MN-major operand stored [K, M|N], box {64 mn, 64 k}
  if (A_MN) MVC_TRY(get_tensor_map(A, K, M, lda, 64, &ma));
  else MVC_TRY(get_tensor_map(A, M, K, lda, TC_BM, &ma));
  if (B_MN) MVC_TRY(get_tensor_map(B, K, N, ldb, 64, &mb));
  else MVC_TRY(get_tensor_map(B, N, K, ldb, PG_BN, &mb));
  static bool configured = false;
  if (!configured) {
    MVC_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_persist_kernel<MODE, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  PG_SMEM));
    configured = true;
  }
  const int64_t tiles = cdiv(M, TC_BM) * cdiv(N, PG_BN);
  // Optional balanced grid (MVC_B200_GEMM_BALANCED=1): the kernel takes ceil(tiles / SMs) rounds whatever the grid; the
  // SMALLEST grid with that many rounds gives every CTA the same number of tiles and leaves the other SMs to the
  // neighbouring streams (192 tiles: 96 CTAs x 2 instead of 148 CTAs of which 44 run two).  Same stand-alone duration;
  // inside the train step it measured 1.7 % SLOWER (137.1 k vs 139.5 k samples/s): the neighbours are cluster launches
  // that do better on the SMs the one-tile CTAs free half-way than on scattered idle SMs.  Off by default.
  static int balanced = -1;
  if (balanced < 0) {
    const char* e = getenv("MVC_B200_GEMM_BALANCED");
    balanced = (e && e[0] == '1') ? 1 : 0;
  }
  const int64_t rounds = cdiv(tiles, kNumSMs);
  const int64_t grid = balanced ? cdiv(tiles, rounds) : (tiles < kNumSMs ? tiles : kNumSMs);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = PG_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  ProfScope prof(PK_GEMM_TC, M, N, K, st);
  MVC_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_persist_kernel<MODE, A_MN, B_MN>, ma, mb, M, N, K, ep));
  MVC_LAUNCH_CHECK();
  return 0;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MVC_B200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int tc_gemm(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, const TcEpilogue& ep_in,
            int flags, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  TcEpilogue ep = ep_in;
  MVC_CHECK(A && B, "tcgen05 GEMM: null operand");
  MVC_CHECK(K > 0, "tcgen05 GEMM: K must be positive");
  MVC_CHECK(lda % 8 == 0 && ldb % 8 == 0, "tcgen05 GEMM: lda (%lld) / ldb (%lld) must be multiples of 8",
            (long long)lda, (long long)ldb);
  MVC_CHECK((reinterpret_cast<uintptr_t>(A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(B) & 15u) == 0,
            "tcgen05 GEMM: operands must be 16-byte aligned");
  const bool pdl = (flags & TC_FLAG_PDL) && pdl_enabled();
  ep.b_const = (flags & TC_FLAG_B_CONST) ? 1 : 0;
  const int64_t mt = cdiv(M, TC_BM);
  const int nkb = (int)cdiv(K, TC_BK);
  const bool can_split = !(flags & TC_FLAG_NO_SPLIT);
  // K-splits per output tile = cluster size: a power of two <= 8 (portable cluster limit), <= #K blocks,
  // such that all clusters are resident in ONE wave (the hardware packs clusters per GPC)
  auto splits_for = [&](int64_t tiles, auto max_clusters) {
    if (!can_split || tiles >= 96) return 1;
    int s = 1;
    while (s < 8 && 2 * s <= nkb && tiles <= max_clusters(2 * s)) s *= 2;
    return s;
  };
  if (ep.mode == TC_MODE_CELL) {
    MVC_CHECK(N % 128 == 0 && N == 4 * ep.H, "fused LSTM-cell epilogue needs N == 4H with H %% 32 == 0 (N=%d H=%d)", N, ep.H);
    const int s = splits_for(mt * (N / 128), max_active_clusters<128, 6, TC_MODE_CELL>);
    // several waves of tiles (beam search: M = width x B): a 3-stage ring leaves room for TWO resident CTAs per SM, so
    // one CTA's cell epilogue overlaps the other's MMAs
    if (s == 1 && mt * (N / 128) > kNumSMs) return launch_tc<128, 3, TC_MODE_CELL>(M, N, K, A, lda, B, ldb, ep, 1, pdl, st);
    return launch_tc<128, 6, TC_MODE_CELL>(M, N, K, A, lda, B, ldb, ep, s, pdl, st);
  }
  // big GEMMs (>= half a wave of 128x256 tiles): persistent kernel with double-buffered TMEM accumulators
  if (flags & (TC_FLAG_A_MN | TC_FLAG_B_MN)) {
    // transposed operands (weight-gradient GEMMs: A stored [K,M] and / or B stored [K,N]): persistent kernel only
    MVC_CHECK(ep.mode == TC_MODE_PLAIN, "tcgen05 GEMM: MN-major operands support the plain epilogue only");
    const bool amn = flags & TC_FLAG_A_MN, bmn = flags & TC_FLAG_B_MN;
    if (amn && bmn) return launch_tc_persist<TC_MODE_PLAIN, 1, 1>(M, N, K, A, lda, B, ldb, ep, pdl, st);
    if (amn) return launch_tc_persist<TC_MODE_PLAIN, 1, 0>(M, N, K, A, lda, B, ldb, ep, pdl, st);
    return launch_tc_persist<TC_MODE_PLAIN, 0, 1>(M, N, K, A, lda, B, ldb, ep, pdl, st);
  }
  if (ep.mode == TC_MODE_ARGMAX) {
    MVC_CHECK(ep.amax_val && ep.amax_idx, "tcgen05 GEMM arg-max epilogue: null partial buffers");
    return launch_tc_persist<TC_MODE_ARGMAX>(M, N, K, A, lda, B, ldb, ep, pdl, st);
  }
  if (ep.mode == TC_MODE_TOPK) {
    MVC_CHECK(ep.topk_val && ep.topk_idx && ep.lse_max && ep.lse_sum, "tcgen05 GEMM top-k epilogue: null partial buffers");
    return launch_tc_persist<TC_MODE_TOPK>(M, N, K, A, lda, B, ldb, ep, pdl, st);
  }
  if (ep.C2) {
    MVC_CHECK(ep.mode == TC_MODE_PLAIN && !(flags & (TC_FLAG_A_MN | TC_FLAG_B_MN)) && ep.split_n % 4 == 0 && ep.split_n > 0 &&
                  ep.split_n < N && (ep.C || ep.Cb) && (!ep.Cb || ep.ldcb >= ep.split_n),
              "tcgen05 GEMM: the two-destination epilogue needs the plain mode, K-major operands and split_n %% 4 == 0");
    return launch_tc_persist<TC_MODE_PLAIN>(M, N, K, A, lda, B, ldb, ep, pdl, st);
  }
  static int persist_min_tiles = -1;
  if (persist_min_tiles < 0) {
    const char* e = getenv("MVC_B200_PERSIST_MIN_TILES");
    persist_min_tiles = e ? atoi(e) : kNumSMs / 2;
  }
  if (mt * cdiv(N, PG_BN) >= persist_min_tiles && K >= 2 * TC_BK)
    return launch_tc_persist<TC_MODE_PLAIN>(M, N, K, A, lda, B, ldb, ep, pdl, st);
  const int64_t t128 = mt * cdiv(N, 128), t64 = mt * cdiv(N, 64), t32 = mt * cdiv(N, 32);
  const int s128 = splits_for(t128, max_active_clusters<128, 6, TC_MODE_PLAIN>);
  const int s64 = splits_for(t64, max_active_clusters<64, 8, TC_MODE_PLAIN>);
  const int s32 = splits_for(t32, max_active_clusters<32, 10, TC_MODE_PLAIN>);
  // The kernel is a serial chain of k-blocks per CTA (latency bound at these sizes), so pick the tile width whose
  // CTAs run the FEWEST k-blocks: waves x ceil(nkb / splits), weighted by the per-k-block cost of the tile width
  // (a narrower B tile is a little cheaper, but re-reads A more often).  Ties go to the wider tile.
  // A split adds the park + DSMEM reduction of the partial tile, proportional to the tile width (measured: ~6 us at
  // BN = 128 vs ~1.5 us at BN = 32, in units of one k-block ~ 0.5 us).
  auto cost = [&](int64_t tiles, int s, double w, double red) {
    const int64_t waves = s > 1 ? 1 : cdiv(tiles, kNumSMs);
    return (double)(waves * cdiv(nkb, s)) * w + (s > 1 ? red : 0.0);
  };
  const double c128 = cost(t128, s128, 1.0, 12.0), c64 = cost(t64, s64, 0.85, 6.0), c32 = cost(t32, s32, 0.75, 3.0);
  if (c128 <= c64 && c128 <= c32) return launch_tc<128, 6, TC_MODE_PLAIN>(M, N, K, A, lda, B, ldb, ep, s128, pdl, st);
  if (c64 <= c32) return launch_tc<64, 8, TC_MODE_PLAIN>(M, N, K, A, lda, B, ldb, ep, s64, pdl, st);
  return launch_tc<32, 10, TC_MODE_PLAIN>(M, N, K, A, lda, B, ldb, ep, s32, pdl, st);
}

// token[m] = column of the best partial of row m (tiles ascend: strict > keeps the lowest index)
__global__ void argmax_partials_kernel(const float* __restrict__ pval, const int* __restrict__ pidx, int M, int ntn,
                                       int64_t* __restrict__ out, int64_t* __restrict__ out2, int64_t out2_ld) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float best = -INFINITY;
  int bi = 0;
  for (int t = 0; t < ntn; ++t) {
    const float v = pval[(int64_t)t * M + m];
    if (v > best) { best = v; bi = pidx[(int64_t)t * M + m]; }
  }
  if (out) out[m] = bi;
  if (out2) out2[(int64_t)m * out2_ld] = bi;
}

int tc_gemm_argmax(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias,
                   float* pval, int* pidx, int64_t* out, int64_t* out2, int64_t out2_ld, int flags, cudaStream_t st,
                   const TcAux* aux) {
  TcEpilogue ep{};
  ep.mode = TC_MODE_ARGMAX;
  ep.bias = bias;
  ep.amax_val = pval;
  ep.amax_idx = pidx;
  int Nt = N;
  if (aux) {
    MVC_CHECK(aux->C && aux->n0 % PG_BN == 0 && aux->n0 >= N && aux->cols > 0, "fused arg-max: bad auxiliary block");
    ep.aux_C = aux->C; ep.aux_ld = aux->ld; ep.aux_n0 = aux->n0; ep.n_main = N;
    Nt = aux->n0 + aux->cols;
  }
  MVC_TRY(tc_gemm(M, Nt, K, A, lda, B, ldb, ep, flags, st));
  argmax_partials_kernel<<<(unsigned)cdiv(M, 128), 128, 0, st>>>(pval, pidx, M, (int)cdiv(N, PG_BN), out, out2, out2_ld);
  MVC_LAUNCH_CHECK();
  return 0;
}
int tc_gemm_argmax_tiles(int N) { return (int)cdiv(N, PG_BN); }

// per row: log-sum-exp from the per-tile (max, sum) pairs and the `width` best log-probs from the per-tile top-8
// lists (tiles ascend, lists are sorted: a strict > keeps the lowest token on ties)
__global__ void topk_partials_kernel(const float* __restrict__ tval, const int* __restrict__ tidx,
                                     const float* __restrict__ lmax, const float* __restrict__ lsum, int M, int ntn,
                                     int width, float* __restrict__ cand_val, int* __restrict__ cand_idx) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float mx = -INFINITY;
  for (int t = 0; t < ntn; ++t) mx = fmaxf(mx, lmax[(int64_t)t * M + m]);
  float sm = 0.f;
  for (int t = 0; t < ntn; ++t) sm += lsum[(int64_t)t * M + m] * expf(lmax[(int64_t)t * M + m] - mx);
  const float lse = mx + logf(sm);
  float bv[8];
  int bi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { bv[i] = -INFINITY; bi[i] = 0x7fffffff; }
  for (int t = 0; t < ntn; ++t) {
    for (int k = 0; k < 8; ++k) {
      const float x = tval[((int64_t)t * 8 + k) * M + m];
      if (!(x > bv[7])) break;                  // the list is sorted: nothing better follows in this tile
      bv[7] = x; bi[7] = tidx[((int64_t)t * 8 + k) * M + m];
#pragma unroll
      for (int i = 7; i > 0; --i) {
        if (bv[i] > bv[i - 1]) {
          const float tf = bv[i]; bv[i] = bv[i - 1]; bv[i - 1] = tf;
          const int tn = bi[i]; bi[i] = bi[i - 1]; bi[i - 1] = tn;
        }
      }
    }
  }
  for (int k = 0; k < width; ++k) {
    cand_val[(int64_t)m * width + k] = bv[k] - lse;
    cand_idx[(int64_t)m * width + k] = bi[k];
  }
}

// K-D: x[r,:] (logits written by the GEMM) -> log-probs in place, the row's log-sum-exp coming from the per-tile
// partials of the GEMM epilogue (one read + one write of the row instead of three reads + one write)
__global__ void logsoftmax_finish_kernel(float* __restrict__ x, int64_t ld, int M, int V, int ntn,
                                         const float* __restrict__ lmax, const float* __restrict__ lsum) {
  __shared__ float s_lse;
  const int64_t m = blockIdx.x;
  if (threadIdx.x < 32) {
    float mx = -INFINITY;
    for (int t = threadIdx.x; t < ntn; t += 32) mx = fmaxf(mx, lmax[(int64_t)t * M + m]);
    mx = warp_max(mx);
    float sm = 0.f;
    for (int t = threadIdx.x; t < ntn; t += 32) sm += lsum[(int64_t)t * M + m] * expf(lmax[(int64_t)t * M + m] - mx);
    sm = warp_sum(sm);
    if (threadIdx.x == 0) s_lse = mx + logf(sm);
  }
  __syncthreads();
  const float lse = s_lse;
  float* row = x + m * ld;
  for (int v = threadIdx.x; v < V; v += blockDim.x) row[v] -= lse;
}

int tc_gemm_logsoftmax(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias,
                       float* C, int64_t ldc, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  const int ntn = (int)cdiv(N, PG_BN);
  const bool fused = cdiv(M, TC_BM) * ntn >= kNumSMs / 2 && K >= 2 * TC_BK && scratch &&
                     scratch_bytes >= (size_t)2 * M * ntn * sizeof(float);
  TcEpilogue ep{};
  ep.mode = TC_MODE_PLAIN;
  ep.C = C; ep.ldc = ldc; ep.bias = bias;
  if (fused) {
    ep.lse_max = static_cast<float*>(scratch);
    ep.lse_sum = ep.lse_max + (size_t)M * ntn;
  }
  MVC_TRY(tc_gemm(M, N, K, A, lda, B, ldb, ep, 0, st));
  if (fused) {
    logsoftmax_finish_kernel<<<(unsigned)M, 256, 0, st>>>(C, ldc, M, N, ntn, ep.lse_max, ep.lse_sum);
    MVC_LAUNCH_CHECK();
    return 0;
  }
  return mvc_log_softmax_rows(C, M, N, nullptr, st);      // small problems: separate row kernel (needs ldc == N)
}

size_t tc_gemm_topk_scratch_bytes(int M, int N) { return (size_t)M * cdiv(N, PG_BN) * (8 + 8 + 2) * 4; }

int tc_gemm_topk(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias,
                 void* scratch, int width, float* cand_val, int* cand_idx, int flags, cudaStream_t st, const TcAux* aux) {
  MVC_CHECK(width >= 1 && width <= 8, "fused top-k: width %d not in [1,8]", width);
  const int ntn = (int)cdiv(N, PG_BN);
  float* tval = static_cast<float*>(scratch);
  int* tidx = reinterpret_cast<int*>(tval + (size_t)M * ntn * 8);
  float* lmax = reinterpret_cast<float*>(tidx + (size_t)M * ntn * 8);
  float* lsum = lmax + (size_t)M * ntn;
  TcEpilogue ep{};
  ep.mode = TC_MODE_TOPK;
  ep.bias = bias;
  ep.topk_val = tval; ep.topk_idx = tidx; ep.lse_max = lmax; ep.lse_sum = lsum;
  ep.topk_k = width;
  int Nt = N;
  if (aux) {
    MVC_CHECK(aux->C && aux->n0 % PG_BN == 0 && aux->n0 >= N && aux->cols > 0, "fused top-k: bad auxiliary block");
    ep.aux_C = aux->C; ep.aux_ld = aux->ld; ep.aux_n0 = aux->n0; ep.n_main = N;
    Nt = aux->n0 + aux->cols;
  }
  MVC_TRY(tc_gemm(M, Nt, K, A, lda, B, ldb, ep, flags, st));
  topk_partials_kernel<<<(unsigned)cdiv(M, 128), 128, 0, st>>>(tval, tidx, lmax, lsum, M, ntn, width, cand_val, cand_idx);
  MVC_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvc

using namespace mvc;

extern "C" int mvc_vocab_argmax_bf16(int M, int V, int K, const void* h, int64_t ldh, const void* out_w, int64_t ldw,
                                     const float* out_b, void* workspace, size_t workspace_bytes, int64_t* ids,
                                     void* stream) {
  MVC_CHECK(h && out_w && workspace && ids, "mvc_vocab_argmax_bf16: null argument");
  const size_t need = (size_t)M * tc_gemm_argmax_tiles(V) * 8;
  MVC_CHECK(workspace_bytes >= need, "mvc_vocab_argmax_bf16: workspace %zu < %zu", workspace_bytes, need);
  float* pval = static_cast<float*>(workspace);
  int* pidx = reinterpret_cast<int*>(pval + (size_t)M * tc_gemm_argmax_tiles(V));
  return tc_gemm_argmax(M, V, K, h, ldh, out_w, ldw, out_b, pval, pidx, ids, nullptr, 0, 0, (cudaStream_t)stream);
}

extern "C" int mvc_vocab_aux_row0(int V) { return tc_aux_row0(V); }

extern "C" int mvc_vocab_argmax_wq_bf16(int M, int V, int K, int A, const void* h, int64_t ldh, const void* w_ext,
                                        int64_t ldw, const float* out_b, void* workspace, size_t workspace_bytes,
                                        int64_t* ids, float* wq, void* stream) {
  MVC_CHECK(h && w_ext && workspace && ids && wq && A > 0, "mvc_vocab_argmax_wq_bf16: bad argument");
  const size_t need = (size_t)M * tc_gemm_argmax_tiles(V) * 8;
  MVC_CHECK(workspace_bytes >= need, "mvc_vocab_argmax_wq_bf16: workspace %zu < %zu", workspace_bytes, need);
  float* pval = static_cast<float*>(workspace);
  int* pidx = reinterpret_cast<int*>(pval + (size_t)M * tc_gemm_argmax_tiles(V));
  const TcAux aux{tc_aux_row0(V), A, wq, A};
  return tc_gemm_argmax(M, V, K, h, ldh, w_ext, ldw, out_b, pval, pidx, ids, nullptr, 0, 0, (cudaStream_t)stream, &aux);
}

extern "C" int mvc_gemm_bf16_ex(int M, int N, int K, const void* A, int64_t lda, int a_transposed, const void* B, int64_t ldb,
                                int b_transposed, float beta, float* C, int64_t ldc, const float* bias, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  MVC_CHECK(A && B && C, "mvc_gemm_bf16_ex: null operand");
  if (!a_transposed && !b_transposed) return mvc_gemm_bf16(M, N, K, A, lda, B, ldb, beta, C, ldc, bias, nullptr, 0, stream);
  TcEpilogue ep{};
  ep.mode = TC_MODE_PLAIN;
  ep.beta = beta; ep.C = C; ep.ldc = ldc; ep.bias = bias;
  return tc_gemm(M, N, K, A, lda, B, ldb, ep, (a_transposed ? TC_FLAG_A_MN : 0) | (b_transposed ? TC_FLAG_B_MN : 0),
                 (cudaStream_t)stream);
}

extern "C" size_t mvc_vocab_topk_workspace_bytes(int M, int V) { return tc_gemm_topk_scratch_bytes(M, V); }

extern "C" int mvc_vocab_topk_bf16(int M, int V, int K, const void* h, int64_t ldh, const void* out_w, int64_t ldw,
                                   const float* out_b, int width, void* workspace, size_t workspace_bytes,
                                   float* cand_logp, int* cand_idx, void* stream) {
  MVC_CHECK(h && out_w && workspace && cand_logp && cand_idx, "mvc_vocab_topk_bf16: null argument");
  MVC_CHECK(workspace_bytes >= tc_gemm_topk_scratch_bytes(M, V), "mvc_vocab_topk_bf16: workspace %zu < %zu",
            workspace_bytes, tc_gemm_topk_scratch_bytes(M, V));
  return tc_gemm_topk(M, V, K, h, ldh, out_w, ldw, out_b, workspace, width, cand_logp, cand_idx, 0, (cudaStream_t)stream);
}

extern "C" int mvc_gemm_bf16(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, float beta,
                             float* C, int64_t ldc, const float* bias, void* Cb, int64_t ldcb, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  MVC_CHECK(A && B && (C || Cb), "mvc_gemm_bf16: null operand");
  TcEpilogue ep{};
  ep.mode = TC_MODE_PLAIN;
  ep.beta = beta; ep.C = C; ep.ldc = ldc; ep.bias = bias; ep.Cb = (__nv_bfloat16*)Cb; ep.ldcb = ldcb;
  return tc_gemm(M, N, K, A, lda, B, ldb, ep, 0, (cudaStream_t)stream);
}

extern "C" int mvc_debug_set_gemm_prof(unsigned long long* dev_buf, int M, int N, int K) {
  mvc::g_gemm_prof = dev_buf;
  mvc::g_gemm_prof_m = M; mvc::g_gemm_prof_n = N; mvc::g_gemm_prof_k = K;
  return 0;
}
