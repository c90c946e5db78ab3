// recur2.cuh -- persistent SA-LSTM recurrence kernels, second generation ("projected keys"): interface + the
// device helpers the forward (recur2_fwd.cu) and backward (recur2_bwd.cu) kernels share.
//
// Algebra (exact): the LSTM input is [emb ; ctx] with ctx_s[b] = sum_t alpha_s[b,t] key[b,t] (features_captioning.py:
// 80-84), so the context half of the gate pre-activation is
//     W_c . ctx_s[b] = sum_t alpha_s[b,t] (W_c . key[b,t]) = sum_t alpha_s[b,t] P[b,t,:],   P = keys . W_c^T  [B*T, 4H]
// P is loop invariant: ONE [B*T, F] x [F, 4H] tcgen05 GEMM per sequence.  Row b's slab P[b] (T x 4H fp16 = 180 KB at
// T = 44) lives in the TENSOR MEMORY of the CTA that owns row b for the whole kernel, so
//   * the forward step never forms ctx: the owner accumulates sum_t alpha_t P[b,t,:] straight out of TMEM -- the same
//     T x 2048 multiply-adds the context sum cost before -- and the per-step tensor-core GEMM shrinks from
//     [128, F+H] x [F+H, 4H] (42 k-blocks) to the recurrent half h_s . W_hh^T (8 k-blocks);
//   * ctx no longer crosses CTAs, which removes the grid barrier between attention and gate GEMM: the recurrent GEMM
//     of step s+1 depends only on h_{s+1}, so it runs (all 128 CTAs, weights resident in shared memory) WHILE the row
//     owners compute the attention of step s+1.  Per step the chain is  cell -> max(attention, h-GEMM) -> cell.
//   * the backward needs no d[ctx;h] GEMM either: dalpha_s[b,t] = dG_s[b,:] . P[b,t,:] is local to the row owner
//     (dG of its own row, P in its TMEM); only dh += dG . W_hh (M=128, N=512, K=2048) stays on the tensor cores.
// ctx itself is only needed for dW_ih[:, E:] = sum_s dG_s^T ctx_s: recomputed after the loop from the saved alpha
// (one small batched kernel, r2_ctx_rows), so the weight-gradient GEMMs are unchanged.
//
// Gate columns use the UNIT-MAJOR order: column 4j + g holds gate g (i,f,g,o) of hidden unit j, so TMEM lane L of the
// owner holds all four gates of units 4L..4L+3 and the LSTM cell of a whole row runs in 128 threads without exchange.
//
// Synchronisation: no CTA-wide barriers inside the loop.  Warp groups communicate through
//   * cluster-scope mbarriers (remote arrive over DSMEM) for the four-row exchanges inside a cluster
//     (h rows -> query projection slices -> row owners; K-split partial tiles of the recurrent GEMM),
//   * monotonically increasing global progress counters (release-increments, four per direction -- one per cluster
//     rank, each in its own cache line): Y = "row owners have published h_s", polled by the TMA producer;
//     X = "the recurrent GEMM of step s is in global memory", polled by the row owners.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mvc {

constexpr int R2_THREADS = 384;      // warps 0-7 attention + cell; warps 8-11 recurrent GEMM (8: TMA, 9: MMA, all: epilogue)
constexpr int R2_CS = 4;             // cluster size = K splits = batch rows per cluster
constexpr int R2_H = 512;            // hidden size the kernels are specialised for (every reference config)
constexpr int R2_A = 256;            // attention bottleneck
constexpr int R2_AV = R2_A / 32;
constexpr int R2_R = 6;              // key-frame rounds per warp held in registers (T <= 48)
constexpr int R2_PCOL = 64;          // first TMEM column of the resident P slab (columns [0, 64) = accumulator)
constexpr int R2_WPT = 8;            // 32-bit words of P per frame per TMEM lane (16 bf16 gate columns)

struct Recur2FwdParams {
  int B, T, F, K, S;                 // K = F + H: row pitch of xh
  const __nv_bfloat16* P;            // [B*T, 4H] projected keys (fp16 bit patterns), unit-major gate columns
  const float* uk;                   // [B*T, A]  U.k (hoisted)
  const __nv_bfloat16* attW;         // [A, H]
  const float* att_b;                // [A]
  const float* att_w;                // [A]
  const float* gx;                   // [S*B, 4H] hoisted embedding projection + biases, unit-major columns
  __nv_bfloat16* xh;                 // [(S+1)*B, K]: h_s is written to slot s, columns [F, F+H) (slot 0 = zeros)
  float* c;                          // [(S+1), B, H]
  float* act;                        // [S, B, 4H] activated gates (unit-major) or null
  float* alpha;                      // [S, B, T]
  float* wq_out;                     // [S, B, A]
  float* out_hid;                    // [S+1, B, H] fp32 or null
  float* gh;                         // [4, 128, 4H] scratch: the four K-slice partials of h_s . W_hh^T of the current step
  unsigned* sync;                    // progress counters: Y[r] at [32 r], X[r] at [128 + 32 r] (r = cluster rank), zeroed by the launcher
  long long* prof;
};

struct Recur2BwdParams {
  int B, T, F, K, S;
  const __nv_bfloat16* P;            // [B*T, 4H]
  const float* uk;                   // [B*T, A]
  const float* att_b;
  const float* att_w;
  const float* act;                  // [S, B, 4H] unit-major
  const float* c;                    // [(S+1), B, H]
  const float* wq;                   // [S, B, A]
  const float* alpha;                // [S, B, T]
  const float* dh_ext;               // [S*B, H] gradient reaching h_{s+1} from outside the recurrence (or null)
  const __nv_bfloat16* attWT;        // [H, A] attention.W transposed
  float* dG;                         // [S*B, 4H] gate pre-activation gradients, unit-major
  __nv_bfloat16* dG_b;               // same in bf16 (A operand of dh += dG . W_hh and of the weight-gradient GEMMs)
  float* dwq;                        // [S*B, A]
  __nv_bfloat16* dwq_b;              // [S*B, A]
  float* duk;                        // [B*T, A] written once at the end
  float* dwpart;                     // [B, A]   written once at the end
  float* ghb;                        // [4, 128, H] scratch: the four K-slice partials of dG_s . W_hh of the current step
  unsigned* sync;                    // progress counters: Y[r] at [32 r], X[r] at [128 + 32 r]
  long long* prof;
};

bool recur2_supported(int B, int T, int F, int H, int A);
// sync_cleared: the caller has zeroed the 256 progress-counter words already (off the critical chain)
int recur2_fwd_launch(const Recur2FwdParams& p, const void* whh_um, int64_t ldw, cudaStream_t st, bool sync_cleared = false);
int recur2_bwd_launch(const Recur2BwdParams& p, const void* whhT_um, cudaStream_t st, bool sync_cleared = false);
// xh[(s*B + b), 0:F] = sum_t alpha[s,b,t] feats[b,t,:]  for s in [0,S): the context vectors the forward never formed
int r2_ctx_rows(const void* feats_bf16, const float* alpha, int B, int T, int F, int S, void* xh_bf16, int64_t ldx,
                cudaStream_t st);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
// Spin-wait bound (SM clock cycles; ~4 s at 1.9 GHz by default).  The kernels synchronise 128 co-resident CTAs by
// spinning, so they need the GPU to themselves: under a profiler that serialises / replays kernels, MPS or
// time-slicing, raise it with MVC_B200_SPIN_TIMEOUT_S=<seconds> (read once per process by the launchers) or take
// the launch chain with MVC_B200_PERSISTENT=0.  One copy per translation unit; set by r2_apply_spin_limit().
static __constant__ long long c_r2_spin_limit = 8000000000LL;
static inline int r2_apply_spin_limit() {
  static int done = 0;
  if (done) return 0;
  done = 1;
  const char* e = getenv("MVC_B200_SPIN_TIMEOUT_S");
  if (e && atof(e) > 0.0) {
    const long long cycles = (long long)(atof(e) * 2.0e9);
    MVC_CUDA(cudaMemcpyToSymbol(c_r2_spin_limit, &cycles, sizeof(cycles)));
  }
  return 0;
}
namespace r2 {

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_dsmem_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_dsmem_f32x2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_dsmem_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_dsmem_u32x2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// arrive (release, cluster scope) on an mbarrier of another CTA of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cl(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded waits: a protocol bug (or a GPU shared with another tenant so that the CTAs are not co-resident) sets the
// kernel's abort flag and traps instead of hanging the device.
__device__ __forceinline__ void mbar_wait_cl(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cl(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cl(bar, parity)) {
    if (clock64() - t0 > c_r2_spin_limit) {
      printf("mvc recur2: cluster mbarrier wait timed out (block %d thread %d bar %u)\n", blockIdx.x, threadIdx.x, bar);
      __trap();
    }
  }
}
// ---- grid-scope progress counters.  Four counters per direction, one per cluster rank and each in its own 128-byte
// line: the 128 release-increments of a step spread over four L2 slices instead of serialising in one (measured:
// one counter with 512 arrivals per step cost 1.5 us more per step than one with 128; release STORES to per-CTA
// flag words polled with 512-byte loads were slower still).  A poller reads the four lines with four loads in flight.
constexpr int R2_CNT_STRIDE = 32;                   // unsigned words between counters (128 bytes)
__device__ __forceinline__ void signal_counter(unsigned* counter) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
// one thread: wait until counter[r] >= target[r] for the four ranks
__device__ __forceinline__ void poll_counters4(const unsigned* counters, unsigned t0, unsigned t1, unsigned t2, unsigned t3) {
  const long long c0 = clock64();
  for (;;) {
    unsigned v0, v1, v2, v3;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v0) : "l"(counters) : "memory");
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v1) : "l"(counters + R2_CNT_STRIDE) : "memory");
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v2) : "l"(counters + 2 * R2_CNT_STRIDE) : "memory");
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v3) : "l"(counters + 3 * R2_CNT_STRIDE) : "memory");
    if (v0 >= t0 && v1 >= t1 && v2 >= t2 && v3 >= t3) return;
    if (clock64() - c0 > c_r2_spin_limit) {
      printf("mvc recur2: progress-counter wait timed out (block %d thread %d: %u %u %u %u of %u %u %u %u)\n", blockIdx.x,
             threadIdx.x, v0, v1, v2, v3, t0, t1, t2, t3);
      __trap();
    }
  }
}
template <int ID, int N>
__device__ __forceinline__ void named_bar() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

}  // namespace r2
#endif

}  // namespace mvc
