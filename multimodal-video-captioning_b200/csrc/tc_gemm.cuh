// tc_gemm.cuh -- internal interface of the tcgen05 GEMM (gemm_tc.cu): epilogue descriptor + launch flags.
#pragma once
#include "common.cuh"

namespace mvc {

enum { TC_MODE_PLAIN = 0, TC_MODE_CELL = 1, TC_MODE_ARGMAX = 2, TC_MODE_TOPK = 3 };
enum {
  TC_FLAG_PDL = 1,       // launch with programmatic stream serialization (the kernel waits on griddepcontrol itself)
  TC_FLAG_B_CONST = 2,   // operand B was complete before the previous kernel started: prefetch it before the dependency wait
  TC_FLAG_NO_SPLIT = 4,  // never split K
  TC_FLAG_A_MN = 8,      // A is stored transposed: [K, M] row-major (ld = lda), consumed as an MN-major UMMA operand
  TC_FLAG_B_MN = 16      // B is stored transposed: [K, N] row-major (ld = ldb)
};

struct TcEpilogue {
  int mode;
  int b_const;
  // plain: C = acc (+ bias[n]) (+ beta*C), optional bf16 copy
  float beta;
  float* C;
  int64_t ldc;
  const float* bias;          // also the (permuted) gate bias in cell mode, may be null
  __nv_bfloat16* Cb;
  int64_t ldcb;
  // plain mode, persistent kernel: optional SECOND fp32 destination for the columns >= split_n (C2[m, n - split_n], ld ldc2;
  // split_n a multiple of 4): one GEMM can feed two tensors that share the A operand.  C / Cb (either may be null) then
  // receive the columns < split_n only.
  float* C2;
  int64_t ldc2;
  int split_n;
  int cb_f16;                 // the 16-bit copy is IEEE fp16 (saturating) instead of bf16: 3 more mantissa bits for
                              // the projected keys of the persistent recurrence kernels (recur2.cuh)
  // fused row arg-max (mode == TC_MODE_ARGMAX): per (row, 256-column tile) partial maximum of acc + bias
  float* amax_val;            // [ceil(N/256), M]
  int* amax_idx;              // [ceil(N/256), M]
  // fused top-8 + online log-sum-exp (mode == TC_MODE_TOPK): per (row, 256-column tile)
  float* topk_val;            // [ceil(N/256), 8, M]
  int* topk_idx;              // [ceil(N/256), 8, M]
  int topk_k;                 // entries of the per-tile lists actually needed (<= 8)
  float* lse_max;             // [ceil(N/256), M]
  float* lse_sum;             // [ceil(N/256), M]
  // arg-max / top-k modes, optional auxiliary column block: B rows [aux_n0, N) (aux_n0 a multiple of 256, >= n_main)
  // are a second weight matrix applied to the same A; its product is stored plainly to aux_C[m, n - aux_n0] while
  // columns [0, n_main) go through the arg-max / top-k reduction (rows [n_main, aux_n0) of B are padding).
  // Used to fold the attention query projection W.h of the NEXT decode step into this step's vocabulary GEMM.
  unsigned long long* prof;   // debug: 8 %globaltimer stamps per CTA of the one-tile kernel (mvc_debug_set_gemm_prof)
  float* aux_C;
  int64_t aux_ld;
  int aux_n0;
  int n_main;
  // fused LSTM cell (mode == TC_MODE_CELL): N = 4H, columns permuted (j/16)*64 + gate*16 + j%16
  int H;
  const float* gx;            // [M,4H] hoisted input projection (permuted columns), ld gx_ld, may be null
  int64_t gx_ld;
  const float* embtab;        // [V,4H] gathered by tokens[m] (permuted columns), may be null
  const int64_t* tokens;
  const float* c_prev;        // [M,H] (null = zeros)
  float* act;                 // [M,4H] activated gates, permuted columns (may be null)
  float* c_out;               // [M,H]
  float* h32;                 // fp32 h (ld h_ld), may be null
  int64_t h_ld;
  float* h32b;                // second fp32 copy (ld h2_ld), may be null
  int64_t h2_ld;
  __nv_bfloat16* hb;          // bf16 h (ld hb_ld), may be null
  int64_t hb_ld;
};

int tc_gemm(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, const TcEpilogue& ep, int flags,
            cudaStream_t st);
// K-E: ids[m] = argmax_n (A[m,:] . B[n,:] + bias[n]) without materialising the logits (lowest index wins ties);
// pval / pidx: [M, tc_gemm_argmax_tiles(N)] scratch; out2 (optional) receives the same ids with stride out2_ld.
struct TcAux {          // auxiliary column block of the arg-max / top-k GEMMs (see TcEpilogue::aux_C)
  int n0, cols;
  float* C;
  int64_t ld;
};
inline int tc_aux_row0(int N) { return (N + 255) / 256 * 256; }   // first B row of the auxiliary block for N main rows
int tc_gemm_argmax(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias,
                   float* pval, int* pidx, int64_t* out, int64_t* out2, int64_t out2_ld, int flags, cudaStream_t st,
                   const TcAux* aux = nullptr);
int tc_gemm_argmax_tiles(int N);
// K-D: C = log_softmax(A . B^T + bias) row-wise: the GEMM epilogue keeps an online (max, sum exp) per row and tile,
// a finishing kernel subtracts the row's log-sum-exp in place.  scratch >= 2 * M * ceil(N/256) floats.
int tc_gemm_logsoftmax(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias,
                       float* C, int64_t ldc, void* scratch, size_t scratch_bytes, cudaStream_t st);
// K-E (beam): cand_val[m, k] / cand_idx[m, k] = the `width` (<= 8) largest log-softmax values of row m of
// A . B^T + bias and their columns, without materialising logits or log-probs.
size_t tc_gemm_topk_scratch_bytes(int M, int N);
int tc_gemm_topk(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias,
                 void* scratch, int width, float* cand_val, int* cand_idx, int flags, cudaStream_t st,
                 const TcAux* aux = nullptr);

// "Tile-interleaved" gate order used by the fused cell epilogues: hidden units are grouped in blocks of 16 and
// each block stores its four gates back to back -- column (j/16)*64 + gate*16 + j%16 is nn.LSTM row gate*H + j.
// Any 64-aligned window of 64k columns therefore holds i,f,g,o of 16k units (64-wide tiles in the persistent
// recurrence kernel, 128-wide tiles in the stand-alone fused GEMM).
constexpr int kGateU = 16;
__host__ __device__ inline int gate_col(int perm, int H, int gate, int j) {
  return perm ? (j >> 4) * 64 + gate * 16 + (j & 15) : gate * H + j;
}
// inverse: natural row (gate*H + j) of permuted column c.  H > 0: the tile-interleaved order above; H < 0 selects the
// UNIT-MAJOR order of the persistent recurrence kernels (recur2.cuh) for hidden size -H: column 4j + gate.
__host__ __device__ inline int gate_unperm(int H, int c) {
  if (H < 0) return (c & 3) * (-H) + (c >> 2);
  const int blk = c >> 6, gate = (c >> 4) & 3, u = c & 15;
  return gate * H + blk * 16 + u;
}
// column, relative to the start of its (64k-aligned) tile, of gate `gate` of the tile's local unit j
__host__ __device__ inline int gate_lcol(int gate, int j) { return (j >> 4) * 64 + gate * 16 + (j & 15); }

}  // namespace mvc
