// recur.cuh -- interface of the persistent decoder-recurrence kernel (recur_fwd.cu).
#pragma once
#include "common.cuh"

namespace mvc {

struct RecurFwdParams {
  int B, T, F, H, A, K;           // K = F + H
  int S;                          // loop steps of the whole sequence (xh has S+1 slots)
  int s0, s1;                     // steps [s0, s1) run in this launch
  const __nv_bfloat16* feats;     // [B*T, F]  keys
  const float* uk;                // [B*T, A]  U.k (hoisted)
  const __nv_bfloat16* attW;      // [A, H]
  const float* att_b;             // [A]
  const float* att_w;             // [A]
  const float* gx;                // [S*B, 4H] hoisted input projection, tile-interleaved columns (or null)
  const float* embtab;            // [V, 4H]   embedding-table projection, tile-interleaved columns (or null)
  const int64_t* tokens;          // [S, B]    rows of embtab (when embtab)
  const float* cell_bias;         // [4H] or null
  __nv_bfloat16* xh;              // [(S+1)*B, K]  slot s = [ctx_s ; h_s]
  float* c;                       // [(S+1), B, H]
  float* act;                     // [S, B, 4H] or null
  float* alpha;                   // [S, B, T]
  float* wq_out;                  // [S, B, A]
  float* out_hid;                 // [S+1, B, H] fp32 (slot s+1 written at step s) or null
  unsigned* sync;                 // grid-barrier counter
  long long* prof;                // optional phase timestamps of CTA 0 (debug)
};

struct RecurBwdParams {
  int B, T, F, H, A, K;
  int S;
  const __nv_bfloat16* feats;     // [B*T, F]
  const float* uk;                // [B*T, A]
  const float* att_b;             // [A]
  const float* att_w;             // [A]
  const float* act;               // [S, B, 4H] activated gates (tile-interleaved columns), saved by the forward
  const float* c;                 // [(S+1), B, H]
  const float* wq;                // [S, B, A]  saved queries
  const float* alpha;             // [S, B, T]  saved attention weights
  const float* dh_ext;            // [S*B, H]   gradient reaching h_{s+1} from outside the recurrence (or null)
  const __nv_bfloat16* attWT;     // [H, A]     attention.W transposed
  float* dG;                      // [S*B, 4H]  gate pre-activation gradients (fp32, tile-interleaved columns)
  __nv_bfloat16* dG_b;            // [S*B, 4H]  same in bf16: the A operand of d[ctx;h] = dG . wcat
  float* dxh;                     // [B, K]     d[ctx_s ; h_s] of the current step
  float* dwq;                     // [S*B, A]
  __nv_bfloat16* dwq_b;           // [S*B, A]
  float* duk;                     // [B*T, A]   written once at the end
  float* dwpart;                  // [B, A]     written once at the end
  unsigned* sync;
  long long* prof;
};

bool recur_fwd_supported(int B, int T, int F, int H, int A);
bool recur_bwd_supported(int B, int T, int F, int H, int A);
int recur_bwd_launch(const RecurBwdParams& p, const void* wcatT, cudaStream_t st);
int recur_fwd_launch(const RecurFwdParams& p, const void* wcat, cudaStream_t st);

}  // namespace mvc
