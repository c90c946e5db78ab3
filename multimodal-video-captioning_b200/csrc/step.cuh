// step.cuh -- one "soft attention -> LSTM cell" recurrence step and its backward,
// shared by the caption decoder (features_captioning.py:77-89) and the local
// reconstructor (reconstructor.py:67-74).  Host-side launch sequences only.
//
// bf16 path (H % 32 == 0): three launches per forward step, chained with programmatic dependent
// launch so that each kernel's loop-invariant operand (weights / keys) streams in while its
// producer is still running:
//     wq GEMM (tcgen05, split-K)  ->  fused soft attention  ->  gate GEMM + LSTM cell epilogue
// fp32 path: FFMA GEMMs + separate cell kernel, fixed reduction order (the exact-ids path).
#pragma once
#include "tc_gemm.cuh"

namespace mvc {

static inline int pad8(int x) { return (x + 7) / 8 * 8; }
static inline const char* cptr(const void* p, int64_t elems, size_t es) { return (const char*)p + elems * es; }
static inline char* mptr(void* p, int64_t elems, size_t es) { return (char*)p + elems * es; }

static inline int gridn(int64_t n) {
  int64_t g = cdiv(n, 256);
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// C[M,N] = A[M,K] . B[N,K]^T (+ bias) (+ beta*C); operands K-contiguous in the compute dtype.
static inline int gemm_nt(int prec, int M, int N, int K, const void* A, int64_t lda, const void* Bm, int64_t ldb,
                          float beta, float* C, int64_t ldc, const float* bias, cudaStream_t st) {
  if (prec == MVC_BF16) return mvc_gemm_bf16(M, N, K, A, lda, Bm, ldb, beta, C, ldc, bias, nullptr, 0, st);
  return mvc_gemm_f32(M, N, K, 1.f, (const float*)A, lda, 1, (const float*)Bm, ldb, 1, beta, C, ldc, bias, st);
}

// plain tcgen05 GEMM inside a PDL chain (B operand loop invariant)
static inline int gemm_tc_chain(int M, int N, int K, const void* A, int64_t lda, const void* Bm, int64_t ldb, float beta,
                                float* C, int64_t ldc, void* Cb, int64_t ldcb, bool b_const, cudaStream_t st) {
  TcEpilogue ep{};
  ep.mode = TC_MODE_PLAIN;
  ep.beta = beta; ep.C = C; ep.ldc = ldc; ep.Cb = (__nv_bfloat16*)Cb; ep.ldcb = ldcb;
  return tc_gemm(M, N, K, A, lda, Bm, ldb, ep, TC_FLAG_PDL | (b_const ? TC_FLAG_B_CONST : 0), st);
}

// internal launchers with layout options (elementwise.cu / attention.cu)
int launch_transpose_bf16(const void* src, int src_bf16, int64_t R, int64_t C, int64_t lds, void* dst, int64_t ldd,
                          int permH, cudaStream_t st);
int launch_cell_bwd(int B, int H, const float* act, const float* c_prev, const float* c_new, const float* dh_a,
                    int64_t dha_ld, const float* dh_b, int64_t dhb_ld, float* dc, float* dgates, void* dg_bf16, int perm,
                    cudaStream_t st);
int launch_colsum(const float* x, int64_t rows, int N, int64_t ld, float* out, int permH, cudaStream_t st);
// three column sums in one launch; job 0 may un-permute gate columns (permH) and write a second copy (out0b)
int launch_colsum3(const float* x0, int64_t rows0, int N0, int64_t ld0, float* out0, float* out0b, int permH0,
                   const float* x1, int64_t rows1, int N1, int64_t ld1, float* out1, const float* x2, int64_t rows2, int N2,
                   int64_t ld2, float* out2, cudaStream_t st);
bool pdl_enabled();

struct AttnFwdArgs {
  int B, T, A, F;
  const float* wq; const float* uk; const float* bias; const float* w;
  const void* keys; int keys_bf16; int keys_batch; int64_t k_sb, k_st;
  const uint8_t* mask; int64_t m_sb, m_st;
  float* ctx_f32; int64_t ctx_ld; void* ctx_bf16; int64_t ctxb_ld; float* alpha;
  int fast_math;
  int keys_f16;               // the 16-bit "keys" are IEEE fp16 (projected keys P = keys . W_c^T): streaming kernel only
  unsigned long long* prof;   // debug: 8 globaltimer stamps per CTA (mvc_debug_set_attn_prof), normally null
};
int launch_attention_fwd(const AttnFwdArgs& a, bool pdl, cudaStream_t st);

struct AttnBwdArgs {
  int B, T, A, F;
  const float* wq; const float* uk; const float* bias; const float* w;
  const void* keys; int keys_bf16; int64_t k_sb, k_st;
  const float* alpha; const float* dctx; int64_t dctx_ld;
  float* dwq; void* dwq_bf16; float* duk; float* dw_partial; float* dkeys; int64_t dk_sb, dk_st;
  int fast_math;
};
int launch_attention_bwd(const AttnBwdArgs& a, bool pdl, cudaStream_t st);

// Static description of the recurrence (constant over the time loop).
struct StepCfg {
  int prec;
  int T, F, H, A;        // keys per row, key/context width, LSTM hidden, attention bottleneck
  int perm;              // gate columns in the tile-interleaved order of the fused cell epilogue (bf16, H % 32 == 0)
  const float* uk;       // [keys_batch, T, A]   U.k (hoisted)
  const void* keys;      // element (b,t,f) at keys[(b % keys_batch)*k_sb + t*k_st + f], compute dtype
  int keys_batch;
  int64_t k_sb, k_st;
  const uint8_t* mask;   // optional [.,.] uint8 at mask[b*m_sb + t*m_st]
  int64_t m_sb, m_st;
  const void* wcat;      // [4H, F+H] = [W_ih(ctx part) | W_hh], compute dtype (rows permuted iff perm)
  const void* wcatT;     // [F+H, 4H] bf16 (backward, bf16 path only)
  const void* attW;      // [A, H] attention.W in the compute dtype
  const void* attWT;     // [H, A] bf16 (backward, bf16 path only)
  const float* attW32;   // fp32 master copy (backward dh += dwq . W, fp32 path)
  const float* att_b;    // [A]
  const float* att_w;    // [A]
  const float* cell_bias;  // [4H] added inside the cell (null when folded into gx / embtab); permuted iff perm
  const float* embtab;   // [V,4H] gathered by tokens inside the cell (or null); columns permuted iff perm
  float* pre;            // [rows, 4H] scratch (unfused path)
};

struct StepFwd {
  int rows;
  void* xh_src;          // [rows, F+H] slot: receives ctx_s, holds h_s
  void* xh_dst;          // slot receiving h_{s+1} (may be null)
  float* wq;             // [rows, A]
  float* alpha;          // [rows, T]
  float* act;            // [rows, 4H] or null
  const float* c_prev;
  float* c_out;
  const float* gx;       // hoisted input projection rows [rows,4H] (ld 4H) or null
  const int64_t* tokens; // rows of embtab to gather, or null
  float* h_out32;        // fp32 h_{s+1} (ld h_ld), may be null
  int64_t h_ld;
  bool first;
  bool wq_ready;       // wq already holds W.h_s (written by the previous vocabulary GEMM), skip the projection            // h_s == 0: skip the wq GEMM
  // Projected-keys form (decode loops at several waves of rows): pkeys = P [keys_batch*T, 4H] fp16 = keys . W_c^T in the
  // gate-column order of wcat; the attention pass then returns gc = sum_t alpha_t P[b,t,:] [rows, 4H] fp32 instead of ctx,
  // and the gate GEMM contracts over h only (K = H instead of F + H), gc riding in as an addend of the cell epilogue.
  const void* pkeys = nullptr;
  float* gc = nullptr;
};
bool attention_stream_eligible(int rows, int keys_batch, int T, int A, int F);

static inline int step_forward(const StepCfg& c, const StepFwd& io, cudaStream_t st) {
  const bool bf = c.prec == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  const int64_t ldx = c.F + c.H;
  const int R = io.rows;
  // wq = h_s . W^T                                   (temporal_attention.py:20)
  if (io.first) {
    MVC_CUDA(cudaMemsetAsync(io.wq, 0, sizeof(float) * (size_t)R * c.A, st));
  } else if (io.wq_ready) {
    // already produced by the auxiliary column block of the previous step's vocabulary GEMM (greedy / beam decode)
  } else if (bf) {
    MVC_TRY(gemm_tc_chain(R, c.A, c.H, cptr(io.xh_src, c.F, es), ldx, c.attW, c.H, 0.f, io.wq, c.A, nullptr, 0, true, st));
  } else {
    MVC_TRY(gemm_nt(c.prec, R, c.A, c.H, cptr(io.xh_src, c.F, es), ldx, c.attW, c.H, 0.f, io.wq, c.A, nullptr, st));
  }
  // ctx_s -> xh_src[:, :F]                            (temporal_attention.py:22-32)
  AttnFwdArgs a{};
  a.B = R; a.T = c.T; a.A = c.A; a.F = c.F;
  a.wq = io.wq; a.uk = c.uk; a.bias = c.att_b; a.w = c.att_w;
  a.keys = c.keys; a.keys_bf16 = bf; a.keys_batch = c.keys_batch; a.k_sb = c.k_sb; a.k_st = c.k_st;
  a.mask = c.mask; a.m_sb = c.m_sb; a.m_st = c.m_st;
  a.ctx_f32 = bf ? nullptr : (float*)io.xh_src; a.ctx_ld = ldx;
  a.ctx_bf16 = bf ? io.xh_src : nullptr; a.ctxb_ld = ldx;
  a.alpha = io.alpha; a.fast_math = bf ? 1 : 0;
  if (io.pkeys) {
    a.F = 4 * c.H; a.keys = io.pkeys; a.keys_bf16 = 1; a.keys_f16 = 1; a.k_sb = (int64_t)c.T * 4 * c.H; a.k_st = 4 * c.H;
    a.ctx_f32 = io.gc; a.ctx_ld = 4 * c.H; a.ctx_bf16 = nullptr;
  }
  MVC_TRY(launch_attention_fwd(a, bf && !io.first, st));
  if (io.pkeys) {
    // gates = gc + h_s . W_hh^T + embedding-table row  -> cell
    TcEpilogue ep{};
    ep.mode = TC_MODE_CELL;
    ep.H = c.H;
    ep.bias = c.cell_bias;
    ep.gx = io.gc; ep.gx_ld = 4 * c.H;
    ep.embtab = io.tokens ? c.embtab : nullptr; ep.tokens = io.tokens;
    ep.c_prev = io.c_prev; ep.act = io.act; ep.c_out = io.c_out;
    ep.h32 = io.h_out32; ep.h_ld = io.h_ld;
    ep.hb = io.xh_dst ? (__nv_bfloat16*)mptr(io.xh_dst, c.F, es) : nullptr; ep.hb_ld = ldx;
    return tc_gemm(R, 4 * c.H, c.H, cptr(io.xh_src, c.F, es), ldx, cptr(c.wcat, c.F, es), ldx, ep,
                   TC_FLAG_PDL | TC_FLAG_B_CONST, st);
  }
  // gates = [ctx_s ; h_s] . wcat^T  -> cell         (nn.LSTM, features_captioning.py:84)
  if (bf && c.perm) {
    TcEpilogue ep{};
    ep.mode = TC_MODE_CELL;
    ep.H = c.H;
    ep.bias = c.cell_bias;
    ep.gx = io.gx; ep.gx_ld = 4 * c.H;
    ep.embtab = io.tokens ? c.embtab : nullptr; ep.tokens = io.tokens;
    ep.c_prev = io.c_prev; ep.act = io.act; ep.c_out = io.c_out;
    ep.h32 = io.h_out32; ep.h_ld = io.h_ld;
    ep.hb = io.xh_dst ? (__nv_bfloat16*)mptr(io.xh_dst, c.F, es) : nullptr; ep.hb_ld = ldx;
    return tc_gemm(R, 4 * c.H, c.F + c.H, io.xh_src, ldx, c.wcat, ldx, ep, TC_FLAG_PDL | TC_FLAG_B_CONST, st);
  }
  MVC_TRY(gemm_nt(c.prec, R, 4 * c.H, c.F + c.H, io.xh_src, ldx, c.wcat, ldx, 0.f, c.pre, 4 * c.H, nullptr, st));
  float* h2 = nullptr;
  void* hb = nullptr;
  if (io.xh_dst) {
    if (bf) hb = mptr(io.xh_dst, c.F, es);
    else h2 = (float*)mptr(io.xh_dst, c.F, es);
  }
  MVC_TRY(mvc_lstm_cell_fwd(R, c.H, c.pre, io.gx, 4 * c.H, io.tokens ? c.embtab : nullptr, io.tokens, c.cell_bias,
                            io.c_prev, io.act, io.c_out, io.h_out32, io.h_ld, h2, ldx, hb, ldx, st));
  return 0;
}

struct StepBwd {
  int rows;
  const float* act;      // [rows,4H] saved activations of this step
  const float* c_prev;
  const float* c_new;
  const float* dh_ext;   // external gradient w.r.t. h_{s+1} (ld dh_ld), may be null
  int64_t dh_ld;
  bool has_carry;        // add dxh[:, F:] (gradient carried from step s+1)
  float* dc;             // [rows,H] in/out
  float* dG;             // [rows,4H] out (fp32)
  void* dG_b;            // bf16 copy (bf16 path)
  float* dxh;            // [rows, F+H] out: d[ctx_s ; h_s]
  const float* wq;       // saved
  const float* alpha;    // saved
  float* dwq;            // [rows, A] out
  void* dwq_b;           // [rows, A] bf16 copy (bf16 path)
  float* duk;            // [rows, T, A] accumulated
  float* dwpart;         // [rows, A] accumulated
  float* dkeys;          // optional accumulate, strides dk_sb/dk_st
  int64_t dk_sb, dk_st;
  bool first;            // s == 0: h_s is the zero initial state, skip dh_s
};

static inline int step_backward(const StepCfg& c, const StepBwd& io, cudaStream_t st) {
  const bool bf = c.prec == MVC_BF16;
  const int64_t ldx = c.F + c.H;
  const int R = io.rows;
  MVC_TRY(launch_cell_bwd(R, c.H, io.act, io.c_prev, io.c_new, io.dh_ext, io.dh_ld,
                          io.has_carry ? io.dxh + c.F : nullptr, ldx, io.dc, io.dG, io.dG_b, c.perm, st));
  // d[ctx_s ; h_s] = dgates . wcat
  if (bf) MVC_TRY(gemm_tc_chain(R, c.F + c.H, 4 * c.H, io.dG_b, 4 * c.H, c.wcatT, 4 * c.H, 0.f, io.dxh, ldx, nullptr, 0,
                                true, st));
  else MVC_TRY(mvc_gemm_f32(R, c.F + c.H, 4 * c.H, 1.f, io.dG, 4 * c.H, 1, (const float*)c.wcat, 1, ldx, 0.f, io.dxh,
                            ldx, nullptr, st));
  AttnBwdArgs a{};
  a.B = R; a.T = c.T; a.A = c.A; a.F = c.F;
  a.wq = io.wq; a.uk = c.uk; a.bias = c.att_b; a.w = c.att_w;
  a.keys = c.keys; a.keys_bf16 = bf; a.k_sb = c.k_sb; a.k_st = c.k_st;
  a.alpha = io.alpha; a.dctx = io.dxh; a.dctx_ld = ldx;
  a.dwq = io.dwq; a.dwq_bf16 = bf ? io.dwq_b : nullptr; a.duk = io.duk; a.dw_partial = io.dwpart;
  a.dkeys = io.dkeys; a.dk_sb = io.dk_sb; a.dk_st = io.dk_st; a.fast_math = bf ? 1 : 0;
  MVC_TRY(launch_attention_bwd(a, bf, st));
  // dh_s += dwq . W        (wq = h_s . W^T)
  if (!io.first) {
    if (bf) MVC_TRY(gemm_tc_chain(R, c.H, c.A, io.dwq_b, c.A, c.attWT, c.A, 1.f, io.dxh + c.F, ldx, nullptr, 0, true, st));
    else MVC_TRY(mvc_gemm_f32(R, c.H, c.A, 1.f, io.dwq, c.A, 1, c.attW32, 1, c.H, 1.f, io.dxh + c.F, ldx, nullptr, st));
  }
  return 0;
}

// ---- small shared kernels (defined in decoder.cu) ----
int launch_pack_wcat(const float* w_x, int64_t wx_ld, const float* w_hh, int F, int H, void* out, int out_bf16, int perm,
                     cudaStream_t st, const float* extra = nullptr, int n_extra = 0);
int launch_cast_pad_bf16(const float* src, int64_t rows, int C, int64_t lds, int Cp, void* out, int permH,
                         cudaStream_t st);
int launch_add_vec(const float* a, const float* b, float* o, int n, int permH, cudaStream_t st);
int launch_fill_i64(int64_t* p, int64_t v, int64_t n, cudaStream_t st);
int launch_iota_i64(int64_t* p, int64_t n, cudaStream_t st);
int launch_add_rows(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int C, int accumulate,
                    cudaStream_t st);

}  // namespace mvc
