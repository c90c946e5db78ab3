/* mvc_b200.h -- C ABI of libmvc_b200.so: the B200 (sm_100a) kernels behind the
 * SA-LSTM video-caption hot path of hmartelb/multimodal-video-captioning.
 *
 * The reference has no FFI: its hot path sits behind Python nn.Module methods
 * (SURVEY.md §8b).  Each entry point below names the reference method whose
 * arithmetic it replaces (file:line relative to the reference repo).  The
 * Python host layer (multimodal-video-captioning_b200/salstm/) binds these
 * with ctypes and re-exposes the reference's own class / method names.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors are dense row-major with the dims given in the comment;
 *   - parameters are fp32 (master copy); `precision` picks the compute type:
 *       MVC_F32  : fp32 FFMA everywhere (the exact-greedy-ids path)
 *       MVC_BF16 : bf16 operands on tcgen05 tensor cores, fp32 accumulate,
 *                  fp32 state (h, c, scores, softmax, log-probs);
 *   - `stream` is a cudaStream_t passed as void*; no entry point synchronises
 *     the device or allocates device memory: callers pass a workspace of
 *     at least *_workspace_bytes();
 *   - return value 0 = ok, non-zero = error; mvc_last_error() describes it
 *     (thread-local).  There is no CPU fallback anywhere.
 */
#ifndef MVC_B200_H
#define MVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVC_F32 0
#define MVC_BF16 1

#define MVC_PAD 0 /* get_loader.py:25-26 */
#define MVC_SOS 1
#define MVC_EOS 2

const char* mvc_last_error(void);
int mvc_version(void);
/* 1 when the current device is sm_100 (tcgen05 available). */
int mvc_device_ok(void);
/* Number of kernels this library has launched so far in this process (bench.py `gpu_launches`). */
long long mvc_launch_count(void);
/* In-situ timing of one kernel class: after mvc_prof_arm(kid, m, n, k) every launch of that class
 * (m/n/k = -1 match anything; for GEMMs they are M,N,K, for attention B,T,F, otherwise rows,width,0)
 * is bracketed with CUDA events on its own stream; mvc_prof_collect() synchronises those events,
 * returns their summed duration and count, and disarms.  Kernel classes: 1 tcgen05 GEMM, 2 fp32 GEMM,
 * 3 attention fwd, 4 attention bwd, 5 LSTM cell fwd, 6 LSTM cell bwd, 7 log-softmax rows,
 * 8 fused recurrence step, 9 caption loss, 10 clip+Adam. */
int mvc_prof_arm(int kid, int m, int n, int k);
int mvc_prof_collect(double* total_ms, long long* launches);
/* Debug: device buffer (>= 10 int64 per loop step) that receives SM-clock timestamps of the phases of the
 * persistent recurrence kernel (CTA 0); NULL switches it off. */
int mvc_debug_set_recur_prof(long long* dev_buf);
int mvc_debug_set_recur_bwd_prof(long long* dev_buf);
/* Debug: device buffer (8 uint64 per CTA of the staged soft-attention forward kernel, CTA index = blockIdx.y *
 * gridDim.x + blockIdx.x) that receives %globaltimer stamps of its phases; NULL switches it off. */
int mvc_debug_set_attn_prof(unsigned long long* dev_buf);
/* Debug: same for the one-tile tcgen05 GEMM kernel, for launches of exactly the shape (M, N, K); CTA index =
 * (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x. */
int mvc_debug_set_gemm_prof(unsigned long long* dev_buf, int M, int N, int K);

/* ------------------------------------------------------------------ */
/* Building-block kernels (each is unit-tested against the oracle)     */
/* ------------------------------------------------------------------ */

/* C[M,N] = alpha * sum_k A(m,k) B(n,k) + beta * C + bias[n] (bias may be NULL)
 * with A(m,k) = A[m*a_rs + k*a_cs], B(n,k) = B[n*b_rs + k*b_cs] (element
 * strides), fp32 FFMA.  Replaces every nn.Linear / torch.mm on the path
 * (temporal_attention.py:20-23, features_captioning.py:87) on the fp32 path. */
int mvc_gemm_f32(int M, int N, int K, float alpha, const float* A, int64_t a_rs, int64_t a_cs,
                 const float* B, int64_t b_rs, int64_t b_cs, float beta, float* C, int64_t ldc,
                 const float* bias, void* stream);

/* C[M,N] (fp32) = sum_k A[m,k] B[n,k] + beta*C + bias[n];  A [M,lda], B [N,ldb]
 * bf16 K-contiguous, lda/ldb multiples of 8.  tcgen05.mma (kind::f16, bf16 in,
 * fp32 accumulate in TMEM), operands staged by TMA with 128B swizzle.
 * If Cb != NULL a bf16 copy of the result is also written (ld = ldcb). */
int mvc_gemm_bf16(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb,
                  float beta, float* C, int64_t ldc, const float* bias, void* Cb, int64_t ldcb,
                  void* stream);

/* Same contraction with either operand stored TRANSPOSED: a_transposed -> A is [K, M] row-major (lda >= M),
 * b_transposed -> B is [K, N] row-major (ldb >= N); consumed in place as MN-major tcgen05 operands (no transpose pass).
 * This is the shape of every weight-gradient GEMM (dW = dY^T . X) of the backward pass. */
int mvc_gemm_bf16_ex(int M, int N, int K, const void* A, int64_t lda, int a_transposed, const void* B, int64_t ldb,
                     int b_transposed, float beta, float* C, int64_t ldc, const float* bias, void* stream);

/* dst (fp32 or bf16) [rows, Fa+Fv] = cat(a [rows,Fa], v [rows,Fv]) ; either
 * source may be NULL with width 0.  captioning.py:109 (torch.cat, audio first). */
int mvc_concat_cast(const float* a, int Fa, const float* v, int Fv, int64_t rows, void* dst,
                    int dst_bf16, void* stream);

/* Feature input format (get_loader.py:242-268 hands fp32 .npy features): MVC_INPUT_BF16 declares that the
 * `audio` / `visual` buffers passed to the following mvc_decoder_forward / _greedy / _beam calls OF THIS HOST THREAD
 * hold bf16 values (pre-packed shards: half the host-to-device bytes, no cast pass); the decoder must then run with
 * precision MVC_BF16.  Rounding fp32 features to bf16 on the host (round-to-nearest-even) and passing them this way
 * is bit-identical to passing the fp32 features.  Default MVC_INPUT_F32. */
#define MVC_INPUT_F32 0
#define MVC_INPUT_BF16 1
int mvc_set_input_format(int fmt);
int mvc_get_input_format(void);
/* dst[r, :] = concat(a[r, :Fa], v[r, :Fv]), all bf16 (Fa, Fv multiples of 8; 16-byte aligned buffers). */
int mvc_concat_bf16(const void* a, int Fa, const void* v, int Fv, int64_t rows, void* dst, void* stream);
int mvc_cast_bf16(const float* src, void* dst, int64_t n, void* stream);
/* dst[c, r] = bf16(src[r, c]); src [R,C] ld=lds fp32 (src_bf16=0) or bf16, dst [C,R] ld=ldd bf16. */
int mvc_transpose_to_bf16(const void* src, int src_bf16, int64_t R, int64_t C, int64_t lds, void* dst,
                          int64_t ldd, void* stream);

/* Fused soft attention forward.  temporal_attention.py:19-33 minus the two
 * projections (W.q is `wq`, U.k is `uk`, both precomputed fp32):
 *   e[b,t]  = sum_a w[a] tanh(wq[b,a] + uk[b,t,a] + bias[a]); e[~mask] = -inf
 *   alpha   = softmax_t(e);  ctx[b,:] = sum_t alpha[b,t] keys[b,t,:]
 * keys element (b,t,f) lives at keys[(b % keys_batch)*k_sb + t*k_st + f]
 * (fp32 or bf16); mask (uint8, may be NULL) at mask[b*m_sb + t*m_st].
 * ctx is written fp32 to ctx_f32 (ld ctx_ld) and/or bf16 to ctx_bf16
 * (ld ctxb_ld); either may be NULL.  alpha [B,T] fp32. */
int mvc_soft_attention_fwd(int B, int T, int A, int F, const float* wq, const float* uk,
                           const float* bias, const float* w, const void* keys, int keys_bf16,
                           int keys_batch, int64_t k_sb, int64_t k_st, const uint8_t* mask,
                           int64_t m_sb, int64_t m_st, float* ctx_f32, int64_t ctx_ld,
                           void* ctx_bf16, int64_t ctxb_ld, float* alpha, int fast_math,
                           void* stream);

/* Backward of the above w.r.t. wq, uk, w (and keys when dkeys != NULL); one CTA
 * per batch row, no atomics (deterministic):
 *   dalpha[t] = dctx . keys[t];  de = alpha*(dalpha - sum alpha dalpha)
 *   dpre[t,a] = de[t] w[a] (1 - tanh^2)
 *   dwq[b,a]        = sum_t dpre            (overwritten)
 *   duk[b,t,a]     += dpre                  (accumulated, may be NULL)
 *   dw_partial[b,a]+= sum_t de[t] tanh(.)   (accumulated; dw = column sum)
 *   dkeys[b,t,f]   += alpha[b,t] dctx[b,f]  (fp32, strides dk_sb/dk_st, may be NULL)
 * dbias is the column sum of dwq (dpre does not depend on which of wq/bias). */
int mvc_soft_attention_bwd(int B, int T, int A, int F, const float* wq, const float* uk,
                           const float* bias, const float* w, const void* keys, int keys_bf16,
                           int64_t k_sb, int64_t k_st, const float* alpha, const float* dctx,
                           int64_t dctx_ld, float* dwq, float* duk, float* dw_partial,
                           float* dkeys, int64_t dk_sb, int64_t dk_st, int fast_math,
                           void* stream);

/* LSTM cell.  features_captioning.py:84 (nn.LSTM, gate order i,f,g,o).
 * pre [B,4H] = recurrent GEMM result; optional addends: gx [B,4H] (hoisted
 * input projection, row stride gx_ld), emb_table [V,4H] gathered by tokens
 * [B] (int64), bias [4H].  Writes activated gates act [B,4H], c_out, h_out
 * (fp32, ld h_ld), optional second fp32 copy h_out2 (ld h2_ld) and bf16 copy
 * h_bf16 (ld hb_ld). */
int mvc_lstm_cell_fwd(int B, int H, const float* pre, const float* gx, int64_t gx_ld,
                      const float* emb_table, const int64_t* tokens, const float* bias,
                      const float* c_prev, float* act, float* c_out, float* h_out, int64_t h_ld,
                      float* h_out2, int64_t h2_ld, void* h_bf16, int64_t hb_ld, void* stream);

/* Fused LSTM step on the tensor cores (K-C): gates = x . W^T on tcgen05 with fp32 accumulators in
 * TMEM, and the cell update (bias / hoisted-projection addends, sigmoid/tanh, c and h) in the GEMM
 * epilogue; K is split across CTAs and reduced deterministically.  features_captioning.py:84.
 *   x [B,K] bf16 (ld ldx);  w_packed [4H,K] bf16 = the gate rows in TILE-INTERLEAVED order: packed
 *   row (j/16)*64 + g*16 + j%16 is nn.LSTM row g*H + j (g = i,f,g,o) -- mvc_pack_gate_rows_bf16
 *   builds it from an fp32 [4H,C] weight (ld ldw), zero-padding C to Cp (multiple of 8);
 *   bias_packed [4H], gx_packed [B,4H] (optional addends) and act_packed [B,4H] (activated gates,
 *   optional output) use the same column order; c_prev/c_out [B,H]; h_out fp32 (ld h_ld) and
 *   h_bf16 (ld hb_ld) optional.  Requires H % 32 == 0. */
int mvc_pack_gate_rows_bf16(const float* w, int H, int C, int64_t ldw, int Cp, void* out, void* stream);
int mvc_lstm_gates_cell_bf16(int B, int H, int K, const void* x, int64_t ldx, const void* w_packed, int64_t ldw,
                             const float* bias_packed, const float* gx_packed, int64_t gx_ld, const float* c_prev,
                             float* act_packed, float* c_out, float* h_out, int64_t h_ld, void* h_bf16,
                             int64_t hb_ld, void* stream);

/* Vocabulary projection fused with the row arg-max (K-E, greedy decoding: features_captioning.py:87-88,:109):
 * ids[m] = argmax_v (h[m,:] . out_w[v,:] + out_b[v]), ties -> lowest v.  h [M,K] bf16 (ld ldh), out_w [V,K] bf16
 * (ld ldw).  The arg-max runs in the tcgen05 epilogue (one partial per row and 256-column tile, then a tiny
 * reduce): the [M,V] logits are never written.  workspace >= M * ceil(V/256) * 8 bytes. */
int mvc_vocab_argmax_bf16(int M, int V, int K, const void* h, int64_t ldh, const void* out_w, int64_t ldw,
                          const float* out_b, void* workspace, size_t workspace_bytes, int64_t* ids, void* stream);

/* Same GEMM with an auxiliary column block (decode loops: features_captioning.py:87-88 of step s together with
 * temporal_attention.py:20 of step s+1).  w_ext [mvc_vocab_aux_row0(V) + A, K] bf16 holds out_w in rows [0, V),
 * padding up to the next multiple of 256, then the attention query weights W in the last A rows; besides ids the call
 * writes wq[m, a] = h[m,:] . W[a,:] (fp32, ld A) -- the query projection the NEXT step's soft attention needs -- so the
 * decode loop launches no separate projection kernel. */
int mvc_vocab_aux_row0(int V);
int mvc_vocab_argmax_wq_bf16(int M, int V, int K, int A, const void* h, int64_t ldh, const void* w_ext, int64_t ldw,
                             const float* out_b, void* workspace, size_t workspace_bytes, int64_t* ids, float* wq,
                             void* stream);

/* Vocabulary projection fused with log-softmax + top-k (K-E, beam search: features_captioning.py:160-189):
 * cand_logp[m,k], cand_idx[m,k] (k < width <= 8) = the largest log_softmax(h[m,:] . out_w^T + out_b) values of row m
 * and their tokens, best first, ties -> lowest token.  Top-8 lists and an online log-sum-exp are kept per row and
 * 256-column tile in the tcgen05 epilogue and merged by a small kernel; logits / log-probs are never written. */
size_t mvc_vocab_topk_workspace_bytes(int M, int V);
int mvc_vocab_topk_bf16(int M, int V, int K, const void* h, int64_t ldh, const void* out_w, int64_t ldw,
                        const float* out_b, int width, void* workspace, size_t workspace_bytes, float* cand_logp,
                        int* cand_idx, void* stream);

/* dgates (pre-activation) from dh (two optional addends dh_a [ld dha_ld], dh_b
 * [ld dhb_ld]) and the carried dc (in/out, [B,H]); dg_bf16 optional copy. */
int mvc_lstm_cell_bwd(int B, int H, const float* act, const float* c_prev, const float* c_new,
                      const float* dh_a, int64_t dha_ld, const float* dh_b, int64_t dhb_ld,
                      float* dc, float* dgates, void* dg_bf16, void* stream);

/* In-place row-wise log-softmax over x [rows, V] (ld = V) + optional argmax
 * (lowest index wins ties) written as int64.  features_captioning.py:88,:109 */
int mvc_log_softmax_rows(float* x, int64_t rows, int V, int64_t* argmax, void* stream);
/* argmax over V of (x [+ y]) rows -> int64 ; captioning.py:140, :283-285 */
int mvc_argmax_rows(const float* x, const float* y, int64_t rows, int V, int64_t* out,
                    void* stream);
/* dlogits = dlogp - exp(logp) * rowsum(dlogp); optional bf16 copy. */
int mvc_log_softmax_bwd(const float* logp, const float* dlogp, int64_t rows, int V, float* dlogits,
                        void* dlogits_bf16, void* stream);

/* out[r,:] = table[idx[r],:] (fp32 -> fp32 or bf16, out ld = out_ld, row width E)
 * features_captioning.py:78 (nn.Embedding, no padding_idx). */
int mvc_embedding_gather(const float* table, int E, const int64_t* idx, int64_t rows, void* out,
                         int64_t out_ld, int out_bf16, void* stream);
/* dtable[idx[r],:] += dx[r,:]  (atomicAdd fp32) */
int mvc_embedding_scatter_add(const float* dx, int64_t dx_ld, int E, const int64_t* idx,
                              int64_t rows, float* dtable, void* stream);
/* out[n] = sum_r x[r, n]   (x [rows,N], ld) ; out overwritten */
int mvc_colsum(const float* x, int64_t rows, int N, int64_t ld, float* out, void* stream);

/* ------------------------------------------------------------------ */
/* Decoder (FeaturesCaptioning) orchestrators                          */
/* ------------------------------------------------------------------ */
typedef struct {
  int B, T, F, H, E, A, V; /* batch, frames, feature, hidden, embed, attn, vocab */
  int L;                   /* max_caption_len: L-1 loop steps                   */
  int precision;           /* MVC_F32 | MVC_BF16                                 */
} MvcDecoderDims;

/* fp32 parameters, names as in FeaturesCaptioning.state_dict()
 * (features_captioning.py:36-56, temporal_attention.py:13-17). */
typedef struct {
  const float* embedding;  /* [V,E]      embedding.weight      */
  const float* att_W;      /* [A,H]      attention.W.weight    */
  const float* att_U;      /* [A,F]      attention.U.weight    */
  const float* att_b;      /* [A]        attention.b           */
  const float* att_w;      /* [1,A]      attention.w.weight    */
  const float* w_ih;       /* [4H,E+F]   rnn.weight_ih_l0      */
  const float* w_hh;       /* [4H,H]     rnn.weight_hh_l0      */
  const float* b_ih;       /* [4H]       rnn.bias_ih_l0        */
  const float* b_hh;       /* [4H]       rnn.bias_hh_l0        */
  const float* out_w;      /* [V,H]      out.weight            */
  const float* out_b;      /* [V]        out.bias              */
} MvcDecoderParams;

typedef struct {
  float* embedding; float* att_W; float* att_U; float* att_b; float* att_w;
  float* w_ih; float* w_hh; float* b_ih; float* b_hh; float* out_w; float* out_b;
} MvcDecoderGrads;

size_t mvc_decoder_fwd_workspace_bytes(const MvcDecoderDims* d, int save_for_backward);
size_t mvc_decoder_bwd_workspace_bytes(const MvcDecoderDims* d);

/* FeaturesCaptioning.decode / forward_sentence (features_captioning.py:91-129).
 *   audio [B,T,Fa] / visual [B,T,Fv] fp32, Fa+Fv == F (either may be NULL/0):
 *     the early-fusion cat of captioning.py:109 is done inside;
 *   captions [L,B] int64 or NULL (free running);
 *   tf_flags_host [L-1] : step t=1+i feeds captions[t] as the NEXT input iff
 *     flag[i] (the reference's `torch.rand(1) < ratio` draw, made by the
 *     caller on the host RNG so the stream of draws is identical, :113-117);
 *   out_logp [L,B,V] log-probs (row 0 zeroed), out_hid [L,B,H] (row 0 zeroed);
 *   tokens_in [L-1,B] int64 receives the token fed at each step. */
int mvc_decoder_forward(const MvcDecoderDims* d, const MvcDecoderParams* p, const float* audio,
                        int Fa, const float* visual, int Fv, const int64_t* captions,
                        const uint8_t* tf_flags_host, float* out_logp, float* out_hid,
                        int64_t* tokens_in, void* workspace, size_t workspace_bytes,
                        int save_for_backward, void* stream);

/* BPTT through the above.  dlogp [L,B,V] / dhid [L,B,H] may be NULL (= 0).
 * `workspace` is the forward workspace (saved activations); grads are
 * OVERWRITTEN (fp32, parameter shapes). */
int mvc_decoder_backward(const MvcDecoderDims* d, const MvcDecoderParams* p, const float* out_logp,
                         const float* dlogp, const float* dhid, const int64_t* tokens_in,
                         const void* fwd_workspace, MvcDecoderGrads* g, void* bwd_workspace,
                         size_t bwd_workspace_bytes, void* stream);

/* AVCaptioning.predict(mode="direct") ids (captioning.py:138-141): greedy
 * free-running decode that never materialises [L,B,V]; ids [B,L] int64 with
 * column 0 = 0.  If logp_sum != NULL ([L,B,V], Dual late fusion,
 * captioning.py:279-285) the per-step log-probs are ADDED into it instead and
 * ids may be NULL. */
size_t mvc_decoder_greedy_workspace_bytes(const MvcDecoderDims* d);
int mvc_decoder_greedy(const MvcDecoderDims* d, const MvcDecoderParams* p, const float* audio,
                       int Fa, const float* visual, int Fv, int64_t* ids, void* workspace,
                       size_t workspace_bytes, void* stream);

/* FeaturesCaptioning.beam_search_predict (features_captioning.py:131-228) with
 * all bookkeeping on the device; ids [B, L+2] int64 = SOS + (L+1) ids of the
 * best beam (L = max_caption_len).  Ties -> lowest flat index (beam*V+token). */
size_t mvc_decoder_beam_workspace_bytes(const MvcDecoderDims* d, int width);
int mvc_decoder_beam(const MvcDecoderDims* d, const MvcDecoderParams* p, const float* audio, int Fa,
                     const float* visual, int Fv, int width, float alpha, int64_t* ids,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ */
/* RecNet reconstructors                                               */
/* ------------------------------------------------------------------ */
typedef struct {
  int B, L, H;   /* decoder hiddens [L,B,H]                       */
  int Fr;        /* reconstructor hidden = feature size           */
  int A;         /* attention bottleneck (local only)             */
  int T;         /* frames to reconstruct (local only)            */
  int precision;
} MvcReconDims;

typedef struct {
  const float* w_ih;  /* global [4Fr,2H] ; local [4Fr,H]  rnn.weight_ih_l0 */
  const float* w_hh;  /* [4Fr,Fr]                          rnn.weight_hh_l0 */
  const float* b_ih;  /* [4Fr] */
  const float* b_hh;  /* [4Fr] */
  const float* att_W; /* local: [A,Fr]  attention.W.weight */
  const float* att_U; /* local: [A,H]   attention.U.weight */
  const float* att_b; /* local: [A] */
  const float* att_w; /* local: [1,A] */
} MvcReconParams;

typedef struct {
  float* w_ih; float* w_hh; float* b_ih; float* b_hh;
  float* att_W; float* att_U; float* att_b; float* att_w;
} MvcReconGrads;

/* caption mask (cap != PAD) & (cap != EOS) as uint8 [L,B]; reconstructor.py:197-206 */
int mvc_caption_mask(const int64_t* captions, int64_t n, uint8_t* mask, void* stream);

size_t mvc_global_recon_workspace_bytes(const MvcReconDims* d);
size_t mvc_global_recon_bwd_workspace_bytes(const MvcReconDims* d);
/* GlobalReconstructor.reconstruct (reconstructor.py:142-194):
 * hid [L,B,H], mask [L,B] -> rec [B,L,Fr] (row t=0 zero). */
int mvc_global_recon_forward(const MvcReconDims* d, const MvcReconParams* p, const float* hid,
                             const uint8_t* mask, float* rec, void* workspace,
                             size_t workspace_bytes, void* stream);
int mvc_global_recon_backward(const MvcReconDims* d, const MvcReconParams* p, const float* hid,
                              const uint8_t* mask, const float* drec, const void* fwd_workspace,
                              float* dhid, MvcReconGrads* g, void* bwd_workspace,
                              size_t bwd_workspace_bytes, void* stream);

size_t mvc_local_recon_workspace_bytes(const MvcReconDims* d);
size_t mvc_local_recon_bwd_workspace_bytes(const MvcReconDims* d);
/* LocalReconstructor.reconstruct (reconstructor.py:67-97): -> rec [B,T,Fr]. */
int mvc_local_recon_forward(const MvcReconDims* d, const MvcReconParams* p, const float* hid,
                            const uint8_t* mask, float* rec, void* workspace,
                            size_t workspace_bytes, void* stream);
int mvc_local_recon_backward(const MvcReconDims* d, const MvcReconParams* p, const float* hid,
                             const uint8_t* mask, const float* drec, const void* fwd_workspace,
                             float* dhid, MvcReconGrads* g, void* bwd_workspace,
                             size_t bwd_workspace_bytes, void* stream);

/* ------------------------------------------------------------------ */
/* Losses (src/losses.py)                                              */
/* ------------------------------------------------------------------ */
/* NLL (ignore_index = PAD, mean over non-PAD) + EntropyLoss (softmax over the
 * BATCH axis of the log-probs -- losses.py:12-17 quirk kept) over
 * logp[1:] ([L,B,V]) and captions[1:].  result[0] = ce, result[1] = entropy,
 * result[2] = number of non-PAD targets.  If dlogp != NULL it receives
 * d(ce_scale*ce + ent_scale*entropy)/dlogp for the whole [L,B,V] (row 0 = 0).
 * workspace: >= mvc_caption_loss_workspace_bytes. losses.py:112-115 */
size_t mvc_caption_loss_workspace_bytes(int L, int B, int V);
int mvc_caption_loss(const float* logp, const int64_t* captions, int L, int B, int V,
                     float* result, float* dlogp, float ce_scale, float ent_scale,
                     void* workspace, void* stream);

/* GlobalReconstructionLoss (losses.py:20-36): x [B,T,F] slice (ld x_ld),
 * xrec [B,L,F] slice (ld r_ld), captions [L,B].  result[0] = mse.
 * dxrec ([B,L,F], row pitch d_ld, may be NULL) = scale * dloss/dxrec (OVERWRITTEN, zeros on PAD rows). */
int mvc_global_recon_loss(const float* x, int64_t x_ld, const float* xrec, int64_t r_ld, int B,
                          int T, int L, int F, const int64_t* captions, float* result,
                          float* dxrec, int64_t d_ld, float scale, void* workspace, void* stream);
size_t mvc_global_recon_loss_workspace_bytes(int B, int F);
/* LocalReconstructionLoss (losses.py:39-40): mse over [B,T,F] slices; dxrec (may be NULL) is OVERWRITTEN. */
int mvc_local_recon_loss(const float* x, int64_t x_ld, const float* xrec, int64_t r_ld,
                         int64_t rows, int F, float* result, float* dxrec, int64_t d_ld,
                         float scale, void* workspace, void* stream);

/* Total of ModalityWiseReconstructionLoss (losses.py:122-124) on the device: result[0..4] = ce, entropy, n_tokens,
 * audio_rec, visual_rec as the calls above left them (a missing reconstruction term is set to 0, :100-101);
 * result[5] = ce + reg*entropy + a_lambda*audio_rec + v_lambda*visual_rec, summed in the reference's order. */
int mvc_loss_combine(float* result, float reg_lambda, float a_lambda, float v_lambda, int have_a, int have_v,
                     void* stream);
/* x[0..n) *= *g_dev (g_dev: one fp32 on the device -- the upstream gradient of the loss scalar; exits at once when
 * it is 1.0, as in `loss.mean().backward()`, train.py:198).  x 16-byte aligned. */
int mvc_scale_by_scalar(float* x, int64_t n, const float* g_dev, void* stream);

/* ------------------------------------------------------------------ */
/* Trainer step tail (train.py:207-210): clip_grad_value_ + Adam(amsgrad, */
/* weight_decay) over one flat fp32 buffer.                             */
/* ------------------------------------------------------------------ */
int mvc_clip_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                       float* max_exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, float clip_value, int step, float grad_scale,
                       void* stream);

/* The same update with its two per-step scalars in DEVICE memory, so that the launch can be recorded in a CUDA graph
 * and replayed: state_dev[0] = number of updates applied so far (a float; incremented by one before the update when
 * tick != 0 -- pass tick only for the first range of a step), state_dev[1] = learning rate (the host rewrites it when
 * an lr scheduler changes it; train.py:90-97). */
int mvc_clip_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                           float* max_exp_avg_sq, int64_t n, float* state_dev, int tick, float beta1, float beta2,
                           float eps, float weight_decay, float clip_value, float grad_scale, void* stream);

/* Data-parallel trainer tail as ONE kernel over NVSwitch multicast (the per-step gradient all-reduce of
 * train.py's data-parallel variant + train.py:207-210): param_mc / grad_mc are the MULTICAST addresses of the flat fp32
 * parameter / gradient buffers, which every rank allocated symmetrically and bound to one multicast object (the host
 * layer does this with torch.distributed._symmetric_memory); param_local and the optimiser state are this rank's own
 * buffers.  The rank owns elements [lo, hi) (multiples of 4): multimem.ld_reduce sums the replicas' gradients in the
 * switch, the clip + Adam(amsgrad) update is applied to the owned slice, multimem.st writes the new parameters into
 * every replica.  grad_scale = 1/world.  The caller brackets the call with cross-rank barriers (all gradients written
 * before / all slices stored after). */
int mvc_clip_adam_multimem(const float* param_local, float* param_mc, const float* grad_mc, float* exp_avg,
                           float* exp_avg_sq, float* max_exp_avg_sq, int64_t lo, int64_t hi, float* state_dev, int tick,
                           float beta1, float beta2, float eps, float weight_decay, float clip_value, float grad_scale,
                           void* stream);

/* Same update with the reduce-scatter done by peer loads: grad_replicas[r] (r = 0 .. world-1, host array of device
 * pointers: every rank's flat gradient buffer as mapped into THIS process, e.g. the buffer_ptrs of a symmetric-memory
 * allocation) are read by the owner of [lo, hi) and summed in rank order; the new parameters still leave by multimem.st.
 * multimem.ld_reduce makes the switch fetch every replica over NVLink, the requester's own included; peer loads export a
 * third fewer bytes per GPU at 2 ranks (11 % fewer at 8).  Same barriers around the call as mvc_clip_adam_multimem. */
int mvc_clip_adam_p2p_multimem(const float* param_local, float* param_mc, const float* const* grad_replicas, int world,
                               float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, int64_t lo, int64_t hi,
                               float* state_dev, int tick, float beta1, float beta2, float eps, float weight_decay,
                               float clip_value, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVC_B200_H */
