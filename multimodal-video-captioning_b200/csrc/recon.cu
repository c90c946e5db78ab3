// recon.cu -- RecNet-style reconstructors (reconstructor.py): host-side
// orchestration of forward and BPTT backward on one stream.
//
// Global (reconstructor.py:142-194): x_t = [h_dec_t ; meanpool(h_dec)] -> LSTM(2H -> Fr).
//   The whole input projection is loop invariant (all decoder hiddens are known),
//   so it is hoisted: Gx = Hdec[1:] . W_ih[:, :H]^T  (one [S*B,H]x[H,4Fr] GEMM) and
//   Gp = pooled . W_ih[:, H:]^T + b_ih + b_hh (one [B,H]x[H,4Fr] GEMM, added per row
//   inside the cell kernel).  Only h_rec . W_hh^T stays in the loop.
// Local (reconstructor.py:67-97): masked soft attention over the decoder hiddens with
//   the reconstructor state as query, LSTM(H -> Fr): the same attention->LSTM step as
//   the decoder (step.cuh) with keys = decoder hiddens [B,L,H].
#include "step.cuh"

namespace mvc {

// pooled[b,:] = sum_l mask[l,b] hid[l,b,:] / sum_l mask[l,b]      (reconstructor.py:142-149)
__global__ void masked_mean_kernel(const float* __restrict__ hid, const uint8_t* __restrict__ mask, int L, int B, int H,
                                   float* __restrict__ pooled) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H) return;
  const int b = (int)(i / H);
  float s = 0.f, n = 0.f;
  for (int l = 0; l < L; ++l)
    if (mask[(int64_t)l * B + b]) {
      s += hid[(int64_t)l * B * H + i];
      n += 1.f;
    }
  pooled[i] = s / n;
}

// dhid[l,b,:] += mask[l,b] * dpooled[b,:] / n_b
__global__ void masked_mean_bwd_kernel(const float* __restrict__ dpooled, const uint8_t* __restrict__ mask, int L, int B,
                                       int H, float* __restrict__ dhid) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H) return;
  const int b = (int)(i / H);
  float n = 0.f;
  for (int l = 0; l < L; ++l) n += mask[(int64_t)l * B + b] ? 1.f : 0.f;
  const float g = dpooled[i] / n;
  for (int l = 0; l < L; ++l)
    if (mask[(int64_t)l * B + b]) dhid[(int64_t)l * B * H + i] += g;
}

// out[r, :] = sum_s x[s, r, :]   (x [S, R*C] contiguous)
__global__ void sum_over_steps_kernel(const float* __restrict__ x, int S, int64_t n, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < S; ++k) s += x[(int64_t)k * n + i];
    out[i] = s;
  }
}

// dst[b, l, :] = src[l, b, :]  ([L,B,H] -> [B,L,H]), fp32 -> fp32 or bf16
template <typename OutT>
__global__ void permute_lbh_kernel(const float* __restrict__ src, int L, int B, int H, OutT* __restrict__ dst) {
  const int64_t total = (int64_t)L * B * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(i % H);
    const int64_t r = i / H;
    const int l = (int)(r % L), b = (int)(r / L);
    const float v = src[((int64_t)l * B + b) * H + h];
    if constexpr (sizeof(OutT) == 2) dst[i] = __float2bfloat16(v);
    else dst[i] = v;
  }
}

// dhid[l, b, :] += src[b, l, :]
__global__ void permute_add_blh_kernel(const float* __restrict__ src, int L, int B, int H, float* __restrict__ dhid) {
  const int64_t total = (int64_t)L * B * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(i % H);
    const int64_t r = i / H;
    const int l = (int)(r % L), b = (int)(r / L);
    dhid[((int64_t)l * B + b) * H + h] += src[i];
  }
}

// ------------------------------------------------------------------ global
struct GlobWs {
  float* pooled;    // [B, H]
  float* gx;        // [S*B, 4Fr]
  float* gp;        // [B, 4Fr]
  float* act;       // [S, B, 4Fr]
  float* c;         // [S+1, B, Fr]
  void* hs;         // [S+1, B, Fr] compute dtype (h_rec_0 .. h_rec_S)
  float* pre;       // [B, 4Fr]
  float* bsum;      // [4Fr]
  int64_t* iota;    // [B]
  void* hid_b;      // bf16 [S*B, H]   (hid[1:])
  void* pooled_b;   // bf16 [B, H]
  void* wih_b;      // bf16 [4Fr, 2H]
  void* whh_b;      // bf16 [4Fr, Fr]
  size_t bytes;
};
static GlobWs glob_layout(const MvcReconDims* d, void* base) {
  const int64_t B = d->B, H = d->H, Fr = d->Fr, S = d->L - 1;
  const bool bf = d->precision == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  Arena ar(base);
  GlobWs w{};
  w.pooled = ar.take<float>(B * H);
  w.gx = ar.take<float>(S * B * 4 * Fr);
  w.gp = ar.take<float>(B * 4 * Fr);
  w.act = ar.take<float>(S * B * 4 * Fr);
  w.c = ar.take<float>((S + 1) * B * Fr);
  w.hs = ar.take<char>((S + 1) * B * Fr * es);
  w.pre = ar.take<float>(B * 4 * Fr);
  w.bsum = ar.take<float>(4 * Fr);
  w.iota = ar.take<int64_t>(B);
  if (bf) {
    w.hid_b = ar.take<char>(S * B * H * 2);
    w.pooled_b = ar.take<char>(B * H * 2);
    w.wih_b = ar.take<char>(4 * Fr * 2 * H * 2);
    w.whh_b = ar.take<char>(4 * Fr * Fr * 2);
  }
  w.bytes = ar.off + 256;
  return w;
}

struct GlobBwdWs {
  float* dG;        // [S*B, 4Fr]
  float* dGsum;     // [B, 4Fr]
  float* dhcar;     // [B, Fr]
  float* dc;        // [B, Fr]
  float* dpooled;   // [B, H]
  void* dG_b;       // bf16 [S*B, 4Fr]
  void* dGT;        // bf16 [4Fr, SBp]
  void* hsT;        // bf16 [Fr, SBp]
  void* hidT;       // bf16 [H, SBp]
  void* whhT;       // bf16 [Fr, 4Fr]
  void* wihT;       // bf16 [2H, 4Fr]
  void* dGsum_b;    // bf16 [B, 4Fr]
  void* dGsumT;     // bf16 [4Fr, Bp]
  void* pooledT;    // bf16 [H, Bp]
  size_t bytes;
};
static GlobBwdWs glob_bwd_layout(const MvcReconDims* d, void* base) {
  const int64_t B = d->B, H = d->H, Fr = d->Fr, S = d->L - 1;
  const bool bf = d->precision == MVC_BF16;
  const int64_t SBp = pad8((int)(S * B)), Bp = pad8((int)B);
  Arena ar(base);
  GlobBwdWs w{};
  w.dG = ar.take<float>(S * B * 4 * Fr);
  w.dGsum = ar.take<float>(B * 4 * Fr);
  w.dhcar = ar.take<float>(B * Fr);
  w.dc = ar.take<float>(B * Fr);
  w.dpooled = ar.take<float>(B * H);
  if (bf) {
    w.dG_b = ar.take<char>(S * B * 4 * Fr * 2);
    w.dGT = ar.take<char>(4 * Fr * SBp * 2);
    w.hsT = ar.take<char>(Fr * SBp * 2);
    w.hidT = ar.take<char>(H * SBp * 2);
    w.whhT = ar.take<char>(Fr * 4 * Fr * 2);
    w.wihT = ar.take<char>(2 * H * 4 * Fr * 2);
    w.dGsum_b = ar.take<char>(B * 4 * Fr * 2);
    w.dGsumT = ar.take<char>(4 * Fr * Bp * 2);
    w.pooledT = ar.take<char>(H * Bp * 2);
  }
  w.bytes = ar.off + 256;
  return w;
}

// tile-interleaved gate order (fused gate GEMM + cell epilogue): bf16 path, reconstructor hidden % 32 == 0
static inline int rec_perm(const MvcReconDims* d) { return d->precision == MVC_BF16 && d->Fr % 32 == 0; }

static int check_recon_dims(const MvcReconDims* d, bool local) {
  MVC_CHECK(d, "reconstructor: null dims");
  MVC_CHECK(d->B > 0 && d->L >= 2 && d->H > 0 && d->Fr > 0, "reconstructor: bad dims B=%d L=%d H=%d Fr=%d", d->B, d->L,
            d->H, d->Fr);
  if (local) MVC_CHECK(d->A > 0 && d->T > 0, "local reconstructor: bad dims A=%d T=%d", d->A, d->T);
  MVC_CHECK(d->precision == MVC_F32 || d->precision == MVC_BF16, "reconstructor: unknown precision %d", d->precision);
  if (d->precision == MVC_BF16)
    MVC_CHECK(d->H % 8 == 0 && d->Fr % 8 == 0 && (!local || d->A % 8 == 0),
              "reconstructor(bf16): H, Fr, A must be multiples of 8; got H=%d Fr=%d A=%d", d->H, d->Fr, d->A);
  return 0;
}

}  // namespace mvc

using namespace mvc;

extern "C" size_t mvc_global_recon_workspace_bytes(const MvcReconDims* d) { return glob_layout(d, nullptr).bytes; }
extern "C" size_t mvc_global_recon_bwd_workspace_bytes(const MvcReconDims* d) { return glob_bwd_layout(d, nullptr).bytes; }

extern "C" int mvc_global_recon_forward(const MvcReconDims* d, const MvcReconParams* p, const float* hid,
                                        const uint8_t* mask, float* rec, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  MVC_TRY(check_recon_dims(d, false));
  MVC_CHECK(p && hid && mask && rec && workspace, "mvc_global_recon_forward: null argument");
  GlobWs w = glob_layout(d, workspace);
  MVC_CHECK(workspace_bytes >= w.bytes, "mvc_global_recon_forward: workspace %zu < %zu", workspace_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, L = d->L, H = d->H, Fr = d->Fr, S = L - 1;
  const bool bf = d->precision == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  const int64_t G4 = 4 * (int64_t)Fr;

  masked_mean_kernel<<<(unsigned)cdiv((int64_t)B * H, 256), 256, 0, st>>>(hid, mask, L, B, H, w.pooled);
  MVC_LAUNCH_CHECK();
  const int perm = rec_perm(d), permH = perm ? Fr : 0;
  MVC_TRY(launch_add_vec(p->b_ih, p->b_hh, w.bsum, 4 * Fr, permH, st));
  MVC_TRY(launch_iota_i64(w.iota, B, st));
  const float* hid1 = hid + (int64_t)B * H;                       // decoder_hiddens[1:]
  if (bf) {
    MVC_TRY(mvc_cast_bf16(hid1, w.hid_b, (int64_t)S * B * H, st));
    MVC_TRY(mvc_cast_bf16(w.pooled, w.pooled_b, (int64_t)B * H, st));
    MVC_TRY(launch_cast_pad_bf16(p->w_ih, G4, 2 * H, 2 * H, 2 * H, w.wih_b, permH, st));
    MVC_TRY(launch_cast_pad_bf16(p->w_hh, G4, Fr, Fr, Fr, w.whh_b, permH, st));
    MVC_TRY(mvc_gemm_bf16(S * B, 4 * Fr, H, w.hid_b, H, w.wih_b, 2 * H, 0.f, w.gx, G4, nullptr, nullptr, 0, st));
    MVC_TRY(mvc_gemm_bf16(B, 4 * Fr, H, w.pooled_b, H, cptr(w.wih_b, H, 2), 2 * H, 0.f, w.gp, G4, w.bsum, nullptr, 0, st));
  } else {
    MVC_TRY(mvc_gemm_f32(S * B, 4 * Fr, H, 1.f, hid1, H, 1, p->w_ih, 2 * H, 1, 0.f, w.gx, G4, nullptr, st));
    MVC_TRY(mvc_gemm_f32(B, 4 * Fr, H, 1.f, w.pooled, H, 1, p->w_ih + H, 2 * H, 1, 0.f, w.gp, G4, w.bsum, st));
  }
  MVC_CUDA(cudaMemsetAsync(rec, 0, sizeof(float) * (size_t)B * L * Fr, st));   // row t=0 stays zero (:175)
  MVC_CUDA(cudaMemsetAsync(w.c, 0, sizeof(float) * (size_t)B * Fr, st));
  MVC_CUDA(cudaMemsetAsync(w.hs, 0, es * (size_t)B * Fr, st));
  for (int s = 0; s < S; ++s) {
    const int t = s + 1;
    if (perm) {
      // gates = h_rec_s . W_hh^T + Gx[s] + Gp[b] -> cell, one tcgen05 launch (Gp rides on the row-gather addend)
      TcEpilogue ep{};
      ep.mode = TC_MODE_CELL;
      ep.H = Fr;
      ep.gx = w.gx + (int64_t)s * B * G4; ep.gx_ld = G4;
      ep.embtab = w.gp; ep.tokens = w.iota;
      ep.c_prev = w.c + (int64_t)s * B * Fr; ep.act = w.act + (int64_t)s * B * G4; ep.c_out = w.c + (int64_t)(s + 1) * B * Fr;
      ep.h32 = rec + (int64_t)t * Fr; ep.h_ld = (int64_t)L * Fr;                                  // feats_recons[t] (:183)
      ep.hb = (__nv_bfloat16*)mptr(w.hs, (int64_t)(s + 1) * B * Fr, es); ep.hb_ld = Fr;
      MVC_TRY(tc_gemm(B, 4 * Fr, Fr, cptr(w.hs, (int64_t)s * B * Fr, es), Fr, w.whh_b, Fr, ep,
                      (s > 0 ? TC_FLAG_PDL : 0) | TC_FLAG_B_CONST, st));
      continue;
    }
    if (s == 0) {
      MVC_CUDA(cudaMemsetAsync(w.pre, 0, sizeof(float) * (size_t)B * G4, st));
    } else {
      MVC_TRY(gemm_nt(d->precision, B, 4 * Fr, Fr, cptr(w.hs, (int64_t)s * B * Fr, es), Fr,
                      bf ? w.whh_b : (const void*)p->w_hh, Fr, 0.f, w.pre, G4, nullptr, st));
    }
    // gates = pre + Gx[s] + Gp[b]  (Gp rides on the cell kernel's row-gather addend)
    MVC_TRY(mvc_lstm_cell_fwd(B, Fr, w.pre, w.gx + (int64_t)s * B * G4, G4, w.gp, w.iota, nullptr,
                              w.c + (int64_t)s * B * Fr, w.act + (int64_t)s * B * G4, w.c + (int64_t)(s + 1) * B * Fr,
                              rec + (int64_t)t * Fr, (int64_t)L * Fr,                        // feats_recons[t] (:183)
                              bf ? nullptr : (float*)mptr(w.hs, (int64_t)(s + 1) * B * Fr, es), Fr,
                              bf ? mptr(w.hs, (int64_t)(s + 1) * B * Fr, es) : nullptr, Fr, st));
  }
  return 0;
}

extern "C" int mvc_global_recon_backward(const MvcReconDims* d, const MvcReconParams* p, const float* hid,
                                         const uint8_t* mask, const float* drec, const void* fwd_workspace, float* dhid,
                                         MvcReconGrads* g, void* bwd_workspace, size_t bwd_workspace_bytes,
                                         void* stream) {
  MVC_TRY(check_recon_dims(d, false));
  MVC_CHECK(p && hid && mask && drec && fwd_workspace && dhid && g && bwd_workspace,
            "mvc_global_recon_backward: null argument");
  GlobWs w = glob_layout(d, const_cast<void*>(fwd_workspace));
  GlobBwdWs q = glob_bwd_layout(d, bwd_workspace);
  MVC_CHECK(bwd_workspace_bytes >= q.bytes, "mvc_global_recon_backward: workspace %zu < %zu", bwd_workspace_bytes, q.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, L = d->L, H = d->H, Fr = d->Fr, S = L - 1, SB = S * B;
  const bool bf = d->precision == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  const int G4 = 4 * Fr;
  const int SBp = pad8(SB), Bp = pad8(B);
  const float* hid1 = hid + (int64_t)B * H;

  const int perm = rec_perm(d), permH = perm ? Fr : 0;
  MVC_CUDA(cudaMemsetAsync(q.dc, 0, sizeof(float) * (size_t)B * Fr, st));
  if (bf) MVC_TRY(mvc_transpose_to_bf16(w.whh_b, 1, G4, Fr, Fr, q.whhT, G4, st));
  for (int s = S - 1; s >= 0; --s) {
    const int t = s + 1;
    float* dG = q.dG + (int64_t)s * B * G4;
    void* dGb = bf ? mptr(q.dG_b, (int64_t)s * B * G4, 2) : nullptr;
    MVC_TRY(launch_cell_bwd(B, Fr, w.act + (int64_t)s * B * G4, w.c + (int64_t)s * B * Fr,
                            w.c + (int64_t)(s + 1) * B * Fr, drec + (int64_t)t * Fr, (int64_t)L * Fr,
                            s == S - 1 ? nullptr : q.dhcar, Fr, q.dc, dG, dGb, perm, st));
    if (s > 0) {   // dh_rec_s = dgates . W_hh
      if (bf) MVC_TRY(gemm_tc_chain(B, Fr, G4, dGb, G4, q.whhT, G4, 0.f, q.dhcar, Fr, nullptr, 0, true, st));
      else MVC_TRY(mvc_gemm_f32(B, Fr, G4, 1.f, dG, G4, 1, p->w_hh, 1, Fr, 0.f, q.dhcar, Fr, nullptr, st));
    }
  }
  sum_over_steps_kernel<<<gridn((int64_t)B * G4), 256, 0, st>>>(q.dG, S, (int64_t)B * G4, q.dGsum);
  MVC_LAUNCH_CHECK();
  MVC_TRY(launch_colsum(q.dGsum, B, G4, G4, g->b_ih, permH, st));
  MVC_CUDA(cudaMemcpyAsync(g->b_hh, g->b_ih, sizeof(float) * (size_t)G4, cudaMemcpyDeviceToDevice, st));
  MVC_CUDA(cudaMemsetAsync(dhid, 0, sizeof(float) * (size_t)L * B * H, st));
  float* dhid1 = dhid + (int64_t)B * H;
  if (!bf) {
    const float* hs = (const float*)w.hs;      // rows (s,b): h_rec_s for s = 0..S-1
    MVC_TRY(mvc_gemm_f32(G4, Fr, SB, 1.f, q.dG, 1, G4, hs, 1, Fr, 0.f, g->w_hh, Fr, nullptr, st));
    MVC_TRY(mvc_gemm_f32(G4, H, SB, 1.f, q.dG, 1, G4, hid1, 1, H, 0.f, g->w_ih, 2 * H, nullptr, st));
    MVC_TRY(mvc_gemm_f32(G4, H, B, 1.f, q.dGsum, 1, G4, w.pooled, 1, H, 0.f, g->w_ih + H, 2 * H, nullptr, st));
    MVC_TRY(mvc_gemm_f32(SB, H, G4, 1.f, q.dG, G4, 1, p->w_ih, 1, 2 * H, 0.f, dhid1, H, nullptr, st));
    MVC_TRY(mvc_gemm_f32(B, H, G4, 1.f, q.dGsum, G4, 1, p->w_ih + H, 1, 2 * H, 0.f, q.dpooled, H, nullptr, st));
  } else {
    MVC_TRY(launch_transpose_bf16(q.dG_b, 1, SB, G4, G4, q.dGT, SBp, permH, st));        // natural gate rows
    MVC_TRY(mvc_transpose_to_bf16(w.hs, 1, SB, Fr, Fr, q.hsT, SBp, st));
    MVC_TRY(mvc_transpose_to_bf16(w.hid_b, 1, SB, H, H, q.hidT, SBp, st));
    MVC_TRY(mvc_transpose_to_bf16(w.wih_b, 1, G4, 2 * H, 2 * H, q.wihT, G4, st));
    MVC_TRY(mvc_cast_bf16(q.dGsum, q.dGsum_b, (int64_t)B * G4, st));
    MVC_TRY(launch_transpose_bf16(q.dGsum_b, 1, B, G4, G4, q.dGsumT, Bp, permH, st));
    MVC_TRY(mvc_transpose_to_bf16(w.pooled_b, 1, B, H, H, q.pooledT, Bp, st));
    MVC_TRY(mvc_gemm_bf16(G4, Fr, SB, q.dGT, SBp, q.hsT, SBp, 0.f, g->w_hh, Fr, nullptr, nullptr, 0, st));
    MVC_TRY(mvc_gemm_bf16(G4, H, SB, q.dGT, SBp, q.hidT, SBp, 0.f, g->w_ih, 2 * H, nullptr, nullptr, 0, st));
    MVC_TRY(mvc_gemm_bf16(G4, H, B, q.dGsumT, Bp, q.pooledT, Bp, 0.f, g->w_ih + H, 2 * H, nullptr, nullptr, 0, st));
    MVC_TRY(mvc_gemm_bf16(SB, H, G4, q.dG_b, G4, q.wihT, G4, 0.f, dhid1, H, nullptr, nullptr, 0, st));
    MVC_TRY(mvc_gemm_bf16(B, H, G4, q.dGsum_b, G4, cptr(q.wihT, (int64_t)H * G4, 2), G4, 0.f, q.dpooled, H, nullptr,
                          nullptr, 0, st));
  }
  masked_mean_bwd_kernel<<<(unsigned)cdiv((int64_t)B * H, 256), 256, 0, st>>>(q.dpooled, mask, L, B, H, dhid);
  MVC_LAUNCH_CHECK();
  (void)es;
  return 0;
}

// ------------------------------------------------------------------ local
namespace mvc {
struct LocWs {
  void* keys;       // [B, L, H] compute dtype (decoder hiddens, batch-major)
  float* uk;        // [B, L, A]
  float* wq;        // [T, B, A]
  float* alpha;     // [T, B, L]
  void* xh;         // [T+1, B, H+Fr] compute dtype: slot t = [ctx_t ; h_rec_t]
  float* act;       // [T, B, 4Fr]
  float* c;         // [T+1, B, Fr]
  void* wcat;       // [4Fr, H+Fr] compute dtype
  void* U;          // bf16 [A, H]
  void* W;          // bf16 [A, Fr]
  float* bsum;      // [4Fr]
  float* pre;       // [B, 4Fr]
  size_t bytes;
};
static LocWs loc_layout(const MvcReconDims* d, void* base) {
  const int64_t B = d->B, L = d->L, H = d->H, Fr = d->Fr, A = d->A, T = d->T;
  const bool bf = d->precision == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  Arena ar(base);
  LocWs w{};
  w.keys = ar.take<char>(B * L * H * es);
  w.uk = ar.take<float>(B * L * A);
  w.wq = ar.take<float>(T * B * A);
  w.alpha = ar.take<float>(T * B * L);
  w.xh = ar.take<char>((T + 1) * B * (H + Fr) * es);
  w.act = ar.take<float>(T * B * 4 * Fr);
  w.c = ar.take<float>((T + 1) * B * Fr);
  w.wcat = ar.take<char>(4 * Fr * (H + Fr) * es);
  if (bf) {
    w.U = ar.take<char>(A * H * 2);
    w.W = ar.take<char>(A * Fr * 2);
  }
  w.bsum = ar.take<float>(4 * Fr);
  w.pre = ar.take<float>(B * 4 * Fr);
  w.bytes = ar.off + 256;
  return w;
}
struct LocBwdWs {
  float* dG;        // [T*B, 4Fr]
  float* dxh;       // [B, H+Fr]
  float* dc;        // [B, Fr]
  float* dwq;       // [T*B, A]
  float* duk;       // [B*L, A]
  float* dwpart;    // [B, A]
  float* dkeys;     // [B, L, H]
  void* dG_b;       // bf16 [T*B, 4Fr]
  void* wcatT;      // bf16 [H+Fr, 4Fr]
  void* dGT;        // bf16 [4Fr, TBp]
  void* xhT;        // bf16 [H+Fr, TBp]
  void* dwqT;       // bf16 [A, TBp]
  void* dukT;       // bf16 [A, BLp]
  void* keysT;      // bf16 [H, BLp]
  void* duk_b;      // bf16 [B*L, A]
  void* UT;         // bf16 [H, A]
  void* dwq_b;      // bf16 [T*B, A]
  void* attWT;      // bf16 [Fr, A]
  size_t bytes;
};
static LocBwdWs loc_bwd_layout(const MvcReconDims* d, void* base) {
  const int64_t B = d->B, L = d->L, H = d->H, Fr = d->Fr, A = d->A, T = d->T;
  const bool bf = d->precision == MVC_BF16;
  const int64_t TBp = pad8((int)(T * B)), BLp = pad8((int)(B * L));
  Arena ar(base);
  LocBwdWs w{};
  w.dG = ar.take<float>(T * B * 4 * Fr);
  w.dxh = ar.take<float>(B * (H + Fr));
  w.dc = ar.take<float>(B * Fr);
  w.dwq = ar.take<float>(T * B * A);
  w.duk = ar.take<float>(B * L * A);
  w.dwpart = ar.take<float>(B * A);
  w.dkeys = ar.take<float>(B * L * H);
  if (bf) {
    w.dG_b = ar.take<char>(T * B * 4 * Fr * 2);
    w.wcatT = ar.take<char>((H + Fr) * 4 * Fr * 2);
    w.dGT = ar.take<char>(4 * Fr * TBp * 2);
    w.xhT = ar.take<char>((H + Fr) * TBp * 2);
    w.dwqT = ar.take<char>(A * TBp * 2);
    w.dukT = ar.take<char>(A * BLp * 2);
    w.keysT = ar.take<char>(H * BLp * 2);
    w.duk_b = ar.take<char>(B * L * A * 2);
    w.UT = ar.take<char>(H * A * 2);
    w.dwq_b = ar.take<char>(T * B * A * 2);
    w.attWT = ar.take<char>(Fr * A * 2);
  }
  w.bytes = ar.off + 256;
  return w;
}
static StepCfg loc_cfg(const MvcReconDims* d, const MvcReconParams* p, const LocWs& w, const uint8_t* mask) {
  const bool bf = d->precision == MVC_BF16;
  StepCfg c{};
  c.prec = d->precision;
  c.T = d->L;            // keys per row = caption positions
  c.F = d->H;            // key / context width = decoder hidden size
  c.H = d->Fr;           // LSTM hidden = reconstructed feature size
  c.A = d->A;
  c.perm = rec_perm(d);
  c.uk = w.uk;
  c.keys = w.keys; c.keys_batch = d->B; c.k_sb = (int64_t)d->L * d->H; c.k_st = d->H;
  c.mask = mask; c.m_sb = 1; c.m_st = d->B;           // caption_masks is [L,B]; transposed view (reconstructor.py:69)
  c.wcat = w.wcat; c.wcatT = nullptr;
  c.attW = bf ? w.W : (const void*)p->att_W;
  c.attW32 = p->att_W;
  c.att_b = p->att_b; c.att_w = p->att_w;
  c.cell_bias = w.bsum;
  c.embtab = nullptr;
  c.pre = w.pre;
  return c;
}
}  // namespace mvc

extern "C" size_t mvc_local_recon_workspace_bytes(const MvcReconDims* d) { return loc_layout(d, nullptr).bytes; }
extern "C" size_t mvc_local_recon_bwd_workspace_bytes(const MvcReconDims* d) { return loc_bwd_layout(d, nullptr).bytes; }

extern "C" int mvc_local_recon_forward(const MvcReconDims* d, const MvcReconParams* p, const float* hid,
                                       const uint8_t* mask, float* rec, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  MVC_TRY(check_recon_dims(d, true));
  MVC_CHECK(p && hid && mask && rec && workspace && p->att_W && p->att_U && p->att_b && p->att_w,
            "mvc_local_recon_forward: null argument");
  LocWs w = loc_layout(d, workspace);
  MVC_CHECK(workspace_bytes >= w.bytes, "mvc_local_recon_forward: workspace %zu < %zu", workspace_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, L = d->L, H = d->H, Fr = d->Fr, A = d->A, T = d->T;
  const bool bf = d->precision == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  const int64_t ldx = H + Fr;

  // decoder hiddens, batch-major (reconstructor.py:79-82)
  if (bf) permute_lbh_kernel<__nv_bfloat16><<<gridn((int64_t)L * B * H), 256, 0, st>>>(hid, L, B, H, (__nv_bfloat16*)w.keys);
  else permute_lbh_kernel<float><<<gridn((int64_t)L * B * H), 256, 0, st>>>(hid, L, B, H, (float*)w.keys);
  MVC_LAUNCH_CHECK();
  const int perm = rec_perm(d);
  MVC_TRY(launch_add_vec(p->b_ih, p->b_hh, w.bsum, 4 * Fr, perm ? Fr : 0, st));
  MVC_TRY(launch_pack_wcat(p->w_ih, H, p->w_hh, H, Fr, w.wcat, bf, perm, st));
  if (bf) {
    MVC_TRY(mvc_cast_bf16(p->att_U, w.U, (int64_t)A * H, st));
    MVC_TRY(mvc_cast_bf16(p->att_W, w.W, (int64_t)A * Fr, st));
  }
  // uk = keys . U^T (hoisted; the reference recomputes it for every frame, temporal_attention.py:21)
  MVC_TRY(gemm_nt(d->precision, B * L, A, H, w.keys, H, bf ? w.U : (const void*)p->att_U, H, 0.f, w.uk, A, nullptr, st));
  MVC_CUDA(cudaMemsetAsync(w.c, 0, sizeof(float) * (size_t)B * Fr, st));
  MVC_CUDA(cudaMemsetAsync(w.xh, 0, es * (size_t)B * ldx, st));
  const StepCfg cfg = loc_cfg(d, p, w, mask);
  for (int t = 0; t < T; ++t) {
    StepFwd io{};
    io.rows = B;
    io.xh_src = mptr(w.xh, (int64_t)t * B * ldx, es);
    io.xh_dst = mptr(w.xh, (int64_t)(t + 1) * B * ldx, es);
    io.wq = w.wq + (int64_t)t * B * A;
    io.alpha = w.alpha + (int64_t)t * B * L;
    io.act = w.act + (int64_t)t * B * 4 * Fr;
    io.c_prev = w.c + (int64_t)t * B * Fr;
    io.c_out = w.c + (int64_t)(t + 1) * B * Fr;
    io.gx = nullptr; io.tokens = nullptr;
    io.h_out32 = rec + (int64_t)t * Fr;                 // feats_recons[t] -> [B,T,Fr] (:90-91)
    io.h_ld = (int64_t)T * Fr;
    io.first = (t == 0);
    MVC_TRY(step_forward(cfg, io, st));
  }
  return 0;
}

extern "C" int mvc_local_recon_backward(const MvcReconDims* d, const MvcReconParams* p, const float* hid,
                                        const uint8_t* mask, const float* drec, const void* fwd_workspace, float* dhid,
                                        MvcReconGrads* g, void* bwd_workspace, size_t bwd_workspace_bytes,
                                        void* stream) {
  MVC_TRY(check_recon_dims(d, true));
  MVC_CHECK(p && hid && mask && drec && fwd_workspace && dhid && g && bwd_workspace,
            "mvc_local_recon_backward: null argument");
  LocWs w = loc_layout(d, const_cast<void*>(fwd_workspace));
  LocBwdWs q = loc_bwd_layout(d, bwd_workspace);
  MVC_CHECK(bwd_workspace_bytes >= q.bytes, "mvc_local_recon_backward: workspace %zu < %zu", bwd_workspace_bytes, q.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, L = d->L, H = d->H, Fr = d->Fr, A = d->A, T = d->T, TB = T * B, G4 = 4 * Fr;
  const bool bf = d->precision == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  const int64_t ldx = H + Fr;
  const int TBp = pad8(TB), BLp = pad8(B * L);

  MVC_CUDA(cudaMemsetAsync(q.dc, 0, sizeof(float) * (size_t)B * Fr, st));
  MVC_CUDA(cudaMemsetAsync(q.duk, 0, sizeof(float) * (size_t)B * L * A, st));
  MVC_CUDA(cudaMemsetAsync(q.dwpart, 0, sizeof(float) * (size_t)B * A, st));
  MVC_CUDA(cudaMemsetAsync(q.dkeys, 0, sizeof(float) * (size_t)B * L * H, st));
  StepCfg cfg = loc_cfg(d, p, w, mask);
  const int permH = cfg.perm ? Fr : 0;
  if (bf) {
    MVC_TRY(mvc_transpose_to_bf16(w.wcat, 1, G4, H + Fr, ldx, q.wcatT, G4, st));
    MVC_TRY(mvc_transpose_to_bf16(w.W, 1, A, Fr, Fr, q.attWT, A, st));
    cfg.wcatT = q.wcatT;
    cfg.attWT = q.attWT;
  }
  for (int t = T - 1; t >= 0; --t) {
    StepBwd io{};
    io.rows = B;
    io.act = w.act + (int64_t)t * B * G4;
    io.c_prev = w.c + (int64_t)t * B * Fr;
    io.c_new = w.c + (int64_t)(t + 1) * B * Fr;
    io.dh_ext = drec + (int64_t)t * Fr;
    io.dh_ld = (int64_t)T * Fr;
    io.has_carry = (t != T - 1);
    io.dc = q.dc;
    io.dG = q.dG + (int64_t)t * B * G4;
    io.dG_b = bf ? mptr(q.dG_b, (int64_t)t * B * G4, 2) : nullptr;
    io.dxh = q.dxh;
    io.wq = w.wq + (int64_t)t * B * A;
    io.alpha = w.alpha + (int64_t)t * B * L;
    io.dwq = q.dwq + (int64_t)t * B * A;
    io.dwq_b = bf ? mptr(q.dwq_b, (int64_t)t * B * A, 2) : nullptr;
    io.duk = q.duk;
    io.dwpart = q.dwpart;
    io.dkeys = q.dkeys; io.dk_sb = (int64_t)L * H; io.dk_st = H;
    io.first = (t == 0);
    MVC_TRY(step_backward(cfg, io, st));
  }
  MVC_TRY(mvc_colsum(q.dwq, TB, A, A, g->att_b, st));
  MVC_TRY(mvc_colsum(q.dwpart, B, A, A, g->att_w, st));
  MVC_TRY(launch_colsum(q.dG, TB, G4, G4, g->b_ih, permH, st));
  MVC_CUDA(cudaMemcpyAsync(g->b_hh, g->b_ih, sizeof(float) * (size_t)G4, cudaMemcpyDeviceToDevice, st));
  const char* hprev = cptr(w.xh, H, es);          // h_rec_t for t = 0..T-1 (h-part of slots 0..T-1)
  if (!bf) {
    MVC_TRY(mvc_gemm_f32(A, Fr, TB, 1.f, q.dwq, 1, A, (const float*)hprev, 1, ldx, 0.f, g->att_W, Fr, nullptr, st));
    MVC_TRY(mvc_gemm_f32(A, H, B * L, 1.f, q.duk, 1, A, (const float*)w.keys, 1, H, 0.f, g->att_U, H, nullptr, st));
    MVC_TRY(mvc_gemm_f32(G4, H, TB, 1.f, q.dG, 1, G4, (const float*)w.xh, 1, ldx, 0.f, g->w_ih, H, nullptr, st));
    MVC_TRY(mvc_gemm_f32(G4, Fr, TB, 1.f, q.dG, 1, G4, (const float*)hprev, 1, ldx, 0.f, g->w_hh, Fr, nullptr, st));
    // dkeys += duk . U   (uk = keys . U^T)
    MVC_TRY(mvc_gemm_f32(B * L, H, A, 1.f, q.duk, A, 1, p->att_U, 1, H, 1.f, q.dkeys, H, nullptr, st));
  } else {
    MVC_TRY(launch_transpose_bf16(q.dG_b, 1, TB, G4, G4, q.dGT, TBp, permH, st));        // natural gate rows
    MVC_TRY(mvc_transpose_to_bf16(w.xh, 1, TB, H + Fr, ldx, q.xhT, TBp, st));
    MVC_TRY(mvc_transpose_to_bf16(q.dwq_b, 1, TB, A, A, q.dwqT, TBp, st));
    MVC_TRY(mvc_transpose_to_bf16(q.duk, 0, B * L, A, A, q.dukT, BLp, st));
    MVC_TRY(mvc_transpose_to_bf16(w.keys, 1, B * L, H, H, q.keysT, BLp, st));
    const char* hprevT = cptr(q.xhT, (int64_t)H * TBp, 2);
    MVC_TRY(mvc_gemm_bf16(A, Fr, TB, q.dwqT, TBp, hprevT, TBp, 0.f, g->att_W, Fr, nullptr, nullptr, 0, st));
    MVC_TRY(mvc_gemm_bf16(A, H, B * L, q.dukT, BLp, q.keysT, BLp, 0.f, g->att_U, H, nullptr, nullptr, 0, st));
    MVC_TRY(mvc_gemm_bf16(G4, H, TB, q.dGT, TBp, q.xhT, TBp, 0.f, g->w_ih, H, nullptr, nullptr, 0, st));
    MVC_TRY(mvc_gemm_bf16(G4, Fr, TB, q.dGT, TBp, hprevT, TBp, 0.f, g->w_hh, Fr, nullptr, nullptr, 0, st));
    MVC_TRY(mvc_cast_bf16(q.duk, q.duk_b, (int64_t)B * L * A, st));
    MVC_TRY(mvc_transpose_to_bf16(w.U, 1, A, H, H, q.UT, A, st));
    MVC_TRY(mvc_gemm_bf16(B * L, H, A, q.duk_b, A, q.UT, A, 1.f, q.dkeys, H, nullptr, nullptr, 0, st));
  }
  MVC_CUDA(cudaMemsetAsync(dhid, 0, sizeof(float) * (size_t)L * B * H, st));
  permute_add_blh_kernel<<<gridn((int64_t)L * B * H), 256, 0, st>>>(q.dkeys, L, B, H, dhid);
  MVC_LAUNCH_CHECK();
  return 0;
}
