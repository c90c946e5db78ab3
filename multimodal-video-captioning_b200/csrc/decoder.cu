// decoder.cu -- host-side orchestration of the SA-LSTM decoder on one stream:
// teacher-forced / scheduled-sampling / free-running forward, BPTT backward,
// greedy ids and beam search.  No device synchronisation, no allocation.
//
// Restructuring relative to the reference's per-step Python loop
// (features_captioning.py:91-119), all exact algebra:
//   * U.feats is loop invariant -> one GEMM before the loop (the reference
//     recomputes it every step, temporal_attention.py:21);
//   * the embedding half of the LSTM input projection is hoisted: under full
//     teacher forcing as one [S*B,E]x[E,4H] GEMM, otherwise as a [V,4H] table
//     (embedding . W_ih[:, :E]^T + b_ih + b_hh) gathered per step inside the
//     cell kernel;
//   * the vocabulary projection + log-softmax runs once over all S*B rows when
//     no step needs its own argmax;
//   * backward: every weight gradient is one GEMM over all S*B rows after the
//     time loop; only d[ctx;h] = dgates.[W_ih[:,E:]|W_hh], the attention
//     backward and dh += dwq.W stay inside the loop.
#include <mutex>

#include <unordered_map>

#include "recur.cuh"
#include "recur2.cuh"
#include "step.cuh"

namespace mvc {

struct DecWs {
  // saved for backward
  void* feats;    // [B*T, F]        compute dtype
  float* uk;      // [B*T, A]
  float* wq;      // [S, B, A]
  float* alpha;   // [S, B, T]
  void* xh;       // [S+1, B, F+H]   compute dtype: slot s = [ctx_s ; h_s]
  float* act;     // [S, B, 4H]
  float* c;       // [S+1, B, H]
  void* xemb;     // [S*B, Ep]       compute dtype
  // weights in compute dtype (rebuilt every forward)
  void* wcat;     // [4H, F+H] = [W_ih[:,E:] | W_hh]
  void* wie;      // [4H, Ep]  = W_ih[:, :E] (bf16: zero padded to Ep)
  void* U;        // [A, F]
  void* W;        // [A, H]
  void* outw;     // [V, H]
  void* embb;     // [V, Ep]
  float* bsum;    // [4H] = b_ih + b_hh
  // scratch
  float* pre;     // [B, 4H]
  float* gx;      // [S*B, 4H]
  float* embtab;  // [V, 4H]
  float* hzero;   // [B, H] zeros (c_0 / fp32 h_0)
  unsigned* sync; // grid-barrier counters of the persistent recurrence kernels
  void* P;        // [B*T, 4H] fp16: keys . W_ih[:, E:]^T, unit-major gate columns (recur2 path)
  float* gh;      // [4, 128, 4H]  K-slice partials of h_s . W_hh^T of the current step (recur2 path)
  // transposed weights of the backward pass, written by a training forward next to the weight casts (recur2 path)
  void* outwT;    // [H, Vp]
  void* attWT;    // [H, A]
  void* whhT;     // [H, 4H]   unit-major gate columns
  void* wieT;     // [Ep, 4H]
  void* featsT;   // [F, BTp]  keys, K-major operand of dU = duk^T . feats
  void* xallT;    // [E + F + H (+8), SBp]  [emb ; ctx ; h_prev]^T, K-major operand of the merged LSTM weight-gradient GEMM
  size_t bytes;
};

// Which time-loop implementation a forward call ran, keyed by its workspace: the backward call must read the saved
// activations in the layout that forward wrote (tile-interleaved vs unit-major gate columns, ctx present or not).
enum { DEC_LOOP_CHAIN = 0, DEC_LOOP_RECUR1 = 1, DEC_LOOP_RECUR2 = 2, DEC_LOOP_MASK = 0xff,
       DEC_FLAG_WT_READY = 0x100 };   // forward also wrote the transposed weights of the backward pass
static std::mutex g_loop_mu;
static std::unordered_map<const void*, int> g_loop_mode;
static void set_loop_mode(const void* ws, int mode) {
  std::lock_guard<std::mutex> lk(g_loop_mu);
  if (g_loop_mode.size() > 4096) g_loop_mode.clear();
  g_loop_mode[ws] = mode;
}
static int get_loop_mode(const void* ws) {
  std::lock_guard<std::mutex> lk(g_loop_mu);
  auto it = g_loop_mode.find(ws);
  return it == g_loop_mode.end() ? -1 : it->second;
}

// A library-owned side stream per (device, caller stream) with its own fork / join events: work that does not feed
// the recurrence (weight packing, the weight gradients of the vocabulary projection, operand transposes) runs there,
// next to the caller's stream.  Keyed by the caller's stream so that two user streams (or a CUDA-graph capture stream)
// never share events; `SideGuard` makes every exit path -- including MVC_TRY / MVC_CUDA early returns between fork and
// join -- order the caller's stream after whatever the side stream was given, so the torch-owned workspaces the side
// stream writes cannot be recycled under it.
struct SideStream {
  cudaStream_t stream = nullptr, stream2 = nullptr;    // stream2: a second leg for three-way splits
  cudaEvent_t fork = nullptr, join = nullptr, join2 = nullptr;
  cudaEvent_t aux[2] = {nullptr, nullptr};             // intermediate dependencies between the two streams
};
static int get_side_stream(cudaStream_t caller, SideStream** out) {
  static std::mutex mu;
  static std::unordered_map<uint64_t, SideStream*> pool;
  int dev = 0;
  MVC_CUDA(cudaGetDevice(&dev));
  const uint64_t key = (reinterpret_cast<uint64_t>(caller) << 8) ^ (uint64_t)(dev & 0xff);
  std::lock_guard<std::mutex> lk(mu);
  auto it = pool.find(key);
  if (it == pool.end()) {
    SideStream* s = new SideStream();
    MVC_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    MVC_CUDA(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));
    MVC_CUDA(cudaEventCreateWithFlags(&s->join2, cudaEventDisableTiming));
    MVC_CUDA(cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming));
    MVC_CUDA(cudaEventCreateWithFlags(&s->join, cudaEventDisableTiming));
    for (auto& e : s->aux) MVC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    it = pool.emplace(key, s).first;
  }
  *out = it->second;
  return 0;
}
struct SideGuard {
  cudaStream_t caller;
  SideStream* side = nullptr;
  bool open = false;                                   // forked and not yet joined
  bool open2 = false;                                  // second leg forked and not yet joined
  explicit SideGuard(cudaStream_t c) : caller(c) {}
  int fork2() {                                        // right after fork(): the second leg starts at the same point
    MVC_CUDA(cudaStreamWaitEvent(side->stream2, side->fork, 0));
    open2 = true;
    return 0;
  }
  int mark2() {
    MVC_CUDA(cudaEventRecord(side->join2, side->stream2));
    return 0;
  }
  int fork() {                                         // side stream continues after everything enqueued on `caller`
    if (!side) MVC_TRY(get_side_stream(caller, &side));
    MVC_CUDA(cudaEventRecord(side->fork, caller));
    MVC_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    open = true;
    return 0;
  }
  int mark() {                                         // "everything given to the side stream so far"
    MVC_CUDA(cudaEventRecord(side->join, side->stream));
    return 0;
  }
  int side_waits(int i) {                              // side stream waits for what `caller` holds now (forked already)
    MVC_CUDA(cudaEventRecord(side->aux[i], caller));
    MVC_CUDA(cudaStreamWaitEvent(side->stream, side->aux[i], 0));
    return 0;
  }
  int caller_waits(int i) {                            // caller waits for what the side stream holds now
    MVC_CUDA(cudaEventRecord(side->aux[i], side->stream));
    MVC_CUDA(cudaStreamWaitEvent(caller, side->aux[i], 0));
    return 0;
  }
  int join() {                                         // caller waits for the last mark()
    if (open) {
      MVC_CUDA(cudaStreamWaitEvent(caller, side->join, 0));
      open = false;
    }
    if (open2) {
      MVC_CUDA(cudaStreamWaitEvent(caller, side->join2, 0));
      open2 = false;
    }
    return 0;
  }
  ~SideGuard() {
    if (open) {                                        // error path: never leave the side stream running un-joined
      cudaEventRecord(side->join, side->stream);
      cudaStreamWaitEvent(caller, side->join, 0);
    }
    if (open2) {
      cudaEventRecord(side->join2, side->stream2);
      cudaStreamWaitEvent(caller, side->join2, 0);
    }
  }
};

// tile-interleaved gate order (fused gate-GEMM + cell epilogue): bf16 path with H a multiple of 32
static inline int dec_perm(const MvcDecoderDims* d) { return d->precision == MVC_BF16 && d->H % 32 == 0; }

static DecWs dec_layout(const MvcDecoderDims* d, void* base) {
  const int64_t B = d->B, T = d->T, F = d->F, H = d->H, E = d->E, A = d->A, V = d->V, S = d->L - 1;
  const bool bf = d->precision == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  const int64_t Ep = bf ? pad8((int)E) : E;
  Arena ar(base);
  DecWs w{};
  w.feats = ar.take<char>(B * T * F * es);
  w.uk = ar.take<float>(B * T * A);
  w.wq = ar.take<float>(S * B * A);
  w.alpha = ar.take<float>(S * B * T);
  w.xh = ar.take<char>((S + 1) * B * (F + H) * es);
  w.act = ar.take<float>(S * B * 4 * H);
  w.c = ar.take<float>((S + 1) * B * H);
  w.xemb = ar.take<char>(S * B * Ep * es);
  w.wcat = ar.take<char>((4 * H + (bf ? A : 0)) * (F + H) * es);   // bf16: + A rows [U | 0] (merged P / U.k GEMM, recur2)
  w.wie = bf ? ar.take<char>(4 * H * Ep * es) : nullptr;
  w.U = bf ? ar.take<char>(A * F * es) : nullptr;
  // bf16 vocabulary weights, followed (after padding to a multiple of 256 rows) by the attention query weights: one
  // B operand [aux0 + A, H] lets the decode loops fold W.h of the next step into the vocabulary GEMM (TcAux)
  w.outw = bf ? ar.take<char>((tc_aux_row0((int)V) + A) * H * es) : nullptr;
  w.W = bf ? (void*)((char*)w.outw + (size_t)tc_aux_row0((int)V) * H * es) : nullptr;
  w.embb = bf ? ar.take<char>(V * Ep * es) : nullptr;
  w.bsum = ar.take<float>(4 * H);
  w.pre = ar.take<float>(B * 4 * H);
  w.gx = ar.take<float>(S * B * 4 * H);
  w.embtab = ar.take<float>(V * 4 * H);
  w.hzero = ar.take<float>(B * H);
  w.sync = ar.take<unsigned>(1024);   // recur1: [0,32); recur2 forward flags: [256,512); recur2 backward flags: [512,768)
  w.P = bf ? ar.take<char>(B * T * 4 * H * 2) : nullptr;
  w.gh = bf ? ar.take<float>(4 * 128 * 4 * H) : nullptr;
  w.outwT = bf ? ar.take<char>(H * pad8((int)V) * 2) : nullptr;
  w.attWT = bf ? ar.take<char>(H * A * 2) : nullptr;
  w.whhT = bf ? ar.take<char>(H * 4 * H * 2) : nullptr;
  w.wieT = bf ? ar.take<char>(Ep * 4 * H * 2) : nullptr;
  w.featsT = bf ? ar.take<char>(F * pad8((int)(B * T)) * 2) : nullptr;
  w.xallT = bf ? ar.take<char>((E + F + H + 8) * pad8((int)(S * B)) * 2) : nullptr;
  w.bytes = ar.off + 256;
  return w;
}

// ------------------------------------------------------------------ small kernels
// out[4H, F+H] = [w_x (4H x F, row pitch wx_ld) | w_hh (4H x H)]
// row r of the output is gate row gate_unperm(H, r) of the weights when perm (tile-interleaved gate order)
template <typename OutT>
__global__ void pack_wcat_kernel(const float* __restrict__ w_x, int64_t wx_ld, const float* __restrict__ w_hh, int F,
                                 int H, OutT* __restrict__ out, int perm) {
  const int64_t K = F + H, total = (int64_t)4 * H * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ro = i / K;
    const int k = (int)(i - ro * K);
    const int64_t r = perm == 2 ? gate_unperm(-H, (int)ro) : (perm ? gate_unperm(H, (int)ro) : ro);   // 2: unit-major
    const float v = k < F ? w_x[r * wx_ld + k] : w_hh[r * H + (k - F)];
    if constexpr (sizeof(OutT) == 2) out[i] = __float2bfloat16(v);
    else out[i] = v;
  }
}

// bf16 variant with 16-byte loads: one block row per output row (no 64-bit division per element), 4 elements per thread
// (rows [4H, 4H + n_extra) of the output, if any: [extra (n_extra x F) | 0])
__global__ void __launch_bounds__(256)
pack_wcat_vec_kernel(const float* __restrict__ w_x, int64_t wx_ld, const float* __restrict__ w_hh, int F, int H,
                     __nv_bfloat16* __restrict__ out, int perm, const float* __restrict__ extra) {
  const int ro = blockIdx.y;
  const int K = F + H;
  const int k = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (k >= K) return;
  float4 v;
  if (ro >= 4 * H) {
    v = k < F ? *reinterpret_cast<const float4*>(extra + (int64_t)(ro - 4 * H) * F + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    const int64_t r = perm == 2 ? gate_unperm(-H, ro) : (perm ? gate_unperm(H, ro) : ro);
    v = k < F ? *reinterpret_cast<const float4*>(w_x + r * wx_ld + k)
              : *reinterpret_cast<const float4*>(w_hh + r * H + (k - F));
  }
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<const uint32_t*>(&lo);
  o.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(out + (int64_t)ro * K + k) = o;
}

// out[r, 0:Cp] = bf16(src[r*lds + 0:C]) zero padded
__global__ void cast_pad_bf16_kernel(const float* __restrict__ src, int64_t rows, int C, int64_t lds, int Cp,
                                     __nv_bfloat16* __restrict__ out, int permH) {
  const int64_t total = rows * Cp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ro = i / Cp;
    const int c = (int)(i - ro * Cp);
    const int64_t r = permH ? gate_unperm(permH, (int)ro) : ro;
    out[i] = __float2bfloat16(c < C ? src[r * lds + c] : 0.f);
  }
}

// same, 4 columns per thread (16-byte loads, 8-byte stores), one block row per output row: no 64-bit division per
// element (the reconstructors' 23 M + 19 M element weight casts took 35 + 68 us with the scalar kernel)
__global__ void __launch_bounds__(256)
cast_pad_bf16_vec_kernel(const float* __restrict__ src, int C, int64_t lds, int Cp, __nv_bfloat16* __restrict__ out,
                         int permH) {
  const int ro = blockIdx.y;
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= Cp) return;
  const int64_t r = permH ? gate_unperm(permH, ro) : ro;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c + 3 < C) {
    v = *reinterpret_cast<const float4*>(src + r * lds + c);
  } else if (c < C) {
    const float* p = src + r * lds + c;
    v.x = p[0];
    if (c + 1 < C) v.y = p[1];
    if (c + 2 < C) v.z = p[2];
  }
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<const uint32_t*>(&lo);
  o.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(out + (int64_t)ro * Cp + c) = o;
}

__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int n,
                               int permH) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int j = permH ? gate_unperm(permH, i) : i;
    o[i] = a[j] + b[j];
  }
}

__global__ void fill_i64_kernel(int64_t* __restrict__ p, int64_t v, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ids[b*ld + col] = src[b]
__global__ void scatter_col_i64_kernel(const int64_t* __restrict__ src, int64_t* __restrict__ dst, int64_t ld, int col,
                                       int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) dst[b * ld + col] = src[b];
}

// dst[r*ldd + c] (+)= src[r*lds + c]
__global__ void add_rows_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd,
                                int64_t rows, int C, int accumulate) {
  const int64_t total = rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const float v = src[r * lds + c];
    if (accumulate) dst[r * ldd + c] += v;
    else dst[r * ldd + c] = v;
  }
}


__global__ void iota_i64_kernel(int64_t* __restrict__ p, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

int launch_pack_wcat(const float* w_x, int64_t wx_ld, const float* w_hh, int F, int H, void* out, int out_bf16, int perm,
                     cudaStream_t st, const float* extra, int n_extra) {
  const int64_t n = (int64_t)4 * H * (F + H);
  if (out_bf16 && F % 4 == 0 && H % 4 == 0 && wx_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(w_x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(w_hh) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0 &&
      (!extra || (reinterpret_cast<uintptr_t>(extra) & 15) == 0)) {
    pack_wcat_vec_kernel<<<dim3((unsigned)cdiv(F + H, 1024), (unsigned)(4 * H + (extra ? n_extra : 0))), 256, 0, st>>>(
        w_x, wx_ld, w_hh, F, H, (__nv_bfloat16*)out, perm, extra);
    MVC_LAUNCH_CHECK();
    return 0;
  }
  if (extra)       // unaligned views: the extra rows by the generic cast (bf16 output only)
    MVC_TRY(launch_cast_pad_bf16(extra, n_extra, F, F, F + H, mptr(out, (int64_t)4 * H * (F + H), 2), 0, st));
  if (out_bf16) pack_wcat_kernel<__nv_bfloat16><<<gridn(n), 256, 0, st>>>(w_x, wx_ld, w_hh, F, H, (__nv_bfloat16*)out, perm);
  else pack_wcat_kernel<float><<<gridn(n), 256, 0, st>>>(w_x, wx_ld, w_hh, F, H, (float*)out, perm);
  MVC_LAUNCH_CHECK();
  return 0;
}
int launch_cast_pad_bf16(const float* src, int64_t rows, int C, int64_t lds, int Cp, void* out, int permH,
                         cudaStream_t st) {
  if (rows > 0 && rows <= 65535 && Cp % 4 == 0 && lds % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 7) == 0) {
    cast_pad_bf16_vec_kernel<<<dim3((unsigned)cdiv(Cp, 1024), (unsigned)rows), 256, 0, st>>>(src, C, lds, Cp,
                                                                                           (__nv_bfloat16*)out, permH);
    MVC_LAUNCH_CHECK();
    return 0;
  }
  cast_pad_bf16_kernel<<<gridn(rows * Cp), 256, 0, st>>>(src, rows, C, lds, Cp, (__nv_bfloat16*)out, permH);
  MVC_LAUNCH_CHECK();
  return 0;
}
int launch_add_vec(const float* a, const float* b, float* o, int n, int permH, cudaStream_t st) {
  add_vec_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(a, b, o, n, permH);
  MVC_LAUNCH_CHECK();
  return 0;
}
int launch_fill_i64(int64_t* p, int64_t v, int64_t n, cudaStream_t st) {
  fill_i64_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(p, v, n);
  MVC_LAUNCH_CHECK();
  return 0;
}
int launch_iota_i64(int64_t* p, int64_t n, cudaStream_t st) {
  iota_i64_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(p, n);
  MVC_LAUNCH_CHECK();
  return 0;
}
int launch_add_rows(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int C, int accumulate,
                    cudaStream_t st) {
  add_rows_kernel<<<gridn(rows * C), 256, 0, st>>>(src, lds, dst, ldd, rows, C, accumulate);
  MVC_LAUNCH_CHECK();
  return 0;
}

// Prepare weights/features in the compute dtype + U.feats.  Shared by forward, greedy and beam.
static int dec_prepare(const MvcDecoderDims* d, const MvcDecoderParams* p, const float* audio, int Fa,
                       const float* visual, int Fv, DecWs& w, bool need_embtab, cudaStream_t st, bool unit_major = false) {
  const int B = d->B, T = d->T, F = d->F, H = d->H, E = d->E, A = d->A, V = d->V;
  const bool bf = d->precision == MVC_BF16;
  const int Ep = bf ? pad8(E) : E;
  const int perm = unit_major ? 2 : dec_perm(d);
  const int permH = perm == 2 ? -H : (perm ? H : 0);
  MVC_CHECK(Fa + Fv == F, "decoder: Fa (%d) + Fv (%d) != in_feature_size (%d)", Fa, Fv, F);
  // weight packing / casting on the side stream, concurrent with the feature cast and the U.k GEMM
  SideGuard sg(st);
  MVC_TRY(sg.fork());
  cudaStream_t ss = sg.side->stream;
  MVC_TRY(launch_add_vec(p->b_ih, p->b_hh, w.bsum, 4 * H, permH, ss));
  MVC_TRY(launch_pack_wcat(p->w_ih + E, E + F, p->w_hh, F, H, w.wcat, bf, perm, ss));
  if (bf) {
    MVC_TRY(launch_cast_pad_bf16(p->w_ih, 4 * H, E, E + F, Ep, w.wie, permH, ss));
    MVC_TRY(mvc_cast_bf16(p->att_W, w.W, (int64_t)A * H, ss));
    MVC_TRY(mvc_cast_bf16(p->out_w, w.outw, (int64_t)V * H, ss));
    if (tc_aux_row0(V) > V)
      MVC_CUDA(cudaMemsetAsync((char*)w.outw + (size_t)V * H * 2, 0, (size_t)(tc_aux_row0(V) - V) * H * 2, ss));
    if (need_embtab) MVC_TRY(launch_cast_pad_bf16(p->embedding, V, E, E, Ep, w.embb, 0, ss));
  }
  MVC_TRY(sg.mark());
  if (mvc_get_input_format() == MVC_INPUT_BF16) {
    // pre-packed bf16 feature shards: half the H2D bytes upstream, no cast pass here
    MVC_CHECK(bf, "decoder: bf16 input features need precision = MVC_BF16");
    MVC_TRY(mvc_concat_bf16(audio, Fa, visual, Fv, (int64_t)B * T, w.feats, st));
  } else {
    MVC_TRY(mvc_concat_cast(audio, Fa, visual, Fv, (int64_t)B * T, w.feats, bf, st));
  }
  if (bf) MVC_TRY(mvc_cast_bf16(p->att_U, w.U, (int64_t)A * F, st));
  // uk = feats . U^T      (temporal_attention.py:21, hoisted)
  MVC_TRY(gemm_nt(d->precision, B * T, A, F, w.feats, F, bf ? w.U : (const void*)p->att_U, F, 0.f, w.uk, A, nullptr, st));
  MVC_TRY(sg.join());
  if (need_embtab) {
    // embtab[v,:] = embedding[v] . W_ih[:, :E]^T + b_ih + b_hh
    if (bf) MVC_TRY(gemm_nt(MVC_BF16, V, 4 * H, Ep, w.embb, Ep, w.wie, Ep, 0.f, w.embtab, 4 * H, w.bsum, st));
    else MVC_TRY(gemm_nt(MVC_F32, V, 4 * H, E, p->embedding, E, p->w_ih, E + F, 0.f, w.embtab, 4 * H, w.bsum, st));
  }
  return 0;
}

// The whole prelude of a fully teacher-forced bf16 forward on the projected-keys path (recur2), scheduled over two
// streams so that the serial chain in front of the persistent kernel is  features -> [P | U.k] GEMM  only:
//   side:   bias sum, [W_c | W_hh ; U | 0] pack (-> event 0), W_ie cast, <SOS> / caption tokens, embedding gather, W cast,
//           state + progress-counter clears, gx GEMM (-> event 1: everything the persistent kernel reads); then, UNDER the
//           persistent kernel (it leaves 20 SMs idle): W_out cast, output clears and (training) the transposed weights /
//           keys the backward pass will need -- joined by the caller after the kernel
//   caller: feature concat / cast, (wait event 0) one GEMM for P = keys . W_c^T and U.k = keys . U^T, (wait event 1) kernel
// (the round-2 profile of the single-stream order: 139 us from step start to the persistent kernel, 80 us of it GEMMs;
// now ~100 us, 64 of it GEMMs.  A captured graph issues its root kernels one after the other in enqueue order, so the
// features leg is enqueued first.)
static int dec_prepare_r2(const MvcDecoderDims* d, const MvcDecoderParams* p, const float* audio, int Fa,
                          const float* visual, int Fv, const int64_t* captions, float* out_logp, float* out_hid,
                          int64_t* tokens_in, DecWs& w, bool train, SideGuard& sg, cudaStream_t st) {
  const int B = d->B, T = d->T, F = d->F, H = d->H, E = d->E, A = d->A, V = d->V, S = d->L - 1;
  const int Ep = pad8(E);
  const int64_t ldx = F + H;
  MVC_CHECK(Fa + Fv == F, "decoder: Fa (%d) + Fv (%d) != in_feature_size (%d)", Fa, Fv, F);
  MVC_TRY(sg.fork());
  cudaStream_t ss = sg.side->stream;
  // (enqueue order = issue order of a captured graph's roots: the features leg first)
  if (mvc_get_input_format() == MVC_INPUT_BF16) MVC_TRY(mvc_concat_bf16(audio, Fa, visual, Fv, (int64_t)B * T, w.feats, st));
  else MVC_TRY(mvc_concat_cast(audio, Fa, visual, Fv, (int64_t)B * T, w.feats, 1, st));
  MVC_TRY(launch_add_vec(p->b_ih, p->b_hh, w.bsum, 4 * H, -H, ss));
  // rows 4H .. 4H+A of the packed weights: [U | 0] -- the hoisted U.k projection rides in the P GEMM as one more N tile
  // (352 -> 396 tiles of 128 x 256: three waves either way), instead of a 17 us GEMM of its own behind it
  MVC_TRY(launch_pack_wcat(p->w_ih + E, E + F, p->w_hh, F, H, w.wcat, 1, 2, ss, p->att_U, A));
  MVC_CUDA(cudaEventRecord(sg.side->aux[0], ss));
  MVC_CUDA(cudaStreamWaitEvent(st, sg.side->aux[0], 0));
  {
    // P = keys . W_ih[:, E:]^T  ([B*T, F] x [F, 4H], 16-bit out, unit-major gate columns): the context half of every
    // step's gate pre-activation becomes sum_t alpha_t P[b,t,:], accumulated out of tensor memory by the row owner
    // (stored as fp16, saturating: |P| beyond 65504 means fully saturated gates anyway, and the 11-bit mantissa keeps
    // the rounding error of the pre-activation 4x below bf16's on unnormalised features)
    // columns [4H, 4H + A): uk = feats . U^T, fp32 (temporal_attention.py:21, hoisted)
    TcEpilogue ep{};
    ep.mode = TC_MODE_PLAIN;
    ep.Cb = (__nv_bfloat16*)w.P; ep.ldcb = 4 * H; ep.cb_f16 = 1;
    ep.C2 = w.uk; ep.ldc2 = A; ep.split_n = 4 * H;
    MVC_TRY(tc_gemm(B * T, 4 * H + A, F, w.feats, F, w.wcat, ldx, ep, 0, st));
  }
  MVC_TRY(launch_cast_pad_bf16(p->w_ih, 4 * H, E, E + F, Ep, w.wie, -H, ss));
  // tokens_in[0] = <SOS>, tokens_in[s] = captions[s] (features_captioning.py:99, :121-125); hoisted embedding GEMM
  MVC_TRY(launch_fill_i64(tokens_in, MVC_SOS, B, ss));
  if (S > 1)
    MVC_CUDA(cudaMemcpyAsync(tokens_in + B, captions + B, sizeof(int64_t) * (size_t)(S - 1) * B, cudaMemcpyDeviceToDevice, ss));
  MVC_TRY(mvc_embedding_gather(p->embedding, E, tokens_in, (int64_t)S * B, w.xemb, Ep, 1, ss));
  MVC_TRY(mvc_cast_bf16(p->att_W, w.W, (int64_t)A * H, ss));
  MVC_CUDA(cudaMemsetAsync(out_hid, 0, sizeof(float) * (size_t)B * H, ss));       // hidden_states[0] = 0 (:98)
  MVC_CUDA(cudaMemsetAsync(w.c, 0, sizeof(float) * (size_t)B * H, ss));           // c_0 = 0 (:66-75)
  MVC_CUDA(cudaMemsetAsync(w.xh, 0, (size_t)2 * B * ldx, ss));                    // h_0 = 0: clear slot 0
  MVC_CUDA(cudaMemsetAsync(w.sync + 256, 0, sizeof(unsigned) * 256, ss));         // progress counters of the kernel
  // (last on this leg: the GEMM queues behind the caller's P GEMM for SMs)
  MVC_TRY(gemm_nt(MVC_BF16, S * B, 4 * H, Ep, w.xemb, Ep, w.wie, Ep, 0.f, w.gx, 4 * H, w.bsum, ss));
  MVC_CUDA(cudaEventRecord(sg.side->aux[1], ss));
  // what only the vocabulary projection / the backward pass need: behind the features (event 0 of the caller's stream),
  // next to the persistent kernel
  MVC_TRY(sg.side_waits(0));
  MVC_TRY(mvc_cast_bf16(p->out_w, w.outw, (int64_t)V * H, ss));
  if (tc_aux_row0(V) > V)
    MVC_CUDA(cudaMemsetAsync((char*)w.outw + (size_t)V * H * 2, 0, (size_t)(tc_aux_row0(V) - V) * H * 2, ss));
  MVC_CUDA(cudaMemsetAsync(out_logp, 0, sizeof(float) * (size_t)B * V, ss));      // sentence[0] = 0  (:96)
  if (train) {
    MVC_TRY(mvc_transpose_to_bf16(cptr(w.wcat, F, 2), 1, 4 * H, H, ldx, w.whhT, 4 * H, ss));
    MVC_TRY(mvc_transpose_to_bf16(w.W, 1, A, H, H, w.attWT, A, ss));
    MVC_TRY(mvc_transpose_to_bf16(w.outw, 1, V, H, H, w.outwT, pad8(V), ss));
    MVC_TRY(mvc_transpose_to_bf16(w.wie, 1, 4 * H, Ep, Ep, w.wieT, 4 * H, ss));
    MVC_TRY(mvc_transpose_to_bf16(w.feats, 1, B * T, F, F, w.featsT, pad8(B * T), ss));
  }
  MVC_TRY(sg.mark());
  MVC_CUDA(cudaStreamWaitEvent(st, sg.side->aux[1], 0));
  return 0;
}

static StepCfg dec_cfg(const MvcDecoderDims* d, const MvcDecoderParams* p, const DecWs& w, float* pre) {
  const bool bf = d->precision == MVC_BF16;
  StepCfg c{};
  c.prec = d->precision;
  c.T = d->T; c.F = d->F; c.H = d->H; c.A = d->A;
  c.perm = dec_perm(d);
  c.uk = w.uk;
  c.keys = w.feats; c.keys_batch = d->B; c.k_sb = (int64_t)d->T * d->F; c.k_st = d->F;
  c.mask = nullptr; c.m_sb = 0; c.m_st = 0;                 // decoder attention is unmasked
  c.wcat = w.wcat; c.wcatT = nullptr;
  c.attW = bf ? w.W : (const void*)p->att_W;
  c.attW32 = p->att_W;
  c.att_b = p->att_b; c.att_w = p->att_w;
  c.cell_bias = nullptr;                                    // b_ih + b_hh is folded into gx / embtab
  c.embtab = w.embtab;
  c.pre = pre ? pre : w.pre;
  return c;
}

}  // namespace mvc

using namespace mvc;

extern "C" int mvc_pack_gate_rows_bf16(const float* w, int H, int C, int64_t ldw, int Cp, void* out, void* stream) {
  MVC_CHECK(w && out && H > 0 && H % 32 == 0 && C > 0 && Cp >= C && Cp % 8 == 0,
            "mvc_pack_gate_rows_bf16: need H %% 32 == 0 and Cp %% 8 == 0 (H=%d C=%d Cp=%d)", H, C, Cp);
  return launch_cast_pad_bf16(w, 4 * (int64_t)H, C, ldw, Cp, out, H, (cudaStream_t)stream);
}

extern "C" int mvc_lstm_gates_cell_bf16(int B, int H, int K, const void* x, int64_t ldx, const void* w_packed, int64_t ldw,
                                        const float* bias_packed, const float* gx_packed, int64_t gx_ld,
                                        const float* c_prev, float* act_packed, float* c_out, float* h_out, int64_t h_ld,
                                        void* h_bf16, int64_t hb_ld, void* stream) {
  MVC_CHECK(x && w_packed && c_out && H % 32 == 0, "mvc_lstm_gates_cell_bf16: bad arguments (H=%d)", H);
  TcEpilogue ep{};
  ep.mode = TC_MODE_CELL;
  ep.H = H;
  ep.bias = bias_packed;
  ep.gx = gx_packed; ep.gx_ld = gx_ld;
  ep.c_prev = c_prev; ep.act = act_packed; ep.c_out = c_out;
  ep.h32 = h_out; ep.h_ld = h_ld;
  ep.hb = (__nv_bfloat16*)h_bf16; ep.hb_ld = hb_ld;
  return tc_gemm(B, 4 * H, K, x, ldx, w_packed, ldw, ep, 0, (cudaStream_t)stream);
}

extern "C" size_t mvc_decoder_fwd_workspace_bytes(const MvcDecoderDims* d, int) { return dec_layout(d, nullptr).bytes; }

static int check_dims(const MvcDecoderDims* d) {
  MVC_CHECK(d, "decoder: null dims");
  MVC_CHECK(d->B > 0 && d->T > 0 && d->F > 0 && d->H > 0 && d->E > 0 && d->A > 0 && d->V > 2 && d->L >= 2,
            "decoder: bad dims B=%d T=%d F=%d H=%d E=%d A=%d V=%d L=%d", d->B, d->T, d->F, d->H, d->E, d->A, d->V, d->L);
  MVC_CHECK(d->precision == MVC_F32 || d->precision == MVC_BF16, "decoder: unknown precision %d", d->precision);
  if (d->precision == MVC_BF16)
    MVC_CHECK(d->F % 8 == 0 && d->H % 8 == 0 && d->A % 8 == 0,
              "decoder(bf16): F, H, A must be multiples of 8 (TMA 16-byte row pitch); got F=%d H=%d A=%d", d->F, d->H, d->A);
  return 0;
}

extern "C" int mvc_decoder_forward(const MvcDecoderDims* d, const MvcDecoderParams* p, const float* audio, int Fa,
                                   const float* visual, int Fv, const int64_t* captions, const uint8_t* tf_flags_host,
                                   float* out_logp, float* out_hid, int64_t* tokens_in, void* workspace,
                                   size_t workspace_bytes, int save_for_backward, void* stream) {
  MVC_TRY(check_dims(d));
  MVC_CHECK(p && out_logp && out_hid && tokens_in && workspace, "mvc_decoder_forward: null argument");
  DecWs w = dec_layout(d, workspace);
  MVC_CHECK(workspace_bytes >= w.bytes, "mvc_decoder_forward: workspace %zu < %zu", workspace_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, T = d->T, F = d->F, H = d->H, E = d->E, A = d->A, V = d->V, S = d->L - 1;
  const bool bf = d->precision == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  const int Ep = bf ? pad8(E) : E;
  const int64_t ldx = F + H;

  // Every step teacher forced?  (captions given and all flags set; the flag of the
  // last step is never consumed: its "next input" is not fed to anything.)
  bool all_tf = captions != nullptr;
  if (captions) {
    MVC_CHECK(tf_flags_host, "mvc_decoder_forward: captions without tf flags");
    for (int i = 0; i + 1 < S; ++i) all_tf = all_tf && tf_flags_host[i];
  }

  // fully teacher-forced bf16 sequences take the projected-keys persistent kernels (recur2.cuh)
  const bool use_r2 = bf && all_tf && recur2_supported(B, T, F, H, A);
  SideGuard sg_r2(st);
  if (use_r2) {
    MVC_TRY(dec_prepare_r2(d, p, audio, Fa, visual, Fv, captions, out_logp, out_hid, tokens_in, w, save_for_backward != 0,
                           sg_r2, st));
  } else {
    MVC_TRY(dec_prepare(d, p, audio, Fa, visual, Fv, w, !all_tf, st, false));
    MVC_CUDA(cudaMemsetAsync(out_logp, 0, sizeof(float) * (size_t)B * V, st));      // sentence[0] = 0  (:96)
    MVC_CUDA(cudaMemsetAsync(out_hid, 0, sizeof(float) * (size_t)B * H, st));       // hidden_states[0] = 0 (:98)
    MVC_CUDA(cudaMemsetAsync(w.c, 0, sizeof(float) * (size_t)B * H, st));           // c_0 = 0 (:66-75)
    // h_0 = 0 in slot 0's h-part (strided): clear the whole slot 0
    MVC_CUDA(cudaMemsetAsync(w.xh, 0, es * (size_t)B * ldx, st));
    MVC_TRY(launch_fill_i64(tokens_in, MVC_SOS, B, st));                              // first input = <SOS> (:99)
  }

  if (all_tf && !use_r2) {
    // tokens_in[s] = captions[s] for s >= 1 ; hoisted embedding GEMM
    if (S > 1)
      MVC_CUDA(cudaMemcpyAsync(tokens_in + B, captions + B, sizeof(int64_t) * (size_t)(S - 1) * B,
                               cudaMemcpyDeviceToDevice, st));
    MVC_TRY(mvc_embedding_gather(p->embedding, E, tokens_in, (int64_t)S * B, w.xemb, Ep, bf, st));
    MVC_TRY(gemm_nt(d->precision, S * B, 4 * H, Ep, w.xemb, Ep, bf ? w.wie : (const void*)p->w_ih, bf ? Ep : E + F, 0.f,
                    w.gx, 4 * H, w.bsum, st));
  }

  const StepCfg cfg = dec_cfg(d, p, w, nullptr);
  const bool persistent = use_r2 || (cfg.perm && recur_fwd_supported(B, T, F, H, A));
  set_loop_mode(workspace, use_r2 ? (DEC_LOOP_RECUR2 | (save_for_backward ? DEC_FLAG_WT_READY : 0))
                                  : (persistent && all_tf ? DEC_LOOP_RECUR1 : DEC_LOOP_CHAIN));
  if (use_r2) {
    Recur2FwdParams rp{};
    rp.B = B; rp.T = T; rp.F = F; rp.K = F + H; rp.S = S;
    rp.P = (const __nv_bfloat16*)w.P; rp.uk = w.uk; rp.attW = (const __nv_bfloat16*)w.W;
    rp.att_b = p->att_b; rp.att_w = p->att_w; rp.gx = w.gx;
    rp.xh = (__nv_bfloat16*)w.xh; rp.c = w.c; rp.act = w.act; rp.alpha = w.alpha; rp.wq_out = w.wq;
    rp.out_hid = out_hid; rp.gh = w.gh; rp.sync = w.sync + 256;
    MVC_TRY(recur2_fwd_launch(rp, cptr(w.wcat, F, 2), ldx, st, /*sync_cleared=*/true));
    MVC_TRY(sg_r2.join());                        // W_out cast, clears, transposed weights: done under the kernel
    if (save_for_backward) {
      // Training: what the backward pass needs from the saved activations alone -- the ctx halves of the xh slots (this
      // path never formed them: ctx_s = sum_t alpha_s,t keys_t) and [emb ; ctx ; h_prev]^T, the K-major operand of the
      // merged LSTM weight-gradient GEMM -- is produced HERE, on the side stream, under the vocabulary projection (FMA /
      // copy work next to a tensor-pipe kernel; 25 us hidden in a 58 us window).  Under the persistent backward kernel
      // the same work had 20 SMs and delayed the kernel's own launch.
      const int SB = S * B, SBp = pad8(SB);
      MVC_TRY(sg_r2.fork());
      cudaStream_t ss = sg_r2.side->stream;
      MVC_TRY(r2_ctx_rows(w.feats, w.alpha, B, T, F, S, w.xh, ldx, ss));
      MVC_TRY(mvc_transpose_to_bf16(w.xemb, 1, SB, Ep, Ep, w.xallT, SBp, ss));       // before the xh block: its Ep - E pad rows
      MVC_TRY(mvc_transpose_to_bf16(w.xh, 1, SB, F + H, ldx, (char*)w.xallT + (size_t)E * SBp * 2, SBp, ss));
      MVC_TRY(sg_r2.mark());
    }
  } else if (persistent && all_tf) {
    // the whole teacher-forced time loop in ONE persistent cluster-cooperative launch (recur_fwd.cu)
    RecurFwdParams rp{};
    rp.B = B; rp.T = T; rp.F = F; rp.H = H; rp.A = A; rp.K = F + H; rp.S = S; rp.s0 = 0; rp.s1 = S;
    rp.feats = (const __nv_bfloat16*)w.feats; rp.uk = w.uk; rp.attW = (const __nv_bfloat16*)w.W;
    rp.att_b = p->att_b; rp.att_w = p->att_w;
    rp.gx = w.gx; rp.embtab = nullptr; rp.tokens = nullptr; rp.cell_bias = nullptr;
    rp.xh = (__nv_bfloat16*)w.xh; rp.c = w.c; rp.act = w.act; rp.alpha = w.alpha; rp.wq_out = w.wq;
    rp.out_hid = out_hid; rp.sync = w.sync;
    MVC_TRY(recur_fwd_launch(rp, w.wcat, st));
  }
  for (int s = 0; s < S && !(persistent && all_tf); ++s) {
    const int t = s + 1;
    StepFwd io{};
    io.rows = B;
    io.xh_src = mptr(w.xh, (int64_t)s * B * ldx, es);
    io.xh_dst = mptr(w.xh, (int64_t)(s + 1) * B * ldx, es);
    io.wq = w.wq + (int64_t)s * B * A;
    io.alpha = w.alpha + (int64_t)s * B * T;
    io.act = w.act + (int64_t)s * B * 4 * H;
    io.c_prev = w.c + (int64_t)s * B * H;
    io.c_out = w.c + (int64_t)(s + 1) * B * H;
    io.gx = all_tf ? w.gx + (int64_t)s * B * 4 * H : nullptr;
    io.tokens = all_tf ? nullptr : tokens_in + (int64_t)s * B;
    io.h_out32 = out_hid + (int64_t)t * B * H;                                      // hidden_states[t] (:107)
    io.h_ld = H;
    io.first = (s == 0);
    MVC_TRY(step_forward(cfg, io, st));
    if (!all_tf) {
      // logits -> log-probs (+ argmax) for this step                           (:87-88, :109)
      float* lp = out_logp + (int64_t)t * B * V;
      MVC_TRY(gemm_nt(d->precision, B, V, H, bf ? cptr(io.xh_dst, F, es) : (const char*)io.h_out32, bf ? ldx : H,
                      bf ? w.outw : (const void*)p->out_w, H, 0.f, lp, V, p->out_b, st));
      const bool feed_caption = captions && tf_flags_host[s];
      const bool last = (s + 1 == S);
      int64_t* nxt = last ? nullptr : tokens_in + (int64_t)(s + 1) * B;
      MVC_TRY(mvc_log_softmax_rows(lp, B, V, (last || feed_caption) ? nullptr : nxt, st));
      if (!last && feed_caption)
        MVC_CUDA(cudaMemcpyAsync(nxt, captions + (int64_t)t * B, sizeof(int64_t) * B, cudaMemcpyDeviceToDevice, st));
    }
  }
  if (all_tf) {
    // all S*B rows at once: log_softmax(h . out_w^T + out_b)
    float* lp = out_logp + (int64_t)B * V;
    if (bf) {
      // K-D: vocabulary projection with the log-sum-exp kept in the tcgen05 epilogue; `pre` is free scratch here
      MVC_TRY(tc_gemm_logsoftmax(S * B, V, H, cptr(w.xh, (int64_t)B * ldx + F, es), ldx, w.outw, H, p->out_b, lp, V, w.pre,
                                 sizeof(float) * (size_t)B * 4 * H, st));
    } else {
      MVC_TRY(gemm_nt(MVC_F32, S * B, V, H, out_hid + (int64_t)B * H, H, p->out_w, H, 0.f, lp, V, p->out_b, st));
      MVC_TRY(mvc_log_softmax_rows(lp, (int64_t)S * B, V, nullptr, st));
    }
  }
  MVC_TRY(sg_r2.join());
  return 0;
}

// ------------------------------------------------------------------ backward
namespace mvc {
struct DecBwdWs {
  float* dlogits;   // [S*B, V]
  float* dhall;     // [S*B, H]
  float* dG;        // [S*B, 4H]
  float* dxh;       // [B, F+H]
  float* dhcar;     // [B, H] carried dh (from step s+1)
  float* dc;        // [B, H]
  float* dwq;       // [S*B, A]
  float* duk;       // [B*T, A]
  float* dwpart;    // [B, A]
  float* dxemb;     // [S*B, E]
  // bf16 operands
  void* dlogits_b;  // [S*B, Vp]
  void* outwT;      // [H, Vp]
  void* dlogitsT;   // [V, SBp]
  void* hallT;      // [H, SBp]
  void* dG_b;       // [S*B, 4H]
  void* wcatT;      // [F+H, 4H]
  void* dGT;        // [4H, SBp]
  void* xhT;        // [F+H, SBp]
  void* xembT;      // [Ep, SBp]
  void* wieT;       // [Ep, 4H]
  void* dukT;       // [A, BTp]
  void* featsT;     // [F, BTp]
  void* dwqT;       // [A, SBp]
  void* dwq_b;      // [S*B, A]
  void* attWT;      // [H, A]
  void* whhT;       // [H, 4H] W_hh^T, unit-major gate columns (recur2 path)
  float* ghb;       // [4, 128, H] K-slice partials of dG_s . W_hh of the current step (recur2 path)
  size_t bytes;
};

static DecBwdWs dec_bwd_layout(const MvcDecoderDims* d, void* base) {
  const int64_t B = d->B, T = d->T, F = d->F, H = d->H, E = d->E, A = d->A, V = d->V, S = d->L - 1;
  const bool bf = d->precision == MVC_BF16;
  const int64_t Vp = pad8((int)V), SBp = pad8((int)(S * B)), BTp = pad8((int)(B * T)), Ep = pad8((int)E);
  Arena ar(base);
  DecBwdWs w{};
  w.dlogits = ar.take<float>(S * B * V);
  w.dhall = ar.take<float>(S * B * H);
  w.dG = ar.take<float>(S * B * 4 * H);
  w.dxh = ar.take<float>(B * (F + H));
  w.dhcar = ar.take<float>(B * H);
  w.dc = ar.take<float>(B * H);
  w.dwq = ar.take<float>(S * B * A);
  w.duk = ar.take<float>(B * T * A);
  w.dwpart = ar.take<float>(B * A);
  w.dxemb = ar.take<float>(S * B * E);
  if (bf) {
    w.dlogits_b = ar.take<char>(S * B * Vp * 2);
    w.outwT = ar.take<char>(H * Vp * 2);
    w.dlogitsT = ar.take<char>(V * SBp * 2);
    w.hallT = ar.take<char>(H * SBp * 2);
    w.dG_b = ar.take<char>(S * B * 4 * H * 2);
    w.wcatT = ar.take<char>((F + H) * 4 * H * 2);
    w.dGT = ar.take<char>(4 * H * SBp * 2);
    // [emb ; ctx ; h]^T in ONE buffer (rows E + F + H, +8 slack rows): the three LSTM weight gradients that share the
    // operand dG^T become one GEMM.  The embedding block is transposed first (its Ep - E zero pad rows land on the first
    // ctx rows), the [ctx ; h] block after it.
    w.xembT = ar.take<char>((E + F + H + 8) * SBp * 2);
    w.xhT = w.xembT ? (char*)w.xembT + (size_t)E * SBp * 2 : nullptr;
    w.wieT = ar.take<char>(Ep * 4 * H * 2);
    w.dukT = ar.take<char>(A * BTp * 2);
    w.featsT = ar.take<char>(F * BTp * 2);
    w.dwqT = ar.take<char>(A * SBp * 2);
    w.dwq_b = ar.take<char>(S * B * A * 2);
    w.attWT = ar.take<char>(H * A * 2);
    w.whhT = ar.take<char>(H * 4 * H * 2);
    w.ghb = ar.take<float>(4 * 128 * H);
  }
  w.bytes = ar.off + 256;
  return w;
}
}  // namespace mvc

extern "C" size_t mvc_decoder_bwd_workspace_bytes(const MvcDecoderDims* d) { return dec_bwd_layout(d, nullptr).bytes; }

extern "C" int mvc_decoder_backward(const MvcDecoderDims* d, const MvcDecoderParams* p, const float* out_logp,
                                    const float* dlogp, const float* dhid, const int64_t* tokens_in,
                                    const void* fwd_workspace, MvcDecoderGrads* g, void* bwd_workspace,
                                    size_t bwd_workspace_bytes, void* stream) {
  MVC_TRY(check_dims(d));
  MVC_CHECK(p && out_logp && tokens_in && fwd_workspace && g && bwd_workspace, "mvc_decoder_backward: null argument");
  DecWs w = dec_layout(d, const_cast<void*>(fwd_workspace));
  DecBwdWs q = dec_bwd_layout(d, bwd_workspace);
  MVC_CHECK(bwd_workspace_bytes >= q.bytes, "mvc_decoder_backward: workspace %zu < %zu", bwd_workspace_bytes, q.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, T = d->T, F = d->F, H = d->H, E = d->E, A = d->A, V = d->V, S = d->L - 1;
  const bool bf = d->precision == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  const int64_t ldx = F + H;
  const int Vp = pad8(V), SBp = pad8(S * B), BTp = pad8(B * T), Ep = pad8(E);
  const int SB = S * B;
  // fp32 h_1..h_S (contiguous [S*B,H]) are not saved separately: they are the h-parts of xh slots 1..S.
  const char* hall = cptr(w.xh, (int64_t)B * ldx + F, es);    // rows = slots 1..S, ld = ldx
  const char* hprev = cptr(w.xh, F, es);                      // rows = slots 0..S-1
  const int loop_flags = get_loop_mode(fwd_workspace);
  MVC_CHECK(loop_flags >= 0, "mvc_decoder_backward: fwd_workspace was not written by mvc_decoder_forward in this process");
  const int loop_mode = loop_flags & DEC_LOOP_MASK;
  const bool use_r2 = loop_mode == DEC_LOOP_RECUR2;
  // transposed weights: written by the training forward (recur2 path) or produced here
  const bool wt_ready = bf && (loop_flags & DEC_FLAG_WT_READY);
  const void* outwT = wt_ready ? w.outwT : q.outwT;
  const void* attWT = wt_ready ? w.attWT : q.attWT;
  const void* whhT = wt_ready ? w.whhT : q.whhT;
  const void* wieT = wt_ready ? w.wieT : q.wieT;
  const void* featsT = wt_ready ? w.featsT : q.featsT;
  const void* xallT = wt_ready ? w.xallT : q.xembT;           // [emb ; ctx ; h_prev]^T
  const void* xhT = cptr(xallT, (int64_t)E * SBp, 2);

  // ---- vocabulary projection backward (all steps at once)
  // Two streams (round-2 timeline, tools/step_timeline.py): everything that needs forward data only -- state clears, the
  // rebuilt ctx rows, the K-major operand copies of the post-loop weight-gradient GEMMs -- starts on the side stream at
  // entry, next to log-softmax backward and dhall; dW_out / db_out follow there once dlogits exists, under the
  // persistent backward kernel (which leaves 20 SMs idle).
  SideGuard sg(st);
  bool forked = false;
  const bool early = bf && dlogp;          // the side stream carries the clears and the operand copies
  if (early) {
    MVC_TRY(sg.fork());
    forked = true;
    cudaStream_t ss = sg.side->stream;
    MVC_CUDA(cudaMemsetAsync(q.dc, 0, sizeof(float) * (size_t)B * H, ss));
    MVC_CUDA(cudaMemsetAsync(q.duk, 0, sizeof(float) * (size_t)B * T * A, ss));
    MVC_CUDA(cudaMemsetAsync(q.dwpart, 0, sizeof(float) * (size_t)B * A, ss));
    if (use_r2) MVC_CUDA(cudaMemsetAsync(w.sync + 512, 0, sizeof(unsigned) * 256, ss));
    MVC_CUDA(cudaEventRecord(sg.side->aux[0], ss));       // "clears done": the time loop waits for this one only
    MVC_CUDA(cudaMemsetAsync(g->embedding, 0, sizeof(float) * (size_t)V * E, ss));
    if (!wt_ready) {
      // (a forward without save_for_backward, or one of the launch-chain / recur1 paths: operand copies made here)
      if (use_r2) MVC_TRY(r2_ctx_rows(w.feats, w.alpha, B, T, F, S, w.xh, ldx, ss));
      // (a teacher-forced forward left the gathered embedding rows in xemb; the token-fed paths did not)
      if (loop_mode == DEC_LOOP_CHAIN) MVC_TRY(mvc_embedding_gather(p->embedding, E, tokens_in, SB, w.xemb, Ep, 1, ss));
      MVC_TRY(mvc_transpose_to_bf16(w.xemb, 1, SB, Ep, Ep, q.xembT, SBp, ss));        // before xhT: see dec_bwd_layout
      MVC_TRY(mvc_transpose_to_bf16(w.xh, 1, SB, F + H, ldx, q.xhT, SBp, ss));
      MVC_TRY(mvc_transpose_to_bf16(w.feats, 1, B * T, F, F, q.featsT, BTp, ss));
      MVC_TRY(mvc_transpose_to_bf16(w.wie, 1, 4 * H, Ep, Ep, q.wieT, 4 * H, ss));
    }
  }
  if (dlogp) {
    const float* lp = out_logp + (int64_t)B * V;
    const float* dl = dlogp + (int64_t)B * V;
    if (bf) {
      MVC_TRY(mvc_log_softmax_bwd(lp, dl, SB, V, q.dlogits, q.dlogits_b, st));      // (zero-fills the Vp - V pad columns)
      // dW_out = dlogits^T . hall ; db_out = colsum(dlogits) (fp32).  Nothing in the recurrence needs them: side stream,
      // under the persistent backward kernel.  Both operands are consumed where they lie as MN-major tcgen05 operands
      // (dlogits [SB, Vp] and the h halves of the xh slots [SB, F+H]): on the 20 SMs the persistent kernel leaves, the two
      // transpose passes of the K-major form cost more than the GEMM.
      // The persistent backward kernel -- a whole-SM cluster launch, next on the caller's stream behind dhall -- wants 128
      // free SMs; whatever the side stream holds at that moment delays it.  A caller stream that outranks the side stream
      // (GraphedTrainStep captures on a priority -1 stream) gets its blocks placed first when both are ready together:
      // db_out (column sum) then runs next to dhall and dW_out is released together with the kernel.  On a caller stream
      // of equal priority (the eager loop) dW_out's 52 persistent CTAs would win that race often enough to cost 6 %, so
      // the column sum is put between them: it is short-lived, and by the time dW_out starts the kernel is resident.
      int prio = 0;
      MVC_CUDA(cudaStreamGetPriority(st, &prio));
      const bool outranks = prio < 0;
      cudaStream_t ss = sg.side->stream;
      if (outranks) {
        MVC_TRY(sg.side_waits(1));
        MVC_TRY(mvc_colsum(q.dlogits, SB, V, V, g->out_b, ss));
      }
      if (!wt_ready) MVC_TRY(mvc_transpose_to_bf16(p->out_w, 0, V, H, H, q.outwT, Vp, st));
      // dhall = dlogits . out_w
      MVC_TRY(mvc_gemm_bf16(SB, H, V, q.dlogits_b, Vp, outwT, Vp, 0.f, q.dhall, H, nullptr, nullptr, 0, st));
      MVC_TRY(sg.side_waits(1));
      if (!outranks) MVC_TRY(mvc_colsum(q.dlogits, SB, V, V, g->out_b, ss));
      {
        TcEpilogue ep{};
        ep.mode = TC_MODE_PLAIN;
        ep.C = g->out_w; ep.ldc = H;
        MVC_TRY(tc_gemm(V, H, SB, q.dlogits_b, Vp, hall, ldx, ep, TC_FLAG_A_MN | TC_FLAG_B_MN, ss));
      }
      MVC_TRY(sg.mark());
    } else {
      MVC_TRY(mvc_log_softmax_bwd(lp, dl, SB, V, q.dlogits, nullptr, st));
      MVC_TRY(mvc_gemm_f32(SB, H, V, 1.f, q.dlogits, V, 1, p->out_w, 1, H, 0.f, q.dhall, H, nullptr, st));
      MVC_TRY(mvc_gemm_f32(V, H, SB, 1.f, q.dlogits, 1, V, (const float*)hall, 1, ldx, 0.f, g->out_w, H, nullptr, st));
      MVC_TRY(mvc_colsum(q.dlogits, SB, V, V, g->out_b, st));
    }
  } else {
    MVC_CUDA(cudaMemsetAsync(q.dhall, 0, sizeof(float) * (size_t)SB * H, st));
    MVC_CUDA(cudaMemsetAsync(g->out_w, 0, sizeof(float) * (size_t)V * H, st));
    MVC_CUDA(cudaMemsetAsync(g->out_b, 0, sizeof(float) * (size_t)V, st));
  }
  if (dhid) {
    add_rows_kernel<<<gridn((int64_t)SB * H), 256, 0, st>>>(dhid + (int64_t)B * H, H, q.dhall, H, SB, H, 1);
    MVC_LAUNCH_CHECK();
  }

  // ---- time loop
  if (early) {
    MVC_CUDA(cudaStreamWaitEvent(st, sg.side->aux[0], 0));
  } else {
    MVC_CUDA(cudaMemsetAsync(q.dc, 0, sizeof(float) * (size_t)B * H, st));
    MVC_CUDA(cudaMemsetAsync(q.duk, 0, sizeof(float) * (size_t)B * T * A, st));
    MVC_CUDA(cudaMemsetAsync(q.dwpart, 0, sizeof(float) * (size_t)B * A, st));
    MVC_CUDA(cudaMemsetAsync(g->embedding, 0, sizeof(float) * (size_t)V * E, st));
  }
  if (bf && !wt_ready) {
    if (use_r2) MVC_TRY(mvc_transpose_to_bf16(cptr(w.wcat, F, 2), 1, 4 * H, H, ldx, q.whhT, 4 * H, st));
    else MVC_TRY(mvc_transpose_to_bf16(w.wcat, 1, 4 * H, F + H, ldx, q.wcatT, 4 * H, st));
    MVC_TRY(mvc_transpose_to_bf16(w.W, 1, A, H, H, q.attWT, A, st));
  }
  StepCfg cfg = dec_cfg(d, p, w, nullptr);
  cfg.wcatT = q.wcatT;
  cfg.attWT = attWT;
  const int permH = use_r2 ? -H : (cfg.perm ? H : 0);
  const bool persistent_bwd = use_r2 || (bf && cfg.perm && recur_bwd_supported(B, T, F, H, A));
  if (use_r2) {
    if (!dlogp && !wt_ready) MVC_TRY(r2_ctx_rows(w.feats, w.alpha, B, T, F, S, w.xh, ldx, st));   // (else: forward / side stream)
    Recur2BwdParams rp{};
    rp.B = B; rp.T = T; rp.F = F; rp.K = F + H; rp.S = S;
    rp.P = (const __nv_bfloat16*)w.P; rp.uk = w.uk; rp.att_b = p->att_b; rp.att_w = p->att_w;
    rp.act = w.act; rp.c = w.c; rp.wq = w.wq; rp.alpha = w.alpha; rp.dh_ext = q.dhall;
    rp.attWT = (const __nv_bfloat16*)attWT;
    rp.dG = q.dG; rp.dG_b = (__nv_bfloat16*)q.dG_b; rp.dwq = q.dwq; rp.dwq_b = (__nv_bfloat16*)q.dwq_b;
    rp.duk = q.duk; rp.dwpart = q.dwpart; rp.ghb = q.ghb; rp.sync = w.sync + 512;
    MVC_TRY(recur2_bwd_launch(rp, whhT, st, /*sync_cleared=*/early));
  } else if (persistent_bwd) {
    // the whole BPTT time loop in ONE persistent cluster-cooperative launch (recur_bwd.cu)
    RecurBwdParams rp{};
    rp.B = B; rp.T = T; rp.F = F; rp.H = H; rp.A = A; rp.K = F + H; rp.S = S;
    rp.feats = (const __nv_bfloat16*)w.feats; rp.uk = w.uk; rp.att_b = p->att_b; rp.att_w = p->att_w;
    rp.act = w.act; rp.c = w.c; rp.wq = w.wq; rp.alpha = w.alpha; rp.dh_ext = q.dhall;
    rp.attWT = (const __nv_bfloat16*)attWT;
    rp.dG = q.dG; rp.dG_b = (__nv_bfloat16*)q.dG_b; rp.dxh = q.dxh; rp.dwq = q.dwq; rp.dwq_b = (__nv_bfloat16*)q.dwq_b;
    rp.duk = q.duk; rp.dwpart = q.dwpart; rp.sync = w.sync + 16;
    MVC_TRY(recur_bwd_launch(rp, q.wcatT, st));
  }
  for (int s = S - 1; s >= 0 && !persistent_bwd; --s) {
    StepBwd io{};
    io.rows = B;
    io.act = w.act + (int64_t)s * B * 4 * H;
    io.c_prev = w.c + (int64_t)s * B * H;
    io.c_new = w.c + (int64_t)(s + 1) * B * H;
    io.dh_ext = q.dhall + (int64_t)s * B * H;     // from the vocabulary projection (+ dhid)
    io.dh_ld = H;
    io.has_carry = (s != S - 1);
    io.dc = q.dc;
    io.dG = q.dG + (int64_t)s * B * 4 * H;
    io.dG_b = bf ? mptr(q.dG_b, (int64_t)s * B * 4 * H, 2) : nullptr;
    io.dxh = q.dxh;
    io.wq = w.wq + (int64_t)s * B * A;
    io.alpha = w.alpha + (int64_t)s * B * T;
    io.dwq = q.dwq + (int64_t)s * B * A;
    io.dwq_b = bf ? mptr(q.dwq_b, (int64_t)s * B * A, 2) : nullptr;
    io.duk = q.duk;
    io.dwpart = q.dwpart;
    io.dkeys = nullptr;                           // features are inputs: no gradient
    io.first = (s == 0);
    MVC_TRY(step_backward(cfg, io, st));
  }

  // ---- hoisted parameter gradients
  // attention: dW = dwq^T . h_prev ; db = colsum(dwq) ; dw = colsum(dwpart) ; dU = duk^T . feats
  auto bias_grads = [&]() -> int {
    // LSTM biases (db_ih = db_hh = colsum dG), attention biases: one launch
    return launch_colsum3(q.dG, SB, 4 * H, 4 * H, g->b_ih, g->b_hh, permH, q.dwq, SB, A, A, g->att_b, q.dwpart, B, A, A,
                          g->att_w, st);
  };
  if (!bf) {
    MVC_TRY(bias_grads());
    const float* xh = (const float*)w.xh;
    MVC_TRY(mvc_gemm_f32(A, H, SB, 1.f, q.dwq, 1, A, (const float*)hprev, 1, ldx, 0.f, g->att_W, H, nullptr, st));
    MVC_TRY(mvc_gemm_f32(A, F, B * T, 1.f, q.duk, 1, A, (const float*)w.feats, 1, F, 0.f, g->att_U, F, nullptr, st));
    // dW_ih[:, E:] = dG^T . ctx ; dW_hh = dG^T . h_prev
    MVC_TRY(mvc_gemm_f32(4 * H, F, SB, 1.f, q.dG, 1, 4 * H, xh, 1, ldx, 0.f, g->w_ih + E, E + F, nullptr, st));
    MVC_TRY(mvc_gemm_f32(4 * H, H, SB, 1.f, q.dG, 1, 4 * H, (const float*)hprev, 1, ldx, 0.f, g->w_hh, H, nullptr, st));
    // embedding side: x_emb rows are gathered again (the teacher-forced path saved them, the table path did not)
    MVC_TRY(mvc_embedding_gather(p->embedding, E, tokens_in, SB, w.xemb, E, 0, st));
    MVC_TRY(mvc_gemm_f32(4 * H, E, SB, 1.f, q.dG, 1, 4 * H, (const float*)w.xemb, 1, E, 0.f, g->w_ih, E + F, nullptr, st));
    MVC_TRY(mvc_gemm_f32(SB, E, 4 * H, 1.f, q.dG, 4 * H, 1, p->w_ih, 1, E + F, 0.f, q.dxemb, E, nullptr, st));
    MVC_TRY(mvc_embedding_scatter_add(q.dxemb, E, E, tokens_in, SB, g->embedding, st));
  } else {
    // transposed bf16 operands (tcgen05 GEMM takes K-contiguous A[M,K], B[N,K])
    const bool pre_t = forked;                 // xhT / featsT / xembT / wieT were produced on the side stream
    if (forked) {
      MVC_TRY(sg.join());
      forked = false;
    }
    if (!pre_t && !wt_ready) {
      if (loop_mode == DEC_LOOP_CHAIN) MVC_TRY(mvc_embedding_gather(p->embedding, E, tokens_in, SB, w.xemb, Ep, 1, st));
      MVC_TRY(mvc_transpose_to_bf16(w.xemb, 1, SB, Ep, Ep, q.xembT, SBp, st));
      MVC_TRY(mvc_transpose_to_bf16(w.xh, 1, SB, F + H, ldx, q.xhT, SBp, st));
      MVC_TRY(mvc_transpose_to_bf16(w.feats, 1, B * T, F, F, q.featsT, BTp, st));
      MVC_TRY(mvc_transpose_to_bf16(w.wie, 1, 4 * H, Ep, Ep, q.wieT, 4 * H, st));
    }
    const char* hprevT = cptr(xhT, (int64_t)F * SBp, 2);
    // the attention-parameter gradients (small GEMMs) run on the side stream next to the LSTM weight gradients
    // three legs: the LSTM weight gradients + bias sums (caller), embedding + dW_att (side), dU (second side leg) -- the
    // small GEMMs are serial k-block chains on few tiles (15-30 us each), six of them in a row were the tail's critical path
    MVC_TRY(sg.fork());
    MVC_TRY(sg.fork2());
    {
      cudaStream_t ss = sg.side->stream, s2 = sg.side->stream2;
      MVC_TRY(mvc_transpose_to_bf16(q.duk, 0, B * T, A, A, q.dukT, BTp, s2));
      MVC_TRY(mvc_gemm_bf16(A, F, B * T, q.dukT, BTp, featsT, BTp, 0.f, g->att_U, F, nullptr, nullptr, 0, s2));
      MVC_TRY(sg.mark2());
      // the embedding gradient (dxemb = dG . W_ie, scattered into the table) needs neither dG^T nor the main stream
      MVC_TRY(mvc_gemm_bf16(SB, E, 4 * H, q.dG_b, 4 * H, wieT, 4 * H, 0.f, q.dxemb, E, nullptr, nullptr, 0, ss));
      MVC_TRY(mvc_embedding_scatter_add(q.dxemb, E, E, tokens_in, SB, g->embedding, ss));
      MVC_TRY(mvc_transpose_to_bf16(q.dwq_b, 1, SB, A, A, q.dwqT, SBp, ss));
      MVC_TRY(mvc_gemm_bf16(A, H, SB, q.dwqT, SBp, hprevT, SBp, 0.f, g->att_W, H, nullptr, nullptr, 0, ss));
      MVC_TRY(sg.mark());
      forked = true;
    }
    MVC_TRY(launch_transpose_bf16(q.dG_b, 1, SB, 4 * H, 4 * H, q.dGT, SBp, permH, st));   // natural gate rows
    if (E % 4 == 0 && (E + F) % 4 == 0 && cdiv(4 * H, 128) * cdiv(E + F + H, 256) >= kNumSMs / 2) {
      // dW_ih = dG^T . [emb ; ctx] and dW_hh = dG^T . h_prev share the operand dG^T: ONE GEMM over [emb ; ctx ; h]^T whose
      // epilogue writes columns [0, E+F) to dW_ih and [E+F, E+F+H) to dW_hh (was three launches: 37 + 25 + 29 us)
      TcEpilogue ep{};
      ep.mode = TC_MODE_PLAIN;
      ep.C = g->w_ih; ep.ldc = E + F;
      ep.C2 = g->w_hh; ep.ldc2 = H; ep.split_n = E + F;
      MVC_TRY(tc_gemm(4 * H, E + F + H, SB, q.dGT, SBp, xallT, SBp, ep, 0, st));
    } else {
      MVC_TRY(mvc_gemm_bf16(4 * H, F, SB, q.dGT, SBp, xhT, SBp, 0.f, g->w_ih + E, E + F, nullptr, nullptr, 0, st));
      MVC_TRY(mvc_gemm_bf16(4 * H, H, SB, q.dGT, SBp, hprevT, SBp, 0.f, g->w_hh, H, nullptr, nullptr, 0, st));
      MVC_TRY(mvc_gemm_bf16(4 * H, E, SB, q.dGT, SBp, xallT, SBp, 0.f, g->w_ih, E + F, nullptr, nullptr, 0, st));
    }
    // bias gradients (column sums) behind the big GEMM on the caller's stream, next to the side stream's small GEMMs
    MVC_TRY(bias_grads());
  }
  if (forked) MVC_TRY(sg.join());
  return 0;
}

// ------------------------------------------------------------------ greedy ids
namespace mvc {
struct GreedyWs {
  float* gc;       // [B, 4H] sum_t alpha_t P[b,t,:] (projected-keys form)
  float* logits;   // [B, V]
  int64_t* tok;    // [2, B]
  float* c;        // [2, B, H]
  float* wq;       // [B, A]
  float* alpha;    // [B, T]
  float* h32;      // [B, H]
  size_t bytes;
};
static GreedyWs greedy_layout(const MvcDecoderDims* d, void* base, size_t dec_bytes) {
  Arena ar(base);
  ar.off = dec_bytes;
  GreedyWs g{};
  g.gc = ar.take<float>((int64_t)d->B * 4 * d->H);
  g.logits = ar.take<float>((int64_t)d->B * d->V);
  g.tok = ar.take<int64_t>(2 * (int64_t)d->B);
  g.c = ar.take<float>(2 * (int64_t)d->B * d->H);
  g.wq = ar.take<float>((int64_t)d->B * d->A);
  g.alpha = ar.take<float>((int64_t)d->B * d->T);
  g.h32 = ar.take<float>((int64_t)d->B * d->H);
  g.bytes = ar.off + 256;
  return g;
}
// the step loop only needs two xh slots: shrink the layout by pretending L = 2
static MvcDecoderDims two_slot_dims(const MvcDecoderDims* d) {
  MvcDecoderDims e = *d;
  e.L = 2;
  return e;
}
}  // namespace mvc

extern "C" size_t mvc_decoder_greedy_workspace_bytes(const MvcDecoderDims* d) {
  MvcDecoderDims e = two_slot_dims(d);
  return greedy_layout(d, nullptr, dec_layout(&e, nullptr).bytes).bytes;
}

extern "C" int mvc_decoder_greedy(const MvcDecoderDims* d, const MvcDecoderParams* p, const float* audio, int Fa,
                                  const float* visual, int Fv, int64_t* ids, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  MVC_TRY(check_dims(d));
  MVC_CHECK(p && ids && workspace, "mvc_decoder_greedy: null argument");
  MvcDecoderDims e = two_slot_dims(d);
  DecWs w = dec_layout(&e, workspace);
  GreedyWs gw = greedy_layout(d, workspace, w.bytes);
  MVC_CHECK(workspace_bytes >= gw.bytes, "mvc_decoder_greedy: workspace %zu < %zu", workspace_bytes, gw.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, F = d->F, H = d->H, V = d->V, S = d->L - 1, L = d->L;
  const bool bf = d->precision == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  const int64_t ldx = F + H;
  MVC_TRY(dec_prepare(d, p, audio, Fa, visual, Fv, w, true, st));
  // Optional (MVC_B200_GREEDY_P=1), several waves of rows (B > SMs): projected keys.  P = keys . W_c^T once per call
  // (fp16, the gate-column order of wcat); every step's attention pass then sums P rows instead of key rows and the gate
  // GEMM contracts over h only (K = 512 instead of 2688): the attention pass moves the same bytes, the GEMM a fifth of
  // them.  Measured at the C3 shape: 285.8 k vs 266.4 k captions/s (+7 %: the P GEMM costs 110 us per call up front),
  // captions identical to the fp32 oracle's at the same rate (0.729 vs 0.730) -- but the free-running `decode` keeps
  // the plain form, and on raw features the two bf16 roundings agree on 78 % of the captions only, so `predict` and
  // `decode(...).argmax` would stop telling the same story.  Off by default for that reason.
  const bool use_p = bf && dec_perm(d) && attention_stream_eligible(B, B, d->T, d->A, 4 * H) &&
                     getenv("MVC_B200_GREEDY_P") && getenv("MVC_B200_GREEDY_P")[0] == '1';
  if (use_p) {
    TcEpilogue ep{};
    ep.mode = TC_MODE_PLAIN;
    ep.Cb = (__nv_bfloat16*)w.P; ep.ldcb = 4 * H; ep.cb_f16 = 1;
    MVC_TRY(tc_gemm(B * d->T, 4 * H, F, w.feats, F, w.wcat, ldx, ep, 0, st));
  }
  MVC_CUDA(cudaMemsetAsync(w.xh, 0, es * (size_t)2 * B * ldx, st));
  MVC_CUDA(cudaMemsetAsync(gw.c, 0, sizeof(float) * (size_t)B * H, st));
  MVC_CUDA(cudaMemsetAsync(ids, 0, sizeof(int64_t) * (size_t)B * L, st));      // column 0 = argmax of zeros = 0
  fill_i64_kernel<<<(unsigned)cdiv(B, 256), 256, 0, st>>>(gw.tok, MVC_SOS, B);
  MVC_LAUNCH_CHECK();
  const StepCfg cfg = dec_cfg(d, p, w, nullptr);
  for (int s = 0; s < S; ++s) {
    StepFwd io{};
    io.rows = B;
    io.xh_src = mptr(w.xh, (int64_t)(s & 1) * B * ldx, es);
    io.xh_dst = mptr(w.xh, (int64_t)((s + 1) & 1) * B * ldx, es);
    io.wq = gw.wq; io.alpha = gw.alpha; io.act = nullptr;
    io.c_prev = gw.c + (int64_t)(s & 1) * B * H;
    io.c_out = gw.c + (int64_t)((s + 1) & 1) * B * H;
    io.tokens = gw.tok + (int64_t)(s & 1) * B;
    io.h_out32 = gw.h32;
    io.h_ld = H;
    io.first = (s == 0);
    io.wq_ready = bf;
    if (use_p) { io.pkeys = w.P; io.gc = gw.gc; }
    MVC_TRY(step_forward(cfg, io, st));
    int64_t* nxt = gw.tok + (int64_t)((s + 1) & 1) * B;
    if (bf) {
      // K-E: vocabulary projection with the row arg-max in the tcgen05 epilogue -- the [B,V] logits never reach
      // memory (arg-max of the logits == arg-max of the log-probs)
      const int ntn = tc_gemm_argmax_tiles(V);
      float* pval = gw.logits;
      int* pidx = reinterpret_cast<int*>(gw.logits + (size_t)B * ntn);
      // the same GEMM also produces wq = W.h_{s+1} for the next step's attention (auxiliary column block)
      const TcAux aux{tc_aux_row0(V), d->A, gw.wq, d->A};
      MVC_TRY(tc_gemm_argmax(B, V, H, cptr(io.xh_dst, F, es), ldx, w.outw, H, p->out_b, pval, pidx, nxt, ids + (s + 1), L,
                             TC_FLAG_PDL | TC_FLAG_B_CONST, st, &aux));
      continue;
    }
    MVC_TRY(gemm_nt(d->precision, B, V, H, (const char*)gw.h32, H, (const void*)p->out_w, H, 0.f, gw.logits, V, p->out_b, st));
    // fp32 path: normalise before the arg-max so that ids agree bit-for-bit with decode(captions=None).argmax(2)
    // (captioning.py:138-140)
    MVC_TRY(mvc_log_softmax_rows(gw.logits, B, V, nxt, st));
    scatter_col_i64_kernel<<<(unsigned)cdiv(B, 256), 256, 0, st>>>(nxt, ids, L, s + 1, B);
    MVC_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------ beam search
namespace mvc {

constexpr int kMaxBeam = 8;

// Per row of logits: log-sum-exp and the `width` best (log-prob, token), ties -> lowest token.
__global__ void beam_row_topk_kernel(const float* __restrict__ logits, int V, int width, float* __restrict__ cand_val,
                                     int* __restrict__ cand_idx) {
  __shared__ float red[32];
  __shared__ int redi[32];
  __shared__ float s_last_val;
  __shared__ int s_last_idx;
  const float* row = logits + (int64_t)blockIdx.x * V;
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, row[v]);
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) sum += expf(row[v] - mx);
  sum = block_sum(sum, red);
  const float lse = mx + logf(sum);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  float last_val = INFINITY;
  int last_idx = -1;
  for (int k = 0; k < width; ++k) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float y = row[v] - lse;
      const bool after = (y < last_val) || (y == last_val && v > last_idx);   // strictly after the previous pick
      if (after && (y > best || (y == best && v < bi))) { best = y; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { red[wid] = best; redi[wid] = bi; }
    __syncthreads();
    if (wid == 0) {
      best = lane < nw ? red[lane] : -INFINITY;
      bi = lane < nw ? redi[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) {
        s_last_val = best; s_last_idx = bi;
        cand_val[(int64_t)blockIdx.x * width + k] = best;
        cand_idx[(int64_t)blockIdx.x * width + k] = bi;
      }
    }
    __syncthreads();
    last_val = s_last_val; last_idx = s_last_idx;
  }
}

// One thread per batch item: merge nb*width candidates, keep `width`
// (features_captioning.py:166-193, 201-209).  State arrays are beam-major [beam, B].
__global__ void beam_merge_kernel(int B, int V, int nb, int width, int t, float alpha, const float* __restrict__ cand_val,
                                  const int* __restrict__ cand_idx, const float* __restrict__ cum,
                                  const uint8_t* __restrict__ done, const int* __restrict__ len,
                                  float* __restrict__ cum_out, uint8_t* __restrict__ done_out, int* __restrict__ len_out,
                                  int* __restrict__ sel_beam, int64_t* __restrict__ tok_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float sc[kMaxBeam * kMaxBeam], rk[kMaxBeam * kMaxBeam];
  int64_t flat[kMaxBeam * kMaxBeam];
  const int n = nb * width;
  for (int k = 0; k < nb; ++k) {
    const int row = k * B + b;
    const bool fin = done[row];
    const float clen = fin ? (float)len[row] : (float)(t + 1);                  // :171-176
    const float norm = alpha == 0.f ? 1.f : powf(5.f + clen, alpha) / powf(6.f, alpha);   // :177
    for (int j = 0; j < width; ++j) {
      // finished beam: EOS_mask zeroes the step scores, so every child scores `cum`; lowest tokens win the tie
      const float v = fin ? 0.f : cand_val[(int64_t)row * width + j];
      const int tok = fin ? j : cand_idx[(int64_t)row * width + j];
      const float s = v + cum[row];                                             // :166-168
      sc[k * width + j] = s;
      rk[k * width + j] = s / norm;
      flat[k * width + j] = (int64_t)k * V + tok;
    }
  }
  bool used[kMaxBeam * kMaxBeam];
  for (int i = 0; i < n; ++i) used[i] = false;
  for (int r = 0; r < width; ++r) {
    int best = -1;
    for (int i = 0; i < n; ++i) {
      if (used[i]) continue;
      if (best < 0 || rk[i] > rk[best] || (rk[i] == rk[best] && flat[i] < flat[best])) best = i;
    }
    used[best] = true;
    const int kb = (int)(flat[best] / V);                                       // :192
    const int tok = (int)(flat[best] % V);                                      // :193
    const int src = kb * B + b, dst = r * B + b;
    const bool was = done[src];
    const bool now = !was && tok == MVC_EOS;
    cum_out[dst] = sc[best];                                                    // :207 (un-normalised)
    done_out[dst] = (was || now) ? 1 : 0;
    len_out[dst] = was ? len[src] : (now ? t + 1 : 0);
    sel_beam[dst] = kb;
    tok_out[dst] = tok;
  }
}

// Reorder beam state: h (inside xh rows), c, sequences; append the chosen token.
template <typename XT>
__global__ void beam_reorder_kernel(int B, int H, int F, int width, int t, int Lb, const int* __restrict__ sel_beam,
                                    const int64_t* __restrict__ tok, const XT* __restrict__ xh_src,
                                    XT* __restrict__ xh_dst, const float* __restrict__ c_src, float* __restrict__ c_dst,
                                    const int* __restrict__ seq_src, int* __restrict__ seq_dst,
                                    const float* __restrict__ wq_src, float* __restrict__ wq_dst, int A) {
  const int dst = blockIdx.x;             // k*B + b
  const int b = dst % B;
  const int src = sel_beam[dst] * B + b;
  const int64_t ldx = F + H;
  if (wq_src)
    for (int j = threadIdx.x; j < A; j += blockDim.x) wq_dst[(int64_t)dst * A + j] = wq_src[(int64_t)src * A + j];
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    xh_dst[(int64_t)dst * ldx + F + j] = xh_src[(int64_t)src * ldx + F + j];
    c_dst[(int64_t)dst * H + j] = c_src[(int64_t)src * H + j];
  }
  for (int j = threadIdx.x; j < t; j += blockDim.x) seq_dst[(int64_t)dst * Lb + j] = seq_src[(int64_t)src * Lb + j];
  if (threadIdx.x == 0) seq_dst[(int64_t)dst * Lb + t] = (int)tok[dst];
}

__global__ void beam_emit_kernel(int B, int Lb, const int* __restrict__ seq, int64_t* __restrict__ ids) {
  const int b = blockIdx.x;
  int64_t* o = ids + (int64_t)b * (Lb + 1);
  if (threadIdx.x == 0) o[0] = MVC_SOS;                                          // :227
  for (int j = threadIdx.x; j < Lb; j += blockDim.x) o[1 + j] = seq[(int64_t)b * Lb + j];   // beam 0
}

struct BeamWs {
  float* logits;     // [W*B, V]
  float* cand_val;   // [W*B, W]
  int* cand_idx;     // [W*B, W]
  float* cum[2];     // [W*B]
  uint8_t* done[2];
  int* len[2];
  int* sel;          // [W*B]
  int64_t* tok[2];   // [W*B]
  float* c[3];       // [W*B, H]  (prev, new-unordered, reordered)
  int* seq[2];       // [W*B, Lb]
  float* wq;         // [W*B, A]
  float* wq2;        // [W*B, A] W.h of the new (not yet reordered) states, from the vocabulary GEMM
  float* alpha;      // [W*B, T]
  float* h32;        // [W*B, H]
  float* pre;        // [W*B, 4H]
  void* xh[3];       // [W*B, F+H]
  size_t bytes;
};
static BeamWs beam_layout(const MvcDecoderDims* d, int width, void* base, size_t dec_bytes) {
  Arena ar(base);
  ar.off = dec_bytes;
  const int64_t R = (int64_t)width * d->B, Lb = d->L + 1;
  const size_t es = d->precision == MVC_BF16 ? 2 : 4;
  BeamWs w{};
  w.logits = ar.take<float>(R * d->V);
  w.cand_val = ar.take<float>(R * width);
  w.cand_idx = ar.take<int>(R * width);
  for (int i = 0; i < 2; ++i) {
    w.cum[i] = ar.take<float>(R);
    w.done[i] = ar.take<uint8_t>(R);
    w.len[i] = ar.take<int>(R);
    w.tok[i] = ar.take<int64_t>(R);
    w.seq[i] = ar.take<int>(R * Lb);
  }
  w.sel = ar.take<int>(R);
  for (int i = 0; i < 3; ++i) {
    w.c[i] = ar.take<float>(R * d->H);
    w.xh[i] = ar.take<char>(R * (d->F + d->H) * es);
  }
  w.wq = ar.take<float>(R * d->A);
  w.wq2 = ar.take<float>(R * d->A);
  w.alpha = ar.take<float>(R * d->T);
  w.h32 = ar.take<float>(R * d->H);
  w.pre = ar.take<float>(R * 4 * d->H);
  w.bytes = ar.off + 256;
  return w;
}
static MvcDecoderDims beam_dec_dims(const MvcDecoderDims* d) {
  MvcDecoderDims e = *d;
  e.L = 2;          // xh slots come from BeamWs instead
  return e;
}
}  // namespace mvc

extern "C" size_t mvc_decoder_beam_workspace_bytes(const MvcDecoderDims* d, int width) {
  MvcDecoderDims e = beam_dec_dims(d);
  return beam_layout(d, width, nullptr, dec_layout(&e, nullptr).bytes).bytes;
}

extern "C" int mvc_decoder_beam(const MvcDecoderDims* d, const MvcDecoderParams* p, const float* audio, int Fa,
                                const float* visual, int Fv, int width, float alpha, int64_t* ids, void* workspace,
                                size_t workspace_bytes, void* stream) {
  MVC_TRY(check_dims(d));
  MVC_CHECK(p && ids && workspace, "mvc_decoder_beam: null argument");
  MVC_CHECK(width >= 1 && width <= kMaxBeam, "mvc_decoder_beam: beam_width %d not in [1,%d]", width, kMaxBeam);
  MVC_CHECK(width <= d->V, "mvc_decoder_beam: beam_width > vocab");
  MvcDecoderDims e = beam_dec_dims(d);
  DecWs w = dec_layout(&e, workspace);
  BeamWs bw = beam_layout(d, width, workspace, w.bytes);
  MVC_CHECK(workspace_bytes >= bw.bytes, "mvc_decoder_beam: workspace %zu < %zu", workspace_bytes, bw.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, F = d->F, H = d->H, V = d->V;
  const int max_len = d->L, Lb = max_len + 1;       // max_caption_len + 1 steps (:149)
  const bool bf = d->precision == MVC_BF16;
  const size_t es = bf ? 2 : 4;
  const int64_t ldx = F + H;
  const int64_t R = (int64_t)width * B;
  MVC_TRY(dec_prepare(d, p, audio, Fa, visual, Fv, w, true, st));
  MVC_CUDA(cudaMemsetAsync(bw.xh[0], 0, es * (size_t)R * ldx, st));
  MVC_CUDA(cudaMemsetAsync(bw.c[0], 0, sizeof(float) * (size_t)R * H, st));
  MVC_CUDA(cudaMemsetAsync(bw.cum[0], 0, sizeof(float) * (size_t)R, st));       // log(1) = 0 (:144-145)
  MVC_CUDA(cudaMemsetAsync(bw.done[0], 0, (size_t)R, st));
  MVC_CUDA(cudaMemsetAsync(bw.len[0], 0, sizeof(int) * (size_t)R, st));
  fill_i64_kernel<<<(unsigned)cdiv(R, 256), 256, 0, st>>>(bw.tok[0], MVC_SOS, R);
  MVC_LAUNCH_CHECK();
  int cur = 0;      // index of the current state buffers
  void* xh_cur = bw.xh[0];
  void* xh_new = bw.xh[1];
  void* xh_nxt = bw.xh[2];
  float* c_cur = bw.c[0];
  float* c_new = bw.c[1];
  float* c_nxt = bw.c[2];
  const StepCfg cfg = dec_cfg(d, p, w, bw.pre);
  for (int t = 0; t < Lb; ++t) {
    const int nb = (t == 0) ? 1 : width;
    const int rows = nb * B;
    StepFwd io{};
    io.rows = rows;
    io.xh_src = xh_cur; io.xh_dst = xh_new;
    io.wq = bw.wq; io.alpha = bw.alpha; io.act = nullptr;
    io.c_prev = c_cur; io.c_out = c_new;
    io.tokens = bw.tok[cur];
    io.h_out32 = bw.h32;
    io.h_ld = H;
    io.first = (t == 0);
    const bool fused_topk = bf && V >= 64;
    io.wq_ready = fused_topk;
    MVC_TRY(step_forward(cfg, io, st));
    if (fused_topk) {
      // K-E: vocabulary projection with top-k + log-sum-exp in the tcgen05 epilogue (no [rows,V] logits); the same
      // GEMM produces W.h of the new states (auxiliary column block), reordered with them below
      const TcAux aux{tc_aux_row0(V), d->A, bw.wq2, d->A};
      MVC_TRY(tc_gemm_topk(rows, V, H, cptr(xh_new, F, es), ldx, w.outw, H, p->out_b, bw.logits, width, bw.cand_val,
                           bw.cand_idx, TC_FLAG_PDL | TC_FLAG_B_CONST, st, &aux));
    } else {
      MVC_TRY(gemm_nt(d->precision, rows, V, H, bf ? cptr(xh_new, F, es) : (const char*)bw.h32, bf ? ldx : H,
                      bf ? w.outw : (const void*)p->out_w, H, 0.f, bw.logits, V, p->out_b, st));
      beam_row_topk_kernel<<<rows, 256, 0, st>>>(bw.logits, V, width, bw.cand_val, bw.cand_idx);
      MVC_LAUNCH_CHECK();
    }
    const int nx = cur ^ 1;
    beam_merge_kernel<<<(unsigned)cdiv(B, 128), 128, 0, st>>>(B, V, nb, width, t, alpha, bw.cand_val, bw.cand_idx,
                                                             bw.cum[cur], bw.done[cur], bw.len[cur], bw.cum[nx],
                                                             bw.done[nx], bw.len[nx], bw.sel, bw.tok[nx]);
    MVC_LAUNCH_CHECK();
    if (bf)
      beam_reorder_kernel<__nv_bfloat16><<<(unsigned)R, 128, 0, st>>>(B, H, F, width, t, Lb, bw.sel, bw.tok[nx],
                                                                     (const __nv_bfloat16*)xh_new, (__nv_bfloat16*)xh_nxt,
                                                                     c_new, c_nxt, bw.seq[cur], bw.seq[nx],
                                                                     fused_topk ? bw.wq2 : nullptr, bw.wq, d->A);
    else
      beam_reorder_kernel<float><<<(unsigned)R, 128, 0, st>>>(B, H, F, width, t, Lb, bw.sel, bw.tok[nx],
                                                             (const float*)xh_new, (float*)xh_nxt, c_new, c_nxt,
                                                             bw.seq[cur], bw.seq[nx], nullptr, nullptr, 0);
    MVC_LAUNCH_CHECK();
    // rotate: reordered state becomes current
    void* tx = xh_cur; xh_cur = xh_nxt; xh_nxt = tx;
    float* tc = c_cur; c_cur = c_nxt; c_nxt = tc;
    cur = nx;
  }
  beam_emit_kernel<<<B, 64, 0, st>>>(B, Lb, bw.seq[cur], ids);
  MVC_LAUNCH_CHECK();
  return 0;
}
