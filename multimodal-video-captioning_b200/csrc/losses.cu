// losses.cu -- the loss bundle of src/losses.py as fused forward+gradient kernels.
//
//  * caption loss = F.nll_loss(ignore_index=PAD) (losses.py:112) + EntropyLoss
//    (losses.py:12-17).  EntropyLoss applies softmax/log_softmax over dim=1 of
//    the [L-1,B,V] log-prob tensor, i.e. over the BATCH axis; that quirk is
//    part of the loss value and of its gradient, so it is kept: one thread owns
//    one (step, vocab) column and walks the B batch entries (coalesced over V).
//  * GlobalReconstructionLoss (losses.py:20-36), LocalReconstructionLoss (:39-40).
#include "common.cuh"

namespace mvc {

// result[0] = mean NLL over non-PAD targets, result[2] = count; zeroes result[1] and the entropy accumulator.
__global__ void nll_kernel(const float* __restrict__ logp, const int64_t* __restrict__ cap, int L, int B, int V,
                           float* __restrict__ result, double* __restrict__ ent_acc) {
  __shared__ float red[32];
  if (threadIdx.x == 0) *ent_acc = 0.0;
  const int64_t n = (int64_t)(L - 1) * B;
  float s = 0.f, c = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t tok = cap[B + i];             // captions[1:]
    if (tok != MVC_PAD) {
      s -= logp[((int64_t)B + i) * V + tok];    // logp[1:]
      c += 1.f;
    }
  }
  s = block_sum(s, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) {
    result[0] = s / c;
    result[1] = 0.f;
    result[2] = c;
  }
}

// One thread per (step s, vocab v) column of x = logp[1:].
__global__ void entropy_kernel(const float* __restrict__ logp, const int64_t* __restrict__ cap, int L, int B, int V,
                               float* __restrict__ result, double* __restrict__ ent_acc, float* __restrict__ dlogp,
                               float ce_scale, float ent_scale) {
  __shared__ float red[32];
  const int s = blockIdx.y + 1;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = v < V;
  const float* col = logp + (int64_t)s * B * V + v;
  const int64_t* caps = cap + (int64_t)s * B;
  float e_sum = 0.f;
  if (live) {
    float mx = -INFINITY;
    for (int b = 0; b < B; ++b) mx = fmaxf(mx, col[(int64_t)b * V]);
    float z = 0.f;
    for (int b = 0; b < B; ++b) z += expf(col[(int64_t)b * V] - mx);
    const float lse = mx + logf(z);
    float q = 0.f;                               // sum_b m_b p_b (log p_b + 1)
    for (int b = 0; b < B; ++b) {
      if (caps[b] == MVC_PAD) continue;
      const float lp = col[(int64_t)b * V] - lse;
      const float p = expf(lp);
      e_sum += p * lp;
      q += p * (lp + 1.f);
    }
    if (dlogp) {
      const float cnt = result[2];
      const float es = -ent_scale / (float)B;
      for (int b = 0; b < B; ++b) {
        const float lp = col[(int64_t)b * V] - lse;
        const float p = expf(lp);
        const int64_t tok = caps[b];
        const float m = tok != MVC_PAD ? 1.f : 0.f;
        float g = es * p * (m * (lp + 1.f) - q);
        if (tok != MVC_PAD && tok == v) g -= ce_scale / cnt;
        dlogp[((int64_t)s * B + b) * V + v] = g;
        if (s == 1) dlogp[(int64_t)b * V + v] = 0.f;                     // row 0 gets no gradient
      }
    }
  }
  e_sum = block_sum(e_sum, red);
  if (threadIdx.x == 0) atomicAdd(ent_acc, (double)e_sum);
}

// Same quantities for B <= ENT_RG * ENT_RPT, ONE pass over the log-probs: block = 32 columns x ENT_RG row groups, thread
// (x, y) keeps rows y, y + ENT_RG, ... of its column in registers, the column-wise max / sum / weighted sums are combined
// across the row groups through shared memory.  (The 4-pass kernel above ran 16 warps per SM with 512 dependent-latency
// loads per thread: 63 us for 2 x 37.7 MB; this one moves the same bytes in one read + one write.)
// (8 row groups x 16 rows per thread; 16 groups x 8 rows -- twice the occupancy, twice the shared-memory reads per
// element -- measured 3 us slower: the kernel is issue bound, not latency bound)
constexpr int ENT_RPT = 16;
constexpr int ENT_RG = 8;
__global__ void __launch_bounds__(32 * ENT_RG)
entropy_tile_kernel(const float* __restrict__ logp, const int64_t* __restrict__ cap, int L, int B, int V,
                    float* __restrict__ result, double* __restrict__ ent_acc, float* __restrict__ dlogp, float ce_scale,
                    float ent_scale) {
  __shared__ float red[ENT_RG][33];
  __shared__ float red2[ENT_RG][33];
  __shared__ float bsum[32];
  __shared__ int stok[ENT_RG * ENT_RPT];          // target token of every row of this step (PAD beyond B)
  const int s = blockIdx.y + 1;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int v = blockIdx.x * 32 + tx;
  const bool live = v < V;
  for (int i = ty * 32 + tx; i < ENT_RG * ENT_RPT; i += 32 * ENT_RG) stok[i] = i < B ? (int)cap[(int64_t)s * B + i] : (int)MVC_PAD;
  // rows ty, ty + ENT_RG, ...: one base pointer, compile-time multiples of the row-group pitch
  const size_t pitch = (size_t)ENT_RG * V;
  const float* xp = logp + ((size_t)s * B + ty) * V + v;
  float x[ENT_RPT], pr[ENT_RPT];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < ENT_RPT; ++i) {
    x[i] = (live && ty + ENT_RG * i < B) ? xp[i * pitch] : -INFINITY;
    mx = fmaxf(mx, x[i]);
  }
  red[ty][tx] = mx;
  __syncthreads();                                // (also publishes stok)
#pragma unroll
  for (int g = 0; g < ENT_RG; ++g) mx = fmaxf(mx, red[g][tx]);
  // one exponential per element: e = exp(x - max); p = e / z; log p = x - (max + log z)
  float z = 0.f;
#pragma unroll
  for (int i = 0; i < ENT_RPT; ++i) {
    pr[i] = live ? __expf(x[i] - mx) : 0.f;       // rows beyond B hold -inf: exp -> 0
    z += pr[i];
  }
  red2[ty][tx] = z;
  __syncthreads();
  z = 0.f;
#pragma unroll
  for (int g = 0; g < ENT_RG; ++g) z += red2[g][tx];
  const float lse = live ? mx + logf(z) : 0.f;
  const float rz = live ? 1.f / z : 0.f;
  float e_sum = 0.f, q = 0.f;                     // q = sum_b m_b p_b (log p_b + 1)
  unsigned keep = 0;                              // bit i: row i of this thread is a non-PAD target row
#pragma unroll
  for (int i = 0; i < ENT_RPT; ++i) {
    const bool in = live && (ty + ENT_RG * i < B);
    const float lp = in ? x[i] - lse : 0.f;
    const float p = pr[i] * rz;
    x[i] = lp;
    pr[i] = p;
    if (in && stok[ty + ENT_RG * i] != MVC_PAD) {
      keep |= 1u << i;
      e_sum += p * lp;
      q += p * (lp + 1.f);
    }
  }
  __syncthreads();                                // red is read above by everyone before it is reused
  red[ty][tx] = q;
  __syncthreads();
  q = 0.f;
#pragma unroll
  for (int g = 0; g < ENT_RG; ++g) q += red[g][tx];
  if (dlogp && live) {
    const float cnt = result[2];
    const float es = -ent_scale / (float)B;
    const float ce = ce_scale / cnt;
    float* gp = dlogp + ((size_t)s * B + ty) * V + v;
    float* g0 = dlogp + (size_t)ty * V + v;       // row 0 of the gradient (no gradient: sentence[0] is a constant)
#pragma unroll
    for (int i = 0; i < ENT_RPT; ++i) {
      if (ty + ENT_RG * i < B) {
        const bool k = (keep >> i) & 1u;
        float g = es * pr[i] * ((k ? x[i] + 1.f : 0.f) - q);
        if (k && stok[ty + ENT_RG * i] == v) g -= ce;
        gp[i * pitch] = g;
        if (s == 1) g0[i * pitch] = 0.f;
      }
    }
  }
  // block sum of e_sum -> one double atomic per block
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e_sum += __shfl_xor_sync(0xffffffffu, e_sum, o);
  if (tx == 0) bsum[ty] = e_sum;
  __syncthreads();
  if (tx == 0 && ty == 0) {
    float t = 0.f;
    for (int g = 0; g < ENT_RG; ++g) t += bsum[g];
    atomicAdd(ent_acc, (double)t);
  }
}

__global__ void entropy_finish_kernel(const double* __restrict__ ent_acc, int B, float* __restrict__ result) {
  result[1] = (float)(-(*ent_acc) / (double)B);
}

// thread per (b, f).  acc[0] += sum (xm - xr)^2 ; optional gradient.
__global__ void global_recon_loss_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ xrec,
                                         int64_t r_ld, int B, int T, int L, int F, const int64_t* __restrict__ cap,
                                         double* __restrict__ acc, float* __restrict__ dxrec, int64_t d_ld, float scale) {
  __shared__ float red[32];
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float sq = 0.f;
  if (i < (int64_t)B * F) {
    const int b = (int)(i / F), f = (int)(i - (int64_t)b * F);
    float xm = 0.f;
    for (int t = 0; t < T; ++t) xm += x[((int64_t)b * T + t) * x_ld + f];
    xm /= (float)T;                                                    // losses.py:25 (padding frames included)
    float xr = 0.f, n = 0.f;
    for (int l = 0; l < L; ++l)
      if (cap[(int64_t)l * B + b] != MVC_PAD) {                        // keep_mask = captions != PAD (losses.py:104)
        xr += xrec[((int64_t)b * L + l) * r_ld + f];
        n += 1.f;
      }
    xr /= n;
    const float d = xm - xr;
    sq = d * d;
    if (dxrec) {
      const float g = scale * 2.f * (xr - xm) / ((float)B * (float)F) / n;
      for (int l = 0; l < L; ++l)
        dxrec[((int64_t)b * L + l) * d_ld + f] = cap[(int64_t)l * B + b] != MVC_PAD ? g : 0.f;
    }
  }
  sq = block_sum(sq, red);
  if (threadIdx.x == 0) atomicAdd(acc, (double)sq);
}

__global__ void local_recon_loss_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ xrec,
                                        int64_t r_ld, int64_t rows, int F, double* __restrict__ acc,
                                        float* __restrict__ dxrec, int64_t d_ld, float gscale) {
  __shared__ float red[32];
  const int64_t n = rows * F;
  float sq = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F;
    const int f = (int)(i - r * F);
    const float d = xrec[r * r_ld + f] - x[r * x_ld + f];
    sq = fmaf(d, d, sq);
    if (dxrec) dxrec[r * d_ld + f] = gscale * d;
  }
  sq = block_sum(sq, red);
  if (threadIdx.x == 0) atomicAdd(acc, (double)sq);
}

__global__ void mse_finish_kernel(const double* __restrict__ acc, double n, float* __restrict__ result) {
  result[0] = (float)(*acc / n);
}

// loss = ce + reg*entropy; loss += a_lambda*a_rec; loss += v_lambda*v_rec  (losses.py:122-124, same fp32 order)
__global__ void loss_combine_kernel(float* __restrict__ r, float reg, float a_lambda, float v_lambda, int have_a,
                                    int have_v) {
  if (!have_a) r[3] = 0.f;
  if (!have_v) r[4] = 0.f;
  float loss = r[0] + reg * r[1];
  loss += a_lambda * r[3];
  loss += v_lambda * r[4];
  r[5] = loss;
}

// x *= *g (g on the device: the upstream gradient of the loss, 1.0 in the reference's train loop -> nothing to do)
__global__ void scale_by_scalar_kernel(float4* __restrict__ x, int64_t n4, float* __restrict__ tail, int ntail,
                                       const float* __restrict__ g) {
  const float s = *g;
  if (s == 1.f) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    x[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] *= s;
}

}  // namespace mvc

using namespace mvc;

extern "C" size_t mvc_caption_loss_workspace_bytes(int, int, int) { return 256; }

extern "C" int mvc_caption_loss(const float* logp, const int64_t* captions, int L, int B, int V, float* result,
                                float* dlogp, float ce_scale, float ent_scale, void* workspace, void* stream) {
  MVC_CHECK(logp && captions && result && workspace, "mvc_caption_loss: null argument");
  MVC_CHECK(L >= 2 && B >= 1 && V >= 1, "mvc_caption_loss: bad dims L=%d B=%d V=%d", L, B, V);
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = static_cast<double*>(workspace);
  // (no memset nodes on this chain: the NLL kernel clears the accumulator, the entropy kernels write the zero
  // gradient of row 0 -- sentence[0] is a constant -- next to row 1)
  nll_kernel<<<1, 1024, 0, st>>>(logp, captions, L, B, V, result, acc);
  MVC_LAUNCH_CHECK();
  dim3 grid((unsigned)cdiv(V, 128), (unsigned)(L - 1));
  ProfScope prof(PK_LOSS, L, B, V, st);
  if (B <= ENT_RG * ENT_RPT) {
    const dim3 tgrid((unsigned)cdiv(V, 32), (unsigned)(L - 1));
    entropy_tile_kernel<<<tgrid, dim3(32, ENT_RG), 0, st>>>(logp, captions, L, B, V, result, acc, dlogp, ce_scale, ent_scale);
  } else {
    entropy_kernel<<<grid, 128, 0, st>>>(logp, captions, L, B, V, result, acc, dlogp, ce_scale, ent_scale);
  }
  MVC_LAUNCH_CHECK();
  entropy_finish_kernel<<<1, 1, 0, st>>>(acc, B, result);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t mvc_global_recon_loss_workspace_bytes(int, int) { return 256; }

extern "C" int mvc_global_recon_loss(const float* x, int64_t x_ld, const float* xrec, int64_t r_ld, int B, int T, int L,
                                     int F, const int64_t* captions, float* result, float* dxrec, int64_t d_ld,
                                     float scale, void* workspace, void* stream) {
  MVC_CHECK(x && xrec && captions && result && workspace, "mvc_global_recon_loss: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = static_cast<double*>(workspace);
  MVC_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  const int64_t n = (int64_t)B * F;
  global_recon_loss_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(x, x_ld, xrec, r_ld, B, T, L, F, captions, acc,
                                                                   dxrec, d_ld, scale);
  MVC_LAUNCH_CHECK();
  mse_finish_kernel<<<1, 1, 0, st>>>(acc, (double)n, result);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_local_recon_loss(const float* x, int64_t x_ld, const float* xrec, int64_t r_ld, int64_t rows, int F,
                                    float* result, float* dxrec, int64_t d_ld, float scale, void* workspace,
                                    void* stream) {
  MVC_CHECK(x && xrec && result && workspace, "mvc_local_recon_loss: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = static_cast<double*>(workspace);
  MVC_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  const int64_t n = rows * F;
  int64_t g = cdiv(n, 256);
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  local_recon_loss_kernel<<<(unsigned)g, 256, 0, st>>>(x, x_ld, xrec, r_ld, rows, F, acc, dxrec, d_ld,
                                                       scale * 2.f / (float)n);
  MVC_LAUNCH_CHECK();
  mse_finish_kernel<<<1, 1, 0, st>>>(acc, (double)n, result);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_loss_combine(float* result, float reg_lambda, float a_lambda, float v_lambda, int have_a, int have_v,
                                void* stream) {
  MVC_CHECK(result, "mvc_loss_combine: null argument");
  loss_combine_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(result, reg_lambda, a_lambda, v_lambda, have_a, have_v);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_scale_by_scalar(float* x, int64_t n, const float* g_dev, void* stream) {
  if (n == 0) return 0;
  MVC_CHECK(x && g_dev, "mvc_scale_by_scalar: null argument");
  MVC_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0, "mvc_scale_by_scalar: x must be 16-byte aligned");
  const int64_t n4 = n / 4;
  int64_t g = cdiv(n4 > 0 ? n4 : 1, 256);
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  scale_by_scalar_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(x), n4, x + n4 * 4,
                                                                        (int)(n - n4 * 4), g_dev);
  MVC_LAUNCH_CHECK();
  return 0;
}
