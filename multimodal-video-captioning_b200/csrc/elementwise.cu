// elementwise.cu -- HBM-bound glue kernels of the SA-LSTM path: feature
// concat/cast, LSTM cell update and its backward, row-wise log-softmax /
// argmax, embedding gather / scatter-add, column sums, Adam.
#include <stdarg.h>

#include <mutex>
#include <utility>
#include <vector>

#include "tc_gemm.cuh"

namespace mvc {

static thread_local char g_err[512] = "";
long long g_launches = 0;

int g_prof_kid = 0, g_prof_m = -1, g_prof_n = -1, g_prof_k = -1;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;
static size_t g_prof_used = 0;
void prof_begin(cudaStream_t st) {
  if (g_prof_used == g_prof_pool.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    g_prof_pool.emplace_back(a, b);
  }
  cudaEventRecord(g_prof_pool[g_prof_used].first, st);
}
void prof_end(cudaStream_t st) {
  if (g_prof_used < g_prof_pool.size()) cudaEventRecord(g_prof_pool[g_prof_used++].second, st);
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------- concat / cast
template <typename OutT>
__global__ void concat_cast_kernel(const float* __restrict__ a, int Fa, const float* __restrict__ v, int Fv,
                                   int64_t rows, OutT* __restrict__ dst) {
  const int F = Fa + Fv;
  const int64_t total = rows * F;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F;
    const int f = (int)(i - r * F);
    const float x = (f < Fa) ? a[r * Fa + f] : v[r * Fv + (f - Fa)];
    if constexpr (sizeof(OutT) == 2) dst[i] = __float2bfloat16(x);
    else dst[i] = x;
  }
}

// 8 outputs per thread: two float4 loads from the right source, one 16-byte bf16 store (Fa, Fv multiples of 8)
__global__ void concat_cast_bf16x8_kernel(const float* __restrict__ a, int Fa, const float* __restrict__ v, int Fv,
                                          int64_t rows, __nv_bfloat16* __restrict__ dst) {
  const int F8 = (Fa + Fv) / 8, Fa8 = Fa / 8;
  const int64_t total = rows * F8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F8;
    const int g = (int)(i - r * F8);
    const float4* src = reinterpret_cast<const float4*>(g < Fa8 ? a + r * Fa + g * 8 : v + r * Fv + (g - Fa8) * 8);
    const float4 x0 = __ldcs(src), x1 = __ldcs(src + 1);
    __nv_bfloat162 q0 = __floats2bfloat162_rn(x0.x, x0.y), q1 = __floats2bfloat162_rn(x0.z, x0.w);
    __nv_bfloat162 q2 = __floats2bfloat162_rn(x1.x, x1.y), q3 = __floats2bfloat162_rn(x1.z, x1.w);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&q0); o.y = *reinterpret_cast<uint32_t*>(&q1);
    o.z = *reinterpret_cast<uint32_t*>(&q2); o.w = *reinterpret_cast<uint32_t*>(&q3);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
}

// bf16 feature shards: both inputs already bf16 -> 16-byte copies into the concatenated layout
__global__ void concat_bf16x8_kernel(const __nv_bfloat16* __restrict__ a, int Fa, const __nv_bfloat16* __restrict__ v,
                                     int Fv, int64_t rows, __nv_bfloat16* __restrict__ dst) {
  const int F8 = (Fa + Fv) / 8, Fa8 = Fa / 8;
  const int64_t total = rows * F8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F8;
    const int g = (int)(i - r * F8);
    const uint4* src = reinterpret_cast<const uint4*>(g < Fa8 ? a + r * Fa + g * 8 : v + r * Fv + (g - Fa8) * 8);
    reinterpret_cast<uint4*>(dst)[i] = __ldcs(src);
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

template <typename ST>
__global__ void transpose_bf16_kernel(const ST* __restrict__ src, int64_t R, int64_t C, int64_t lds,
                                      __nv_bfloat16* __restrict__ dst, int64_t ldd, int permH) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = __float2bfloat16((r < R && c < C) ? ld_as_float(src + r * lds + c) : 0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) dst[(permH ? (int64_t)gate_unperm(permH, (int)c) : c) * ldd + r] = tile[threadIdx.x][i];
  }
}

// bf16 [R,C] -> bf16 [C,R], 64x64 tiles, 16-byte global accesses on both sides (C, lds, ldd multiples of 8).
__global__ void __launch_bounds__(256)
transpose_bf16_tile64_kernel(const __nv_bfloat16* __restrict__ src, int64_t R, int64_t C, int64_t lds,
                             __nv_bfloat16* __restrict__ dst, int64_t ldd, int permH) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int64_t c0 = (int64_t)blockIdx.x * 64, r0 = (int64_t)blockIdx.y * 64;
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tid + i * 256;
    const int r = idx >> 3, cv = idx & 7;
    uint4 x = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + r < R && c0 + cv * 8 < C) x = *reinterpret_cast<const uint4*>(src + (r0 + r) * lds + c0 + cv * 8);
    uint32_t* t32 = reinterpret_cast<uint32_t*>(&tile[r][cv * 8]);
    t32[0] = x.x; t32[1] = x.y; t32[2] = x.z; t32[3] = x.w;
  }
  __syncthreads();
  const int cl = tid >> 2, part = tid & 3;       // 4 lanes cover 64 consecutive r of one output row
  const int64_t c = c0 + cl;
  if (c >= C) return;
  __nv_bfloat16* drow = dst + (permH ? (int64_t)gate_unperm(permH, (int)c) : c) * ldd + r0 + part * 16;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int rl = part * 16 + h * 8;
    if (r0 + rl + 7 < R) {
      __nv_bfloat16 e[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) e[k] = tile[rl + k][cl];
      *reinterpret_cast<uint4*>(drow + h * 8) = *reinterpret_cast<const uint4*>(e);
    } else {
      for (int k = 0; k < 8; ++k)
        if (r0 + rl + k < R) drow[h * 8 + k] = tile[rl + k][cl];
    }
  }
}

// ------------------------------------------------------------- LSTM cell
// gates pre-activation = pre + gx + emb_table[token] + bias, PyTorch order i,f,g,o.
__global__ void lstm_cell_fwd_kernel(int B, int H, const float* __restrict__ pre, const float* __restrict__ gx,
                                     int64_t gx_ld, const float* __restrict__ emb_table,
                                     const int64_t* __restrict__ tokens, const float* __restrict__ bias,
                                     const float* __restrict__ c_prev, float* __restrict__ act,
                                     float* __restrict__ c_out, float* __restrict__ h_out, int64_t h_ld,
                                     float* __restrict__ h_out2, int64_t h2_ld, __nv_bfloat16* __restrict__ h_bf16,
                                     int64_t hb_ld) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H) return;
  const int b = (int)(i / H), j = (int)(i - (int64_t)b * H);
  float g[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col = q * H + j;
    float v = pre[(int64_t)b * 4 * H + col];
    if (gx) v += gx[b * gx_ld + col];
    if (emb_table) v += emb_table[tokens[b] * (int64_t)(4 * H) + col];
    if (bias) v += bias[col];
    g[q] = v;
  }
  const float ig = sigmoid_f(g[0]), fg = sigmoid_f(g[1]), gg = tanhf(g[2]), og = sigmoid_f(g[3]);
  const float cp = c_prev ? c_prev[i] : 0.f;
  const float c = fg * cp + ig * gg;
  const float h = og * tanhf(c);
  if (act) {
    float* a = act + (int64_t)b * 4 * H;
    a[j] = ig; a[H + j] = fg; a[2 * H + j] = gg; a[3 * H + j] = og;
  }
  c_out[i] = c;
  if (h_out) h_out[b * h_ld + j] = h;
  if (h_out2) h_out2[b * h2_ld + j] = h;
  if (h_bf16) h_bf16[b * hb_ld + j] = __float2bfloat16(h);
}

__global__ void lstm_cell_bwd_kernel(int B, int H, const float* __restrict__ act, const float* __restrict__ c_prev,
                                     const float* __restrict__ c_new, const float* __restrict__ dh_a, int64_t dha_ld,
                                     const float* __restrict__ dh_b, int64_t dhb_ld, float* __restrict__ dc,
                                     float* __restrict__ dgates, __nv_bfloat16* __restrict__ dg_bf16, int perm) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H) return;
  const int b = (int)(i / H), j = (int)(i - (int64_t)b * H);
  const float* a = act + (int64_t)b * 4 * H;
  const int ci = gate_col(perm, H, 0, j), cf = gate_col(perm, H, 1, j), cg = gate_col(perm, H, 2, j),
            co = gate_col(perm, H, 3, j);
  const float ig = a[ci], fg = a[cf], gg = a[cg], og = a[co];
  float dh = 0.f;
  if (dh_a) dh += dh_a[b * dha_ld + j];
  if (dh_b) dh += dh_b[b * dhb_ld + j];
  const float tc = tanhf(c_new[i]);
  const float dct = dc[i] + dh * og * (1.f - tc * tc);
  const float cp = c_prev ? c_prev[i] : 0.f;
  const float d_i = dct * gg * ig * (1.f - ig);
  const float d_f = dct * cp * fg * (1.f - fg);
  const float d_g = dct * ig * (1.f - gg * gg);
  const float d_o = dh * tc * og * (1.f - og);
  dc[i] = dct * fg;
  float* d = dgates + (int64_t)b * 4 * H;
  d[ci] = d_i; d[cf] = d_f; d[cg] = d_g; d[co] = d_o;
  if (dg_bf16) {
    __nv_bfloat16* q = dg_bf16 + (int64_t)b * 4 * H;
    q[ci] = __float2bfloat16(d_i); q[cf] = __float2bfloat16(d_f);
    q[cg] = __float2bfloat16(d_g); q[co] = __float2bfloat16(d_o);
  }
}

// ------------------------------------------------------------- log-softmax / argmax over rows
__global__ void log_softmax_rows_kernel(float* __restrict__ x, int V, int64_t* __restrict__ argmax) {
  __shared__ float red[32];
  __shared__ int redi[32];
  float* row = x + (int64_t)blockIdx.x * V;
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, row[v]);
  mx = block_max(mx, red);
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(row[v] - mx);
  s = block_sum(s, red);
  const float lse = mx + logf(s);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float y = row[v] - lse;
    row[v] = y;
    if (y > best) { best = y; bi = v; }   // ascending v per thread: first max kept
  }
  if (argmax) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
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
      if (lane == 0) argmax[blockIdx.x] = (bi == 0x7fffffff) ? 0 : bi;
    }
  }
}

__global__ void argmax_rows_kernel(const float* __restrict__ x, const float* __restrict__ y, int V,
                                   int64_t* __restrict__ out) {
  __shared__ float red[32];
  __shared__ int redi[32];
  const float* rx = x + (int64_t)blockIdx.x * V;
  const float* ry = y ? y + (int64_t)blockIdx.x * V : nullptr;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float t = ry ? rx[v] + ry[v] : rx[v];
    if (t > best) { best = t; bi = v; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
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
    if (lane == 0) out[blockIdx.x] = (bi == 0x7fffffff) ? 0 : bi;
  }
}

// REG = values of the row each thread keeps in registers between the row sum and the update (0: re-read the row)
template <int REG>
__global__ void __launch_bounds__(256)
log_softmax_bwd_kernel(const float* __restrict__ logp, const float* __restrict__ dlogp, int V,
                       float* __restrict__ dlogits, __nv_bfloat16* __restrict__ dl_bf16, int64_t ldb) {
  __shared__ float red[32];
  const int64_t off = (int64_t)blockIdx.x * V;
  float s = 0.f;
  if constexpr (REG > 0) {
    float g[REG], lp[REG];
#pragma unroll
    for (int i = 0; i < REG; ++i) {
      const int v = threadIdx.x + i * 256;
      g[i] = v < V ? dlogp[off + v] : 0.f;
      lp[i] = v < V ? logp[off + v] : 0.f;
      s += g[i];
    }
    s = block_sum(s, red);
#pragma unroll
    for (int i = 0; i < REG; ++i) {
      const int v = threadIdx.x + i * 256;
      if (v < V) {
        const float d = g[i] - __expf(lp[i]) * s;
        if (dlogits) dlogits[off + v] = d;
        if (dl_bf16) dl_bf16[(int64_t)blockIdx.x * ldb + v] = __float2bfloat16(d);
      }
    }
    // the pad columns [V, ldb) of the 16-bit copy (row pitch of the TMA operand) are zero-filled here, not by a memset
    if (dl_bf16 && (int)threadIdx.x < (int)(ldb - V)) dl_bf16[(int64_t)blockIdx.x * ldb + V + threadIdx.x] = __float2bfloat16(0.f);
  } else {
    for (int v = threadIdx.x; v < V; v += blockDim.x) s += dlogp[off + v];
    s = block_sum(s, red);
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float d = dlogp[off + v] - expf(logp[off + v]) * s;
      if (dlogits) dlogits[off + v] = d;
      if (dl_bf16) dl_bf16[(int64_t)blockIdx.x * ldb + v] = __float2bfloat16(d);
    }
    if (dl_bf16 && (int)threadIdx.x < (int)(ldb - V)) dl_bf16[(int64_t)blockIdx.x * ldb + V + threadIdx.x] = __float2bfloat16(0.f);
  }
}

// ------------------------------------------------------------- embedding
template <typename OutT>
__global__ void embedding_gather_kernel(const float* __restrict__ table, int E, const int64_t* __restrict__ idx,
                                        OutT* __restrict__ out, int64_t out_ld, int width) {
  const int64_t r = blockIdx.x;
  const float* src = table + idx[r] * (int64_t)E;
  for (int e = threadIdx.x; e < width; e += blockDim.x) {
    const float v = e < E ? src[e] : 0.f;
    if constexpr (sizeof(OutT) == 2) out[r * out_ld + e] = __float2bfloat16(v);
    else out[r * out_ld + e] = v;
  }
}

__global__ void embedding_scatter_add_kernel(const float* __restrict__ dx, int64_t dx_ld, int E,
                                             const int64_t* __restrict__ idx, float* __restrict__ dtable) {
  const int64_t r = blockIdx.x;
  float* dst = dtable + idx[r] * (int64_t)E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) atomicAdd(dst + e, dx[r * dx_ld + e]);
}

// out[n] = sum_r x[r,n]: one thread per column, fixed ascending-r order (deterministic).
__global__ void colsum_kernel(const float* __restrict__ x, int64_t rows, int N, int64_t ld, float* __restrict__ out,
                              int permH) {
  // blockDim = (32, 8): 32 columns per CTA, 8 row-lanes reduced through shared memory.
  __shared__ float part[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (n < N)
    for (int64_t r = threadIdx.y; r < rows; r += 8) s += x[r * ld + n];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    out[permH ? gate_unperm(permH, n) : n] = t;
  }
}

// Two-stage column sum for tall matrices: stage 1 sums row slices (grid.y of them) into partial[slice][N],
// stage 2 adds the slices in order (deterministic).
__global__ void __launch_bounds__(256)
colsum_stage1_kernel(const float* __restrict__ x, int64_t rows, int N, int64_t ld, int64_t rows_per, float* __restrict__ partial) {
  __shared__ float red[2][128];
  const int col = blockIdx.x * 128 + (threadIdx.x & 127), rlane = threadIdx.x >> 7;
  const int64_t ra = (int64_t)blockIdx.y * rows_per, rb = ra + rows_per < rows ? ra + rows_per : rows;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (col < N) {
    int64_t r = ra + rlane;
    for (; r + 6 < rb; r += 8) {
      s0 += __ldcs(x + r * ld + col); s1 += __ldcs(x + (r + 2) * ld + col);
      s2 += __ldcs(x + (r + 4) * ld + col); s3 += __ldcs(x + (r + 6) * ld + col);
    }
    for (; r < rb; r += 2) s0 += __ldcs(x + r * ld + col);
  }
  red[rlane][threadIdx.x & 127] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (rlane == 0 && col < N) partial[(int64_t)blockIdx.y * N + col] = red[0][threadIdx.x] + red[1][threadIdx.x];
}
__global__ void colsum_stage2_kernel(const float* __restrict__ partial, int slices, int N, float* __restrict__ out, int permH) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int i = 0; i < slices; ++i) s += partial[(int64_t)i * N + n];
  out[permH ? gate_unperm(permH, n) : n] = s;
}

// Up to three column sums in ONE launch (the bias gradients after the time loop: dG -> b_ih = b_hh, dwq -> att_b,
// dwpart -> att_w): block = 32 columns x 32 row lanes, every thread sums rows y, y + 32, ... with 8 loads in flight, the
// lanes are added in a fixed order through shared memory (deterministic).  Three two-stage launches + a copy cost 36 us of
// serial small kernels behind the weight-gradient GEMM; this is one ~8 us launch.
struct ColsumJob {
  const float* x;
  int64_t rows, ld;
  int N, permH, blocks;
  float* out;
  float* out2;               // optional second copy of the result
};
struct ColsumJobs {
  ColsumJob j[3];
  int n;
};
__global__ void __launch_bounds__(1024)
colsum_multi_kernel(const ColsumJobs jobs) {
  __shared__ float red[32][33];
  int blk = blockIdx.x, k = 0;
  while (k + 1 < jobs.n && blk >= jobs.j[k].blocks) { blk -= jobs.j[k].blocks; ++k; }
  const ColsumJob& jb = jobs.j[k];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blk * 32 + tx;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  if (col < jb.N) {
    const float* p = jb.x + col;
    int64_t r = ty;
    for (; r + 7 * 32 < jb.rows; r += 8 * 32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += __ldcs(p + (r + i * 32) * jb.ld);
    }
    for (; r < jb.rows; r += 32) s[0] += __ldcs(p + r * jb.ld);
  }
  red[ty][tx] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  __syncthreads();
  if (ty == 0 && col < jb.N) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 32; ++g) t += red[g][tx];
    const int o = jb.permH ? gate_unperm(jb.permH, col) : col;
    jb.out[o] = t;
    if (jb.out2) jb.out2[o] = t;
  }
}

__global__ void caption_mask_kernel(const int64_t* __restrict__ cap, int64_t n, uint8_t* __restrict__ mask) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) mask[i] = (cap[i] != MVC_PAD && cap[i] != MVC_EOS) ? 1 : 0;
}

// ------------------------------------------------------------- clip + Adam(amsgrad)
__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, float* __restrict__ vmax, int64_t n, float lr, float b1,
                                 float b2, float eps, float wd, float clip, float bc1, float bc2_sqrt,
                                 float grad_scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);       // clip_grad_value_  train.py:207-208
    const float pi = p[i];
    gi = fmaf(wd, pi, gi);                                      // Adam weight_decay (L2)
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    const float vm = fmaxf(vmax[i], vi);                        // amsgrad
    m[i] = mi; v[i] = vi; vmax[i] = vm;
    const float denom = sqrtf(vm) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// Same update with the step count and the learning rate read from DEVICE memory, so that the launch can sit inside a
// CUDA graph: state[0] = step count (float, incremented by adam_tick_kernel before this kernel), state[1] = lr.
__global__ void clip_adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                     float* __restrict__ v, float* __restrict__ vmax, int64_t n,
                                     const float* __restrict__ state, float b1, float b2, float eps, float wd, float clip,
                                     float grad_scale) {
  const float step = state[0], lr = state[1];
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
    const float pi = p[i];
    gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    const float vm = fmaxf(vmax[i], vi);
    m[i] = mi; v[i] = vi; vmax[i] = vm;
    const float denom = sqrtf(vm) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}
__global__ void adam_tick_kernel(float* __restrict__ state) { state[0] += 1.f; }

// ---- gradient exchange fused with the optimiser over NVSwitch multicast (NVLS), one kernel:
//   reduce-scatter : this rank owns elements [lo, hi); multimem.ld_reduce.add on the MULTICAST address of the gradient
//                    buffer makes the switch sum the eight replicas' gradients and deliver ONE value per element;
//   update         : clip + Adam(amsgrad) on the owned slice only (1/world of the optimiser work and state traffic);
//   all-gather     : multimem.st of the updated PARAMETERS to the multicast address of the parameter buffer: the
//                    switch writes them into every replica.
// Per GPU that is n*4 bytes out + n*4 bytes in over NVLink (~42 us for 9.4 M parameters at 900 GB/s per direction)
// instead of an NCCL all-reduce (175 us measured, NVLS algorithm) followed by a full-size Adam pass (52 us).
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__global__ void clip_adam_multimem_kernel(const float* __restrict__ p_local, float* __restrict__ p_mc,
                                          const float* __restrict__ g_mc, float* __restrict__ m, float* __restrict__ v,
                                          float* __restrict__ vmax, int64_t lo, int64_t hi, const float* __restrict__ state,
                                          float b1, float b2, float eps, float wd, float clip, float grad_scale) {
  const float step = state[0], lr = state[1];
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  const float lr1 = lr / bc1;
  const int64_t n4 = (hi - lo) >> 2;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n4; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = lo + 4 * j;
    const float4 g4 = multimem_ld_reduce_add(g_mc + i);
    const float4 p4 = *reinterpret_cast<const float4*>(p_local + i);
    float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i),
           x4 = *reinterpret_cast<const float4*>(vmax + i);
    float gs[4] = {g4.x, g4.y, g4.z, g4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w}, ms[4] = {m4.x, m4.y, m4.z, m4.w},
          vs[4] = {v4.x, v4.y, v4.z, v4.w}, xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float gi = gs[e] * grad_scale;
      if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
      gi = fmaf(wd, ps[e], gi);
      ms[e] = b1 * ms[e] + (1.f - b1) * gi;
      vs[e] = b2 * vs[e] + (1.f - b2) * gi * gi;
      xs[e] = fmaxf(xs[e], vs[e]);
      ps[e] = ps[e] - lr1 * (ms[e] / (sqrtf(xs[e]) / bc2_sqrt + eps));
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(ms[0], ms[1], ms[2], ms[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vs[0], vs[1], vs[2], vs[3]);
    *reinterpret_cast<float4*>(vmax + i) = make_float4(xs[0], xs[1], xs[2], xs[3]);
    multimem_st(p_mc + i, make_float4(ps[0], ps[1], ps[2], ps[3]));
  }
  __threadfence_system();
}

// Same update with the reduce-scatter done by peer loads: the owner of a slice reads the N replicas' gradients itself
// (its own from local memory, the others through their P2P mappings, summed in rank order) instead of asking the switch
// for multimem.ld_reduce -- which fetches EVERY replica over NVLink, the requester's own included.  Per GPU and step the
// export drops from G + G/N to G (G = the gradient buffer): a third fewer NVLink bytes at 2 GPUs, 11 % at 8.
struct PeerGrads {
  const float* g[8];
  int n;
};
__global__ void clip_adam_p2p_kernel(const float* __restrict__ p_local, float* __restrict__ p_mc, const PeerGrads pg,
                                     float* __restrict__ m, float* __restrict__ v, float* __restrict__ vmax, int64_t lo,
                                     int64_t hi, const float* __restrict__ state, float b1, float b2, float eps, float wd,
                                     float clip, float grad_scale) {
  const float step = state[0], lr = state[1];
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  const float lr1 = lr / bc1;
  const int64_t n4 = (hi - lo) >> 2;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n4; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = lo + 4 * j;
    float4 t[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < pg.n) t[r] = __ldcg(reinterpret_cast<const float4*>(pg.g[r] + i));     // all replicas in flight together
    const float4 p4 = *reinterpret_cast<const float4*>(p_local + i);
    float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i),
           x4 = *reinterpret_cast<const float4*>(vmax + i);
    float4 g4 = t[0];
#pragma unroll
    for (int r = 1; r < 8; ++r)
      if (r < pg.n) { g4.x += t[r].x; g4.y += t[r].y; g4.z += t[r].z; g4.w += t[r].w; }
    float gs[4] = {g4.x, g4.y, g4.z, g4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w}, ms[4] = {m4.x, m4.y, m4.z, m4.w},
          vs[4] = {v4.x, v4.y, v4.z, v4.w}, xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float gi = gs[e] * grad_scale;
      if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
      gi = fmaf(wd, ps[e], gi);
      ms[e] = b1 * ms[e] + (1.f - b1) * gi;
      vs[e] = b2 * vs[e] + (1.f - b2) * gi * gi;
      xs[e] = fmaxf(xs[e], vs[e]);
      ps[e] = ps[e] - lr1 * (ms[e] / (sqrtf(xs[e]) / bc2_sqrt + eps));
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(ms[0], ms[1], ms[2], ms[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vs[0], vs[1], vs[2], vs[3]);
    *reinterpret_cast<float4*>(vmax + i) = make_float4(xs[0], xs[1], xs[2], xs[3]);
    multimem_st(p_mc + i, make_float4(ps[0], ps[1], ps[2], ps[3]));
  }
  __threadfence_system();
}

static inline int grid_for(int64_t n, int block = 256) {
  int64_t g = cdiv(n, block);
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace mvc

using namespace mvc;

extern "C" const char* mvc_last_error(void) { return mvc::g_err; }
extern "C" int mvc_version(void) { return 100; }
extern "C" long long mvc_launch_count(void) { return mvc::g_launches; }
extern "C" int mvc_prof_arm(int kid, int m, int n, int k) {
  mvc::g_prof_kid = kid; mvc::g_prof_m = m; mvc::g_prof_n = n; mvc::g_prof_k = k;
  mvc::g_prof_used = 0;
  return 0;
}
extern "C" int mvc_prof_collect(double* total_ms, long long* launches) {
  double tot = 0.0;
  for (size_t i = 0; i < mvc::g_prof_used; ++i) {
    MVC_CUDA(cudaEventSynchronize(mvc::g_prof_pool[i].second));
    float ms = 0.f;
    MVC_CUDA(cudaEventElapsedTime(&ms, mvc::g_prof_pool[i].first, mvc::g_prof_pool[i].second));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (long long)mvc::g_prof_used;
  mvc::g_prof_kid = 0;
  mvc::g_prof_used = 0;
  return 0;
}
extern "C" int mvc_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

extern "C" int mvc_concat_cast(const float* a, int Fa, const float* v, int Fv, int64_t rows, void* dst,
                               int dst_bf16, void* stream) {
  MVC_CHECK((Fa == 0 || a) && (Fv == 0 || v) && dst && Fa + Fv > 0, "mvc_concat_cast: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = rows * (Fa + Fv);
  if (n == 0) return 0;
  const bool al16 = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
  if (dst_bf16 && Fa % 8 == 0 && Fv % 8 == 0 && al16)
    concat_cast_bf16x8_kernel<<<grid_for(n / 8), 256, 0, st>>>(a, Fa, v, Fv, rows, (__nv_bfloat16*)dst);
  else if (dst_bf16) concat_cast_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, st>>>(a, Fa, v, Fv, rows, (__nv_bfloat16*)dst);
  else concat_cast_kernel<float><<<grid_for(n), 256, 0, st>>>(a, Fa, v, Fv, rows, (float*)dst);
  MVC_LAUNCH_CHECK();
  return 0;
}

static thread_local int g_input_format = MVC_INPUT_F32;

extern "C" int mvc_set_input_format(int fmt) {
  MVC_CHECK(fmt == MVC_INPUT_F32 || fmt == MVC_INPUT_BF16, "mvc_set_input_format: unknown format %d", fmt);
  g_input_format = fmt;
  return 0;
}
extern "C" int mvc_get_input_format(void) { return g_input_format; }

extern "C" int mvc_concat_bf16(const void* a, int Fa, const void* v, int Fv, int64_t rows, void* dst, void* stream) {
  MVC_CHECK((Fa == 0 || a) && (Fv == 0 || v) && dst && Fa + Fv > 0, "mvc_concat_bf16: bad arguments");
  MVC_CHECK(Fa % 8 == 0 && Fv % 8 == 0, "mvc_concat_bf16: feature widths must be multiples of 8 (Fa=%d Fv=%d)", Fa, Fv);
  const bool al16 = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
  MVC_CHECK(al16, "mvc_concat_bf16: buffers must be 16-byte aligned");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  concat_bf16x8_kernel<<<grid_for(rows * (Fa + Fv) / 8), 256, 0, st>>>((const __nv_bfloat16*)a, Fa, (const __nv_bfloat16*)v, Fv,
                                                                       rows, (__nv_bfloat16*)dst);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n == 0) return 0;
  cast_bf16_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  MVC_LAUNCH_CHECK();
  return 0;
}

namespace mvc {
int launch_transpose_bf16(const void* src, int src_bf16, int64_t R, int64_t C, int64_t lds, void* dst, int64_t ldd,
                          int permH, cudaStream_t st) {
  if (R == 0 || C == 0) return 0;
  // fast path: 16-byte accesses.  C need not be a multiple of 8 as long as the source rows are padded to the pitch
  // (the last vector of a row then reads pad columns, which are never written out: `c >= C` rows are skipped)
  if (src_bf16 && lds % 8 == 0 && ldd % 8 == 0 && (C % 8 == 0 || ((C + 7) / 8) * 8 <= lds) &&
      ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
    dim3 g64((unsigned)cdiv(C, 64), (unsigned)cdiv(R, 64));
    transpose_bf16_tile64_kernel<<<g64, 256, 0, st>>>((const __nv_bfloat16*)src, R, C, lds, (__nv_bfloat16*)dst, ldd, permH);
    MVC_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid((unsigned)cdiv(C, 32), (unsigned)cdiv(R, 32)), block(32, 8);
  if (src_bf16)
    transpose_bf16_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)src, R, C, lds, (__nv_bfloat16*)dst,
                                                                ldd, permH);
  else
    transpose_bf16_kernel<float><<<grid, block, 0, st>>>((const float*)src, R, C, lds, (__nv_bfloat16*)dst, ldd, permH);
  MVC_LAUNCH_CHECK();
  return 0;
}
int launch_cell_bwd(int B, int H, const float* act, const float* c_prev, const float* c_new, const float* dh_a,
                    int64_t dha_ld, const float* dh_b, int64_t dhb_ld, float* dc, float* dgates, void* dg_bf16, int perm,
                    cudaStream_t st) {
  const int64_t n = (int64_t)B * H;
  if (n == 0) return 0;
  ProfScope prof(PK_CELL_BWD, B, H, 0, st);
  lstm_cell_bwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(B, H, act, c_prev, c_new, dh_a, dha_ld, dh_b, dhb_ld, dc,
                                                              dgates, (__nv_bfloat16*)dg_bf16, perm);
  MVC_LAUNCH_CHECK();
  return 0;
}
static int colsum_scratch(cudaStream_t st, float** out, size_t* cap) {
  static std::mutex mu;
  static std::vector<std::pair<uint64_t, float*>> pool;
  constexpr size_t kBytes = 4u << 20;
  int dev = 0;
  MVC_CUDA(cudaGetDevice(&dev));
  const uint64_t key = (reinterpret_cast<uint64_t>(st) << 8) ^ (uint64_t)dev;
  std::lock_guard<std::mutex> lk(mu);
  for (auto& e : pool)
    if (e.first == key) { *out = e.second; *cap = kBytes; return 0; }
  float* pbuf = nullptr;
  MVC_CUDA(cudaMalloc(&pbuf, kBytes));
  pool.emplace_back(key, pbuf);
  *out = pbuf; *cap = kBytes;
  return 0;
}

int launch_colsum3(const float* x0, int64_t rows0, int N0, int64_t ld0, float* out0, float* out0b, int permH0,
                   const float* x1, int64_t rows1, int N1, int64_t ld1, float* out1, const float* x2, int64_t rows2, int N2,
                   int64_t ld2, float* out2, cudaStream_t st) {
  ColsumJobs jobs{};
  jobs.n = 3;
  jobs.j[0] = ColsumJob{x0, rows0, ld0, N0, permH0, (int)cdiv(N0, 32), out0, out0b};
  jobs.j[1] = ColsumJob{x1, rows1, ld1, N1, 0, (int)cdiv(N1, 32), out1, nullptr};
  jobs.j[2] = ColsumJob{x2, rows2, ld2, N2, 0, (int)cdiv(N2, 32), out2, nullptr};
  const int blocks = jobs.j[0].blocks + jobs.j[1].blocks + jobs.j[2].blocks;
  if (blocks == 0) return 0;
  colsum_multi_kernel<<<(unsigned)blocks, 1024, 0, st>>>(jobs);
  MVC_LAUNCH_CHECK();
  return 0;
}

int launch_colsum(const float* x, int64_t rows, int N, int64_t ld, float* out, int permH, cudaStream_t st) {
  if (N == 0) return 0;
  if (rows >= 512) {
    float* part = nullptr;
    size_t cap = 0;
    MVC_TRY(colsum_scratch(st, &part, &cap));
    int slices = (int)(rows / 64 < 64 ? rows / 64 : 64);
    while ((size_t)slices * N * sizeof(float) > cap && slices > 1) slices /= 2;
    const int64_t rows_per = cdiv(rows, slices);
    dim3 g1((unsigned)cdiv(N, 128), (unsigned)slices);
    colsum_stage1_kernel<<<g1, 256, 0, st>>>(x, rows, N, ld, rows_per, part);
    MVC_LAUNCH_CHECK();
    colsum_stage2_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(part, slices, N, out, permH);
    MVC_LAUNCH_CHECK();
    return 0;
  }
  dim3 block(32, 8);
  colsum_kernel<<<(unsigned)cdiv(N, 32), block, 0, st>>>(x, rows, N, ld, out, permH);
  MVC_LAUNCH_CHECK();
  return 0;
}
}  // namespace mvc

extern "C" int mvc_transpose_to_bf16(const void* src, int src_bf16, int64_t R, int64_t C, int64_t lds, void* dst,
                                     int64_t ldd, void* stream) {
  return launch_transpose_bf16(src, src_bf16, R, C, lds, dst, ldd, 0, (cudaStream_t)stream);
}

extern "C" int mvc_lstm_cell_fwd(int B, int H, const float* pre, const float* gx, int64_t gx_ld,
                                 const float* emb_table, const int64_t* tokens, const float* bias,
                                 const float* c_prev, float* act, float* c_out, float* h_out, int64_t h_ld,
                                 float* h_out2, int64_t h2_ld, void* h_bf16, int64_t hb_ld, void* stream) {
  MVC_CHECK(pre && c_out, "mvc_lstm_cell_fwd: null pre/c_out");
  MVC_CHECK(!emb_table || tokens, "mvc_lstm_cell_fwd: emb_table without tokens");
  const int64_t n = (int64_t)B * H;
  if (n == 0) return 0;
  ProfScope prof(PK_CELL_FWD, B, H, 0, (cudaStream_t)stream);
  lstm_cell_fwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
      B, H, pre, gx, gx_ld, emb_table, tokens, bias, c_prev, act, c_out, h_out, h_ld, h_out2, h2_ld,
      (__nv_bfloat16*)h_bf16, hb_ld);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_lstm_cell_bwd(int B, int H, const float* act, const float* c_prev, const float* c_new,
                                 const float* dh_a, int64_t dha_ld, const float* dh_b, int64_t dhb_ld, float* dc,
                                 float* dgates, void* dg_bf16, void* stream) {
  MVC_CHECK(act && c_new && dc && dgates, "mvc_lstm_cell_bwd: null argument");
  return launch_cell_bwd(B, H, act, c_prev, c_new, dh_a, dha_ld, dh_b, dhb_ld, dc, dgates, dg_bf16, 0,
                         (cudaStream_t)stream);
}

extern "C" int mvc_log_softmax_rows(float* x, int64_t rows, int V, int64_t* argmax, void* stream) {
  if (rows == 0) return 0;
  MVC_CHECK(x && V > 0, "mvc_log_softmax_rows: bad arguments");
  ProfScope prof(PK_LOGSOFTMAX, (int)rows, V, 0, (cudaStream_t)stream);
  log_softmax_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, V, argmax);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_argmax_rows(const float* x, const float* y, int64_t rows, int V, int64_t* out, void* stream) {
  if (rows == 0) return 0;
  MVC_CHECK(x && out && V > 0, "mvc_argmax_rows: bad arguments");
  argmax_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, y, V, out);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_log_softmax_bwd(const float* logp, const float* dlogp, int64_t rows, int V, float* dlogits,
                                   void* dlogits_bf16, void* stream) {
  if (rows == 0) return 0;
  MVC_CHECK(logp && dlogp && (dlogits || dlogits_bf16), "mvc_log_softmax_bwd: bad arguments");
  const int64_t ldb = (V + 7) / 8 * 8;
  // bf16 path (training hot path): row kept in registers, one read of each operand
  if (dlogits_bf16 && V <= 16 * 256)
    log_softmax_bwd_kernel<16><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(logp, dlogp, V, dlogits,
                                                                              (__nv_bfloat16*)dlogits_bf16, ldb);
  else
    log_softmax_bwd_kernel<0><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(logp, dlogp, V, dlogits,
                                                                             (__nv_bfloat16*)dlogits_bf16, ldb);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_embedding_gather(const float* table, int E, const int64_t* idx, int64_t rows, void* out,
                                    int64_t out_ld, int out_bf16, void* stream) {
  if (rows == 0) return 0;
  MVC_CHECK(table && idx && out && out_ld >= E, "mvc_embedding_gather: bad arguments");
  const int width = out_bf16 ? (int)out_ld : E;   // bf16 rows are zero-padded to out_ld (TMA K padding)
  if (out_bf16) embedding_gather_kernel<__nv_bfloat16><<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(
      table, E, idx, (__nv_bfloat16*)out, out_ld, width);
  else embedding_gather_kernel<float><<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(table, E, idx, (float*)out,
                                                                                     out_ld, width);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_embedding_scatter_add(const float* dx, int64_t dx_ld, int E, const int64_t* idx, int64_t rows,
                                         float* dtable, void* stream) {
  if (rows == 0) return 0;
  embedding_scatter_add_kernel<<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(dx, dx_ld, E, idx, dtable);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_colsum(const float* x, int64_t rows, int N, int64_t ld, float* out, void* stream) {
  return launch_colsum(x, rows, N, ld, out, 0, (cudaStream_t)stream);
}

extern "C" int mvc_caption_mask(const int64_t* captions, int64_t n, uint8_t* mask, void* stream) {
  if (n == 0) return 0;
  caption_mask_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(captions, n, mask);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_clip_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                  float* max_exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                                  float weight_decay, float clip_value, int step, float grad_scale, void* stream) {
  if (n == 0) return 0;
  MVC_CHECK(step >= 1, "mvc_clip_adam_step: step counts from 1");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  ProfScope prof(PK_ADAM, 0, 0, 0, (cudaStream_t)stream);
  clip_adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, n,
                                                               lr, beta1, beta2, eps, weight_decay, clip_value, bc1,
                                                               bc2_sqrt, grad_scale);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_clip_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                      float* max_exp_avg_sq, int64_t n, float* state_dev, int tick, float beta1, float beta2,
                                      float eps, float weight_decay, float clip_value, float grad_scale, void* stream) {
  if (n == 0) return 0;
  MVC_CHECK(state_dev, "mvc_clip_adam_step_dev: null state");
  if (tick) {
    adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state_dev);
    MVC_LAUNCH_CHECK();
  }
  ProfScope prof(PK_ADAM, 0, 0, 0, (cudaStream_t)stream);
  clip_adam_dev_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, n,
                                                                   state_dev, beta1, beta2, eps, weight_decay, clip_value,
                                                                   grad_scale);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_clip_adam_p2p_multimem(const float* param_local, float* param_mc, const float* const* grad_replicas,
                                          int world, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, int64_t lo,
                                          int64_t hi, float* state_dev, int tick, float beta1, float beta2, float eps,
                                          float weight_decay, float clip_value, float grad_scale, void* stream) {
  MVC_CHECK(param_local && param_mc && grad_replicas && exp_avg && exp_avg_sq && max_exp_avg_sq && state_dev,
            "mvc_clip_adam_p2p_multimem: null argument");
  MVC_CHECK(world >= 1 && world <= 8, "mvc_clip_adam_p2p_multimem: world %d not in [1, 8]", world);
  MVC_CHECK(lo >= 0 && hi >= lo && lo % 4 == 0 && hi % 4 == 0, "mvc_clip_adam_p2p_multimem: [lo, hi) must be multiples of 4");
  PeerGrads pg{};
  pg.n = world;
  uintptr_t bits = reinterpret_cast<uintptr_t>(param_mc) | reinterpret_cast<uintptr_t>(param_local);
  for (int r = 0; r < world; ++r) {
    MVC_CHECK(grad_replicas[r], "mvc_clip_adam_p2p_multimem: null replica pointer %d", r);
    pg.g[r] = grad_replicas[r];
    bits |= reinterpret_cast<uintptr_t>(grad_replicas[r]);
  }
  MVC_CHECK((bits & 15u) == 0, "mvc_clip_adam_p2p_multimem: buffers must be 16-byte aligned");
  if (tick) {
    adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state_dev);
    MVC_LAUNCH_CHECK();
  }
  if (hi == lo) return 0;
  ProfScope prof(PK_ADAM, 0, 0, 0, (cudaStream_t)stream);
  clip_adam_p2p_kernel<<<grid_for((hi - lo) / 4), 256, 0, (cudaStream_t)stream>>>(
      param_local, param_mc, pg, exp_avg, exp_avg_sq, max_exp_avg_sq, lo, hi, state_dev, beta1, beta2, eps, weight_decay,
      clip_value, grad_scale);
  MVC_LAUNCH_CHECK();
  return 0;
}

extern "C" int mvc_clip_adam_multimem(const float* param_local, float* param_mc, const float* grad_mc, float* exp_avg,
                                      float* exp_avg_sq, float* max_exp_avg_sq, int64_t lo, int64_t hi, float* state_dev,
                                      int tick, float beta1, float beta2, float eps, float weight_decay, float clip_value,
                                      float grad_scale, void* stream) {
  MVC_CHECK(param_local && param_mc && grad_mc && exp_avg && exp_avg_sq && max_exp_avg_sq && state_dev,
            "mvc_clip_adam_multimem: null argument");
  MVC_CHECK(lo >= 0 && hi >= lo && lo % 4 == 0 && hi % 4 == 0, "mvc_clip_adam_multimem: [lo, hi) must be multiples of 4");
  MVC_CHECK(((reinterpret_cast<uintptr_t>(param_mc) | reinterpret_cast<uintptr_t>(grad_mc) |
              reinterpret_cast<uintptr_t>(param_local)) & 15u) == 0, "mvc_clip_adam_multimem: buffers must be 16-byte aligned");
  if (tick) {
    adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state_dev);
    MVC_LAUNCH_CHECK();
  }
  if (hi == lo) return 0;
  ProfScope prof(PK_ADAM, 0, 0, 0, (cudaStream_t)stream);
  clip_adam_multimem_kernel<<<grid_for((hi - lo) / 4), 256, 0, (cudaStream_t)stream>>>(
      param_local, param_mc, grad_mc, exp_avg, exp_avg_sq, max_exp_avg_sq, lo, hi, state_dev, beta1, beta2, eps,
      weight_decay, clip_value, grad_scale);
  MVC_LAUNCH_CHECK();
  return 0;
}
