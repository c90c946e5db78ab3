// common.cuh -- shared helpers for libmvc_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mvc_b200.h"

namespace mvc {

void set_error(const char* fmt, ...);
extern long long g_launches;

// In-situ kernel timing (mvc_prof_*): when a kernel class is armed, its launch sites bracket the
// launch with CUDA events on the launching stream, so bench.py can read the average duration of the
// dominant kernel over the very steps it times.  Disarmed cost: one predictable branch.
enum ProfKid {
  PK_NONE = 0, PK_GEMM_TC = 1, PK_GEMM_F32 = 2, PK_ATTN_FWD = 3, PK_ATTN_BWD = 4, PK_CELL_FWD = 5, PK_CELL_BWD = 6,
  PK_LOGSOFTMAX = 7, PK_STEP_FUSED = 8, PK_LOSS = 9, PK_ADAM = 10
};
extern int g_prof_kid, g_prof_m, g_prof_n, g_prof_k;
void prof_begin(cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
  bool on;
  cudaStream_t st;
  ProfScope(int kid, int m, int n, int k, cudaStream_t s) : st(s) {
    on = g_prof_kid == kid && (g_prof_m < 0 || g_prof_m == m) && (g_prof_n < 0 || g_prof_n == n) &&
         (g_prof_k < 0 || g_prof_k == k);
    if (on) prof_begin(st);
  }
  ~ProfScope() {
    if (on) prof_end(st);
  }
};

#define MVC_CHECK(cond, ...)            \
  do {                                  \
    if (!(cond)) {                      \
      mvc::set_error(__VA_ARGS__);      \
      return 1;                         \
    }                                   \
  } while (0)

#define MVC_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      mvc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

// Every kernel launch site ends with this: counts the launch (mvc_launch_count) and surfaces errors.
#define MVC_LAUNCH_CHECK()        \
  do {                            \
    ++mvc::g_launches;            \
    MVC_CUDA(cudaGetLastError()); \
  } while (0)

#define MVC_TRY(expr)        \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

constexpr int kNumSMs = 148;

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Bump allocator over a caller-provided workspace.  Running it with a null
// base measures the bytes a layout needs; forward and backward run the same
// layout function so they agree on where every saved tensor lives.
struct Arena {
  char* base;
  size_t off;
  explicit Arena(void* b) : base(static_cast<char*>(b)), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `red` is >= 32 floats of shared memory.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// bf16-path LSTM cell activations: ex2.approx + fast reciprocal (abs. error ~1e-7, far below the bf16 rounding of the
// operands) -- a handful of instructions instead of the ~40-instruction IEEE expf / tanhf sequences, which made the
// fused cell epilogues instruction-fetch bound (11 K SASS instructions executed once per CTA).
__device__ __forceinline__ float sigmoid_ex2(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_ex2(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }

}  // namespace mvc
