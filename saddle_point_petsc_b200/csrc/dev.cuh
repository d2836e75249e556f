// dev.cuh -- device-side helpers shared by the kernels of libb200sp (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "core.h"

namespace b200sp {

constexpr unsigned FULL = 0xffffffffu;

// streaming 128-bit loads that do not allocate in L1 (matrix values / indices / vectors read once)
__device__ __forceinline__ double2 ld_stream_f64x2(const double *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int2 ld_stream_s32x2(const int *p) {
  int2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ld_stream_s32x4(const int *p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_stream_f64(const double *p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ int ld_stream_s32(const int *p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// Deterministic grid-wide sum of K per-thread accumulators (blockDim.x == 256).
//   stage 1: warp shuffle tree -> shared -> per-block partial written to partials[block][j]
//   stage 2: the last block to arrive (atomic ticket) sums the partials in a fixed order and writes out[j].
// The result depends only on (n, gridDim), never on scheduling.
template <int K>
__device__ __forceinline__ void grid_reduce_sum(double (&acc)[K], int k, double *partials, unsigned *ticket, double *out) {
  __shared__ double s_w[8][K];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    double v = warp_sum(acc[j]);
    if (lane == 0) s_w[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < k) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_w[w][threadIdx.x];
    partials[(size_t)blockIdx.x * RED_MAX_OUT + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int j = warp; j < k; j += 8) {
    double s = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(partials + (size_t)b * RED_MAX_OUT + j);
    s = warp_sum(s);
    if (lane == 0) out[j] = s;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

} // namespace b200sp
