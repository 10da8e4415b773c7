// Shared helpers for the clipguide_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/clipguide_b200.h"

#ifndef CG_NUM_SMS
#define CG_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this
#endif

void cg_set_error(const char* fmt, ...);

#define CG_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      cg_set_error(__VA_ARGS__);       \
      return CG_EINVAL;                \
    }                                  \
  } while (0)

#define CG_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      cg_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                  \
    }                                                                                  \
  } while (0)

#define CG_LAUNCH_CHECK()                                                              \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      cg_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                  \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; every thread gets the result.  `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

static inline cudaStream_t cg_stream(void* s) { return (cudaStream_t)s; }

// Programmatic dependent launch.  The guidance step is a long chain of short dependent kernels (per ViT layer: 9 GEMMs, 3 attention
// kernels, 4 LayerNorms; many run 10-40 us), so launch latency and kernel prologues (barrier init, TMEM allocation, descriptor
// prefetch) are a visible share of the step.  Kernels launched through cg_launch_pdl() carry
// cudaLaunchAttributeProgrammaticStreamSerialization: the grid may start while its predecessor in the stream is still draining; it
// runs its prologue and then blocks in cg_griddep_wait() until the predecessor has completed and its writes are visible.  Every
// such kernel calls cg_griddep_launch() at its top (the dependent is scheduled once ALL CTAs of this grid have started) and
// cg_griddep_wait() in every thread before its first global-memory access.  CG_PDL=0 falls back to plain stream order.
__device__ __forceinline__ void cg_griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void cg_griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool cg_pdl_enabled();  // capi.cu
template <typename... KArgs, typename... Args>
cudaError_t cg_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cg_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
