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
// current device ordinal, clamped to the size of the per-device caches (function attributes and SM counts are per device)
constexpr int CG_MAX_DEVICES = 64;
static inline int cg_device_index() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : (dev >= CG_MAX_DEVICES ? CG_MAX_DEVICES - 1 : dev);
}

// Deterministic cross-block reductions (loss values, sum of squares of the guidance gradient): every block stores its partial in a
// library-owned scratch buffer and the LAST block to arrive (atomic ticket -- the only atomic, and order independent) adds the
// partials in index order.  No float atomicAdd anywhere: values are bit-reproducible run to run.  The scratch is per device and shared
// by all reductions: like the reference (module globals, one job per process) the library is not re-entrant across streams.
constexpr int CG_SCRATCH_FLOATS = 1 << 20;  // partials
constexpr int CG_SCRATCH_COUNTERS = 16;     // one ticket counter per reduction kind
struct CgScratch {
  float* partials;
  unsigned* counters;
};
int cg_get_scratch(CgScratch* out);  // capi.cu: lazily allocated (and zeroed) per device

// Sum `v` of every block of the grid over the contiguous block range [first, first + count) -- called by ALL threads of ALL blocks of the
// grid.  Returns true in the last block to arrive, where `total` then holds the sum over blocks [0, nblocks) split as the caller wants:
// the caller loops over its ranges with cg_sum_partials().
__device__ __forceinline__ bool cg_last_block(float v, float* partials, unsigned* counter, unsigned block_linear, unsigned nblocks) {
  __shared__ bool is_last;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    partials[block_linear] = v;
    __threadfence();
    const unsigned ticket = atomicAdd(counter, 1u);
    is_last = ticket == nblocks - 1;
    if (is_last) *counter = 0;  // ready for the next launch
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}
// fixed-order sum of partials[first .. first+count) by one whole block; every thread gets the result.  `red` >= 32 floats of smem.
__device__ __forceinline__ float cg_sum_partials(const float* partials, unsigned first, unsigned count, float* red) {
  const unsigned tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
  float s = 0.f;
  for (unsigned i = tid; i < count; i += nt) s += __ldcg(partials + first + i);
  return block_sum(s, red);
}

// the same by ONE warp (lane-strided, then a fixed shuffle tree): lets the warps of the last block finish different images in parallel
__device__ __forceinline__ float cg_sum_partials_warp(const float* partials, unsigned first, unsigned count) {
  float s = 0.f;
  for (unsigned i = threadIdx.x & 31; i < count; i += 32) s += __ldcg(partials + first + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

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
