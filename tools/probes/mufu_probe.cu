// Probe: throughput of MUFU.EX2 (ex2.approx.ftz.f32), F2FP.BF16 pack and an FMA-pipe polynomial exp2 per SM sub-partition on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe mufu_probe.cu ; prints cycles per warp-instruction.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MODE == 0) v[j] = ex2f(v[j]);
      else if (MODE == 1) { __nv_bfloat162 b = __floats2bfloat162_rn(v[j], v[(j + 1) & 15]); v[j] = __uint_as_float(*reinterpret_cast<unsigned*>(&b)) * 0.f + v[j] * 0.999f; }
      else if (MODE == 2) {  // degree-3 polynomial 2^f on [0,1) + exponent insert (FA4-style), x <= 0
        float x = fmaxf(v[j], -126.f);
        float fl = floorf(x); float f = x - fl;
        float pz = fmaf(fmaf(fmaf(0.0555054f, f, 0.2402265f), f, 0.6931472f), f, 1.0f);
        v[j] = __int_as_float(__float_as_int(pz) + (((int)fl) << 23)) * -0.5f;
      } else v[j] = fmaf(v[j], 0.999f, -0.001f);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const char* names[4] = {"MUFU.EX2", "F2FP.BF16 pack (+FMUL,FFMA)", "poly exp2 (FMA pipe)", "FFMA"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps = 1; warps <= 16; warps *= 2) {  // warps per CTA, one CTA per SM: warps/4 per sub-partition (>= 4: one per SMSP each)
      const int iters = 2000;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<1, warps * 32>>>(out, cyc, iters);
        if (mode == 1) probe<1><<<1, warps * 32>>>(out, cyc, iters);
        if (mode == 2) probe<2><<<1, warps * 32>>>(out, cyc, iters);
        if (mode == 3) probe<3><<<1, warps * 32>>>(out, cyc, iters);
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double per = (double)c / (iters * 16.0);
      printf("%-30s warps/CTA %2d (per SMSP %.2f): %.2f cycles per warp-op of one warp  => %.2f cycles per op per SMSP\n", names[mode], warps, warps / 4.0, per,
             per / (warps >= 4 ? warps / 4.0 : 1.0));
    }
  return 0;
}
