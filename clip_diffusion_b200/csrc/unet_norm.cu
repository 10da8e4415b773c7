// NHWC fp16 kernels for everything BETWEEN the cuDNN convolutions of the replicated guided-diffusion UNet (the
// UNet is built at clip_diffusion/models.py:87-131, its blocks are restated in App. A.3: ResBlock = GN32 -> SiLU -> conv,
// GN32 -> *(1+scale)+shift -> SiLU -> conv, + skip; AttentionBlock = GN32 -> qkv; resblock_updown resamplers; skip concatenations).
//
// Why: after the CLIP side ran on tensor cores, a 512x512 guidance step spent ~25 of 60 ms in stock element-wise kernels around
// cuDNN (NCHW<->NHWC transposes 10 ms, addcmul 7.6 ms, Welford/sum reductions 5 ms, SiLU 2 ms, conv-bias passes 2.9 ms ...;
// profiles/r01_c2_step_kernel_table_*).  Keeping the trunk NHWC removes the transposes; these kernels make a normalisation
// 2-3 (forward) / 5-6 (backward) HBM passes over the activation instead of 10-14:
//
//   forward : partial (per-channel sum / sum of squares per row chunk; skipped when the producer of x already wrote them)
//             -> finalize (mean, rstd per group in fp64; per-channel affine a_c, b_c with gamma/beta, the timestep scale-shift and the
//             producing convolution's deferred bias folded in)  ->  apply y = act(a_c*x + b_c)
//   backward: partial (per-channel sum dv and sum dv*x, dv = dy*silu'(a_c*x+b_c) recomputed)  ->  finalize (per-group B, C)
//             ->  apply dx = a_c*dv + B_g*x + C_g (+ the gradient reaching x through the block's skip path)
//   others  : bias + residual add, 2x average-pool / nearest-upsample (each the other's gradient), skip concat / split; the add and
//             the concat can emit the next normalisation's partials while they stream their result out.
//
// All are HBM-bound streaming kernels: 128-bit accesses of 8 channels per thread, a thread keeps ONE channel octet for its
// whole life (coefficients live in registers), rows are split into chunks (~8 CTAs per SM at the large levels), four independent
// loads in flight per thread (volatile ld.global.nc -- see ldg_stream).  No atomics: partials are reduced in a fixed order =>
// deterministic.  Every launch is a programmatic dependent launch (griddepcontrol), which also survives CUDA-graph capture.
// Weights are frozen (models.py:120-127 only re-enables grads that never reach the sampler state) => no dgamma/dbeta.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int kMaxThreads = 256;
constexpr int kUnroll = 4;

struct Geo {
  int cvecs;           // C / 8: channel octets per row
  int rows_per_iter;   // rows a CTA covers per loop iteration
  int threads;         // cvecs * rows_per_iter  (<= 256)
  int chunks;          // row chunks per sample (gridDim.x)
  int rows_per_chunk;
};

static Geo geometry(int N, int HW, int C) {
  Geo g;
  g.cvecs = C / 8;
  g.rows_per_iter = kMaxThreads / g.cvecs;
  if (g.rows_per_iter < 1) g.rows_per_iter = 1;
  g.threads = g.cvecs * g.rows_per_iter;
  // ~8 CTAs per SM (ncu: at 4 the kernels sat at 50 % occupancy and 53 % of DRAM throughput, latency bound), each chunk a whole
  // number of unrolled iterations so that (almost) no thread runs the one-load-at-a-time tail loop
  int target = (8 * CG_NUM_SMS) / (N > 0 ? N : 1);
  if (target < 1) target = 1;
  const int min_rows = g.rows_per_iter * kUnroll;
  int max_chunks = HW / min_rows;
  if (max_chunks < 1) max_chunks = 1;
  int chunks = target < max_chunks ? target : max_chunks;
  g.rows_per_chunk = (HW + chunks - 1) / chunks;
  g.rows_per_chunk = ((g.rows_per_chunk + min_rows - 1) / min_rows) * min_rows;
  g.chunks = (HW + g.rows_per_chunk - 1) / g.rows_per_chunk;
  return g;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
  return u;
}

// sigmoid(v) = 0.5*tanh(0.5 v) + 0.5 : one MUFU instead of exp + rcp (the kernels must stay HBM-bound, not MUFU-bound)
__device__ __forceinline__ float sigmoid_fast(float v) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
  return fmaf(0.5f, t, 0.5f);
}

// 16-byte read-only load as a volatile asm: the compiler may neither sink it to its first use nor reorder it against its
// siblings, so the kUnroll loads a thread issues per iteration are really in flight together.  (Written as plain __ldg the
// loads were re-serialised -- load, convert, accumulate, next address, load ... -- and the reduction kernels ran at 2.7 TB/s.)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <typename T>
struct Raw8;  // the registers of 8 consecutive channels of one pixel, before conversion
template <>
struct Raw8<__half> {
  uint4 u;
  __device__ __forceinline__ void load(const __half* p) { u = ldg_stream(p); }
  __device__ __forceinline__ void get(float* f) const { unpack8(u, f); }
};
template <>
struct Raw8<float> {
  uint4 u0, u1;
  __device__ __forceinline__ void load(const float* p) {
    u0 = ldg_stream(p);
    u1 = ldg_stream(p + 4);
  }
  __device__ __forceinline__ void get(float* f) const {
    f[0] = __uint_as_float(u0.x); f[1] = __uint_as_float(u0.y); f[2] = __uint_as_float(u0.z); f[3] = __uint_as_float(u0.w);
    f[4] = __uint_as_float(u1.x); f[5] = __uint_as_float(u1.y); f[6] = __uint_as_float(u1.z); f[7] = __uint_as_float(u1.w);
  }
};

template <typename T>
struct Row8;  // 8 consecutive channels of one pixel
template <>
struct Row8<__half> {
  static __device__ __forceinline__ void load(const __half* p, float* f) { unpack8(__ldg(reinterpret_cast<const uint4*>(p)), f); }
  static __device__ __forceinline__ void store(__half* p, const float* f) { *reinterpret_cast<uint4*>(p) = pack8(f); }
};
template <>
struct Row8<float> {
  static __device__ __forceinline__ void load(const float* p, float* f) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
};

__device__ __forceinline__ void load_coef8(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Reduce the per-thread octet accumulators of the CTA's rows_per_iter row lanes and write one (s0, s1) pair per channel.
__device__ __forceinline__ void reduce_rows_and_store(const float* s0, const float* s1, int C, int cvecs, int rows_per_iter, float* red /*[2][2048]*/,
                                                      float2* __restrict__ out /*[C]*/) {
  const int col = threadIdx.x % cvecs, r = threadIdx.x / cvecs;
  float* r0 = red + r * C + col * 8;
  float* r1 = red + 2048 + r * C + col * 8;
  reinterpret_cast<float4*>(r0)[0] = make_float4(s0[0], s0[1], s0[2], s0[3]);
  reinterpret_cast<float4*>(r0)[1] = make_float4(s0[4], s0[5], s0[6], s0[7]);
  reinterpret_cast<float4*>(r1)[0] = make_float4(s1[0], s1[1], s1[2], s1[3]);
  reinterpret_cast<float4*>(r1)[1] = make_float4(s1[4], s1[5], s1[6], s1[7]);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < rows_per_iter; ++k) {
      a += red[k * C + c];
      b += red[2048 + k * C + c];
    }
    out[c] = make_float2(a, b);
  }
}

// ---- programmatic dependent launch ----------------------------------------------------------------------------------------------------
// A normalisation is three dependent launches (partial -> finalize -> apply) and most of the UNet's 115 + 115 of them work on
// maps of a few MB where each kernel runs 3-5 us: launch latency is a third of the op.  The finalize and apply kernels are
// launched with cudaLaunchAttributeProgrammaticStreamSerialization and start with griddepcontrol.wait; their predecessor issues
// griddepcontrol.launch_dependents at its top, so the dependent grid is scheduled (and its prologue runs) while the
// predecessor is still streaming, and only the data dependency remains.  CG_PDL=0 falls back to plain stream order.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- forward -----------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kMaxThreads) gn_stats_partial_kernel(const __half* __restrict__ x, int HW, int C, int cvecs, int rows_per_iter,
                                                                        int rows_per_chunk, float2* __restrict__ partial) {
  __shared__ __align__(16) float red[2 * 2048];
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y, p = blockIdx.x;
  const int col = threadIdx.x % cvecs, r = threadIdx.x / cvecs;
  const int row_end = min(HW, (p + 1) * rows_per_chunk);
  const __half* base = x + (size_t)n * HW * C + col * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  // main loop: kUnroll unconditional loads in flight per thread, pointers advance by a constant stride; then a <kUnroll-row tail
  const size_t rstride = (size_t)rows_per_iter * C;
  int row = p * rows_per_chunk + r;
  const __half* ptr = base + (size_t)row * C;
  auto accumulate = [&](const Raw8<__half>& raw) {
    float v[8];
    raw.get(v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += v[j];
      q[j] = fmaf(v[j], v[j], q[j]);
    }
  };
  for (; row + (kUnroll - 1) * rows_per_iter < row_end; row += kUnroll * rows_per_iter, ptr += kUnroll * rstride) {
    Raw8<__half> raw[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) raw[k].load(ptr + k * rstride);
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) accumulate(raw[k]);
  }
  for (; row < row_end; row += rows_per_iter, ptr += rstride) {
    Raw8<__half> raw;
    raw.load(ptr);
    accumulate(raw);
  }
  reduce_rows_and_store(s, q, C, cvecs, rows_per_iter, red, partial + ((size_t)n * gridDim.x + p) * C);
}

// both sums of a finalize kernel in ONE pass (two barriers instead of four)
__device__ __forceinline__ void block_sum2_f64(double& a, double& b, double* red /*>=64*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    red[2 * wid] = a;
    red[2 * wid + 1] = b;
  }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  a = b = 0.0;
  for (int k = 0; k < nw; ++k) {
    a += red[2 * k];
    b += red[2 * k + 1];
  }
}

constexpr int kFinalizeThreads = 512;

// one CTA per (sample, group): statistics in fp64 from the chunk partials, then the per-channel affine of the whole
// normalisation: y = act(a_c * x + b_c),  a_c = rstd*gamma_c*(1+scale_c),  b_c = (beta_c - mean*rstd*gamma_c)*(1+scale_c) + shift_c.
// pre_bias (the producing convolution's bias, deferred): the normalised tensor is x' = x + pb_c.  Its statistics follow from the
// per-channel sums of x (sum x' = s + HW*pb, sum x'^2 = q + 2*pb*s + HW*pb^2) and y = a_c*x + (a_c*pb_c + b_c), so the
// streaming kernels never see it.
__global__ void __launch_bounds__(kFinalizeThreads) gn_finalize_fwd_kernel(const float2* __restrict__ partial, int chunks, int C, int G, int HW,
                                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                            const float* __restrict__ scale_shift, const float* __restrict__ pre_bias,
                                                                            float eps, float* __restrict__ stats, float* __restrict__ coefA,
                                                                            float* __restrict__ coefB) {
  __shared__ double red[64];
  const int n = blockIdx.x / G, g = blockIdx.x % G, cg = C / G;
  const int total = chunks * cg;
  const float2* base = partial + (size_t)n * chunks * C + g * cg;
  double s = 0.0, q = 0.0;
  pdl_launch_dependents();
  pdl_wait();  // the chunk partials of the preceding kernel are complete and visible
  // the loads of one batch are independent: issue four before touching the fp64 accumulators (latency-bound otherwise)
  for (int idx0 = threadIdx.x; idx0 < total; idx0 += 4 * kFinalizeThreads) {
    float2 v[4];
    int cc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = idx0 + k * kFinalizeThreads;
      const int p = idx / cg;
      cc[k] = idx - p * cg;
      v[k] = idx < total ? __ldg(base + (size_t)p * C + cc[k]) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s += (double)v[k].x;
      q += (double)v[k].y;
      if (pre_bias && idx0 + k * kFinalizeThreads < total) q += 2.0 * (double)pre_bias[g * cg + cc[k]] * (double)v[k].x;
    }
  }
  if (pre_bias)
    for (int cc = threadIdx.x; cc < cg; cc += blockDim.x) {
      const double pb = (double)pre_bias[g * cg + cc];
      s += (double)HW * pb;
      q += (double)HW * pb * pb;
    }
  block_sum2_f64(s, q, red);
  const double m = (double)HW * (double)cg;
  const double mean = s / m;
  double var = q / m - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float meanf = (float)mean;
  if (threadIdx.x == 0) {
    stats[2 * blockIdx.x] = meanf;
    stats[2 * blockIdx.x + 1] = rstd;
  }
  for (int cc = threadIdx.x; cc < cg; cc += blockDim.x) {
    const int c = g * cg + cc;
    float a = gamma[c] * rstd;
    float b = beta[c] - meanf * a;
    if (scale_shift) {
      // the reference casts the embedding projection to the activation dtype first (`.type(h.dtype)`, guided-diffusion ResBlock)
      const float sc = 1.f + __half2float(__float2half_rn(scale_shift[(size_t)n * 2 * C + c]));
      const float sh = __half2float(__float2half_rn(scale_shift[(size_t)n * 2 * C + C + c]));
      a *= sc;
      b = fmaf(b, sc, sh);
    }
    if (pre_bias) b = fmaf(a, pre_bias[c], b);
    coefA[(size_t)n * C + c] = a;
    coefB[(size_t)n * C + c] = b;
  }
}

template <bool SILU, typename TOut>
__global__ void __launch_bounds__(kMaxThreads) gn_apply_fwd_kernel(const __half* __restrict__ x, int HW, int C, int cvecs, int rows_per_iter,
                                                                    int rows_per_chunk, const float* __restrict__ coefA,
                                                                    const float* __restrict__ coefB, TOut* __restrict__ y) {
  const int n = blockIdx.y, p = blockIdx.x;
  const int col = threadIdx.x % cvecs, r = threadIdx.x / cvecs;
  const int row_end = min(HW, (p + 1) * rows_per_chunk);
  float a[8], b[8];
  pdl_launch_dependents();
  pdl_wait();  // coefficients come from the finalize kernel this one may have been launched ahead of
  load_coef8(coefA + (size_t)n * C + col * 8, a);
  load_coef8(coefB + (size_t)n * C + col * 8, b);
  const size_t off = (size_t)n * HW * C + col * 8;
  const size_t rstride = (size_t)rows_per_iter * C;
  int row = p * rows_per_chunk + r;
  const __half* ptr = x + off + (size_t)row * C;
  TOut* optr = y + off + (size_t)row * C;
  auto apply = [&](const Raw8<__half>& raw, TOut* dst) {
    float v[8];
    raw.get(v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float u = fmaf(a[j], v[j], b[j]);
      v[j] = SILU ? u * sigmoid_fast(u) : u;
    }
    Row8<TOut>::store(dst, v);
  };
  for (; row + (kUnroll - 1) * rows_per_iter < row_end; row += kUnroll * rows_per_iter, ptr += kUnroll * rstride, optr += kUnroll * rstride) {
    Raw8<__half> raw[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) raw[k].load(ptr + k * rstride);
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) apply(raw[k], optr + k * rstride);
  }
  for (; row < row_end; row += rows_per_iter, ptr += rstride, optr += rstride) {
    Raw8<__half> raw;
    raw.load(ptr);
    apply(raw, optr);
  }
}

// ---- backward ----------------------------------------------------------------------------------------------------

template <bool SILU>
__device__ __forceinline__ float dact(float dy, float u) {
  if (!SILU) return dy;
  const float sg = sigmoid_fast(u);
  return dy * sg * fmaf(u, 1.f - sg, 1.f);  // silu'(u) = s(u) * (1 + u*(1 - s(u)))
}

template <bool SILU, typename TDy>
__global__ void __launch_bounds__(kMaxThreads) gn_bwd_partial_kernel(const TDy* __restrict__ dy, const __half* __restrict__ x, int HW, int C, int cvecs,
                                                                      int rows_per_iter, int rows_per_chunk, const float* __restrict__ coefA,
                                                                      const float* __restrict__ coefB, float2* __restrict__ partial) {
  __shared__ __align__(16) float red[2 * 2048];
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y, p = blockIdx.x;
  const int col = threadIdx.x % cvecs, r = threadIdx.x / cvecs;
  const int row_end = min(HW, (p + 1) * rows_per_chunk);
  float a[8], b[8];
  if (SILU) {
    load_coef8(coefA + (size_t)n * C + col * 8, a);
    load_coef8(coefB + (size_t)n * C + col * 8, b);
  }
  const size_t off = (size_t)n * HW * C + col * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  constexpr int U = 2;  // two (dy, x) row pairs in flight: same bytes in flight as the forward's four
  const size_t rstride = (size_t)rows_per_iter * C;
  int row = p * rows_per_chunk + r;
  const __half* xp = x + off + (size_t)row * C;
  const TDy* gp = dy + off + (size_t)row * C;
  auto accumulate = [&](const Raw8<__half>& xr, const Raw8<TDy>& gr) {
    float xv[8], gv[8];
    xr.get(xv);
    gr.get(gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dv = dact<SILU>(gv[j], SILU ? fmaf(a[j], xv[j], b[j]) : 0.f);
      s1[j] += dv;
      s2[j] = fmaf(dv, xv[j], s2[j]);
    }
  };
  for (; row + (U - 1) * rows_per_iter < row_end; row += U * rows_per_iter, xp += U * rstride, gp += U * rstride) {
    Raw8<__half> xr[U];
    Raw8<TDy> gr[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      xr[k].load(xp + k * rstride);
      gr[k].load(gp + k * rstride);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) accumulate(xr[k], gr[k]);
  }
  for (; row < row_end; row += rows_per_iter, xp += rstride, gp += rstride) {
    Raw8<__half> xr;
    Raw8<TDy> gr;
    xr.load(xp);
    gr.load(gp);
    accumulate(xr, gr);
  }
  reduce_rows_and_store(s1, s2, C, cvecs, rows_per_iter, red, partial + ((size_t)n * gridDim.x + p) * C);
}

// with w_c = a_c / rstd (= gamma_c*(1+scale_c)), x' = x + pb_c, x^ = (x'-mean)*rstd and dx^ = w_c*dv:
//   dx = rstd*(dx^ - mean_g(dx^) - x^ * mean_g(dx^ x^)) = a_c*dv + B_g*x' + C_g = a_c*dv + B_g*x + (C_g + B_g*pb_c)
// (the partial sums are over the raw x: sum dv*x' = S2 + pb_c*S1)
__global__ void __launch_bounds__(kFinalizeThreads) gn_finalize_bwd_kernel(const float2* __restrict__ partial, int chunks, int C, int G, int HW,
                                                                            const float* __restrict__ stats, const float* __restrict__ coefA,
                                                                            const float* __restrict__ pre_bias, float* __restrict__ coefBx,
                                                                            float* __restrict__ coefCx) {
  __shared__ double red[64];
  const int n = blockIdx.x / G, g = blockIdx.x % G, cg = C / G;
  pdl_launch_dependents();
  pdl_wait();
  const float mean = stats[2 * blockIdx.x], rstd = stats[2 * blockIdx.x + 1];
  const int total = chunks * cg;
  const float2* base = partial + (size_t)n * chunks * C + g * cg;
  const float* wa = coefA + (size_t)n * C + g * cg;
  const double inv_rstd = 1.0 / (double)rstd;
  double t1 = 0.0, t2 = 0.0;
  for (int idx0 = threadIdx.x; idx0 < total; idx0 += 4 * kFinalizeThreads) {
    float2 v[4];
    float w[4], pb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = idx0 + k * kFinalizeThreads;
      const int p = idx / cg, cc = idx - p * cg;
      const bool ok = idx < total;
      v[k] = ok ? __ldg(base + (size_t)p * C + cc) : make_float2(0.f, 0.f);
      w[k] = ok ? __ldg(wa + cc) : 0.f;
      pb[k] = (ok && pre_bias) ? __ldg(pre_bias + g * cg + cc) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double wk = (double)w[k] * inv_rstd;
      t1 += wk * (double)v[k].x;
      t2 += wk * ((double)v[k].y + (double)pb[k] * (double)v[k].x);
    }
  }
  block_sum2_f64(t1, t2, red);
  const double m = (double)HW * (double)cg;
  const double c1 = t1 / m;
  const double c2 = (double)rstd * (t2 - (double)mean * t1) / m;
  const double Bg = -(double)rstd * (double)rstd * c2;
  const double Cg = -(double)rstd * c1 - Bg * (double)mean;
  for (int cc = threadIdx.x; cc < cg; cc += blockDim.x) {
    const int c = g * cg + cc;
    coefBx[(size_t)n * C + c] = (float)Bg;
    coefCx[(size_t)n * C + c] = (float)(Cg + (pre_bias ? Bg * (double)pre_bias[c] : 0.0));
  }
}

template <bool SILU, typename TDy, bool HAS_RES>
__global__ void __launch_bounds__(kMaxThreads, HAS_RES ? 3 : 4) gn_apply_bwd_kernel(const TDy* __restrict__ dy, const __half* __restrict__ x, int HW, int C, int cvecs,
                                                                    int rows_per_iter, int rows_per_chunk, const float* __restrict__ coefA,
                                                                    const float* __restrict__ coefB, const float* __restrict__ coefBx,
                                                                    const float* __restrict__ coefCx, const __half* __restrict__ dres,
                                                                    __half* __restrict__ dx) {
  const int n = blockIdx.y, p = blockIdx.x;
  const int col = threadIdx.x % cvecs, r = threadIdx.x / cvecs;
  const int row_end = min(HW, (p + 1) * rows_per_chunk);
  float a[8], b[8], bx[8], cx[8];
  pdl_launch_dependents();
  pdl_wait();
  load_coef8(coefA + (size_t)n * C + col * 8, a);
  if (SILU) load_coef8(coefB + (size_t)n * C + col * 8, b);
  load_coef8(coefBx + (size_t)n * C + col * 8, bx);
  load_coef8(coefCx + (size_t)n * C + col * 8, cx);
  const size_t off = (size_t)n * HW * C + col * 8;
  constexpr int U = 2;
  const size_t rstride = (size_t)rows_per_iter * C;
  int row = p * rows_per_chunk + r;
  const __half* xp = x + off + (size_t)row * C;
  const TDy* gp = dy + off + (size_t)row * C;
  __half* op = dx + off + (size_t)row * C;
  // dres: gradient reaching x through its OTHER consumer (the block's skip path), added here instead of in a separate pass
  const ptrdiff_t res_off = HAS_RES ? (dres - dx) : 0;
  auto apply = [&](const Raw8<__half>& xr, const Raw8<TDy>& gr, __half* dst) {
    float xv[8], gv[8], rv[8];
    if (HAS_RES) Row8<__half>::load(dst + res_off, rv);
    xr.get(xv);
    gr.get(gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dv = dact<SILU>(gv[j], SILU ? fmaf(a[j], xv[j], b[j]) : 0.f);
      gv[j] = fmaf(a[j], dv, fmaf(bx[j], xv[j], cx[j]));
      if (HAS_RES) gv[j] += rv[j];
    }
    Row8<__half>::store(dst, gv);
  };
  for (; row + (U - 1) * rows_per_iter < row_end; row += U * rows_per_iter, xp += U * rstride, gp += U * rstride, op += U * rstride) {
    Raw8<__half> xr[U];
    Raw8<TDy> gr[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      xr[k].load(xp + k * rstride);
      gr[k].load(gp + k * rstride);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) apply(xr[k], gr[k], op + k * rstride);
  }
  for (; row < row_end; row += rows_per_iter, xp += rstride, gp += rstride, op += rstride) {
    Raw8<__half> xr;
    Raw8<TDy> gr;
    xr.load(xp);
    gr.load(gp);
    apply(xr, gr, op);
  }
}

// ---- the other element-wise passes between the convolutions ----------------------------------------------------------------------

// out = a + b + bias_c : the ResBlock's `skip(x) + out_conv(h)` with both convolutions' biases deferred into this one pass
__global__ void __launch_bounds__(256) bias_residual_add_kernel(const __half* __restrict__ a, const __half* __restrict__ b, const float* __restrict__ bias,
                                                                 long long nvec, int cvecs, __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8], bb[8];
    Row8<__half>::load(a + i * 8, x);
    Row8<__half>::load(b + i * 8, y);
    load_coef8(bias + (size_t)(i % cvecs) * 8, bb);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = x[j] + y[j] + bb[j];
    Row8<__half>::store(out + i * 8, x);
  }
}

// 2x2 average pooling (UP == false: y[i,j] = scale * sum of the 4 inputs) / nearest 2x upsampling (UP == true: y[i,j] = scale * x[i/2,j/2]).
// The pair is closed under differentiation: d(avg_pool) = up(scale 0.25), d(upsample) = down(scale 1).  fp32 accumulation like ATen.
template <bool UP>
__global__ void __launch_bounds__(256) resample2x_kernel(const __half* __restrict__ x, int Ho, int Wo, int cvecs, long long nvec, float scale,
                                                          __half* __restrict__ y) {
  const int Wi = UP ? Wo / 2 : Wo * 2, Hi = UP ? Ho / 2 : Ho * 2;
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvecs);
    long long pix = i / cvecs;
    const int xo = (int)(pix % Wo);
    pix /= Wo;
    const int yo = (int)(pix % Ho);
    const long long n = pix / Ho;
    float acc[8];
    if (UP) {
      Row8<__half>::load(x + (((n * Hi + yo / 2) * Wi + xo / 2) * cvecs + cv) * 8, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= scale;
    } else {
      const __half* p = x + (((n * Hi + 2 * yo) * Wi + 2 * xo) * cvecs + cv) * 8;
      float v[4][8];
      Row8<__half>::load(p, v[0]);
      Row8<__half>::load(p + (size_t)cvecs * 8, v[1]);
      Row8<__half>::load(p + (size_t)Wi * cvecs * 8, v[2]);
      Row8<__half>::load(p + (size_t)(Wi + 1) * cvecs * 8, v[3]);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = ((v[0][j] + v[1][j]) + (v[2][j] + v[3][j])) * scale;
    }
    Row8<__half>::store(y + i * 8, acc);
  }
}

// torch.cat([a, b], dim=1) on NHWC rows (SPLIT == false: out[r] = a[r] | b[r]) and its gradient (SPLIT == true: the two
// contiguous halves of one row-interleaved tensor).  Pure 128-bit copies; ATen's generic cat / strided-slice copies ran at
// ~1/4 of HBM speed on these shapes.
template <bool SPLIT>
__global__ void __launch_bounds__(256) concat2_kernel(uint4* __restrict__ a, int ca, uint4* __restrict__ b, int cb, long long nvec,
                                                       uint4* __restrict__ cat) {
  const int ct = ca + cb;
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ct;
    const int cv = (int)(i - r * ct);
    uint4* part = cv < ca ? a + r * ca + cv : b + r * cb + (cv - ca);
    if (SPLIT) *part = __ldg(cat + i);
    else cat[i] = __ldg(part);
  }
}

// ---- producers that also emit the NEXT normalisation's chunk partials -----------------------------------------------------------------
// The tensor a ResBlock returns (bias + residual add) and the skip concatenation are exactly what the following block's in_norm
// reduces first.  These variants run in the GroupNorm geometry (CTA = row chunk, thread = channel octet) and, while streaming the
// result out, accumulate sum / sum of squares of the fp16-ROUNDED values they store -- the same numbers gn_stats_partial_kernel
// would read back -- so that normalisation starts at its finalize kernel (one full read of the activation saved per block).
template <bool CONCAT>
__global__ void __launch_bounds__(kMaxThreads) producer_stats_kernel(const __half* __restrict__ a, const __half* __restrict__ b,
                                                                      const float* __restrict__ bias, int ca, int HW, int C, int cvecs,
                                                                      int rows_per_iter, int rows_per_chunk, __half* __restrict__ out,
                                                                      float2* __restrict__ partial) {
  __shared__ __align__(16) float red[2 * 2048];
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y, p = blockIdx.x;
  const int col = threadIdx.x % cvecs, r = threadIdx.x / cvecs;
  const int row_end = min(HW, (p + 1) * rows_per_chunk);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  auto account = [&](const uint4& packed) {
    float v[8];
    unpack8(packed, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += v[j];
      q[j] = fmaf(v[j], v[j], q[j]);
    }
  };
  int row = p * rows_per_chunk + r;
  __half* optr = out + ((size_t)n * HW + row) * C + col * 8;
  const size_t ostride = (size_t)rows_per_iter * C;
  if (CONCAT) {
    // a [N,HW,ca] | b [N,HW,C-ca]: this thread's octet comes from one of them for its whole life
    const bool from_a = col * 8 < ca;
    const int cs = from_a ? ca : C - ca;
    const __half* src = (from_a ? a + col * 8 : b + (col * 8 - ca)) + ((size_t)n * HW + row) * cs;
    const size_t sstride = (size_t)rows_per_iter * cs;
    for (; row + (kUnroll - 1) * rows_per_iter < row_end; row += kUnroll * rows_per_iter, src += kUnroll * sstride, optr += kUnroll * ostride) {
      uint4 raw[kUnroll];
#pragma unroll
      for (int k = 0; k < kUnroll; ++k) raw[k] = ldg_stream(src + k * sstride);
#pragma unroll
      for (int k = 0; k < kUnroll; ++k) {
        *reinterpret_cast<uint4*>(optr + k * ostride) = raw[k];
        account(raw[k]);
      }
    }
    for (; row < row_end; row += rows_per_iter, src += sstride, optr += ostride) {
      const uint4 raw = ldg_stream(src);
      *reinterpret_cast<uint4*>(optr) = raw;
      account(raw);
    }
  } else {
    float bb[8];
    load_coef8(bias + col * 8, bb);
    const __half* ap = a + ((size_t)n * HW + row) * C + col * 8;
    const __half* bp = b + ((size_t)n * HW + row) * C + col * 8;
    constexpr int U = 2;
    auto emit = [&](const uint4& xa, const uint4& xb, __half* dst) {
      float x[8], y[8];
      unpack8(xa, x);
      unpack8(xb, y);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = x[j] + y[j] + bb[j];
      const uint4 packed = pack8(x);
      *reinterpret_cast<uint4*>(dst) = packed;
      account(packed);
    };
    for (; row + (U - 1) * rows_per_iter < row_end; row += U * rows_per_iter, ap += U * ostride, bp += U * ostride, optr += U * ostride) {
      uint4 xa[U], xb[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        xa[k] = ldg_stream(ap + k * ostride);
        xb[k] = ldg_stream(bp + k * ostride);
      }
#pragma unroll
      for (int k = 0; k < U; ++k) emit(xa[k], xb[k], optr + k * ostride);
    }
    for (; row < row_end; row += rows_per_iter, ap += ostride, bp += ostride, optr += ostride) emit(ldg_stream(ap), ldg_stream(bp), optr);
  }
  reduce_rows_and_store(s, q, C, cvecs, rows_per_iter, red, partial + ((size_t)n * gridDim.x + p) * C);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CG_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

// Launch `kernel` as a programmatic dependent of the previous kernel in the stream (plain launch when CG_PDL=0).
template <typename... KArgs, typename... Args>
cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int check_shape(int N, int HW, int C, int G) {
  CG_REQUIRE(N >= 1 && HW >= 1 && C >= 8 && G >= 1, "groupnorm_nhwc: bad sizes N=%d HW=%d C=%d G=%d", N, HW, C, G);
  CG_REQUIRE(C % 8 == 0 && C <= 2048, "groupnorm_nhwc: C=%d must be a multiple of 8 and <= 2048", C);
  CG_REQUIRE(C % G == 0, "groupnorm_nhwc: C=%d not divisible by groups=%d", C, G);
  CG_REQUIRE(N <= 65535, "groupnorm_nhwc: N=%d too large", N);
  return 0;
}

}  // namespace

extern "C" size_t cg_groupnorm_nhwc_workspace_bytes(int N, int HW, int C) {
  if (N < 1 || HW < 1 || C < 8 || C % 8) return 0;
  const Geo g = geometry(N, HW, C);
  // chunk partials (float2 per channel) + the backward's two per-channel coefficient rows
  return (size_t)N * g.chunks * C * sizeof(float2) + (size_t)2 * N * C * sizeof(float);
}

extern "C" int cg_groupnorm_nhwc_geometry(int N, int HW, int C, int* out5) {
  if (int rc = check_shape(N, HW, C, 1)) return rc;
  CG_REQUIRE(out5, "groupnorm_nhwc_geometry: null pointer");
  const Geo g = geometry(N, HW, C);
  out5[0] = g.cvecs;
  out5[1] = g.rows_per_iter;
  out5[2] = g.threads;
  out5[3] = g.chunks;
  out5[4] = g.rows_per_chunk;
  return 0;
}

extern "C" int cg_groupnorm_nhwc_fwd(const void* x, int N, int HW, int C, int G, const float* gamma, const float* beta, const float* scale_shift,
                                     const float* pre_bias, const void* input_partial, float eps, int silu, int out_f32, void* y, float* stats,
                                     float* coef, void* workspace, void* stream) {
  if (int rc = check_shape(N, HW, C, G)) return rc;
  CG_REQUIRE(x && y && gamma && beta && stats && coef && workspace, "groupnorm_nhwc_fwd: null pointer");
  CG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)coef & 15) == 0, "groupnorm_nhwc_fwd: buffers must be 16-byte aligned");
  const Geo g = geometry(N, HW, C);
  cudaStream_t st = cg_stream(stream);
  // input_partial: the chunk partials of x, already written by the kernel that produced x (cg_*_stats_nhwc) -> no statistics pass
  float2* partial = input_partial ? (float2*)input_partial : (float2*)workspace;
  float* coefA = coef;
  float* coefB = coef + (size_t)N * C;
  const dim3 grid(g.chunks, N);
  const __half* xh = (const __half*)x;
  if (!input_partial)
    CG_CUDA(launch_dependent(gn_stats_partial_kernel, grid, dim3(g.threads), st, xh, HW, C, g.cvecs, g.rows_per_iter, g.rows_per_chunk, partial));
  CG_CUDA(launch_dependent(gn_finalize_fwd_kernel, dim3(N * G), dim3(kFinalizeThreads), st, (const float2*)partial, g.chunks, C, G, HW, gamma, beta,
                           scale_shift, pre_bias, eps, stats, coefA, coefB));
#define CG_GN_APPLY(S, T)                                                                                                            \
  CG_CUDA(launch_dependent(gn_apply_fwd_kernel<S, T>, grid, dim3(g.threads), st, xh, HW, C, g.cvecs, g.rows_per_iter, g.rows_per_chunk, \
                           (const float*)coefA, (const float*)coefB, (T*)y))
  if (silu) {
    if (out_f32) CG_GN_APPLY(true, float);
    else CG_GN_APPLY(true, __half);
  } else {
    if (out_f32) CG_GN_APPLY(false, float);
    else CG_GN_APPLY(false, __half);
  }
#undef CG_GN_APPLY
  CG_LAUNCH_CHECK();
  return 0;
}

extern "C" int cg_groupnorm_nhwc_bwd(const void* dy, int dy_f32, const void* x, int N, int HW, int C, int G, const float* stats, const float* coef,
                                     const float* pre_bias, int silu, const void* dres, void* dx, void* workspace, void* stream) {
  if (int rc = check_shape(N, HW, C, G)) return rc;
  CG_REQUIRE(dy && x && stats && coef && dx && workspace, "groupnorm_nhwc_bwd: null pointer");
  CG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0 && ((uintptr_t)workspace & 15) == 0 &&
                 ((uintptr_t)dres & 15) == 0,
             "groupnorm_nhwc_bwd: buffers must be 16-byte aligned");
  const Geo g = geometry(N, HW, C);
  cudaStream_t st = cg_stream(stream);
  float2* partial = (float2*)workspace;
  float* coefBx = (float*)(partial + (size_t)N * g.chunks * C);
  float* coefCx = coefBx + (size_t)N * C;
  const float* coefA = coef;
  const float* coefB = coef + (size_t)N * C;
  const dim3 grid(g.chunks, N);
  const __half* xh = (const __half*)x;
#define CG_GN_BWD(S, T)                                                                                                                        \
  do {                                                                                                                                         \
    CG_CUDA(launch_dependent(gn_bwd_partial_kernel<S, T>, grid, dim3(g.threads), st, (const T*)dy, xh, HW, C, g.cvecs, g.rows_per_iter,          \
                             g.rows_per_chunk, coefA, coefB, partial));                                                                        \
    CG_CUDA(launch_dependent(gn_finalize_bwd_kernel, dim3(N * G), dim3(kFinalizeThreads), st, (const float2*)partial, g.chunks, C, G, HW, stats, \
                             coefA, pre_bias, coefBx, coefCx));                                                                                \
    if (dres)                                                                                                                                  \
      CG_CUDA(launch_dependent(gn_apply_bwd_kernel<S, T, true>, grid, dim3(g.threads), st, (const T*)dy, xh, HW, C, g.cvecs, g.rows_per_iter,  \
                               g.rows_per_chunk, coefA, coefB, (const float*)coefBx, (const float*)coefCx, (const __half*)dres, (__half*)dx)); \
    else                                                                                                                                       \
      CG_CUDA(launch_dependent(gn_apply_bwd_kernel<S, T, false>, grid, dim3(g.threads), st, (const T*)dy, xh, HW, C, g.cvecs, g.rows_per_iter, \
                               g.rows_per_chunk, coefA, coefB, (const float*)coefBx, (const float*)coefCx, (const __half*)nullptr,             \
                               (__half*)dx));                                                                                                  \
  } while (0)
  if (silu) {
    if (dy_f32) CG_GN_BWD(true, float);
    else CG_GN_BWD(true, __half);
  } else {
    if (dy_f32) CG_GN_BWD(false, float);
    else CG_GN_BWD(false, __half);
  }
#undef CG_GN_BWD
  return 0;
}

extern "C" int cg_bias_residual_add_nhwc(const void* a, const void* b, const float* bias, int64_t rows, int C, void* out, void* stream) {
  CG_REQUIRE(a && b && bias && out, "bias_residual_add_nhwc: null pointer");
  CG_REQUIRE(rows >= 1 && C >= 8 && C % 8 == 0, "bias_residual_add_nhwc: rows=%lld C=%d (C must be a multiple of 8)", (long long)rows, C);
  CG_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)bias & 15) == 0,
             "bias_residual_add_nhwc: buffers must be 16-byte aligned");
  const long long nvec = (long long)rows * (C / 8);
  long long blocks = (nvec + 255) / 256;
  if (blocks > 16 * CG_NUM_SMS) blocks = 16 * CG_NUM_SMS;
  CG_CUDA(launch_dependent(bias_residual_add_kernel, dim3((unsigned)blocks), dim3(256), cg_stream(stream), (const __half*)a, (const __half*)b, bias, nvec,
                           C / 8, (__half*)out));
  return 0;
}

extern "C" int cg_resample2x_nhwc(const void* x, int N, int H, int W, int C, int up, float scale, void* y, void* stream) {
  CG_REQUIRE(x && y, "resample2x_nhwc: null pointer");
  CG_REQUIRE(N >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "resample2x_nhwc: bad sizes N=%d H=%d W=%d C=%d (C must be a multiple of 8)", N, H, W, C);
  CG_REQUIRE(up || (H % 2 == 0 && W % 2 == 0), "resample2x_nhwc: 2x2 average pooling needs even H, W (got %d x %d)", H, W);
  CG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "resample2x_nhwc: buffers must be 16-byte aligned");
  const int Ho = up ? 2 * H : H / 2, Wo = up ? 2 * W : W / 2;
  const long long nvec = (long long)N * Ho * Wo * (C / 8);
  long long blocks = (nvec + 255) / 256;
  if (blocks > 16 * CG_NUM_SMS) blocks = 16 * CG_NUM_SMS;
  if (up)
    CG_CUDA(launch_dependent(resample2x_kernel<true>, dim3((unsigned)blocks), dim3(256), cg_stream(stream), (const __half*)x, Ho, Wo, C / 8, nvec, scale,
                             (__half*)y));
  else
    CG_CUDA(launch_dependent(resample2x_kernel<false>, dim3((unsigned)blocks), dim3(256), cg_stream(stream), (const __half*)x, Ho, Wo, C / 8, nvec, scale,
                             (__half*)y));
  return 0;
}

extern "C" int cg_concat2_nhwc(void* a, int Ca, void* b, int Cb, int64_t rows, void* cat, int split, void* stream) {
  CG_REQUIRE(a && b && cat, "concat2_nhwc: null pointer");
  CG_REQUIRE(rows >= 1 && Ca >= 8 && Cb >= 8 && Ca % 8 == 0 && Cb % 8 == 0, "concat2_nhwc: rows=%lld Ca=%d Cb=%d (channel counts must be multiples of 8)",
             (long long)rows, Ca, Cb);
  CG_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)cat & 15) == 0, "concat2_nhwc: buffers must be 16-byte aligned");
  const long long nvec = (long long)rows * ((Ca + Cb) / 8);
  long long blocks = (nvec + 255) / 256;
  if (blocks > 16 * CG_NUM_SMS) blocks = 16 * CG_NUM_SMS;
  if (split)
    CG_CUDA(launch_dependent(concat2_kernel<true>, dim3((unsigned)blocks), dim3(256), cg_stream(stream), (uint4*)a, Ca / 8, (uint4*)b, Cb / 8, nvec,
                             (uint4*)cat));
  else
    CG_CUDA(launch_dependent(concat2_kernel<false>, dim3((unsigned)blocks), dim3(256), cg_stream(stream), (uint4*)a, Ca / 8, (uint4*)b, Cb / 8, nvec,
                             (uint4*)cat));
  return 0;
}

extern "C" int cg_bias_residual_add_stats_nhwc(const void* a, const void* b, const float* bias, int N, int HW, int C, void* out, void* partial,
                                               void* stream) {
  if (int rc = check_shape(N, HW, C, 1)) return rc;
  CG_REQUIRE(a && b && bias && out && partial, "bias_residual_add_stats_nhwc: null pointer");
  CG_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)bias & 15) == 0 &&
                 ((uintptr_t)partial & 15) == 0,
             "bias_residual_add_stats_nhwc: buffers must be 16-byte aligned");
  const Geo g = geometry(N, HW, C);
  CG_CUDA(launch_dependent(producer_stats_kernel<false>, dim3(g.chunks, N), dim3(g.threads), cg_stream(stream), (const __half*)a, (const __half*)b, bias, 0,
                           HW, C, g.cvecs, g.rows_per_iter, g.rows_per_chunk, (__half*)out, (float2*)partial));
  return 0;
}

extern "C" int cg_concat2_stats_nhwc(const void* a, int Ca, const void* b, int Cb, int N, int HW, void* cat, void* partial, void* stream) {
  CG_REQUIRE(Ca >= 8 && Cb >= 8 && Ca % 8 == 0 && Cb % 8 == 0, "concat2_stats_nhwc: Ca=%d Cb=%d (channel counts must be multiples of 8)", Ca, Cb);
  if (int rc = check_shape(N, HW, Ca + Cb, 1)) return rc;
  CG_REQUIRE(a && b && cat && partial, "concat2_stats_nhwc: null pointer");
  CG_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)cat & 15) == 0 && ((uintptr_t)partial & 15) == 0,
             "concat2_stats_nhwc: buffers must be 16-byte aligned");
  const int C = Ca + Cb;
  const Geo g = geometry(N, HW, C);
  CG_CUDA(launch_dependent(producer_stats_kernel<true>, dim3(g.chunks, N), dim3(g.threads), cg_stream(stream), (const __half*)a, (const __half*)b,
                           (const float*)nullptr, Ca, HW, C, g.cvecs, g.rows_per_iter, g.rows_per_chunk, (__half*)cat, (float2*)partial));
  return 0;
}
