// Dynamic thresholding of the sampler's x0 prediction (clip_diffusion/sample.py:116-132):
//   thr = max(quantile(|x|, q), 1);  out = clamp(x, -thr, thr) / thr          per sample.
// torch.quantile sorts the 786k-1.8M values every step; here the order statistic comes from a 3-level radix select on
// the float bit patterns (|x| >= 0 => unsigned order == float order): 3 histogram passes + 1 count/min pass + 1 apply
// pass over an L2-resident tensor.  Interpolation ("linear") follows torch: rank = q*(n-1) in fp32, lerp(v_lo, v_hi, frac).
#include "common.cuh"

namespace {

constexpr int BINS = 2048;  // 11 + 11 + 10 bits
struct SelState {            // per sample, in workspace after the histograms
  unsigned int prefix[3];    // selected high bits after level 0 / 1 (slot L is written by the level-L kernel, read by L+1:
  unsigned int rank[3];      // residual rank inside the selected bin          no kernel reads a slot it writes)
  unsigned int count_le;     // #(|x| <= v_lo)
  unsigned int next_bits;    // min{|x| > v_lo} as bits (0xFFFFFFFF if none)
};

__device__ __forceinline__ unsigned int abs_bits(float v) { return __float_as_uint(v) & 0x7FFFFFFFu; }

// Block-parallel search of the bin holding `rank` in a 2048-bin histogram; every thread returns (bin, rank inside it).
__device__ __forceinline__ void locate(const unsigned int* __restrict__ hist, unsigned int rank, unsigned int* bin, unsigned int* inner, unsigned int* sh) {
  const int tid = threadIdx.x;  // 256 threads x 8 bins
  unsigned int local[8], sum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { local[i] = hist[tid * 8 + i]; sum += local[i]; }
  sh[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    unsigned int run = 0;
    int t = 0;
    for (; t < 256; ++t) { if (run + sh[t] > rank) break; run += sh[t]; }
    sh[256] = (unsigned int)t; sh[257] = run;
  }
  __syncthreads();
  const int owner = (int)sh[256];
  if (tid == owner) {
    unsigned int run = sh[257];
    int i = 0;
    for (; i < 8; ++i) { if (run + local[i] > rank) break; run += local[i]; }
    sh[258] = (unsigned int)(owner * 8 + i); sh[259] = rank - run;
  }
  __syncthreads();
  *bin = sh[258]; *inner = sh[259];
  __syncthreads();
}

// LEVEL 0: bits [31:21] (sign cleared => top bit 0, 11 bits [30:20] used: shift 20), LEVEL 1: [19:9], LEVEL 2: [8:0] (512 bins)
template <int LEVEL>
__global__ void __launch_bounds__(256) hist_kernel(const float* __restrict__ x, long long n, unsigned int rank0, unsigned int* __restrict__ hists,
                                                   SelState* __restrict__ states) {
  __shared__ unsigned int sh_hist[BINS];
  __shared__ unsigned int sh[260];
  const int b = blockIdx.y;
  unsigned int* hist = hists + ((size_t)b * 3 + LEVEL) * BINS;
  unsigned int prefix = 0;
  if (LEVEL > 0) {
    // finish the previous level: which bin held the rank?
    const unsigned int* prev = hists + ((size_t)b * 3 + LEVEL - 1) * BINS;
    const unsigned int prev_rank = LEVEL == 1 ? rank0 : states[b].rank[LEVEL - 1];
    const unsigned int prev_prefix = LEVEL == 1 ? 0u : states[b].prefix[LEVEL - 1];
    unsigned int bin, inner;
    locate(prev, prev_rank, &bin, &inner, sh);
    prefix = LEVEL == 1 ? (bin << 20) : (prev_prefix | (bin << 9));
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      states[b].rank[LEVEL] = inner; states[b].prefix[LEVEL] = prefix;  // for the NEXT kernel
    }
  }
  for (int i = threadIdx.x; i < BINS; i += blockDim.x) sh_hist[i] = 0;
  __syncthreads();
  const float* xb = x + (size_t)b * n;
  const unsigned int mask = LEVEL == 0 ? 0u : (LEVEL == 1 ? 0x7FF00000u : 0x7FFFFE00u);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int u = abs_bits(xb[i]);
    if ((u & mask) == prefix) {
      const unsigned int bin = LEVEL == 0 ? (u >> 20) : (LEVEL == 1 ? ((u >> 9) & 0x7FFu) : (u & 0x1FFu));
      atomicAdd(&sh_hist[bin], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BINS; i += blockDim.x)
    if (sh_hist[i]) atomicAdd(&hist[i], sh_hist[i]);
}

// after level 2: exact bits of v_lo; count elements <= v_lo and the smallest element above it
__global__ void __launch_bounds__(256) count_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ hists, SelState* __restrict__ states,
                                                    unsigned int* __restrict__ vlo_bits_out) {
  __shared__ unsigned int sh[260];
  __shared__ unsigned int s_cnt, s_min;
  const int b = blockIdx.y;
  unsigned int bin, inner;
  locate(hists + ((size_t)b * 3 + 2) * BINS, states[b].rank[2], &bin, &inner, sh);
  const unsigned int vbits = states[b].prefix[2] | bin;
  if (threadIdx.x == 0) { s_cnt = 0; s_min = 0xFFFFFFFFu; }
  __syncthreads();
  const float* xb = x + (size_t)b * n;
  unsigned int cnt = 0, mn = 0xFFFFFFFFu;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int u = abs_bits(xb[i]);
    if (u <= vbits) ++cnt; else mn = min(mn, u);
  }
  atomicAdd(&s_cnt, cnt); atomicMin(&s_min, mn);
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&states[b].count_le, s_cnt);
    atomicMin(&states[b].next_bits, s_min);
    if (blockIdx.x == 0) vlo_bits_out[b] = vbits;
  }
}

__device__ __forceinline__ float torch_lerp(float a, float b, float w) {
  return w < 0.5f ? a + w * (b - a) : b - (b - a) * (1.f - w);
}

__global__ void __launch_bounds__(256) apply_kernel(const float* __restrict__ x, long long n, unsigned int lo, float frac, float min_thr,
                                                    const SelState* __restrict__ states, const unsigned int* __restrict__ vlo_bits, float* __restrict__ out,
                                                    float* __restrict__ thr_out) {
  const int b = blockIdx.y;
  const float vlo = __uint_as_float(vlo_bits[b]);
  // v_hi = element of rank lo+1: equal to v_lo if at least lo+2 elements are <= v_lo, else the next larger value
  const float vhi = (states[b].count_le >= lo + 2u || states[b].next_bits == 0xFFFFFFFFu) ? vlo : __uint_as_float(states[b].next_bits);
  const float q = frac > 0.f ? torch_lerp(vlo, vhi, frac) : vlo;
  const float thr = fmaxf(q, min_thr);
  if (blockIdx.x == 0 && threadIdx.x == 0 && thr_out) thr_out[b] = thr;
  const float* xb = x + (size_t)b * n;
  float* ob = out + (size_t)b * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    ob[i] = fminf(fmaxf(xb[i], -thr), thr) / thr;
}

}  // namespace

extern "C" size_t cg_dynamic_threshold_workspace_bytes(int B) {
  return B <= 0 ? 0 : (size_t)B * (3 * BINS * sizeof(unsigned int) + sizeof(SelState) + sizeof(unsigned int)) + 256;
}

extern "C" int cg_dynamic_threshold(const float* x, int B, int64_t n, float q, float min_thr, float* out, float* thr_out, void* workspace,
                                    void* stream) {
  CG_REQUIRE(x && out && workspace && B > 0 && n > 0 && n < (1LL << 31), "cg_dynamic_threshold: bad arguments");
  CG_REQUIRE(q >= 0.f && q <= 1.f, "cg_dynamic_threshold: q=%f outside [0,1]", (double)q);
  cudaStream_t s = cg_stream(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  unsigned int* hists = reinterpret_cast<unsigned int*>(ws);
  SelState* states = reinterpret_cast<SelState*>(ws + (size_t)B * 3 * BINS * sizeof(unsigned int));
  unsigned int* vlo = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(states) + (size_t)B * sizeof(SelState));
  CG_CUDA(cudaMemsetAsync(ws, 0, (size_t)B * 3 * BINS * sizeof(unsigned int) + (size_t)B * sizeof(SelState), s));
  // next_bits must start at 0xFFFFFFFF: set by a tiny strided memset
  CG_CUDA(cudaMemset2DAsync(reinterpret_cast<char*>(states) + offsetof(SelState, next_bits), sizeof(SelState), 0xFF, sizeof(unsigned int), B, s));
  // torch.quantile: ranks = q * (n - 1) in the input dtype (fp32); below = floor; frac = ranks - below
  const float rank_f = q * (float)(n - 1);
  const float below = floorf(rank_f);
  const unsigned int lo = (unsigned int)below;
  const float frac = rank_f - below;
  int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (blocks > CG_NUM_SMS * 4) blocks = CG_NUM_SMS * 4;
  if (blocks < 1) blocks = 1;
  dim3 grid(blocks, B);
  hist_kernel<0><<<grid, 256, 0, s>>>(x, n, lo, hists, states);
  CG_LAUNCH_CHECK();
  hist_kernel<1><<<grid, 256, 0, s>>>(x, n, lo, hists, states);
  CG_LAUNCH_CHECK();
  hist_kernel<2><<<grid, 256, 0, s>>>(x, n, lo, hists, states);
  CG_LAUNCH_CHECK();
  count_kernel<<<grid, 256, 0, s>>>(x, n, hists, states, vlo);
  CG_LAUNCH_CHECK();
  apply_kernel<<<grid, 256, 0, s>>>(x, n, lo, frac, min_thr, states, vlo, out, thr_out);
  CG_LAUNCH_CHECK();
  return 0;
}
