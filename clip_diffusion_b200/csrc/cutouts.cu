// Fused Disco-style cutouts (clip_diffusion/cutouts.py:47-134) + CLIP_NORMALIZE (utils/functional.py:16-18)
// forward and backward for sm_100a.
//
// Forward  : tables -> resample (separable antialiased cubic, H then W) -> augment (flip, noise, affine,
//            noise, gray, noise) -> colour jitter + normalise + layout epilogue.
// Backward : jitter/normalise backward (two passes because of the contrast mean) -> affine backward (gather)
//            -> transposed resample as a GATHER over source pixels: every pixel of d(loss)/d(x_in) is written
//            exactly once, no atomics, deterministic ("fused scatter-add" in gather form).
//
// All the randomness arrives as plain numbers (cg_cut_t / cg_aug_t) drawn on the host in the reference's
// order (clip_diffusion_b200/rng_record.py), which is what makes crop sizes/offsets bit-exact.
#include <string.h>
#include <vector>
#include "common.cuh"

namespace {

constexpr int TAPS_MAX = 32;  // taps per output = ceil(4*size/cs): downscales up to 8x (1792 -> 224, 2048 -> 336 ...)
constexpr int TT = 24;        // ELL row width of the transposed table: outputs touching one source pixel <= ceil(4 * max(cs/size, 1)) + 2
                              // (6 when downsampling; an image SMALLER than the cut size is upsampled -- min_size = min(W, H, cut_size),
                              // cutouts.py:52,84-86 -- up to 336/64 = 5.25x for the sizes Config.update allows)
constexpr int RT = 8;         // output rows per resample CTA
constexpr int MAX_SIZE = 2048;
constexpr float GW0 = 0.2989f, GW1 = 0.587f, GW2 = 0.114f;  // torchvision rgb_to_grayscale

struct WsLayout {
  // metadata block (filled by ONE host-to-device copy): header | cuts[N] | aug | uidx[N] | ucuts[N] | dupoff[N+1] | duplist[N]
  size_t cuts, aug, uidx, ucuts, dupoff, duplist, meta_end;
  size_t taps, left, wfw, tstart, wtr, base, z, scratch, partial, spartial, total;
  int nblk;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

WsLayout ws_layout(int N, int cs, int max_size) {
  WsLayout L;
  size_t o = 256;  // header
  L.cuts = o; o = align_up(o + sizeof(cg_cut_t) * N, 256);
  L.aug = o; o = align_up(o + sizeof(cg_aug_t), 256);
  L.uidx = o; o = align_up(o + sizeof(int) * N, 256);
  L.ucuts = o; o = align_up(o + sizeof(cg_cut_t) * N, 256);
  L.dupoff = o; o = align_up(o + sizeof(int) * (N + 1), 256);
  L.duplist = o; o = align_up(o + sizeof(int) * N, 256);
  L.meta_end = o;
  // per UNIQUE crop geometry (<= N): resize tables and the base (resampled) cutout
  L.taps = o; o = align_up(o + sizeof(int) * N, 256);
  L.left = o; o = align_up(o + sizeof(int) * (size_t)N * cs, 256);
  L.wfw = o; o = align_up(o + sizeof(float) * (size_t)N * cs * TAPS_MAX, 256);
  L.tstart = o; o = align_up(o + sizeof(int) * (size_t)N * max_size, 256);
  L.wtr = o; o = align_up(o + sizeof(float) * (size_t)N * max_size * TT, 256);
  const size_t img = sizeof(float) * (size_t)N * 3 * cs * cs;
  L.base = o; o = align_up(o + img, 256);
  L.z = o; o = align_up(o + img, 256);
  L.scratch = o; o = align_up(o + img, 256);
  L.nblk = (cs * cs + 255) / 256;
  L.partial = o; o = align_up(o + sizeof(float) * (size_t)N * L.nblk, 256);
  L.spartial = o; o = align_up(o + sizeof(float) * (size_t)N * L.nblk, 256);
  L.total = o;
  return L;
}

struct WsHeader {
  int magic, N, cs, max_size, H, W, U, input01;  // U = number of unique crop geometries
  int tt;                                        // ELL entries in use (6 when every crop is downsampled)
};

// ------------------------------------------------------------------------------------------------
// Resize tables (ResizeRight, antialiased cubic, zero "constant" padding).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cubic_w(float x) {
  const float ax = fabsf(x);
  const float ax2 = __fmul_rn(ax, ax);
  const float ax3 = __fmul_rn(ax2, ax);
  float r = 0.f;
  if (ax <= 1.f) r = __fadd_rn(__fsub_rn(__fmul_rn(1.5f, ax3), __fmul_rn(2.5f, ax2)), 1.f);
  else if (ax <= 2.f) r = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(-0.5f, ax3), __fmul_rn(2.5f, ax2)), __fmul_rn(4.f, ax)), 2.f);
  return r;
}

// one block per cutout
__global__ void __launch_bounds__(256) tables_kernel(const cg_cut_t* __restrict__ cuts, int N, int cs, int max_size,
                                                     int* __restrict__ taps_out, int* __restrict__ left, float* __restrict__ wfw,
                                                     int* __restrict__ tstart, float* __restrict__ wtr) {
  extern __shared__ int s_left[];  // cs ints
  const int n = blockIdx.x;
  if (n >= N) return;
  const int size = cuts[n].size;
  const double scale_d = (double)cs / (double)size;
  const float scale = (float)scale_d;
  const bool down = scale_d < 1.0;
  const double support_d = down ? 4.0 / scale_d : 4.0;
  const float eps = 1.1920928955078125e-07f;
  int taps = (int)ceil(support_d - (double)eps);
  if (size == cs) taps = 1;  // scale == 1: the dimension is skipped by ResizeRight (identity)
  if (threadIdx.x == 0) taps_out[n] = taps;
  const float a = (float)((size - 1) / 2.0);
  const float b = (float)((cs - 1) / (2.0 * scale_d));
  const float half_sup = (float)(support_d / 2.0);
  // ResizeRight shifts grid and field of view by the left pad (= -left[0]) BEFORE taking their difference,
  // which re-rounds the fp32 grid; reproduce it so the weights agree to the last bits.
  const float g0 = __fsub_rn(__fadd_rn(__fdiv_rn(0.f, scale), a), b);
  const int pad0 = -(int)ceilf(__fsub_rn(__fsub_rn(g0, half_sup), eps));
  int* lf = left + (size_t)n * cs;
  float* wf = wfw + (size_t)n * cs * TAPS_MAX;
  for (int o = threadIdx.x; o < cs; o += blockDim.x) {
    float* w = wf + (size_t)o * TAPS_MAX;
    if (size == cs) {
      lf[o] = o; s_left[o] = o;
      w[0] = 1.f;
      for (int t = 1; t < TAPS_MAX; ++t) w[t] = 0.f;
      continue;
    }
    float g = __fdiv_rn((float)o, scale);
    g = __fadd_rn(g, a);
    g = __fsub_rn(g, b);
    const int l = (int)ceilf(__fsub_rn(__fsub_rn(g, half_sup), eps));
    lf[o] = l; s_left[o] = l;
    float sum = 0.f;
    for (int t = 0; t < TAPS_MAX; ++t) {
      float v = 0.f;
      if (t < taps) {
        const float d = __fsub_rn(__fadd_rn(g, (float)pad0), (float)(l + t + pad0));
        v = down ? __fmul_rn(scale, cubic_w(__fmul_rn(scale, d))) : cubic_w(d);
      }
      w[t] = v;
      sum = __fadd_rn(sum, v);
    }
    if (sum == 0.f) sum = 1.f;
    for (int t = 0; t < taps; ++t) w[t] = __fdiv_rn(w[t], sum);
  }
  __syncthreads();
  // transposed (ELL) table: for every source coordinate l, the first output that reads it and <= TT weights
  int* ts = tstart + (size_t)n * max_size;
  float* wt = wtr + (size_t)n * max_size * TT;
  for (int l = threadIdx.x; l < size; l += blockDim.x) {
    // first o with left[o] + taps > l   (left is non-decreasing)
    int lo = 0, hi = cs;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_left[mid] + taps > l) hi = mid; else lo = mid + 1;
    }
    ts[l] = lo;
    for (int j = 0; j < TT; ++j) {
      const int o = lo + j;
      float v = 0.f;
      if (o < cs) {
        const int t = l - s_left[o];
        if (t >= 0 && t < taps) v = wf[(size_t)o * TAPS_MAX + t];
      }
      wt[(size_t)l * TT + j] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// F1: resample.  grid (ceil(cs/RT), N), 256 threads.  Vertical pass (global -> smem), horizontal pass
// (smem -> base).  x_in is tiny (3 MB at 512^2) and stays L2 resident; reads are coalesced along x.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resample_fwd_kernel(const float* __restrict__ x_in, int H, int W,
                                                           const cg_cut_t* __restrict__ cuts, int U, const int* __restrict__ tapsN,
                                                           const int* __restrict__ left, const float* __restrict__ wfw, int cs,
                                                           int input01, int tmp_ld, float* __restrict__ base) {
  extern __shared__ float tmp_dyn[];  // [RT][tmp_ld]: the vertically resampled rows of this CTA (tmp_ld = largest crop size of the call)
  __shared__ float s_w[RT][TAPS_MAX];
  __shared__ int s_l[RT];
  const float ka = input01 ? 0.f : 1.f, km = input01 ? 1.f : 0.5f;  // (x + ka) * km
  const int n = blockIdx.y;  // unique crop index
  if (n >= U) return;
  const int r0 = blockIdx.x * RT;
  const cg_cut_t cut = cuts[n];
  const int size = cut.size, taps = tapsN[n];
  const int* lf = left + (size_t)n * cs;
  const float* wf = wfw + (size_t)n * cs * TAPS_MAX;
  const bool gray_pre = cut.flags & CG_CUT_GRAY_PRE;
  const bool gray_post = cut.flags & CG_CUT_GRAY_POST;
  const bool hflip = cut.flags & CG_CUT_HFLIP;
  const size_t plane = (size_t)H * W;
  float* bn = base + (size_t)n * 3 * cs * cs;
  // weights / first tap of this CTA's RT output rows
  for (int i = threadIdx.x; i < RT * TAPS_MAX; i += blockDim.x) {
    const int r = i / TAPS_MAX, t = i - r * TAPS_MAX;
    const int oy = r0 + r;
    s_w[r][t] = (oy < cs && t < taps) ? wf[(size_t)oy * TAPS_MAX + t] : 0.f;
    if (t == 0) s_l[r] = oy < cs ? lf[oy] : 0;
  }
  __syncthreads();
  const int row_lo = s_l[0];
  const int last_r = min(RT, cs - r0) - 1;
  const int row_hi = s_l[last_r] + taps;  // exclusive, crop coordinates
  const int nch = gray_pre ? 1 : 3;
  for (int ch = 0; ch < nch; ++ch) {
    // vertical pass: one thread owns a crop column, walks the source rows ONCE (coalesced across the warp) and
    // feeds the <= RT output-row accumulators that tap each row
    for (int lx = threadIdx.x; lx < size; lx += blockDim.x) {
      float acc[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) acc[r] = 0.f;
      const int sx = cut.x0 + lx;
      if (sx >= 0 && sx < W) {
        for (int ly = max(row_lo, 0); ly < min(row_hi, size); ++ly) {
          const int sy = cut.y0 + ly;
          if (sy < 0 || sy >= H) continue;
          const size_t off = (size_t)sy * W + sx;
          float v;
          if (gray_pre) {
            const float rr = (__ldg(x_in + off) + ka) * km;
            const float gg = (__ldg(x_in + plane + off) + ka) * km;
            const float bb = (__ldg(x_in + 2 * plane + off) + ka) * km;
            v = __fadd_rn(__fadd_rn(__fmul_rn(GW0, rr), __fmul_rn(GW1, gg)), __fmul_rn(GW2, bb));
          } else {
            v = (__ldg(x_in + ch * plane + off) + ka) * km;
          }
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            const int t = ly - s_l[r];
            if (t >= 0 && t < taps) acc[r] = fmaf(s_w[r][t], v, acc[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < RT; ++r) tmp_dyn[r * tmp_ld + lx] = acc[r];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < RT * cs; idx += blockDim.x) {
      const int r = idx / cs, ox = idx - r * cs;
      const int oy = r0 + r;
      if (oy >= cs) continue;
      const int lx0 = lf[ox];
      const float* w = wf + (size_t)ox * TAPS_MAX;
      float acc = 0.f;
      for (int t = 0; t < taps; ++t) {
        const int lx = lx0 + t;
        if (lx >= 0 && lx < size) acc = fmaf(w[t], tmp_dyn[r * tmp_ld + lx], acc);
      }
      const int oxs = hflip ? cs - 1 - ox : ox;
      const size_t o = (size_t)oy * cs + oxs;
      if (gray_pre) {
        bn[o] = acc; bn[(size_t)cs * cs + o] = acc; bn[2 * (size_t)cs * cs + o] = acc;
      } else {
        bn[(size_t)ch * cs * cs + o] = acc;
      }
    }
    __syncthreads();
  }
  if (gray_post) {
    for (int idx = threadIdx.x; idx < RT * cs; idx += blockDim.x) {
      const int r = idx / cs, ox = idx - r * cs;
      const int oy = r0 + r;
      if (oy >= cs) continue;
      const size_t o = (size_t)oy * cs + ox;
      const float g = __fadd_rn(__fadd_rn(__fmul_rn(GW0, bn[o]), __fmul_rn(GW1, bn[(size_t)cs * cs + o])),
                                __fmul_rn(GW2, bn[2 * (size_t)cs * cs + o]));
      bn[o] = g; bn[(size_t)cs * cs + o] = g; bn[2 * (size_t)cs * cs + o] = g;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Noise: explicit tensors (parity mode) or counter-based Philox4x32-10 + Box-Muller (one call -> 3 channels).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0; k.y += W1;
  }
  return c;
}

struct NoiseSrc {
  const float* ptr;  // [3][N,3,cs,cs] or nullptr
  uint64_t seed, cut0;
  int N, cs;
  float std;
  int mode;            // 1: torch's CUDA randn stream (cg_aug_t.noise_mode)
  uint32_t nthreads;   // grid * 256 of torch's launch for the [N_total,3,cs,cs] tensor
  uint64_t offset[3];  // generator offset at each of the three randn_like calls
};

// Element `li` of torch.randn on CUDA (float32): ATen's distribution_nullary_kernel gives thread idx = li % nthreads its k-th
// curand_normal4 call for li / nthreads = 4 k + ii; curand_init(seed, subsequence = idx, offset) puts offset / 4 (+ k per call) in the
// low and idx in the high half of the 128-bit Philox counter; curand_normal4 = two Box-Muller pairs (x = s sin, y = s cos).
__device__ __forceinline__ float torch_randn_elem(uint64_t seed, uint64_t offset, uint32_t nthreads, uint64_t li) {
  const uint64_t q = li / nthreads;
  const uint32_t idx = (uint32_t)(li - q * nthreads);
  const uint64_t c = (offset >> 2) + (q >> 2);
  const int ii = (int)(q & 3);
  const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), idx, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const unsigned int xx = ii < 2 ? r.x : r.z, yy = ii < 2 ? r.y : r.w;
  // _curand_box_muller (curand_normal.h), same expressions so that the compiler contracts them the same way
  const float CURAND_2POW32_INV_ = 2.3283064e-10f, CURAND_2POW32_INV_2PI_ = 2.3283064e-10f * 6.2831855f;
  const float u = xx * CURAND_2POW32_INV_ + (CURAND_2POW32_INV_ / 2);
  const float v = yy * CURAND_2POW32_INV_2PI_ + (CURAND_2POW32_INV_2PI_ / 2);
  const float s = sqrtf(-2.0f * logf(u));
  float sn, cs;
  __sincosf(v, &sn, &cs);
  return (ii & 1) ? s * cs : s * sn;
}

__global__ void __launch_bounds__(256) randn_like_torch_kernel(float* __restrict__ out, int64_t numel, uint64_t seed, uint64_t offset, uint32_t nthreads) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = torch_randn_elem(seed, offset, nthreads, (uint64_t)i);
}

__device__ __forceinline__ void noise3(const NoiseSrc& ns, int stage, int n, int y, int x, float out[3]) {
  const size_t pp = (size_t)ns.cs * ns.cs;
  if (ns.ptr) {
    const float* p = ns.ptr + ((size_t)stage * ns.N + n) * 3 * pp + (size_t)y * ns.cs + x;
    out[0] = ns.std * p[0]; out[1] = ns.std * p[pp]; out[2] = ns.std * p[2 * pp];
    return;
  }
  if (ns.mode == 1) {
    const uint64_t li = (ns.cut0 + (uint64_t)n) * 3 * pp + (uint64_t)y * ns.cs + x;  // element (n, 0, y, x) of the whole [N_total,3,cs,cs] batch
    out[0] = ns.std * torch_randn_elem(ns.seed, ns.offset[stage], ns.nthreads, li);
    out[1] = ns.std * torch_randn_elem(ns.seed, ns.offset[stage], ns.nthreads, li + pp);
    out[2] = ns.std * torch_randn_elem(ns.seed, ns.offset[stage], ns.nthreads, li + 2 * pp);
    return;
  }
  const uint64_t e = (ns.cut0 + (uint64_t)n) * pp + (uint64_t)y * ns.cs + x;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)stage, 0x636c6970u),
                                make_uint2((uint32_t)ns.seed, (uint32_t)(ns.seed >> 32)));
  const float u1 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f, u2 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
  const float u3 = ((float)r.z + 0.5f) * 2.3283064365386963e-10f, u4 = ((float)r.w + 0.5f) * 2.3283064365386963e-10f;
  const float ra = sqrtf(-2.f * __logf(u1)), rb = sqrtf(-2.f * __logf(u3));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  out[0] = ns.std * ra * c; out[1] = ns.std * ra * s;
  out[2] = ns.std * rb * __cosf(6.283185307179586f * u4);
}

// ------------------------------------------------------------------------------------------------
// Colour jitter (torchvision ColorJitter on float tensors, _functional_tensor.py:171-345), per pixel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ float gray_of(const float v[3]) {
  return __fadd_rn(__fadd_rn(__fmul_rn(GW0, v[0]), __fmul_rn(GW1, v[1])), __fmul_rn(GW2, v[2]));
}
__device__ __forceinline__ float blend1(float ratio, float a, float b) {  // _blend before the clamp
  return __fadd_rn(__fmul_rn(ratio, a), __fmul_rn(1.f - ratio, b));
}

struct HueCtx {  // forward intermediates the backward needs
  float M, mn, s, f, dv_, den_s, rc, gc, bc, p_raw, q_raw, t_raw;
  int imax, imin, branch, sext;
  bool eq;
};

__device__ __forceinline__ void hue_fwd(const float in[3], float hf, float out[3], HueCtx* ctx) {
  const float r = in[0], g = in[1], b = in[2];
  const float M = fmaxf(r, fmaxf(g, b)), mn = fminf(r, fminf(g, b));
  const bool eq = (M == mn);
  const float cr = M - mn;
  const float den_s = eq ? 1.f : M;
  const float s = cr / den_s;
  const float dv_ = eq ? 1.f : cr;
  const float rc = (M - r) / dv_, gc = (M - g) / dv_, bc = (M - b) / dv_;
  float hraw; int branch;
  if (M == r) { hraw = bc - gc; branch = 0; }
  else if (M == g) { hraw = 2.f + rc - bc; branch = 1; }
  else { hraw = 4.f + gc - rc; branch = 2; }
  float h = fmodf(hraw / 6.f + 1.f, 1.f);
  float h2 = h + hf;
  h2 = h2 - floorf(h2);  // python-style % 1.0
  if (h2 >= 1.f) h2 = 0.f;
  const float h6 = h2 * 6.f;
  const float fi = floorf(h6);
  const float f = h6 - fi;
  int i = ((int)fi) % 6;
  if (i < 0) i += 6;
  const float v = M;
  const float p_raw = v * (1.f - s), q_raw = v * (1.f - s * f), t_raw = v * (1.f - s * (1.f - f));
  const float p = clamp01(p_raw), q = clamp01(q_raw), t = clamp01(t_raw);
  switch (i) {
    case 0: out[0] = v; out[1] = t; out[2] = p; break;
    case 1: out[0] = q; out[1] = v; out[2] = p; break;
    case 2: out[0] = p; out[1] = v; out[2] = t; break;
    case 3: out[0] = p; out[1] = q; out[2] = v; break;
    case 4: out[0] = t; out[1] = p; out[2] = v; break;
    default: out[0] = v; out[1] = p; out[2] = q; break;
  }
  if (ctx) {
    ctx->M = M; ctx->mn = mn; ctx->s = s; ctx->f = f; ctx->dv_ = dv_; ctx->den_s = den_s;
    ctx->rc = rc; ctx->gc = gc; ctx->bc = bc; ctx->p_raw = p_raw; ctx->q_raw = q_raw; ctx->t_raw = t_raw;
    ctx->branch = branch; ctx->sext = i; ctx->eq = eq;
    ctx->imax = (r == M) ? 0 : ((g == M) ? 1 : 2);
    ctx->imin = (r == mn) ? 0 : ((g == mn) ? 1 : 2);
  }
}

// reverse-mode derivative of hue_fwd as torch autograd computes it (max/min send the gradient to one
// index, comparison masks / floor are constants, clamp passes gradient on the closed interval).
__device__ __forceinline__ void hue_bwd(const HueCtx& c, const float dout[3], float din[3]) {
  float dv = 0.f, dp = 0.f, dq = 0.f, dt = 0.f;
  switch (c.sext) {
    case 0: dv += dout[0]; dt += dout[1]; dp += dout[2]; break;
    case 1: dq += dout[0]; dv += dout[1]; dp += dout[2]; break;
    case 2: dp += dout[0]; dv += dout[1]; dt += dout[2]; break;
    case 3: dp += dout[0]; dq += dout[1]; dv += dout[2]; break;
    case 4: dt += dout[0]; dp += dout[1]; dv += dout[2]; break;
    default: dv += dout[0]; dp += dout[1]; dq += dout[2]; break;
  }
  const float v = c.M, s = c.s, f = c.f;
  float ds = 0.f, df = 0.f;
  if (c.p_raw >= 0.f && c.p_raw <= 1.f) { dv += dp * (1.f - s); ds -= dp * v; }
  if (c.q_raw >= 0.f && c.q_raw <= 1.f) { dv += dq * (1.f - s * f); ds -= dq * v * f; df -= dq * v * s; }
  if (c.t_raw >= 0.f && c.t_raw <= 1.f) { dv += dt * (1.f - s * (1.f - f)); ds -= dt * v * (1.f - f); df += dt * v * s; }
  const float dhraw = df;  // f = 6*h2 - i, h2 = h + hf (mod 1), h = hraw/6 + 1 (mod 1)
  float drc = 0.f, dgc = 0.f, dbc = 0.f;
  if (c.branch == 0) { dbc += dhraw; dgc -= dhraw; }
  else if (c.branch == 1) { drc += dhraw; dbc -= dhraw; }
  else { dgc += dhraw; drc -= dhraw; }
  float dM = dv, dmn = 0.f, ddv = 0.f;
  din[0] = din[1] = din[2] = 0.f;
  const float idv = 1.f / c.dv_;
  dM += (drc + dgc + dbc) * idv;
  din[0] -= drc * idv; din[1] -= dgc * idv; din[2] -= dbc * idv;
  ddv -= (drc * c.rc + dgc * c.gc + dbc * c.bc) * idv;
  float dcr = 0.f;
  if (!c.eq) dcr += ddv;
  dcr += ds / c.den_s;
  if (!c.eq) dM -= ds * s / c.den_s;
  dM += dcr; dmn -= dcr;
  din[c.imax] += dM;
  din[c.imin] += dmn;
}

struct JitterParams {
  int perm[4];
  float b, c, s, h;
};

// Applies ops perm[first..last) to v.  mean is the contrast mean (only read if the contrast op is in range).
__device__ __forceinline__ void jitter_apply(const JitterParams& jp, int first, int last, float mean, float v[3]) {
  for (int k = first; k < last; ++k) {
    const int op = jp.perm[k];
    if (op == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) v[i] = clamp01(blend1(jp.b, v[i], 0.f));
    } else if (op == 1) {
#pragma unroll
      for (int i = 0; i < 3; ++i) v[i] = clamp01(blend1(jp.c, v[i], mean));
    } else if (op == 2) {
      const float g = gray_of(v);
#pragma unroll
      for (int i = 0; i < 3; ++i) v[i] = clamp01(blend1(jp.s, v[i], g));
    } else {
      float o[3];
      hue_fwd(v, jp.h, o, nullptr);
      v[0] = o[0]; v[1] = o[1]; v[2] = o[2];
    }
  }
}

__device__ __forceinline__ int contrast_pos(const JitterParams& jp) {
  for (int k = 0; k < 4; ++k) if (jp.perm[k] == 1) return k;
  return 4;
}

// Backward through ops perm[first..last) (applied in reverse); `in` is the value entering op `first`.
// For the contrast op: dx = c*dy*mask (+ the mean path added by the caller through S_total when have_S).
__device__ __forceinline__ void jitter_bwd_range(const JitterParams& jp, int first, int last, float mean, const float in[3],
                                                 float d[3], bool have_S, float S_over_n) {
  // recompute the inputs of every op in range
  float stage_in[4][3];
  float v[3] = {in[0], in[1], in[2]};
  for (int k = first; k < last; ++k) {
    stage_in[k][0] = v[0]; stage_in[k][1] = v[1]; stage_in[k][2] = v[2];
    jitter_apply(jp, k, k + 1, mean, v);
  }
  for (int k = last - 1; k >= first; --k) {
    const int op = jp.perm[k];
    const float* x = stage_in[k];
    if (op == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { const float y = blend1(jp.b, x[i], 0.f); d[i] = (y >= 0.f && y <= 1.f) ? jp.b * d[i] : 0.f; }
    } else if (op == 1) {
      float g[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) { const float y = blend1(jp.c, x[i], mean); g[i] = (y >= 0.f && y <= 1.f) ? d[i] : 0.f; }
      const float extra = have_S ? (1.f - jp.c) * S_over_n : 0.f;
      d[0] = jp.c * g[0] + GW0 * extra; d[1] = jp.c * g[1] + GW1 * extra; d[2] = jp.c * g[2] + GW2 * extra;
    } else if (op == 2) {
      const float gr = gray_of(x);
      float g[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) { const float y = blend1(jp.s, x[i], gr); g[i] = (y >= 0.f && y <= 1.f) ? d[i] : 0.f; }
      const float sg = (1.f - jp.s) * (g[0] + g[1] + g[2]);
      d[0] = jp.s * g[0] + GW0 * sg; d[1] = jp.s * g[1] + GW1 * sg; d[2] = jp.s * g[2] + GW2 * sg;
    } else {
      HueCtx ctx; float o[3], di[3];
      hue_fwd(x, jp.h, o, &ctx);
      hue_bwd(ctx, d, di);
      d[0] = di[0]; d[1] = di[1]; d[2] = di[2];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Affine sampling coordinates, torchvision tensor path (_functional_tensor.py:579-640) + grid_sample
// (bilinear, zeros, align_corners=False).
// ------------------------------------------------------------------------------------------------
struct AffineCoef { float r00, r10, r20, r01, r11, r21; };

__device__ __forceinline__ AffineCoef affine_coef(const float* theta, int cs) {
  const float hw = 0.5f * (float)cs;
  AffineCoef a;
  a.r00 = theta[0] / hw; a.r10 = theta[1] / hw; a.r20 = theta[2] / hw;
  a.r01 = theta[3] / hw; a.r11 = theta[4] / hw; a.r21 = theta[5] / hw;
  return a;
}
__device__ __forceinline__ void affine_src(const AffineCoef& a, int cs, int ox, int oy, float* ix, float* iy) {
  const float bx = (float)ox - 0.5f * (float)cs + 0.5f, by = (float)oy - 0.5f * (float)cs + 0.5f;
  const float gx = fmaf(1.f, a.r20, fmaf(by, a.r10, __fmul_rn(bx, a.r00)));
  const float gy = fmaf(1.f, a.r21, fmaf(by, a.r11, __fmul_rn(bx, a.r01)));
  *ix = ((gx + 1.f) * (float)cs - 1.f) * 0.5f;
  *iy = ((gy + 1.f) * (float)cs - 1.f) * 0.5f;
}

// F2: augment up to the pre-jitter image z, plus the partial sums of the contrast mean.
// grid (nblk, N), 256 threads, one output pixel (3 channels) per thread.
__global__ void __launch_bounds__(256) augment_fwd_kernel(const float* __restrict__ base, const int* __restrict__ uidx,
                                                          const cg_aug_t* __restrict__ augp, NoiseSrc ns, int cs, float* __restrict__ z,
                                                          float* __restrict__ partial) {
  __shared__ float red[32];
  __shared__ cg_aug_t aug;
  if (threadIdx.x < sizeof(cg_aug_t) / 4) reinterpret_cast<int*>(&aug)[threadIdx.x] = reinterpret_cast<const int*>(augp)[threadIdx.x];
  __syncthreads();
  const int n = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t pp = (size_t)cs * cs;
  const float* bn = base + (size_t)uidx[n] * 3 * pp;  // identical crops (e.g. > 4 overview cuts) share one base
  float v[3] = {0.f, 0.f, 0.f};
  float gsum = 0.f;
  if (pix < cs * cs) {
    const int oy = pix / cs, ox = pix - oy * cs;
    const AffineCoef ac = affine_coef(aug.theta, cs);
    float ix, iy;
    affine_src(ac, cs, ox, oy, &ix, &iy);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
    const float wgt[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};  // nw, ne, sw, se
    float mask = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
      if (xx >= 0 && xx < cs && yy >= 0 && yy < cs) {
        float nz[3];
        noise3(ns, 0, n, yy, xx, nz);
        const int xb = aug.flip ? cs - 1 - xx : xx;
        const size_t o = (size_t)yy * cs + xb;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = fmaf(bn[c * pp + o] + nz[c], wgt[k], v[c]);
        mask += wgt[k];
      }
    }
    float n2[3], n3[3];
    noise3(ns, 1, n, oy, ox, n2);
    noise3(ns, 2, n, oy, ox, n3);
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = v[c] * mask + n2[c];
    if (aug.gray) { const float g = gray_of(v); v[0] = v[1] = v[2] = g; }
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] += n3[c];
    float* zn = z + (size_t)n * 3 * pp + pix;
    zn[0] = v[0]; zn[pp] = v[1]; zn[2 * pp] = v[2];
    JitterParams jp = {{aug.perm[0], aug.perm[1], aug.perm[2], aug.perm[3]}, aug.brightness, aug.contrast, aug.saturation, aug.hue};
    jitter_apply(jp, 0, contrast_pos(jp), 0.f, v);
    gsum = gray_of(v);
  }
  const float t = block_sum(gsum, red);
  if (threadIdx.x == 0) partial[(size_t)n * gridDim.x + blockIdx.x] = t;
}

__device__ __forceinline__ float cut_mean(const float* partial, int n, int nblk, int cs, float* red) {
  float a = 0.f;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) a += partial[(size_t)n * nblk + i];
  return block_sum(a, red) / (float)(cs * cs);
}

__device__ __forceinline__ size_t patch_offset(int n, int oy, int ox, int c, int cs, int patch, int kpad) {
  const int g = cs / patch;
  const int py = oy / patch, px = ox / patch;
  const int iy = oy - py * patch, ixx = ox - px * patch;
  return ((size_t)n * g * g + (size_t)py * g + px) * kpad + (size_t)c * patch * patch + iy * patch + ixx;
}

// F3: jitter + normalise + layout.  src = z (augment on) or base (augment off).
__global__ void __launch_bounds__(256) jitter_fwd_kernel(const float* __restrict__ src, const int* __restrict__ uidx,
                                                         const cg_aug_t* __restrict__ augp, const float* __restrict__ partial, int nblk,
                                                         int cs, void* __restrict__ out, int fmt, int patch, int kpad) {
  __shared__ float red[32];
  __shared__ cg_aug_t aug;
  if (threadIdx.x < sizeof(cg_aug_t) / 4) reinterpret_cast<int*>(&aug)[threadIdx.x] = reinterpret_cast<const int*>(augp)[threadIdx.x];
  __syncthreads();
  const int n = blockIdx.y;
  const size_t pp = (size_t)cs * cs;
  float mean = 0.f;
  if (aug.augment) mean = cut_mean(partial, n, nblk, cs, red);
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= cs * cs) return;
  const float* sn = src + (size_t)(aug.augment ? n : uidx[n]) * 3 * pp + pix;
  float v[3] = {sn[0], sn[pp], sn[2 * pp]};
  if (aug.augment) {
    JitterParams jp = {{aug.perm[0], aug.perm[1], aug.perm[2], aug.perm[3]}, aug.brightness, aug.contrast, aug.saturation, aug.hue};
    jitter_apply(jp, 0, 4, mean, v);
  }
  if (aug.normalize) {
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (v[c] - aug.mean[c]) / aug.stdv[c];
  }
  if (fmt == CG_FMT_F32_NCHW) {
    float* on = reinterpret_cast<float*>(out) + (size_t)n * 3 * pp + pix;
    on[0] = v[0]; on[pp] = v[1]; on[2 * pp] = v[2];
  } else {
    const int oy = pix / cs, ox = pix - oy * cs;
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out);
#pragma unroll
    for (int c = 0; c < 3; ++c) ob[patch_offset(n, oy, ox, c, cs, patch, kpad)] = __float2bfloat16(v[c]);
    if ((oy % patch) == 0 && (ox % patch) == 0) {
      const size_t rowb = patch_offset(n, oy, ox, 0, cs, patch, kpad);
      for (int k = 3 * patch * patch; k < kpad; ++k) ob[rowb + k] = __float2bfloat16(0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_dout(const void* dout, int fmt, int n, int pix, int cs, int patch, int kpad, float d[3]) {
  const size_t pp = (size_t)cs * cs;
  if (fmt == CG_FMT_F32_NCHW) {
    const float* p = reinterpret_cast<const float*>(dout) + (size_t)n * 3 * pp + pix;
    d[0] = p[0]; d[1] = p[pp]; d[2] = p[2 * pp];
  } else if (fmt == CG_FMT_F32_PATCH) {
    const int oy = pix / cs, ox = pix - oy * cs;
    const float* p = reinterpret_cast<const float*>(dout);
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = p[patch_offset(n, oy, ox, c, cs, patch, kpad)];
  } else {
    const int oy = pix / cs, ox = pix - oy * cs;
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(dout);
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = __bfloat162float(p[patch_offset(n, oy, ox, c, cs, patch, kpad)]);
  }
}

// B1 (pass 0): S[n] partials = sum over pixels/channels of (gradient at the contrast output) * clamp mask.
// B1 (pass 1): full jitter backward + RandomGrayscale backward -> d(affine output) into `da`.
template <int PASS>
__global__ void __launch_bounds__(256) jitter_bwd_kernel(const void* __restrict__ dout, const float* __restrict__ z,
                                                         const cg_aug_t* __restrict__ augp, const float* __restrict__ partial,
                                                         float* __restrict__ spartial, int nblk, int cs, int fmt, int patch, int kpad,
                                                         float* __restrict__ da) {
  __shared__ float red[32];
  __shared__ cg_aug_t aug;
  if (threadIdx.x < sizeof(cg_aug_t) / 4) reinterpret_cast<int*>(&aug)[threadIdx.x] = reinterpret_cast<const int*>(augp)[threadIdx.x];
  __syncthreads();
  const int n = blockIdx.y;
  const size_t pp = (size_t)cs * cs;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = pix < cs * cs;
  float d[3] = {0.f, 0.f, 0.f};
  if (active) {
    load_dout(dout, fmt, n, pix, cs, patch, kpad, d);
    if (aug.normalize) {
#pragma unroll
      for (int c = 0; c < 3; ++c) d[c] = d[c] / aug.stdv[c];
    }
  }
  if (!aug.augment) {  // base cutouts only: gradient passes straight through
    if (PASS == 1 && active) { float* o = da + (size_t)n * 3 * pp + pix; o[0] = d[0]; o[pp] = d[1]; o[2 * pp] = d[2]; }
    return;
  }
  const float mean = cut_mean(partial, n, nblk, cs, red);
  JitterParams jp = {{aug.perm[0], aug.perm[1], aug.perm[2], aug.perm[3]}, aug.brightness, aug.contrast, aug.saturation, aug.hue};
  const int cp = contrast_pos(jp);
  float zin[3] = {0.f, 0.f, 0.f};
  if (active) { const float* zn = z + (size_t)n * 3 * pp + pix; zin[0] = zn[0]; zin[1] = zn[pp]; zin[2] = zn[2 * pp]; }
  if (PASS == 0) {
    float s = 0.f;
    if (active) {
      float v[3] = {zin[0], zin[1], zin[2]};
      jitter_apply(jp, 0, cp, mean, v);           // value entering the contrast op
      float vc[3] = {v[0], v[1], v[2]};
      jitter_apply(jp, cp, cp + 1, mean, vc);     // value leaving it
      jitter_bwd_range(jp, cp + 1, 4, mean, vc, d, false, 0.f);
#pragma unroll
      for (int i = 0; i < 3; ++i) { const float y = blend1(jp.c, v[i], mean); if (y >= 0.f && y <= 1.f) s += d[i]; }
    }
    const float t = block_sum(s, red);
    if (threadIdx.x == 0) spartial[(size_t)n * nblk + blockIdx.x] = t;
  } else {
    float a = 0.f;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) a += spartial[(size_t)n * nblk + i];
    const float S_over_n = block_sum(a, red) / (float)(cs * cs);
    if (!active) return;
    jitter_bwd_range(jp, 0, 4, mean, zin, d, true, S_over_n);
    if (aug.gray) { const float g = d[0] + d[1] + d[2]; d[0] = GW0 * g; d[1] = GW1 * g; d[2] = GW2 * g; }
    float* o = da + (size_t)n * 3 * pp + pix;
    o[0] = d[0]; o[pp] = d[1]; o[2 * pp] = d[2];
  }
}

// B1 (pass 1, de-duplicated): grid (nblk, U).  The affine map and the resampling are linear and identical for cutouts
// that share a crop geometry (the "> 4 overview cuts" case: N identical crops, cutouts.py:77-79), so their gradients
// are summed HERE, right after the non-linear jitter backward, and everything downstream runs once per unique crop.
__global__ void __launch_bounds__(256) jitter_bwd_sum_kernel(const void* __restrict__ dout, const float* __restrict__ z,
                                                             const cg_aug_t* __restrict__ augp, const float* __restrict__ partial,
                                                             const float* __restrict__ spartial, const int* __restrict__ hdr,
                                                             const int* __restrict__ dupoff, const int* __restrict__ duplist, int nblk, int cs,
                                                             int fmt, int patch, int kpad, float* __restrict__ da) {
  __shared__ float red[32];
  __shared__ cg_aug_t aug;
  if (threadIdx.x < sizeof(cg_aug_t) / 4) reinterpret_cast<int*>(&aug)[threadIdx.x] = reinterpret_cast<const int*>(augp)[threadIdx.x];
  __syncthreads();
  const int u = blockIdx.y;
  if (u >= hdr[6]) return;  // WsHeader.U
  const size_t pp = (size_t)cs * cs;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = pix < cs * cs;
  JitterParams jp = {{aug.perm[0], aug.perm[1], aug.perm[2], aug.perm[3]}, aug.brightness, aug.contrast, aug.saturation, aug.hue};
  float acc[3] = {0.f, 0.f, 0.f};
  for (int k = dupoff[u]; k < dupoff[u + 1]; ++k) {
    const int n = duplist[k];
    float d[3] = {0.f, 0.f, 0.f};
    if (active) {
      load_dout(dout, fmt, n, pix, cs, patch, kpad, d);
      if (aug.normalize) {
#pragma unroll
        for (int c = 0; c < 3; ++c) d[c] = d[c] / aug.stdv[c];
      }
    }
    if (aug.augment) {
      const float mean = cut_mean(partial, n, nblk, cs, red);
      float a = 0.f;
      for (int i = threadIdx.x; i < nblk; i += blockDim.x) a += spartial[(size_t)n * nblk + i];
      const float S_over_n = block_sum(a, red) / (float)(cs * cs);
      if (active) {
        const float* zn = z + (size_t)n * 3 * pp + pix;
        const float zin[3] = {zn[0], zn[pp], zn[2 * pp]};
        jitter_bwd_range(jp, 0, 4, mean, zin, d, true, S_over_n);
        if (aug.gray) { const float g = d[0] + d[1] + d[2]; d[0] = GW0 * g; d[1] = GW1 * g; d[2] = GW2 * g; }
      }
    }
    acc[0] += d[0]; acc[1] += d[1]; acc[2] += d[2];
  }
  if (active) {
    float* o = da + (size_t)u * 3 * pp + pix;
    o[0] = acc[0]; o[pp] = acc[1]; o[2 * pp] = acc[2];
  }
}

// B2: affine backward as a gather: base pixel (y, xb) collects from the <= 3x3 output pixels whose bilinear
// footprint contains it.  Writes d(base) (flip undone).
__global__ void __launch_bounds__(256) affine_bwd_kernel(const float* __restrict__ da, const cg_aug_t* __restrict__ augp,
                                                         const int* __restrict__ hdr, int cs, float* __restrict__ dbase) {
  __shared__ cg_aug_t aug;
  if (threadIdx.x < sizeof(cg_aug_t) / 4) reinterpret_cast<int*>(&aug)[threadIdx.x] = reinterpret_cast<const int*>(augp)[threadIdx.x];
  __syncthreads();
  const int n = blockIdx.y;  // unique crop index
  if (n >= hdr[6]) return;
  const size_t pp = (size_t)cs * cs;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= cs * cs) return;
  const float* dn = da + (size_t)n * 3 * pp;
  float* on = dbase + (size_t)n * 3 * pp;
  if (!aug.augment) { on[pix] = dn[pix]; on[pp + pix] = dn[pp + pix]; on[2 * pp + pix] = dn[2 * pp + pix]; return; }
  const int py = pix / cs, px = pix - py * cs;  // coordinates in the affine INPUT image
  const AffineCoef ac = affine_coef(aug.theta, cs);
  const float c0 = 0.5f * (float)(cs - 1);
  const float qx = (float)px - c0, qy = (float)py - c0;
  const float osx = aug.theta_fwd[0] * qx + aug.theta_fwd[1] * qy + aug.theta_fwd[2] + c0;
  const float osy = aug.theta_fwd[3] * qx + aug.theta_fwd[4] * qy + aug.theta_fwd[5] + c0;
  const int cx = (int)floorf(osx + 0.5f), cy = (int)floorf(osy + 0.5f);
  float acc[3] = {0.f, 0.f, 0.f};
  for (int dy = -1; dy <= 1; ++dy) {
    const int oy = cy + dy;
    if (oy < 0 || oy >= cs) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int ox = cx + dx;
      if (ox < 0 || ox >= cs) continue;
      float ix, iy;
      affine_src(ac, cs, ox, oy, &ix, &iy);
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = (int)fx, y0 = (int)fy;
      float wx, wy;
      if (px == x0) wx = (fx + 1.f) - ix; else if (px == x0 + 1) wx = ix - fx; else continue;
      if (py == y0) wy = (fy + 1.f) - iy; else if (py == y0 + 1) wy = iy - fy; else continue;
      // sampled mask of this output pixel
      const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
      float mask = 0.f;
      if (x0 >= 0 && x0 < cs && y0 >= 0 && y0 < cs) mask += wx0 * wy0;
      if (x0 + 1 >= 0 && x0 + 1 < cs && y0 >= 0 && y0 < cs) mask += wx1 * wy0;
      if (x0 >= 0 && x0 < cs && y0 + 1 >= 0 && y0 + 1 < cs) mask += wx0 * wy1;
      if (x0 + 1 >= 0 && x0 + 1 < cs && y0 + 1 >= 0 && y0 + 1 < cs) mask += wx1 * wy1;
      const float w = wx * wy * mask;
      const size_t o = (size_t)oy * cs + ox;
      acc[0] = fmaf(w, dn[o], acc[0]); acc[1] = fmaf(w, dn[pp + o], acc[1]); acc[2] = fmaf(w, dn[2 * pp + o], acc[2]);
    }
  }
  const int xb = aug.flip ? cs - 1 - px : px;
  const size_t o = (size_t)py * cs + xb;
  on[o] = acc[0]; on[pp + o] = acc[1]; on[2 * pp + o] = acc[2];
}

// B3: transposed resample, gather over SOURCE pixels.  grid (ceil(W/32), ceil(H/8)), block (32, 8).
// Every thread owns one pixel of d(x_in) (3 channels) and loops over the cutouts covering it.
template <int TTN>
__global__ void __launch_bounds__(256) resample_bwd_kernel(const float* __restrict__ dbase, const cg_cut_t* __restrict__ cuts, const int* __restrict__ hdr,
                                                           const int* __restrict__ tstart, const float* __restrict__ wtr, int cs,
                                                           int max_size, int H, int W, float coef, int accumulate,
                                                           float* __restrict__ dx_in) {
  const int sx = blockIdx.x * 32 + threadIdx.x, sy = blockIdx.y * 8 + threadIdx.y;
  const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * 8;
  const size_t pp = (size_t)cs * cs;
  float acc[3] = {0.f, 0.f, 0.f};
  const int N = hdr[6];  // unique crops only: duplicates were summed in jitter_bwd_sum_kernel
  for (int n = 0; n < N; ++n) {
    const cg_cut_t cut = cuts[n];
    // tile-uniform rejection
    if (tx0 + 32 <= cut.x0 || tx0 >= cut.x0 + cut.size || ty0 + 8 <= cut.y0 || ty0 >= cut.y0 + cut.size) continue;
    const int lx = sx - cut.x0, ly = sy - cut.y0;
    if (lx < 0 || lx >= cut.size || ly < 0 || ly >= cut.size || sx >= W || sy >= H) continue;
    const int* ts = tstart + (size_t)n * max_size;
    const float* wt = wtr + (size_t)n * max_size * TT;
    const int oy0 = ts[ly], ox0 = ts[lx];
    float wy[TTN], wx[TTN];
#pragma unroll
    for (int j = 0; j < TTN; ++j) { wy[j] = wt[(size_t)ly * TT + j]; wx[j] = wt[(size_t)lx * TT + j]; }
    const float* dn = dbase + (size_t)n * 3 * pp;
    const bool hflip = cut.flags & CG_CUT_HFLIP;
    float a[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int jy = 0; jy < TTN; ++jy) {
      const int oy = oy0 + jy;
      if (wy[jy] == 0.f || oy >= cs) continue;
      float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int jx = 0; jx < TTN; ++jx) {
        const int ox = ox0 + jx;
        if (wx[jx] == 0.f || ox >= cs) continue;
        const size_t o = (size_t)oy * cs + (hflip ? cs - 1 - ox : ox);
        r[0] = fmaf(wx[jx], dn[o], r[0]); r[1] = fmaf(wx[jx], dn[pp + o], r[1]); r[2] = fmaf(wx[jx], dn[2 * pp + o], r[2]);
      }
      a[0] = fmaf(wy[jy], r[0], a[0]); a[1] = fmaf(wy[jy], r[1], a[1]); a[2] = fmaf(wy[jy], r[2], a[2]);
    }
    if (cut.flags & (CG_CUT_GRAY_PRE | CG_CUT_GRAY_POST)) {
      const float g = a[0] + a[1] + a[2];
      a[0] = GW0 * g; a[1] = GW1 * g; a[2] = GW2 * g;
    }
    acc[0] += a[0]; acc[1] += a[1]; acc[2] += a[2];
  }
  if (sx < W && sy < H) {
    const size_t plane = (size_t)H * W, o = (size_t)sy * W + sx;
    const float k = coef;  // the caller folds d((x+1)/2)/dx = 0.5 in
#pragma unroll
    for (int c = 0; c < 3; ++c) dx_in[c * plane + o] = (accumulate ? dx_in[c * plane + o] : 0.f) + k * acc[c];
  }
}

}  // namespace

// tt of the latest forward per workspace (the backward entry point does not see the host structs)
struct TtEntry { const void* ws; int tt; };
static TtEntry g_tt_cache[64] = {};
static int g_tt_next = 0;
static void remember_tt(const void* ws, int tt) {
  for (auto& e : g_tt_cache)
    if (e.ws == ws) { e.tt = tt; return; }
  g_tt_cache[g_tt_next] = {ws, tt};
  g_tt_next = (g_tt_next + 1) % 64;
}
static int recall_tt(const void* ws) {
  for (auto& e : g_tt_cache)
    if (e.ws == ws) return e.tt;
  return TT;  // unknown workspace: the widest (always correct) variant
}

extern "C" size_t cg_cutouts_workspace_bytes(int N, int cs, int max_size) {
  if (N <= 0 || cs <= 0 || max_size <= 0) return 0;
  return ws_layout(N, cs, max_size).total;
}

extern "C" int cg_cutouts_fwd(const float* x_in, int H, int W, const cg_cut_t* cuts_h, int N, int cs, const cg_aug_t* aug_h,
                              const float* noise, void* out, int fmt, int patch, int kpad, void* workspace, void* stream) {
  CG_REQUIRE(x_in && cuts_h && aug_h && out && workspace, "cg_cutouts_fwd: null pointer");
  CG_REQUIRE(N > 0 && cs > 0 && H > 0 && W > 0, "cg_cutouts_fwd: bad sizes");
  CG_REQUIRE(fmt == CG_FMT_F32_NCHW || fmt == CG_FMT_BF16_PATCH, "cg_cutouts_fwd: unknown format %d", fmt);
  if (fmt == CG_FMT_BF16_PATCH)
    CG_REQUIRE(patch > 0 && cs % patch == 0 && kpad >= 3 * patch * patch, "cg_cutouts_fwd: bad patch layout (cs=%d patch=%d kpad=%d)", cs, patch, kpad);
  int max_size = 0, min_sz = 1 << 30;
  for (int i = 0; i < N; ++i) {
    const cg_cut_t& c = cuts_h[i];
    CG_REQUIRE(c.size >= 1 && c.size <= MAX_SIZE, "cg_cutouts_fwd: cut %d size %d outside [1, %d]", i, c.size, MAX_SIZE);
    // inner cuts lie inside the image (cutouts.py:84-92 draws x in [0, W - size], y in [0, H - size]); only the overview cut is padded
    CG_REQUIRE((c.flags & CG_CUT_OVERVIEW) || (c.x0 >= 0 && c.y0 >= 0 && c.x0 + c.size <= W && c.y0 + c.size <= H),
               "cg_cutouts_fwd: inner cut %d (y0=%d x0=%d size=%d) leaves the %dx%d image", i, c.y0, c.x0, c.size, H, W);
    if (c.size > max_size) max_size = c.size;
    if (c.size < min_sz) min_sz = c.size;
  }
  CG_REQUIRE((max_size * 4 + cs - 1) / cs <= TAPS_MAX, "cg_cutouts_fwd: downscale ratio above %dx (size %d -> %d)", TAPS_MAX / 4, max_size, cs);
  // ELL entries needed by the transposed resample: ceil(4 * max(cs / size, 1)) + 2
  const int tt_need = min_sz >= cs ? 6 : (4 * cs + min_sz - 1) / min_sz + 2;
  CG_REQUIRE(tt_need <= TT, "cg_cutouts_fwd: upscale ratio too large (size %d -> %d)", min_sz, cs);
  const int tmp_ld = max_size;
  // the workspace was sized by the caller for max(H,W): use that bound so fwd/bwd agree
  max_size = H > W ? H : W;
  if (max_size > MAX_SIZE) max_size = MAX_SIZE;
  const WsLayout L = ws_layout(N, cs, max_size);
  char* ws = reinterpret_cast<char*>(workspace);
  cudaStream_t s = cg_stream(stream);
  // ---- metadata block, built on the host and shipped with ONE copy: header, cuts, aug, unique-crop tables
  std::vector<char> meta(L.meta_end, 0);
  cg_cut_t* ucuts_h = reinterpret_cast<cg_cut_t*>(meta.data() + L.ucuts);
  int* uidx_h = reinterpret_cast<int*>(meta.data() + L.uidx);
  int* dupoff_h = reinterpret_cast<int*>(meta.data() + L.dupoff);
  int* duplist_h = reinterpret_cast<int*>(meta.data() + L.duplist);
  int U = 0;
  for (int i = 0; i < N; ++i) {
    int u = 0;
    for (; u < U; ++u)
      if (ucuts_h[u].y0 == cuts_h[i].y0 && ucuts_h[u].x0 == cuts_h[i].x0 && ucuts_h[u].size == cuts_h[i].size &&
          ((ucuts_h[u].flags ^ cuts_h[i].flags) & (CG_CUT_GRAY_PRE | CG_CUT_GRAY_POST | CG_CUT_HFLIP)) == 0)
        break;
    if (u == U) ucuts_h[U++] = cuts_h[i];
    uidx_h[i] = u;
  }
  {  // CSR of the duplicates of every unique crop
    std::vector<int> cnt(U + 1, 0);
    for (int i = 0; i < N; ++i) cnt[uidx_h[i] + 1]++;
    for (int u = 0; u < U; ++u) cnt[u + 1] += cnt[u];
    for (int u = 0; u <= U; ++u) dupoff_h[u] = cnt[u];
    std::vector<int> fill(cnt.begin(), cnt.end() - 1);
    for (int i = 0; i < N; ++i) duplist_h[fill[uidx_h[i]]++] = i;
  }
  WsHeader hdr = {0x43475753, N, cs, max_size, H, W, U, aug_h->input01, tt_need <= 6 ? 6 : (tt_need <= 12 ? 12 : 24)};
  memcpy(meta.data(), &hdr, sizeof(hdr));
  remember_tt(workspace, hdr.tt);
  memcpy(meta.data() + L.cuts, cuts_h, sizeof(cg_cut_t) * N);
  memcpy(meta.data() + L.aug, aug_h, sizeof(cg_aug_t));
  CG_CUDA(cudaMemcpyAsync(ws, meta.data(), L.meta_end, cudaMemcpyHostToDevice, s));  // pageable source: staged before returning
  const cg_cut_t* ucuts = reinterpret_cast<const cg_cut_t*>(ws + L.ucuts);
  const int* uidx = reinterpret_cast<const int*>(ws + L.uidx);
  const cg_aug_t* aug = reinterpret_cast<const cg_aug_t*>(ws + L.aug);
  int* taps = reinterpret_cast<int*>(ws + L.taps);
  int* left = reinterpret_cast<int*>(ws + L.left);
  float* wfw = reinterpret_cast<float*>(ws + L.wfw);
  int* tstart = reinterpret_cast<int*>(ws + L.tstart);
  float* wtr = reinterpret_cast<float*>(ws + L.wtr);
  float* base = reinterpret_cast<float*>(ws + L.base);
  float* z = reinterpret_cast<float*>(ws + L.z);
  float* partial = reinterpret_cast<float*>(ws + L.partial);

  tables_kernel<<<U, 256, sizeof(int) * cs, s>>>(ucuts, U, cs, max_size, taps, left, wfw, tstart, wtr);
  CG_LAUNCH_CHECK();
  {
    const size_t tmp_bytes = sizeof(float) * RT * (size_t)tmp_ld;
    static size_t configured[CG_MAX_DEVICES] = {};  // per device; static shared memory (weights, first taps) shares the 48 KB default limit
    const int dev = cg_device_index();
    if (tmp_bytes > 40 * 1024 && tmp_bytes > configured[dev]) {
      CG_CUDA(cudaFuncSetAttribute(resample_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tmp_bytes));
      configured[dev] = tmp_bytes;
    }
    resample_fwd_kernel<<<dim3((cs + RT - 1) / RT, U), 256, tmp_bytes, s>>>(x_in, H, W, ucuts, U, taps, left, wfw, cs, aug_h->input01, tmp_ld, base);
    CG_LAUNCH_CHECK();
  }
  dim3 grid(L.nblk, N);
  if (aug_h->augment) {
    NoiseSrc ns = {noise, aug_h->noise_seed, aug_h->cut_index0, N, cs, aug_h->noise_std, aug_h->noise_mode, aug_h->noise_threads,
                   {aug_h->noise_offset[0], aug_h->noise_offset[1], aug_h->noise_offset[2]}};
    CG_REQUIRE(noise || aug_h->noise_mode != 1 || aug_h->noise_threads > 0, "cg_cutouts_fwd: noise_mode 1 needs noise_threads (cg_randn_like_torch_geometry)");
    augment_fwd_kernel<<<grid, 256, 0, s>>>(base, uidx, aug, ns, cs, z, partial);
    CG_LAUNCH_CHECK();
  }
  jitter_fwd_kernel<<<grid, 256, 0, s>>>(aug_h->augment ? z : base, uidx, aug, partial, L.nblk, cs, out, fmt, patch, kpad);
  CG_LAUNCH_CHECK();
  return 0;
}

extern "C" int cg_cutouts_bwd(const void* dout, int H, int W, int N, int cs, int fmt, int patch, int kpad, float coef, int accumulate,
                              int input01, float* dx_in, void* workspace, void* stream) {
  CG_REQUIRE(dout && dx_in && workspace, "cg_cutouts_bwd: null pointer");
  CG_REQUIRE(N > 0 && cs > 0 && H > 0 && W > 0, "cg_cutouts_bwd: bad sizes");
  int max_size = H > W ? H : W;
  if (max_size > MAX_SIZE) max_size = MAX_SIZE;
  const WsLayout L = ws_layout(N, cs, max_size);
  char* ws = reinterpret_cast<char*>(workspace);
  cudaStream_t s = cg_stream(stream);
  const cg_cut_t* ucuts = reinterpret_cast<const cg_cut_t*>(ws + L.ucuts);
  const int* hdr = reinterpret_cast<const int*>(ws);
  const int* dupoff = reinterpret_cast<const int*>(ws + L.dupoff);
  const int* duplist = reinterpret_cast<const int*>(ws + L.duplist);
  const cg_aug_t* aug = reinterpret_cast<const cg_aug_t*>(ws + L.aug);
  int* tstart = reinterpret_cast<int*>(ws + L.tstart);
  float* wtr = reinterpret_cast<float*>(ws + L.wtr);
  float* base = reinterpret_cast<float*>(ws + L.base);
  float* z = reinterpret_cast<float*>(ws + L.z);
  float* scratch = reinterpret_cast<float*>(ws + L.scratch);
  float* partial = reinterpret_cast<float*>(ws + L.partial);
  float* spartial = reinterpret_cast<float*>(ws + L.spartial);
  // The number of unique crops U lives in the device-side header (this call does not see the host structs): the
  // per-unique kernels are launched over N slots and slots >= U exit immediately.
  dim3 grid(L.nblk, N);
  jitter_bwd_kernel<0><<<grid, 256, 0, s>>>(dout, z, aug, partial, spartial, L.nblk, cs, fmt, patch, kpad, scratch);
  CG_LAUNCH_CHECK();
  jitter_bwd_sum_kernel<<<grid, 256, 0, s>>>(dout, z, aug, partial, spartial, hdr, dupoff, duplist, L.nblk, cs, fmt, patch, kpad, scratch);
  CG_LAUNCH_CHECK();
  affine_bwd_kernel<<<grid, 256, 0, s>>>(scratch, aug, hdr, cs, base);
  CG_LAUNCH_CHECK();
  const dim3 rg((W + 31) / 32, (H + 7) / 8), rb(32, 8);
  const float rcoef = input01 ? coef : 0.5f * coef;
  const int tt = recall_tt(workspace);
  if (tt <= 6) resample_bwd_kernel<6><<<rg, rb, 0, s>>>(base, ucuts, hdr, tstart, wtr, cs, max_size, H, W, rcoef, accumulate, dx_in);
  else if (tt <= 12) resample_bwd_kernel<12><<<rg, rb, 0, s>>>(base, ucuts, hdr, tstart, wtr, cs, max_size, H, W, rcoef, accumulate, dx_in);
  else resample_bwd_kernel<24><<<rg, rb, 0, s>>>(base, ucuts, hdr, tstart, wtr, cs, max_size, H, W, rcoef, accumulate, dx_in);
  CG_LAUNCH_CHECK();
  return 0;
}

// torch's launch geometry for a nullary distribution kernel over `numel` elements (ATen/native/cuda/DistributionTemplates.h,
// calc_execution_policy with block size 256 and unroll factor 4)
extern "C" uint32_t cg_randn_like_torch_geometry(int64_t numel, uint64_t* offset_increment) {
  if (numel <= 0) { if (offset_increment) *offset_increment = 0; return 0; }
  int dev = 0, sms = CG_NUM_SMS, max_threads = 2048;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_threads, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
  }
  uint64_t grid = (uint64_t)((numel + 255) / 256);
  const uint64_t cap = (uint64_t)sms * (uint64_t)(max_threads / 256);
  if (grid > cap) grid = cap;
  if (offset_increment) *offset_increment = ((uint64_t)(numel - 1) / (256ull * grid * 4ull) + 1ull) * 4ull;
  return (uint32_t)(grid * 256ull);
}

extern "C" int cg_randn_like_torch(float* out, int64_t numel, uint64_t seed, uint64_t offset, void* stream) {
  CG_REQUIRE(out && numel > 0, "cg_randn_like_torch: bad arguments");
  CG_REQUIRE(offset % 4 == 0, "cg_randn_like_torch: torch's Philox offsets are multiples of 4");
  const uint32_t nthreads = cg_randn_like_torch_geometry(numel, nullptr);
  int64_t blocks = (numel + 255) / 256;
  if (blocks > CG_NUM_SMS * 16) blocks = CG_NUM_SMS * 16;
  randn_like_torch_kernel<<<(unsigned)blocks, 256, 0, cg_stream(stream)>>>(out, numel, seed, offset, nthreads);
  CG_LAUNCH_CHECK();
  return 0;
}
