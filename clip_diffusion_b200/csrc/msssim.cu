// Multi-scale SSIM dissimilarity of the init-image branch of conditon_function (clip_diffusion/sample.py:220-225 ->
// structural_dissimilarity_loss, losses.py:48-54 -> pytorch_msssim.MS_SSIM(win_size=11, win_sigma=1.5, data_range=1,
// size_average=True, channel=3), losses.py:7): value AND analytic gradient with respect to the first image.
//
//   loss = 1 - mean_c prod_l relu(v[l,c]) ^ w[l],   v[l,c] = mean over the valid pixels of cs_l (l < 4) or ssim_4 (l = 4),
//   level l+1 = 2x2 average pool of level l, maps from 11-tap separable Gaussian moments (valid region only).
//
// Kernels (all tiny: 512^2 x 3 floats per map):
//   pool      level l -> l+1 for both images (level 0 applies denormalize_image_zero_to_one, (x+1)/2, image_utils.py:40-42)
//   maps      per level: Gaussian moments through shared-memory tiles, the cs / ssim map, its block partial sums (deterministic,
//             summed in index order later) and the three coefficient maps of the gradient:
//               d m(p) / d x(q) = g(p-q) [ A(p) + 2 x(q) B(p) + y(q) C(p) ],  A = dm/dmu1, B = dm/dE[x^2], C = dm/dE[xy]
//   scalars   one block: v[l,c], the loss, and k[l,c] = d loss / d v[l,c] / (valid pixels of level l)
//   grad      per level: transposed ("full") Gaussian correlation of A, B, C, scaled by k -> d loss / d X_l
//   up        d loss / d x = 0.5 sum_l 4^-l (d loss / d X_l)(q >> l): the transposes of the average pools and of (x+1)/2
#include <math.h>
#include "common.cuh"

namespace {

constexpr int WIN = 11, HALO = WIN - 1, LEVELS = 5;
constexpr int TX = 32, TY = 16;  // output tile of the maps / grad kernels
__constant__ float c_win[WIN];
__constant__ float c_weights[LEVELS] = {0.0448f, 0.2856f, 0.3001f, 0.2363f, 0.1333f};

struct Level {
  int H, W;         // size of the level
  int vh, vw;       // valid map size = H - 10, W - 10
  size_t img_off;   // offset (floats) of X_l in the pyramid buffer (Y_l at + pyr_half)
  size_t map_off;   // offset of the coefficient maps [3 coef][C][vh*vw]
  size_t grad_off;  // offset of d loss / d X_l [C][H*W]
  int part_off;     // first block partial of this level (per channel contiguous)
  int blocks_per_c; // tiles per channel
};
struct Pyramid {
  Level lv[LEVELS];
  size_t pyr_half;  // floats of one image's pyramid
  int C;
};

__global__ void __launch_bounds__(256) ms_pool_kernel(const float* __restrict__ xin, const float* __restrict__ yin, float* pyr, Pyramid P,
                                                      int l) {
  // l == -1: level 0 from the inputs in [-1,1]; otherwise level l -> l+1
  const int C = P.C;
  const Level& dst = P.lv[l + 1];
  const size_t n = (size_t)C * dst.H * dst.W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (size_t)gridDim.x * blockDim.x) {
    const int which = i >= n;
    const size_t j = which ? i - n : i;
    float v;
    if (l < 0) {
      v = ((which ? yin : xin)[j] + 1.f) * 0.5f;
    } else {
      const Level& src = P.lv[l];
      const int c = (int)(j / ((size_t)dst.H * dst.W));
      const int r = (int)(j - (size_t)c * dst.H * dst.W);
      const int y = r / dst.W, x = r - y * dst.W;
      const float* s = pyr + (which ? P.pyr_half : 0) + src.img_off + ((size_t)c * src.H + 2 * y) * src.W + 2 * x;
      v = 0.25f * ((s[0] + s[1]) + (s[src.W] + s[src.W + 1]));
    }
    pyr[(which ? P.pyr_half : 0) + dst.img_off + j] = v;
  }
}

// grid (tiles_x, tiles_y, C)
__global__ void __launch_bounds__(TX * TY) ms_maps_kernel(const float* pyr, float* maps, float* __restrict__ partials, Pyramid P,
                                                          int l) {
  __shared__ float sx[TY + HALO][TX + HALO], sy[TY + HALO][TX + HALO];
  __shared__ float h[5][TY + HALO][TX];  // horizontally filtered x, y, xx, yy, xy
  __shared__ float red[32];
  const Level& L = P.lv[l];
  const int c = blockIdx.z;
  const int ox0 = blockIdx.x * TX, oy0 = blockIdx.y * TY;
  const float* X = pyr + L.img_off + (size_t)c * L.H * L.W;
  const float* Y = X + P.pyr_half;
  const int tid = threadIdx.x, tx = tid & (TX - 1), ty = tid / TX;  // 1-D block of TX*TY threads (block_sum assumes one)
  for (int i = tid; i < (TY + HALO) * (TX + HALO); i += TX * TY) {
    const int r = i / (TX + HALO), q = i - r * (TX + HALO);
    const int gy = oy0 + r, gx = ox0 + q;
    const bool ok = gy < L.H && gx < L.W;
    sx[r][q] = ok ? X[(size_t)gy * L.W + gx] : 0.f;
    sy[r][q] = ok ? Y[(size_t)gy * L.W + gx] : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < (TY + HALO) * TX; i += TX * TY) {
    const int r = i / TX, q = i - r * TX;
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int t = 0; t < WIN; ++t) {
      const float w = c_win[t], u = sx[r][q + t], v = sy[r][q + t];
      a = fmaf(w, u, a); b = fmaf(w, v, b); aa = fmaf(w, u * u, aa); bb = fmaf(w, v * v, bb); ab = fmaf(w, u * v, ab);
    }
    h[0][r][q] = a; h[1][r][q] = b; h[2][r][q] = aa; h[3][r][q] = bb; h[4][r][q] = ab;
  }
  __syncthreads();
  const int oy = oy0 + ty, ox = ox0 + tx;
  float m = 0.f;
  if (oy < L.vh && ox < L.vw) {
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int t = 0; t < WIN; ++t) {
      const float w = c_win[t];
      mu1 = fmaf(w, h[0][ty + t][tx], mu1);
      mu2 = fmaf(w, h[1][ty + t][tx], mu2);
      e11 = fmaf(w, h[2][ty + t][tx], e11);
      e22 = fmaf(w, h[3][ty + t][tx], e22);
      e12 = fmaf(w, h[4][ty + t][tx], e12);
    }
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float s11 = e11 - mu1 * mu1, s22 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
    const float N = 2.f * s12 + C2, D = s11 + s22 + C2;
    const float cs = N / D;
    // d cs / d (mu1, E11, E12)
    float dA = (-2.f * mu2 + 2.f * mu1 * cs) / D, dB = -cs / D, dC = 2.f / D;
    m = cs;
    if (l == LEVELS - 1) {
      const float Nl = 2.f * mu1 * mu2 + C1, Dl = mu1 * mu1 + mu2 * mu2 + C1;
      const float lum = Nl / Dl;
      const float dl = (2.f * mu2 - 2.f * mu1 * lum) / Dl;
      dA = cs * dl + lum * dA; dB *= lum; dC *= lum;
      m = lum * cs;
    }
    const size_t plane = (size_t)L.vh * L.vw, o = (size_t)oy * L.vw + ox;
    float* mp = maps + L.map_off + (size_t)c * plane + o;
    mp[0] = dA; mp[(size_t)P.C * plane] = dB; mp[2 * (size_t)P.C * plane] = dC;
  }
  const float t = block_sum(m, red);
  if (tid == 0) partials[L.part_off + c * L.blocks_per_c + blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// one block: v[l,c] (fixed-order sums), loss, k[l,c]
__global__ void __launch_bounds__(256) ms_scalars_kernel(const float* __restrict__ partials, Pyramid P, float scale, float* __restrict__ loss_out,
                                                         float* __restrict__ kcoef) {
  __shared__ float red[32];
  __shared__ float v[LEVELS][8];
  const int C = P.C;
  for (int l = 0; l < LEVELS; ++l)
    for (int c = 0; c < C; ++c) {
      const float t = cg_sum_partials(partials, (unsigned)(P.lv[l].part_off + c * P.lv[l].blocks_per_c), (unsigned)P.lv[l].blocks_per_c, red);
      if (threadIdx.x == 0) v[l][c] = t / ((float)P.lv[l].vh * (float)P.lv[l].vw);
      __syncthreads();
    }
  if (threadIdx.x == 0) {
    float mean = 0.f;
    for (int c = 0; c < C; ++c) {
      float ms = 1.f;
      for (int l = 0; l < LEVELS; ++l) ms *= powf(fmaxf(v[l][c], 0.f), c_weights[l]);
      mean += ms;
      for (int l = 0; l < LEVELS; ++l) {
        // d(1 - mean_c ms_c)/d v[l,c] = -(1/C) ms_c w_l / v[l,c]  (0 where relu clipped), then the mean over the level's valid pixels
        const float g = v[l][c] > 0.f ? -(1.f / (float)C) * ms * c_weights[l] / v[l][c] : 0.f;
        kcoef[l * C + c] = scale * g / ((float)P.lv[l].vh * (float)P.lv[l].vw);
      }
    }
    if (loss_out) loss_out[0] = 1.f - mean / (float)C;
  }
}

// d loss / d X_l(q) = k[l,c] sum_p g(py-qy... ) : p = q - t, t in [0, 10], p valid.  grid (tiles_x, tiles_y, C) over the LEVEL (not the map)
__global__ void __launch_bounds__(TX * TY) ms_grad_kernel(const float* pyr, const float* maps, const float* __restrict__ kcoef,
                                                          float* lgrad, Pyramid P, int l) {
  __shared__ float sm[3][TY + HALO][TX + HALO];
  __shared__ float hv[3][TY + HALO][TX];
  const Level& L = P.lv[l];
  const int c = blockIdx.z;
  const int qx0 = blockIdx.x * TX, qy0 = blockIdx.y * TY;
  const size_t plane = (size_t)L.vh * L.vw;
  const int tid = threadIdx.x, tx = tid & (TX - 1), ty = tid / TX;
  // map pixels p with py in [qy0 - 10, qy0 + TY), px in [qx0 - 10, qx0 + TX)
  for (int i = tid; i < (TY + HALO) * (TX + HALO); i += TX * TY) {
    const int r = i / (TX + HALO), q = i - r * (TX + HALO);
    const int py = qy0 - HALO + r, px = qx0 - HALO + q;
    const bool ok = py >= 0 && py < L.vh && px >= 0 && px < L.vw;
    const float* mp = maps + L.map_off + (size_t)c * plane + (size_t)py * L.vw + px;
    sm[0][r][q] = ok ? mp[0] : 0.f;
    sm[1][r][q] = ok ? mp[(size_t)P.C * plane] : 0.f;
    sm[2][r][q] = ok ? mp[2 * (size_t)P.C * plane] : 0.f;
  }
  __syncthreads();
  // horizontal: out(qx) = sum_t g[t] map(qx - t) = sum_t g[t] sm[.][r][qx_local + 10 - t]
  for (int i = tid; i < (TY + HALO) * TX; i += TX * TY) {
    const int r = i / TX, q = i - r * TX;
    float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
    for (int t = 0; t < WIN; ++t) {
      const float w = c_win[t];
      a = fmaf(w, sm[0][r][q + HALO - t], a); b = fmaf(w, sm[1][r][q + HALO - t], b); d = fmaf(w, sm[2][r][q + HALO - t], d);
    }
    hv[0][r][q] = a; hv[1][r][q] = b; hv[2][r][q] = d;
  }
  __syncthreads();
  const int qy = qy0 + ty, qx = qx0 + tx;
  if (qy < L.H && qx < L.W) {
    float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
    for (int t = 0; t < WIN; ++t) {
      const float w = c_win[t];
      a = fmaf(w, hv[0][ty + HALO - t][tx], a);
      b = fmaf(w, hv[1][ty + HALO - t][tx], b);
      d = fmaf(w, hv[2][ty + HALO - t][tx], d);
    }
    const size_t o = (size_t)c * L.H * L.W + (size_t)qy * L.W + qx;
    const float xv = pyr[L.img_off + o], yv = pyr[P.pyr_half + L.img_off + o];
    lgrad[L.grad_off + o] = kcoef[l * P.C + c] * (a + 2.f * xv * b + yv * d);
  }
}

__global__ void __launch_bounds__(256) ms_up_kernel(const float* __restrict__ lgrad, Pyramid P, int accumulate, float* __restrict__ grad) {
  const Level& L0 = P.lv[0];
  const size_t n = (size_t)P.C * L0.H * L0.W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / ((size_t)L0.H * L0.W));
    const int r = (int)(i - (size_t)c * L0.H * L0.W);
    const int y = r / L0.W, x = r - y * L0.W;
    float g = 0.f, f = 1.f;
#pragma unroll
    for (int l = 0; l < LEVELS; ++l) {
      const Level& L = P.lv[l];
      g = fmaf(f, lgrad[L.grad_off + ((size_t)c * L.H + (y >> l)) * L.W + (x >> l)], g);
      f *= 0.25f;
    }
    grad[i] = (accumulate ? grad[i] : 0.f) + 0.5f * g;  // d((x+1)/2)/dx
  }
}

int build_pyramid(int C, int H, int W, Pyramid* P, size_t* total_floats, int* total_partials) {
  CG_REQUIRE(C >= 1 && C <= 8, "cg_ms_ssim: C=%d outside [1, 8]", C);
  CG_REQUIRE(H % 16 == 0 && W % 16 == 0, "cg_ms_ssim: H=%d and W=%d must be multiples of 16 (four exact 2x2 poolings; Config.update floors both to multiples of 64)", H, W);
  CG_REQUIRE((H >> 4) >= WIN && (W >> 4) >= WIN && (H < W ? H : W) > HALO * 16, "cg_ms_ssim: the smaller side must exceed %d (pytorch_msssim's own check)", HALO * 16);
  P->C = C;
  size_t img = 0, maps = 0, grads = 0;
  int parts = 0;
  for (int l = 0; l < LEVELS; ++l) {
    Level& L = P->lv[l];
    L.H = H >> l; L.W = W >> l; L.vh = L.H - HALO; L.vw = L.W - HALO;
    L.img_off = img; img += (size_t)C * L.H * L.W;
    L.blocks_per_c = ((L.vw + TX - 1) / TX) * ((L.vh + TY - 1) / TY);
    L.part_off = parts; parts += C * L.blocks_per_c;
  }
  P->pyr_half = img;
  // layout: [X pyramid | Y pyramid | coefficient maps | level gradients]
  size_t o = 2 * img;
  for (int l = 0; l < LEVELS; ++l) { P->lv[l].map_off = o - 0; o += 3 * (size_t)C * P->lv[l].vh * P->lv[l].vw; }
  for (int l = 0; l < LEVELS; ++l) { P->lv[l].grad_off = o; o += (size_t)C * P->lv[l].H * P->lv[l].W; }
  (void)maps; (void)grads;
  *total_floats = o + 64;
  *total_partials = parts;
  return 0;
}

}  // namespace

extern "C" size_t cg_ms_ssim_workspace_bytes(int C, int H, int W) {
  Pyramid P;
  size_t floats = 0;
  int parts = 0;
  if (build_pyramid(C, H, W, &P, &floats, &parts)) return 0;
  return sizeof(float) * (floats + (size_t)parts + LEVELS * 8);
}

extern "C" int cg_ms_ssim_dissimilarity_fwd_bwd(const float* x, const float* y, int C, int H, int W, float grad_scale, int accumulate, float* loss,
                                                float* grad, void* workspace, void* stream) {
  CG_REQUIRE(x && y && workspace && (loss || grad), "cg_ms_ssim_dissimilarity_fwd_bwd: null pointer");
  Pyramid P;
  size_t floats = 0;
  int parts = 0;
  int rc = build_pyramid(C, H, W, &P, &floats, &parts);
  if (rc) return rc;
  static bool win_ready = false;
  if (!win_ready) {  // pytorch_msssim._fspecial_gauss_1d(11, 1.5) in float32
    float w[WIN], s = 0.f;
    for (int i = 0; i < WIN; ++i) { const float d = (float)(i - WIN / 2); w[i] = expf(-(d * d) / (2.f * 1.5f * 1.5f)); s += w[i]; }
    for (int i = 0; i < WIN; ++i) w[i] /= s;
    CG_CUDA(cudaMemcpyToSymbol(c_win, w, sizeof(w)));
    win_ready = true;
  }
  cudaStream_t s = cg_stream(stream);
  float* ws = reinterpret_cast<float*>(workspace);
  float* pyr = ws;            // pyramids, maps and level gradients share one offset space
  float* partials = ws + floats;
  float* kcoef = partials + parts;
  auto blocks = [](size_t n) { size_t b = (n + 255) / 256; return (unsigned)(b > 148 * 8 ? 148 * 8 : b); };
  ms_pool_kernel<<<blocks(2 * (size_t)C * H * W), 256, 0, s>>>(x, y, pyr, P, -1);
  CG_LAUNCH_CHECK();
  for (int l = 0; l + 1 < LEVELS; ++l) {
    ms_pool_kernel<<<blocks(2 * (size_t)C * P.lv[l + 1].H * P.lv[l + 1].W), 256, 0, s>>>(nullptr, nullptr, pyr, P, l);
    CG_LAUNCH_CHECK();
  }
  for (int l = 0; l < LEVELS; ++l) {
    const Level& L = P.lv[l];
    ms_maps_kernel<<<dim3((L.vw + TX - 1) / TX, (L.vh + TY - 1) / TY, C), dim3(TX * TY), 0, s>>>(pyr, pyr, partials, P, l);
    CG_LAUNCH_CHECK();
  }
  ms_scalars_kernel<<<1, 256, 0, s>>>(partials, P, grad_scale, loss, kcoef);
  CG_LAUNCH_CHECK();
  if (grad) {
    for (int l = 0; l < LEVELS; ++l) {
      const Level& L = P.lv[l];
      ms_grad_kernel<<<dim3((L.W + TX - 1) / TX, (L.H + TY - 1) / TY, C), dim3(TX * TY), 0, s>>>(pyr, pyr, kcoef, pyr, P, l);
      CG_LAUNCH_CHECK();
    }
    ms_up_kernel<<<blocks((size_t)C * H * W), 256, 0, s>>>(pyr, P, accumulate, grad);
    CG_LAUNCH_CHECK();
  }
  return 0;
}
