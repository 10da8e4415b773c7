// Single-pass loss kernels: value + analytic gradient (clip_diffusion/losses.py:10-35).
// HBM-bound: read x once (neighbours come from L1/L2), write grad once.
#include "common.cuh"

namespace {

// Deterministic loss values: the blocks of one image are a contiguous range of the linear block index; the last block of the grid adds
// each image's partials in index order (common.cuh) -- no float atomics.
__device__ __forceinline__ void per_image_loss(float block_total, float inv, float* loss, const CgScratch& ws, int counter, unsigned block_linear,
                                               unsigned nblocks, unsigned blocks_per_image, unsigned images, float* red) {
  if (cg_last_block(block_total, ws.partials, ws.counters + counter, block_linear, nblocks)) {
    if (images == 1) {
      const float t = cg_sum_partials(ws.partials, 0, blocks_per_image, red);
      if (threadIdx.x == 0) loss[0] = t * inv;
    } else {  // one warp per image (fixed lane-strided order: still deterministic), the warps of the block work in parallel
      for (unsigned b = threadIdx.x >> 5; b < images; b += blockDim.x >> 5) {
        const float t = cg_sum_partials_warp(ws.partials, b * blocks_per_image, blocks_per_image);
        if ((threadIdx.x & 31) == 0) loss[b] = t * inv;
      }
    }
  }
}

// ---- total variation (losses.py:20-28) ---------------------------------------------------------
// One thread per 4 consecutive pixels of a row.  VEC path (W % 4 == 0, 16-byte aligned): three 128-bit loads (row
// above / this row / row below) + two scalar halo loads per 4 outputs; neighbours come from L1/L2, HBM sees each pixel once.
template <bool VEC>
__global__ void __launch_bounds__(256) tv_kernel(const float* __restrict__ x, int C, int H, int W, float gscale,
                                                 int accumulate, float* __restrict__ loss, float* __restrict__ grad, CgScratch ws) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  const int64_t plane = (int64_t)H * W;
  const int64_t per_img = plane * C;
  const float* xb = x + b * per_img;
  float* gb = grad ? grad + b * per_img : nullptr;
  const float inv = 1.0f / (float)per_img;
  const float gs = 2.0f * gscale * inv;
  const int W4 = (W + 3) >> 2;
  const int64_t nvec = (int64_t)C * H * W4;
  float acc = 0.f;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    const int xv = (int)(v % W4);
    const int64_t row = v / W4;  // c*H + h
    const int h = (int)(row % H);
    const int x0 = xv << 2;
    const float* r = xb + row * W;
    float cur[6], up[4], dn[4];  // cur = x[x0-1 .. x0+4]
    if (VEC) {
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(r + x0));
      cur[1] = c4.x; cur[2] = c4.y; cur[3] = c4.z; cur[4] = c4.w;
      cur[0] = x0 > 0 ? __ldg(r + x0 - 1) : 0.f;
      cur[5] = x0 + 4 < W ? __ldg(r + x0 + 4) : 0.f;
      float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = u4;
      if (h > 0) u4 = __ldg(reinterpret_cast<const float4*>(r - W + x0));
      if (h + 1 < H) d4 = __ldg(reinterpret_cast<const float4*>(r + W + x0));
      up[0] = u4.x; up[1] = u4.y; up[2] = u4.z; up[3] = u4.w;
      dn[0] = d4.x; dn[1] = d4.y; dn[2] = d4.z; dn[3] = d4.w;
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int xx = x0 - 1 + i;
        cur[i] = (xx >= 0 && xx < W) ? __ldg(r + xx) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int xx = x0 + i;
        const bool ok = xx < W;
        up[i] = (ok && h > 0) ? __ldg(r - W + xx) : 0.f;
        dn[i] = (ok && h + 1 < H) ? __ldg(r + W + xx) : 0.f;
      }
    }
    float g[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int xx = x0 + i;
      if (xx >= W) { g[i] = 0.f; continue; }
      const float c = cur[i + 1];
      const float dx = (xx + 1 < W) ? cur[i + 2] - c : 0.f;   // replicate pad => 0 at the last column
      const float dy = (h + 1 < H) ? dn[i] - c : 0.f;
      const float dxl = (xx > 0) ? c - cur[i] : 0.f;           // dx of the left neighbour
      const float dyu = (h > 0) ? c - up[i] : 0.f;             // dy of the upper neighbour
      acc += dx * dx + dy * dy;
      g[i] = gs * (dxl + dyu - dx - dy);
    }
    if (gb) {
      float* o = gb + row * W + x0;
      if (VEC) {
        float4 w = make_float4(g[0], g[1], g[2], g[3]);
        if (accumulate) { const float4 p = *reinterpret_cast<float4*>(o); w.x += p.x; w.y += p.y; w.z += p.z; w.w += p.w; }
        *reinterpret_cast<float4*>(o) = w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (x0 + i < W) o[i] = (accumulate ? o[i] : 0.f) + g[i];
      }
    }
  }
  if (loss) per_image_loss(block_sum(acc, red), inv, loss, ws, 1, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, gridDim.x, gridDim.y, red);
}

// 2-D tiled variant (W % 128 == 0): a block owns a 32-row x 128-column tile of one plane; warp w walks rows w, w+8, ...
// so the rows above/below a warp's row are the rows its neighbour warps load at the same time (L1 hits) -- the row-major
// grid-stride kernel above re-read every row three times from L2 and stalled at ~50% of HBM peak.
__global__ void __launch_bounds__(256) tv_tiled_kernel(const float* __restrict__ x, int C, int H, int W, float gscale, int accumulate,
                                                       float* __restrict__ loss, float* __restrict__ grad, CgScratch ws) {
  __shared__ float red[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int plane_id = blockIdx.z;  // b*C + c
  const int b = plane_id / C;
  const int64_t per_img = (int64_t)C * H * W;
  const float inv = 1.0f / (float)per_img;
  const float gs = 2.0f * gscale * inv;
  const float* xp = x + (int64_t)plane_id * H * W;
  float* gp = grad ? grad + (int64_t)plane_id * H * W : nullptr;
  const int x0 = blockIdx.x * 128 + lane * 4;
  const int h0 = blockIdx.y * 32;
  float acc = 0.f;
#pragma unroll 1
  for (int r = warp; r < 32; r += 8) {
    const int h = h0 + r;
    if (h >= H) break;
    const float* row = xp + (int64_t)h * W;
    const float4 c4 = __ldg(reinterpret_cast<const float4*>(row + x0));
    const float left = x0 > 0 ? __ldg(row + x0 - 1) : 0.f;
    const float right = x0 + 4 < W ? __ldg(row + x0 + 4) : 0.f;
    float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = u4;
    if (h > 0) u4 = __ldg(reinterpret_cast<const float4*>(row - W + x0));
    if (h + 1 < H) d4 = __ldg(reinterpret_cast<const float4*>(row + W + x0));
    const float cur[6] = {left, c4.x, c4.y, c4.z, c4.w, right};
    const float up[4] = {u4.x, u4.y, u4.z, u4.w}, dn[4] = {d4.x, d4.y, d4.z, d4.w};
    float g[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int xx = x0 + i;
      const float c = cur[i + 1];
      const float dx = (xx + 1 < W) ? cur[i + 2] - c : 0.f;
      const float dy = (h + 1 < H) ? dn[i] - c : 0.f;
      const float dxl = (xx > 0) ? c - cur[i] : 0.f;
      const float dyu = (h > 0) ? c - up[i] : 0.f;
      acc += dx * dx + dy * dy;
      g[i] = gs * (dxl + dyu - dx - dy);
    }
    if (gp) {
      float4* o = reinterpret_cast<float4*>(gp + (int64_t)h * W + x0);
      float4 w = make_float4(g[0], g[1], g[2], g[3]);
      if (accumulate) { const float4 p = *o; w.x += p.x; w.y += p.y; w.z += p.z; w.w += p.w; }
      *o = w;
    }
  }
  if (loss)
    per_image_loss(block_sum(acc, red), inv, loss, ws, 2, (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x, gridDim.x * gridDim.y * gridDim.z,
                   gridDim.x * gridDim.y * C, gridDim.z / C, red);
  (void)b;
}

// ---- rgb range (losses.py:31-35) -----------------------------------------------------------------
__global__ void __launch_bounds__(256) range_kernel(const float* __restrict__ x, int64_t per_img, float gscale, int accumulate,
                                                    float* __restrict__ loss, float* __restrict__ grad, CgScratch ws) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  const float* xb = x + b * per_img;
  float* gb = grad ? grad + b * per_img : nullptr;
  const float inv = 1.0f / (float)per_img;
  const float gs = 2.0f * gscale * inv;
  float acc = 0.f;
  const int64_t n4 = per_img >> 2;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(xb) + v);
    float e[4] = {p.x, p.y, p.z, p.w};
    float g[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float d = e[i] - fminf(fmaxf(e[i], -1.f), 1.f);
      acc += d * d;
      g[i] = gs * d;
    }
    if (gb) {
      float4 w = make_float4(g[0], g[1], g[2], g[3]);
      float4* o = reinterpret_cast<float4*>(gb) + v;
      if (accumulate) { const float4 q = *o; w.x += q.x; w.y += q.y; w.z += q.z; w.w += q.w; }
      *o = w;
    }
  }
  // tail (per_img not a multiple of 4)
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = xb[i] - fminf(fmaxf(xb[i], -1.f), 1.f);
    acc += d * d;
    if (gb) gb[i] = (accumulate ? gb[i] : 0.f) + gs * d;
  }
  if (loss) per_image_loss(block_sum(acc, red), inv, loss, ws, 3, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, gridDim.x, gridDim.y, red);
}

// ---- fused image losses: TV + range value and gradient in ONE pass, plus the NaN flag of the finished guidance gradient -------------
// (sample.py:217-228: tv_loss * denoise_scale [+ range_loss * range_scale] is differentiated, added to grad_tensor and the sum is tested
// for NaN.)  W % 128 == 0: a block owns a 32-row x 128-column tile of one plane, like tv_tiled_kernel.  grad (+)= d(tv_scale * TV +
// range_scale * range)/dx; flag[0] = 1 if any element of the finished grad is NaN (the caller zeroes flag[0..1] is NOT needed: the last
// block writes it).  loss2 (optional) receives [B][2] = (TV, range) values, deterministic.
template <int TR>  // rows per tile: 32 for big batches, 8 (one row per warp, no loop: all loads of the CTA in flight at once) when the grid would not fill the GPU
__global__ void __launch_bounds__(256) image_losses_kernel(const float* __restrict__ x, int C, int H, int W, float tv_scale, float range_scale,
                                                           int accumulate, float* __restrict__ loss2, float* __restrict__ grad, float* __restrict__ flag,
                                                           CgScratch ws) {
  __shared__ float red[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int plane_id = blockIdx.z;  // b*C + c
  const int64_t per_img = (int64_t)C * H * W;
  const float inv = 1.0f / (float)per_img;
  const float gt = 2.0f * tv_scale * inv, gr = 2.0f * range_scale * inv;
  const float* xp = x + (int64_t)plane_id * H * W;
  float* gp = grad + (int64_t)plane_id * H * W;
  const int x0 = blockIdx.x * 128 + lane * 4;
  const int h0 = blockIdx.y * TR;
  float acc_tv = 0.f, acc_rg = 0.f, bad = 0.f;
#pragma unroll 1
  for (int r = warp; r < TR; r += 8) {
    const int h = h0 + r;
    if (h >= H) break;
    const float* row = xp + (int64_t)h * W;
    const float4 c4 = __ldg(reinterpret_cast<const float4*>(row + x0));
    const float left = x0 > 0 ? __ldg(row + x0 - 1) : 0.f;
    const float right = x0 + 4 < W ? __ldg(row + x0 + 4) : 0.f;
    float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = u4;
    if (h > 0) u4 = __ldg(reinterpret_cast<const float4*>(row - W + x0));
    if (h + 1 < H) d4 = __ldg(reinterpret_cast<const float4*>(row + W + x0));
    const float cur[6] = {left, c4.x, c4.y, c4.z, c4.w, right};
    const float up[4] = {u4.x, u4.y, u4.z, u4.w}, dn[4] = {d4.x, d4.y, d4.z, d4.w};
    float4* o = reinterpret_cast<float4*>(gp + (int64_t)h * W + x0);
    float prev[4] = {0.f, 0.f, 0.f, 0.f};
    if (accumulate) { const float4 p = *o; prev[0] = p.x; prev[1] = p.y; prev[2] = p.z; prev[3] = p.w; }
    float g[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int xx = x0 + i;
      const float c = cur[i + 1];
      const float dx = (xx + 1 < W) ? cur[i + 2] - c : 0.f;
      const float dy = (h + 1 < H) ? dn[i] - c : 0.f;
      const float dxl = (xx > 0) ? c - cur[i] : 0.f;
      const float dyu = (h > 0) ? c - up[i] : 0.f;
      acc_tv += dx * dx + dy * dy;
      const float d = c - fminf(fmaxf(c, -1.f), 1.f);
      acc_rg += d * d;
      g[i] = prev[i] + gt * (dxl + dyu - dx - dy) + gr * d;
      if (g[i] != g[i]) bad = 1.f;
    }
    *o = make_float4(g[0], g[1], g[2], g[3]);
  }
  // NaN flag: order independent (any block that saw one stores "bad"); flag[1] is the raw indicator, flag[0] the 0/1 flag of the whole grid
  const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
  const unsigned lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const float tb = block_sum(bad, red);
  const float t_tv = block_sum(acc_tv, red), t_rg = block_sum(acc_rg, red);
  if (threadIdx.x == 0) {
    ws.partials[nblocks + lin] = t_rg;
    ws.partials[2 * nblocks + lin] = tb;
  }
  if (cg_last_block(t_tv, ws.partials, ws.counters + 4, lin, nblocks)) {
    const float anybad = cg_sum_partials(ws.partials, 2 * nblocks, nblocks, red);
    if (threadIdx.x == 0) { flag[0] = anybad > 0.f ? 1.f : 0.f; flag[1] = anybad; }
    if (loss2) {
      const unsigned per = gridDim.x * gridDim.y * C, images = gridDim.z / C;
      for (unsigned b = threadIdx.x >> 5; b < images; b += blockDim.x >> 5) {
        const float a = cg_sum_partials_warp(ws.partials, b * per, per);
        const float c = cg_sum_partials_warp(ws.partials, nblocks + b * per, per);
        if ((threadIdx.x & 31) == 0) { loss2[2 * b] = a * inv; loss2[2 * b + 1] = c * inv; }
      }
    }
  }
}

// ---- squared spherical distance (losses.py:10-16) -----------------------------------------------
// One warp per (n, p) pair for fwd / per n for bwd.  E <= 4096.
__device__ __forceinline__ float sph_from_r(float r) {
  const float a = asinf(0.5f * r);
  return 2.f * a * a;
}
// d dist / d r  divided by r (r > 0)
__device__ __forceinline__ float sph_dr_over_r(float r) {
  const float h = 0.5f * r;
  const float a = asinf(h);
  return 2.f * a / (sqrtf(fmaxf(1.f - h * h, 0.f)) * r);
}

__global__ void __launch_bounds__(128) sph_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ txt, int N, int P, int E,
                                                      float* __restrict__ dist) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N * P) return;
  const int n = warp / P, p = warp % P;
  const float* x = emb + (int64_t)n * E;
  const float* y = txt + (int64_t)p * E;
  float sx = 0.f, sy = 0.f;
  for (int e = lane; e < E; e += 32) { sx += x[e] * x[e]; sy += y[e] * y[e]; }
  sx = warp_sum(sx); sy = warp_sum(sy);
  const float ix = 1.f / fmaxf(sqrtf(sx), 1e-12f), iy = 1.f / fmaxf(sqrtf(sy), 1e-12f);
  float r2 = 0.f;
  for (int e = lane; e < E; e += 32) { const float u = x[e] * ix - y[e] * iy; r2 += u * u; }
  r2 = warp_sum(r2);
  if (lane == 0) dist[warp] = sph_from_r(sqrtf(r2));
}

// demb[n,:] = sum_p g[n,p] * d dist[n,p] / d emb[n,:];  g given explicitly (gdist) or as coef*w[p].
// Optionally accumulates loss = sum g*dist into loss_out (fused form).
__global__ void __launch_bounds__(128) sph_bwd_kernel(const float* __restrict__ emb, const float* __restrict__ txt,
                                                      const float* __restrict__ gdist, const float* __restrict__ w, float coef,
                                                      int N, int P, int E, float* __restrict__ demb, float* __restrict__ loss_out, CgScratch ws) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* x = emb + (int64_t)n * E;
  float sx = 0.f;
  for (int e = lane; e < E; e += 32) sx += x[e] * x[e];
  sx = warp_sum(sx);
  const float nx = sqrtf(sx);
  const float ix = 1.f / fmaxf(nx, 1e-12f);
  const bool clampd = nx < 1e-12f;  // F.normalize's eps branch: x/eps is linear in x
  float* o = demb + (int64_t)n * E;
  for (int e = lane; e < E; e += 32) o[e] = 0.f;
  float lsum = 0.f;
  for (int p = 0; p < P; ++p) {
    const float* y = txt + (int64_t)p * E;
    float sy = 0.f;
    for (int e = lane; e < E; e += 32) sy += y[e] * y[e];
    sy = warp_sum(sy);
    const float iy = 1.f / fmaxf(sqrtf(sy), 1e-12f);
    float r2 = 0.f, xu = 0.f;
    for (int e = lane; e < E; e += 32) {
      const float xh = x[e] * ix;
      const float u = xh - y[e] * iy;
      r2 += u * u;
      xu += xh * u;
    }
    r2 = warp_sum(r2); xu = warp_sum(xu);
    const float r = sqrtf(r2);
    const float g = gdist ? gdist[(int64_t)n * P + p] : coef * (w ? w[p] : 1.f);
    lsum += g * sph_from_r(r);
    if (r != 0.f) {  // NaN must propagate (the reference's NaN guard, sample.py:228, relies on it)
      const float k = g * sph_dr_over_r(r) * ix;
      for (int e = lane; e < E; e += 32) {
        const float xh = x[e] * ix;
        const float u = xh - y[e] * iy;
        o[e] += k * (clampd ? u : (u - xh * xu));
      }
    }
  }
  if (loss_out && lane == 0) ws.partials[n] = lsum;  // summed in index order by sph_loss_sum_kernel
}

__global__ void __launch_bounds__(256) sph_loss_sum_kernel(CgScratch ws, int N, float* __restrict__ loss_out) {
  __shared__ float red[32];
  const float t = cg_sum_partials(ws.partials, 0, (unsigned)N, red);
  if (threadIdx.x == 0) loss_out[0] = t;
}

}  // namespace

static int grid_for(int64_t work_items, int threads) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)CG_NUM_SMS * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

extern "C" int cg_tv_loss_fwd_bwd(const float* x, int B, int C, int H, int W, float grad_scale, int accumulate, float* loss,
                                  float* grad, void* stream) {
  CG_REQUIRE(x && B > 0 && C > 0 && H > 0 && W > 0, "cg_tv_loss_fwd_bwd: bad arguments");
  cudaStream_t s = cg_stream(stream);
  CgScratch ws;
  int rc = cg_get_scratch(&ws);
  if (rc) return rc;
  const int64_t nvec = (int64_t)C * H * ((W + 3) / 4);
  dim3 grid(grid_for(nvec, 256), B);
  const bool vec = (W % 4 == 0) && (((uintptr_t)x | (uintptr_t)grad) & 15) == 0;
  const long long tiles = (long long)(W / 128) * ((H + 31) / 32) * B * C;
  if (vec && W % 128 == 0 && (long long)B * C <= 65535 && tiles <= CG_SCRATCH_FLOATS) {
    tv_tiled_kernel<<<dim3(W / 128, (H + 31) / 32, B * C), 256, 0, s>>>(x, C, H, W, grad_scale, accumulate, loss, grad, ws);
    CG_LAUNCH_CHECK();
    return 0;
  }
  if (vec) tv_kernel<true><<<grid, 256, 0, s>>>(x, C, H, W, grad_scale, accumulate, loss, grad, ws);
  else tv_kernel<false><<<grid, 256, 0, s>>>(x, C, H, W, grad_scale, accumulate, loss, grad, ws);
  CG_LAUNCH_CHECK();
  return 0;
}

extern "C" int cg_range_loss_fwd_bwd(const float* x, int B, int C, int H, int W, float grad_scale, int accumulate, float* loss,
                                     float* grad, void* stream) {
  CG_REQUIRE(x && B > 0 && C > 0 && H > 0 && W > 0, "cg_range_loss_fwd_bwd: bad arguments");
  const int64_t per_img = (int64_t)C * H * W;
  CG_REQUIRE((per_img & 3) == 0 || B == 1, "cg_range_loss_fwd_bwd: C*H*W must be a multiple of 4 when B > 1");
  CG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)grad & 15) == 0, "cg_range_loss_fwd_bwd: pointers must be 16-byte aligned");
  cudaStream_t s = cg_stream(stream);
  CgScratch ws;
  int rc = cg_get_scratch(&ws);
  if (rc) return rc;
  // batches: keep the whole grid at ~8 CTAs per SM (the last block's ordered sum grows with the block count)
  int gx = grid_for(per_img / 4 + 1, 256);
  if (B > 1) gx = max(1, min(gx, (8 * 148 * 2 + B - 1) / B));
  dim3 grid(gx, B);
  range_kernel<<<grid, 256, 0, s>>>(x, per_img, grad_scale, accumulate, loss, grad, ws);
  CG_LAUNCH_CHECK();
  return 0;
}

extern "C" int cg_spherical_dist_fwd(const float* emb, const float* txt, int N, int P, int E, float* dist, void* stream) {
  CG_REQUIRE(emb && txt && dist && N > 0 && P > 0 && E > 0, "cg_spherical_dist_fwd: bad arguments");
  const int warps = N * P;
  sph_fwd_kernel<<<(warps + 3) / 4, 128, 0, cg_stream(stream)>>>(emb, txt, N, P, E, dist);
  CG_LAUNCH_CHECK();
  return 0;
}

extern "C" int cg_spherical_dist_bwd(const float* emb, const float* txt, const float* gdist, int N, int P, int E, float* demb,
                                     void* stream) {
  CG_REQUIRE(emb && txt && gdist && demb && N > 0 && P > 0 && E > 0, "cg_spherical_dist_bwd: bad arguments");
  sph_bwd_kernel<<<(N + 3) / 4, 128, 0, cg_stream(stream)>>>(emb, txt, gdist, nullptr, 0.f, N, P, E, demb, nullptr, CgScratch{});
  CG_LAUNCH_CHECK();
  return 0;
}

extern "C" int cg_spherical_loss_fwd_bwd(const float* emb, const float* txt, const float* w, int N, int P, int E, float coef,
                                         float* loss_out, float* demb, void* stream) {
  CG_REQUIRE(emb && txt && demb && N > 0 && P > 0 && E > 0, "cg_spherical_loss_fwd_bwd: bad arguments");
  CgScratch ws = {};
  if (loss_out) {
    CG_REQUIRE(N <= CG_SCRATCH_FLOATS, "cg_spherical_loss_fwd_bwd: N=%d too large for the loss value (pass loss_out = NULL)", N);
    int rc = cg_get_scratch(&ws);
    if (rc) return rc;
  }
  sph_bwd_kernel<<<(N + 3) / 4, 128, 0, cg_stream(stream)>>>(emb, txt, nullptr, w, coef, N, P, E, demb, loss_out, ws);
  CG_LAUNCH_CHECK();
  if (loss_out) {  // deterministic value: per-row contributions added in index order
    sph_loss_sum_kernel<<<1, 256, 0, cg_stream(stream)>>>(ws, N, loss_out);
    CG_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int cg_image_losses_fwd_bwd(const float* x, int B, int C, int H, int W, float tv_scale, float range_scale, int accumulate,
                                       float* loss2, float* grad, float* nan_flag, void* stream) {
  CG_REQUIRE(x && grad && nan_flag && B > 0 && C > 0 && H > 0 && W > 0, "cg_image_losses_fwd_bwd: bad arguments");
  CG_REQUIRE(W % 128 == 0 && (((uintptr_t)x | (uintptr_t)grad) & 15) == 0, "cg_image_losses_fwd_bwd: W=%d must be a multiple of 128 and the pointers 16-byte aligned (use cg_tv_loss_fwd_bwd / cg_range_loss_fwd_bwd / cg_any_nan otherwise)", W);
  long long tiles = (long long)(W / 128) * ((H + 31) / 32) * B * C;
  const bool small = tiles < 4 * 148;
  if (small) tiles = (long long)(W / 128) * ((H + 7) / 8) * B * C;
  CG_REQUIRE((long long)B * C <= 65535 && 3 * tiles <= CG_SCRATCH_FLOATS, "cg_image_losses_fwd_bwd: too many tiles (%lld)", tiles);
  CgScratch ws;
  int rc = cg_get_scratch(&ws);
  if (rc) return rc;
  if (small)
    image_losses_kernel<8><<<dim3(W / 128, (H + 7) / 8, B * C), 256, 0, cg_stream(stream)>>>(x, C, H, W, tv_scale, range_scale, accumulate, loss2, grad, nan_flag, ws);
  else
    image_losses_kernel<32><<<dim3(W / 128, (H + 31) / 32, B * C), 256, 0, cg_stream(stream)>>>(x, C, H, W, tv_scale, range_scale, accumulate, loss2, grad, nan_flag, ws);
  CG_LAUNCH_CHECK();
  return 0;
}
