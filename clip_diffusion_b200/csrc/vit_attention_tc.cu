// Fused multi-head self-attention of the CLIP ViT on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), forward AND
// backward, for the short sequences of the 224-pixel towers (T <= 272: ViT-B/32 T=50, ViT-B/16 T=197, ViT-L/14 T=257).
// Replaces nn.MultiheadAttention inside clip_model.encode_image (clip_diffusion/utils/functional.py:97-102) and its
// autograd backward (sample.py:199-214).  Scores never leave the SM.
//
// One CTA per (image, head); everything the head needs is loaded ONCE by TMA into 128B-swizzled shared memory
// (forward: Q, K, V; backward: Q, K, V, dO; <= 272 x 64 bf16 each), one mbarrier per 64-row block so the first MMAs
// start while the rest is still in flight.  The work is cut into BLOCKS of 128 rows x 64 columns of the score matrix
// and streamed through a pipeline with double-buffered TMEM and staging buffers:
//
//   warp 1 (one thread)   tcgen05.mma   S_blk = A_tile B_blk^T (and dP_blk in the backward) into TMEM buffer b = g & 1,
//                                       then, when the block's staged operand is ready, the accumulator MMA
//                                       (O += P V | dQ += dS K | dV += P^T dO, dK += dS^T Q) with the B operand in MN-major form
//   warps 4-7 / 8-11      softmax       two warpgroups, block g goes to warpgroup g & 1; a thread owns one row
//                                       (= one TMEM lane): tcgen05.ld, exp2 / dS arithmetic in registers, bf16 result
//                                       written to the staging buffer in the K-major 128B-swizzled operand layout
//   warps 12-15           epilogue      tcgen05.ld of the finished accumulator tile (double buffered) -> bf16 -> HBM
//   warp 0                TMA producer, warp 2 TMEM allocation
//
// so the tensor pipe computes block g+1 (and the accumulator MMA of block g-1) while a warpgroup is busy with the
// MUFU-bound exponentials of block g.  All smem / TMEM operand forms are the ones vit_gemm.cu and the round-1 forward
// validated: K-major SW128 A and B, MN-major SW128 B with N = 64.
//
//   forward   per 128-query tile: pass 0 streams S blocks for the row maximum, pass 1 recomputes them (the tensor pipe is
//             idle otherwise), P = exp2((s - m) c) and O accumulates in TMEM with no rescaling; epilogue O / l, lse.
//   backward  phase A (lane = query):  S = Q K^T, dP = dO V^T, dS = P (dP - delta) scale, dQ += dS K
//             phase B (lane = key):    S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
//             P is recomputed from the saved log-sum-exp; two orientations instead of a transposed smem operand, no atomics,
//             deterministic.  Rows / columns beyond T are masked (tail block only) or never stored.
#include <stdlib.h>
#include "common.cuh"
#include "tcgen05.cuh"

using namespace tc;

namespace {

constexpr int AT_THREADS = 512;
constexpr int MAX_BLK = 5;  // 64-row blocks per operand: T <= 272 -> <= 5
constexpr float LOG2E_F = 1.4426950408889634f;
constexpr uint32_t BLK_BYTES = 64 * 128;     // one 64-row operand block
constexpr uint32_t TILE_BYTES = 128 * 128;   // one 128-row tile of an operand = one [128 x 64] staging tile

__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// instruction descriptor, kind::f16: D fp32 (bit 4), A/B bf16 (bits 7, 10), b_major (bit 16: MN-major B), N>>3 at 17, M>>4 at 24
__device__ __forceinline__ uint32_t idesc_bf16(int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void st_shared_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// 32 consecutive columns [c0, c0+32) (c0 = 0 or 32) of row r of a [128 x 64] bf16 K-major 128B-swizzled tile: w = 16 packed pairs
__device__ __forceinline__ void stage_store32(uint32_t tile_base, int r, int c0, const uint32_t* w) {
  const uint32_t rowb = tile_base + (uint32_t)r * 128u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t chunk = (uint32_t)(((c0 >> 3) + j) ^ (r & 7));
    st_shared_v4(rowb + (chunk << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
  }
}
__device__ __forceinline__ void st256g(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]),
               "r"(w[6]), "r"(w[7])
               : "memory");
}
// 64 fp32 accumulator values of one row (two tcgen05.ld x32) * mul -> 64 bf16 = 128 contiguous bytes
__device__ __forceinline__ void store_row64_bf16(__nv_bfloat16* dst, const uint32_t* a, const uint32_t* b, float mul) {
  uint32_t w[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    w[j] = pack_bf2(__uint_as_float(a[2 * j]) * mul, __uint_as_float(a[2 * j + 1]) * mul);
    w[16 + j] = pack_bf2(__uint_as_float(b[2 * j]) * mul, __uint_as_float(b[2 * j + 1]) * mul);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) st256g(dst + 16 * j, w + 8 * j);
}

// barrier block (8 bytes each)
struct Bars {
  uint32_t base;
  __device__ __forceinline__ uint32_t sfull(int b) const { return base + 8u * b; }          // S (dP) block b is in TMEM           (tcgen05.commit)
  __device__ __forceinline__ uint32_t sfree(int b) const { return base + 16u + 8u * b; }    // warpgroup b has read it             (4 warps)
  __device__ __forceinline__ uint32_t pready(int b) const { return base + 32u + 8u * b; }   // staging buffer b is written          (4 warps)
  __device__ __forceinline__ uint32_t pfree(int b) const { return base + 48u + 8u * b; }    // accumulator MMAs have consumed it    (tcgen05.commit)
  __device__ __forceinline__ uint32_t accfull(int a) const { return base + 64u + 8u * a; }  // accumulator tile a is complete       (tcgen05.commit)
  __device__ __forceinline__ uint32_t accfree(int a) const { return base + 80u + 8u * a; }  // epilogue has read it                 (4 warps)
  __device__ __forceinline__ uint32_t lready(int a) const { return base + 96u + 8u * a; }   // forward: row sums / maxima in smem   (8 warps)
  __device__ __forceinline__ uint32_t op(int o, int blk) const { return base + 112u + 8u * (o * MAX_BLK + blk); }  // operand o, 64-row block
  __device__ __forceinline__ uint32_t tmem_slot() const { return base + 112u + 8u * (4 * MAX_BLK); }
};
constexpr uint32_t BARS_BYTES = 112 + 8 * 4 * MAX_BLK + 16;

// The block stream.  Every role (MMA issuer, the two softmax warpgroups, the epilogue warps) walks the SAME enumeration, so
// buffer indices and mbarrier parities are derived identically everywhere.
//   backward: phase-major  (phase 0 = dQ with lane = query, phase 1 = dK/dV with lane = key) -> tile -> 64-wide block
//   forward:  tile-major   tile -> pass (0 = row maximum, 1 = exponentials + O) -> block
template <bool FWD>
struct BlkIt {
  int phase = 0, tile = 0, blk = 0;
  int g = 0;       // running block index: TMEM buffer = g & 1, use = g >> 1
  int tcount = 0;  // running accumulator-tile index: accumulator buffer = tcount & 1, use = tcount >> 1
  int su0 = 0, su1 = 0;  // staging-buffer use counters (forward: pass-0 blocks do not stage anything)
  __device__ __forceinline__ bool valid(int ntiles) const { return FWD ? tile < ntiles : phase < 2; }
  __device__ __forceinline__ bool stages() const { return !FWD || phase == 1; }
  __device__ __forceinline__ int su() const { return (g & 1) ? su1 : su0; }
  __device__ __forceinline__ void next(int ntiles, int nblk) {
    if (stages()) { if (g & 1) ++su1; else ++su0; }
    ++g;
    if (++blk < nblk) return;
    blk = 0;
    if (FWD) {
      if (phase == 0) { phase = 1; } else { phase = 0; ++tile; ++tcount; }
    } else {
      ++tcount;
      if (++tile == ntiles) { tile = 0; ++phase; }
    }
  }
};

struct AttnParams {
  int T, heads, ntiles, nblk, tail_rows;  // tail_rows: rows of the last 64-row block, rounded up to 16
  float scale;
  __nv_bfloat16* ctx;        // forward out [Nimg*T, D]
  float* lse;                // [Nimg, heads, T]  (forward out, backward in)
  const float* delta;        // backward in [Nimg, heads, T]
  __nv_bfloat16* dqkv;       // backward out [Nimg*T, 3D]
};

__device__ __forceinline__ int blk_cols(const AttnParams& p, int blk) { return blk == p.nblk - 1 ? p.T - blk * 64 : 64; }
__device__ __forceinline__ int blk_cols16(const AttnParams& p, int blk) { return blk == p.nblk - 1 ? p.tail_rows : 64; }

// operands in shared memory: 0 = Q, 1 = K, 2 = V, 3 = dO
template <bool FWD>
__global__ void __launch_bounds__(AT_THREADS, 1)
    attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmQKVtail, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmDOtail, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int NOPS = FWD ? 3 : 4;
  constexpr uint32_t STAGE_BYTES = FWD ? TILE_BYTES : 2 * TILE_BYTES;  // backward phase B stages P^T and dS^T
  const int T = p.T, D = p.heads * 64;
  const int h = blockIdx.x, n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t opb = (uint32_t)((p.nblk - 1) * 64 + p.tail_rows) * 128u;  // bytes per operand buffer (multiple of 2048)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sOp0 = base;
  const uint32_t sStage = base + NOPS * opb;                    // 1024-aligned; also absorbs the A-operand over-read of the last tile
  const uint32_t sF = sStage + 2 * STAGE_BYTES;                 // float scratch
  float* fscr = reinterpret_cast<float*>(smem_raw + (sF - smem_u32(smem_raw)));
  // forward: smax[2][128] (per warpgroup partial row maxima), sl[3][2][128] (tile, warpgroup: partial row sums), smf[3][128] (final row
  // maxima per tile; T <= 272 -> <= 3 tiles, so nothing is ever overwritten while the epilogue may still read it)
  // backward: nlse[320], delta[320]
  constexpr uint32_t FSCR_FLOATS = FWD ? (256 + 768 + 384) : 640;
  Bars bars{sF + FSCR_FLOATS * 4};
  const int row0 = n * T;  // first row of this image in the packed [Nimg*T, 3D] qkv matrix

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKVtail) : "memory");
    if (!FWD) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDO) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDOtail) : "memory");
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars.sfull(b), 1); mbar_init(bars.sfree(b), 4); mbar_init(bars.pready(b), 4); mbar_init(bars.pfree(b), 1);
      mbar_init(bars.accfull(b), 1); mbar_init(bars.accfree(b), 4); mbar_init(bars.lready(b), 8);
    }
    for (int o = 0; o < 4; ++o)
      for (int k = 0; k < MAX_BLK; ++k) mbar_init(bars.op(o, k), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars.tmem_slot()), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (!FWD) {
    // per-query constants of phase B (lane = key, column = query): -lse * log2(e) (-inf beyond T => P = 0 there) and delta
    const float* lb = p.lse + ((long long)n * p.heads + h) * T;
    const float* db = p.delta + ((long long)n * p.heads + h) * T;
    for (int i = threadIdx.x; i < 320; i += AT_THREADS) {
      fscr[i] = i < T ? -lb[i] * LOG2E_F : -INFINITY;
      fscr[320 + i] = i < T ? db[i] : 0.f;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bars.tmem_slot()) : "memory");
  // TMEM columns: S (dP) buffers b at b*128 (+64); accumulators from column 256: forward O[a] at 256 + 64a; backward phase A dQ[a] at
  // 256 + 128a, phase B dV[a] at 256 + 128a and dK[a] at 256 + 128a + 64
  const int ntiles = p.ntiles, nblk = p.nblk;

  if (warp == 0) {
    // ================= TMA producer: every operand block once, in the order the MMA issuer needs them
    if (lane == 0) {
      auto load = [&](int o, int blk) {
        const bool tail = blk == nblk - 1;
        const uint32_t bytes = (uint32_t)(tail ? p.tail_rows : 64) * 128u;
        const uint32_t dst = sOp0 + o * opb + blk * BLK_BYTES;
        const uint32_t bar = bars.op(o, blk);
        mbar_arrive_expect_tx(bar, bytes);
        if (o < 3) tma_load_2d(dst, tail ? &tmQKVtail : &tmQKV, bar, o * D + h * 64, row0 + blk * 64);
        else tma_load_2d(dst, tail ? &tmDOtail : &tmDO, bar, h * 64, row0 + blk * 64);
      };
      const int first = nblk < 2 ? nblk : 2;
      load(1, 0);
      for (int k = 0; k < first; ++k) load(0, k);
      if (FWD) {
        for (int k = 1; k < nblk; ++k) load(1, k);
        for (int k = 0; k < nblk; ++k) load(2, k);
        for (int k = first; k < nblk; ++k) load(0, k);
      } else {
        load(2, 0);
        for (int k = 0; k < first; ++k) load(3, k);
        for (int k = 1; k < nblk; ++k) { load(1, k); load(2, k); }
        for (int k = first; k < nblk; ++k) { load(0, k); load(3, k); }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer
    if (lane == 0) {
      uint32_t loaded = 0;  // operand blocks already seen complete (bit o*MAX_BLK + blk)
      auto need = [&](int o, int blk) {
        const uint32_t bit = 1u << (o * MAX_BLK + blk);
        if (!(loaded & bit)) { mbar_wait(bars.op(o, blk), 0); loaded |= bit; }
      };
      auto need_tile = [&](int o, int tile) { need(o, 2 * tile); if (2 * tile + 1 < nblk) need(o, 2 * tile + 1); };
      const uint32_t id_acc = idesc_bf16(64, true);
      // S (and dP) of one block into TMEM buffer b
      auto issue_s = [&](const BlkIt<FWD>& it) {
        const int b = it.g & 1;
        mbar_wait(bars.sfree(b), (uint32_t)(((it.g >> 1) & 1) ^ 1));
        const bool transposed = !FWD && it.phase == 1;
        const int oa = transposed ? 1 : 0, ob = transposed ? 0 : 1;  // S: A rows (tile) x B rows (block)
        need_tile(oa, it.tile); need(ob, it.blk);
        if (!FWD) { need_tile(transposed ? 2 : 3, it.tile); need(transposed ? 3 : 2, it.blk); }
        tcgen05_fence_after();
        const uint32_t idesc = idesc_bf16(blk_cols16(p, it.blk), false);
        const uint32_t d_s = tmem_base + (uint32_t)(b * 128);
        const uint64_t ad = make_smem_desc(sOp0 + oa * opb + it.tile * TILE_BYTES), bd = make_smem_desc(sOp0 + ob * opb + it.blk * BLK_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d_s, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
        if (!FWD) {
          const int oa2 = transposed ? 2 : 3, ob2 = transposed ? 3 : 2;  // dP = dO V^T | dP^T = V dO^T
          const uint64_t ad2 = make_smem_desc(sOp0 + oa2 * opb + it.tile * TILE_BYTES), bd2 = make_smem_desc(sOp0 + ob2 * opb + it.blk * BLK_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_s + 64u, ad2 + (uint64_t)(2 * k), bd2 + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(bars.sfull(b));
      };
      // accumulator MMAs of one block from staging buffer b
      auto issue_acc = [&](const BlkIt<FWD>& it) {
        if (!it.stages()) return;
        const int b = it.g & 1, a = it.tcount & 1;
        mbar_wait(bars.pready(b), (uint32_t)(it.su() & 1));
        if (it.blk == 0) mbar_wait(bars.accfree(a), (uint32_t)(((it.tcount >> 1) & 1) ^ 1));
        tcgen05_fence_after();
        const int ksteps = blk_cols16(p, it.blk) >> 4;
        const uint32_t stage = sStage + b * STAGE_BYTES;
        if (FWD) {
          const uint32_t d_o = tmem_base + 256u + (uint32_t)(a * 64);
          for (int s = 0; s < ksteps; ++s)
            umma_bf16(d_o, make_smem_desc(stage) + (uint64_t)(2 * s), make_smem_desc(sOp0 + 2 * opb + it.blk * BLK_BYTES + s * 2048), id_acc,
                      (it.blk | s) != 0 ? 1u : 0u);
        } else if (it.phase == 0) {
          const uint32_t d_q = tmem_base + 256u + (uint32_t)(a * 128);
          for (int s = 0; s < ksteps; ++s)  // dQ += dS K
            umma_bf16(d_q, make_smem_desc(stage) + (uint64_t)(2 * s), make_smem_desc(sOp0 + 1 * opb + it.blk * BLK_BYTES + s * 2048), id_acc,
                      (it.blk | s) != 0 ? 1u : 0u);
        } else {
          const uint32_t d_v = tmem_base + 256u + (uint32_t)(a * 128), d_k = d_v + 64u;
          for (int s = 0; s < ksteps; ++s) {  // dV += P^T dO ; dK += dS^T Q
            umma_bf16(d_v, make_smem_desc(stage) + (uint64_t)(2 * s), make_smem_desc(sOp0 + 3 * opb + it.blk * BLK_BYTES + s * 2048), id_acc,
                      (it.blk | s) != 0 ? 1u : 0u);
            umma_bf16(d_k, make_smem_desc(stage + TILE_BYTES) + (uint64_t)(2 * s), make_smem_desc(sOp0 + 0 * opb + it.blk * BLK_BYTES + s * 2048), id_acc,
                      (it.blk | s) != 0 ? 1u : 0u);
          }
        }
        umma_commit(bars.pfree(b));
        if (it.blk == nblk - 1) umma_commit(bars.accfull(a));
      };
      BlkIt<FWD> is, ia;
      for (int k = 0; k < 2 && is.valid(ntiles); ++k) { issue_s(is); is.next(ntiles, nblk); }
      while (ia.valid(ntiles)) {
        if (is.valid(ntiles)) { issue_s(is); is.next(ntiles, nblk); }
        issue_acc(ia);
        ia.next(ntiles, nblk);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ================= softmax warpgroups: block g belongs to warpgroup g & 1; thread = one row (TMEM lane) of the tile
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;
    const int rl = q * 32 + lane;  // row inside the 128-row tile
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sl2 = p.scale * LOG2E_F;
    float rc0 = 0.f, rc1 = 0.f;     // backward phase A: -lse*log2e and delta of this row; forward: m*c and the running sum
    float mx = -INFINITY;            // forward pass 0: running maximum of this warpgroup's blocks
    float* smax = fscr;              // [2][128]
    float* sl = fscr + 256;          // [3][2][128]
    float* smf = fscr + 1024;        // [3][128]
    int cur_tile = -1, cur_phase = -1;
    for (BlkIt<FWD> it; it.valid(ntiles); it.next(ntiles, nblk)) {
      const int row = it.tile * 128 + rl;
      const bool wvalid = it.tile * 128 + q * 32 < T;  // warp-uniform: this warp has at least one real row
      if (FWD) {
        if (it.blk == 0 && it.phase == 0) { mx = -INFINITY; }
        if (it.blk == 0 && it.phase == 1) {
          // both warpgroups have finished pass 0 of this tile: combine the partial maxima
          smax[wg * 128 + rl] = mx;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const float m = fmaxf(smax[rl], smax[128 + rl]);
          asm volatile("bar.sync 1, 256;" ::: "memory");  // smax may be overwritten by the next tile only after everybody has read it
          rc0 = m * sl2;
          rc1 = 0.f;
          if (wg == 0) smf[it.tile * 128 + rl] = m;
        }
      } else if (it.phase == 0 && (it.tile != cur_tile || it.phase != cur_phase)) {
        const bool ok = row < T;
        rc0 = ok ? -p.lse[((long long)n * p.heads + h) * T + row] * LOG2E_F : 0.f;
        rc1 = ok ? p.delta[((long long)n * p.heads + h) * T + row] : 0.f;
      }
      cur_tile = it.tile; cur_phase = it.phase;
      const bool last_of_tile_pass = it.blk == nblk - 1;
      if ((it.g & 1) == wg) {
        const int b = wg;
        const int ncols = blk_cols(p, it.blk), ncols16 = blk_cols16(p, it.blk);
        const bool tail = it.blk == nblk - 1;
        mbar_wait(bars.sfull(b), (uint32_t)((it.g >> 1) & 1));
        tcgen05_fence_after();
        const uint32_t t_s = t_lane + (uint32_t)(b * 128);
        if (FWD && it.phase == 0) {
          if (wvalid) {
            for (int c0 = 0; c0 < ncols16; c0 += 32) {
              uint32_t sv[32];
              tmem_ld32(t_s + (uint32_t)c0, sv);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (!tail || c0 + j < ncols) mx = fmaxf(mx, __uint_as_float(sv[j]));
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars.sfree(b));
        } else {
          const uint32_t stage = sStage + b * STAGE_BYTES;
          bool waited = false;
          if (wvalid) {
            for (int c0 = 0; c0 < ncols16; c0 += 32) {
              uint32_t sv[32], dv[32], w0[16], w1[16];
              tmem_ld32(t_s + (uint32_t)c0, sv);
              if (!FWD) tmem_ld32(t_s + 64u + (uint32_t)c0, dv);
              tmem_ld_wait();
              if (c0 + 32 >= ncols16) {  // last TMEM read of this block: release the buffer before the arithmetic
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars.sfree(b));
              }
              if (FWD) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float p0 = ex2f(fmaf(__uint_as_float(sv[2 * j]), sl2, -rc0)), p1 = ex2f(fmaf(__uint_as_float(sv[2 * j + 1]), sl2, -rc0));
                  if (tail) { if (c0 + 2 * j >= ncols) p0 = 0.f; if (c0 + 2 * j + 1 >= ncols) p1 = 0.f; }
                  rc1 += p0 + p1;
                  w0[j] = pack_bf2(p0, p1);
                }
              } else if (it.phase == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float p0 = ex2f(fmaf(__uint_as_float(sv[2 * j]), sl2, rc0)), p1 = ex2f(fmaf(__uint_as_float(sv[2 * j + 1]), sl2, rc0));
                  if (tail) { if (c0 + 2 * j >= ncols) p0 = 0.f; if (c0 + 2 * j + 1 >= ncols) p1 = 0.f; }
                  w0[j] = pack_bf2(p0 * p.scale * (__uint_as_float(dv[2 * j]) - rc1), p1 * p.scale * (__uint_as_float(dv[2 * j + 1]) - rc1));
                }
              } else {
                const float* nl = fscr + it.blk * 64 + c0;  // per-query constants: broadcast reads
                const float* dl = fscr + 320 + it.blk * 64 + c0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float2 l2 = *reinterpret_cast<const float2*>(nl + 2 * j), d2 = *reinterpret_cast<const float2*>(dl + 2 * j);
                  const float p0 = ex2f(fmaf(__uint_as_float(sv[2 * j]), sl2, l2.x)), p1 = ex2f(fmaf(__uint_as_float(sv[2 * j + 1]), sl2, l2.y));
                  w0[j] = pack_bf2(p0, p1);
                  w1[j] = pack_bf2(p0 * p.scale * (__uint_as_float(dv[2 * j]) - d2.x), p1 * p.scale * (__uint_as_float(dv[2 * j + 1]) - d2.y));
                }
              }
              if (!waited) { mbar_wait(bars.pfree(b), (uint32_t)((it.su() & 1) ^ 1)); waited = true; }  // staging buffer free again
              stage_store32(stage, rl, c0, w0);
              if (!FWD && it.phase == 1) stage_store32(stage + TILE_BYTES, rl, c0, w1);
            }
          } else {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars.sfree(b));
          }
          fence_proxy_async_smem();  // generic-proxy writes of the staged operand -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(bars.pready(b));
        }
      }
      if (FWD && it.phase == 1 && last_of_tile_pass) {
        // this warpgroup is done with the tile (its last block of pass 1 is staged, or it had none): publish its partial row sums
        sl[(it.tile * 2 + wg) * 128 + rl] = rc1;
        __syncwarp();
        if (lane == 0) mbar_arrive(bars.lready(it.tcount & 1));
      }
    }
  } else if (warp >= 12) {
    // ================= epilogue: finished accumulator tiles -> HBM
    const int q = warp & 3;
    const int rl = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* sl = fscr + 256;
    float* smf = fscr + 1024;
    const int total = FWD ? ntiles : 2 * ntiles;
    for (int tc = 0; tc < total; ++tc) {
      const int a = tc & 1;
      const uint32_t par = (uint32_t)((tc >> 1) & 1);
      const int phase = FWD ? 0 : tc / ntiles, tile = FWD ? tc : tc - phase * ntiles;
      const int row = tile * 128 + rl;
      mbar_wait(bars.accfull(a), par);
      if (FWD) mbar_wait(bars.lready(a), par);
      tcgen05_fence_after();
      if (tile * 128 + q * 32 < T) {
        uint32_t v0[32], v1[32];
        const uint32_t col = 256u + (uint32_t)(FWD ? a * 64 : a * 128);
        tmem_ld32(t_lane + col, v0);
        tmem_ld32(t_lane + col + 32u, v1);
        tmem_ld_wait();
        if (FWD) {
          const float l = sl[(tile * 2 + 0) * 128 + rl] + sl[(tile * 2 + 1) * 128 + rl];
          const float m = smf[tile * 128 + rl];
          if (row < T) {
            store_row64_bf16(p.ctx + ((long long)(row0 + row)) * D + h * 64, v0, v1, 1.f / l);
            p.lse[((long long)n * p.heads + h) * T + row] = m * p.scale + logf(l);
          }
        } else if (phase == 0) {
          if (row < T) store_row64_bf16(p.dqkv + ((long long)(row0 + row)) * 3 * D + h * 64, v0, v1, 1.f);
        } else {
          if (row < T) store_row64_bf16(p.dqkv + ((long long)(row0 + row)) * 3 * D + 2 * D + h * 64, v0, v1, 1.f);  // dV
          tmem_ld32(t_lane + col + 64u, v0);
          tmem_ld32(t_lane + col + 96u, v1);
          tmem_ld_wait();
          if (row < T) store_row64_bf16(p.dqkv + ((long long)(row0 + row)) * 3 * D + D + h * 64, v0, v1, 1.f);  // dK
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars.accfree(a));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int tc_enabled() {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("CG_ATTN_TC");
    enabled = e ? (atoi(e) != 0) : 1;  // default ON; CG_ATTN_TC=0 selects the mma.sync kernels of vit_attention.cu (A/B comparisons)
  }
  return enabled;
}

template <bool FWD>
int launch_attn_tc(const void* qkv, const void* dctx, int Nimg, int T, int heads, const AttnParams& p0, cudaStream_t s) {
  AttnParams p = p0;
  p.T = T; p.heads = heads; p.scale = 0.125f;
  p.ntiles = (T + 127) / 128;
  p.nblk = (T + 63) / 64;
  p.tail_rows = ((T - (p.nblk - 1) * 64) + 15) & ~15;
  const int D = heads * 64;
  CUtensorMap tq, tqt, td, tdt;
  int rc = cg_make_tensor_map_bf16(&tq, qkv, (long long)Nimg * T, 3LL * D, 3LL * D, 64);
  if (rc) return rc;
  rc = cg_make_tensor_map_bf16(&tqt, qkv, (long long)Nimg * T, 3LL * D, 3LL * D, p.tail_rows);
  if (rc) return rc;
  td = tq; tdt = tqt;
  if (!FWD) {
    rc = cg_make_tensor_map_bf16(&td, dctx, (long long)Nimg * T, D, D, 64);
    if (rc) return rc;
    rc = cg_make_tensor_map_bf16(&tdt, dctx, (long long)Nimg * T, D, D, p.tail_rows);
    if (rc) return rc;
  }
  const size_t opb = (size_t)((p.nblk - 1) * 64 + p.tail_rows) * 128;
  const size_t smem = 1024 + (FWD ? 3 : 4) * opb + 2 * (FWD ? TILE_BYTES : 2 * TILE_BYTES) + (FWD ? 1408 : 640) * 4 + BARS_BYTES;
  static size_t configured = 0;
  if (smem > configured) {
    CG_CUDA(cudaFuncSetAttribute(attn_tc_kernel<FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  attn_tc_kernel<FWD><<<dim3(heads, Nimg), AT_THREADS, smem, s>>>(tq, tqt, td, tdt, p);
  CG_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Return 0 when the tcgen05 path handled the call, 1 when the caller should use the mma.sync kernel (T > 272 or CG_ATTN_TC=0), or an error.
int cg_attention_fwd_tc(const void* qkv, int Nimg, int T, int heads, void* ctx, float* lse, cudaStream_t s) {
  if (T > 272 || !tc_enabled()) return 1;
  AttnParams p = {};
  p.ctx = reinterpret_cast<__nv_bfloat16*>(ctx);
  p.lse = lse;
  return launch_attn_tc<true>(qkv, nullptr, Nimg, T, heads, p, s);
}

// delta[n,h,q] = rowsum(dO * O) must already be in `delta` (attn_delta_kernel, vit_attention.cu)
int cg_attention_bwd_tc(const void* qkv, const void* dctx, const float* lse, const float* delta, int Nimg, int T, int heads, void* dqkv, cudaStream_t s) {
  if (T > 272 || !tc_enabled()) return 1;
  AttnParams p = {};
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  return launch_attn_tc<false>(qkv, dctx, Nimg, T, heads, p, s);
}
