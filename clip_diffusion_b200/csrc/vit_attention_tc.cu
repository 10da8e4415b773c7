// Attention FORWARD on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) for the short CLIP sequences (T <= 272:
// ViT-B/32 T=50, ViT-B/16 T=197, ViT-L/14 T=257).  One CTA per (image, head):
//
//   TMA        K, V (whole head, <= 272 x 64 bf16 each) and one 128-row Q tile at a time -> 128B-swizzled smem
//   MMA 1      S[128 x Tk] = Q K^T      tcgen05.mma M=128, N=Tk (256 + 16 for Tk=272), K=64   -> TMEM columns [0, Tk)
//   softmax    8 warps, thread = (row, half of the columns): tcgen05.ld the row, max, exp2, sum; P (bf16) is written
//              to smem in the K-major 128B-swizzled operand layout, so it is the A operand of the second MMA
//   MMA 2      O[128 x 64] = P V        tcgen05.mma M=128, N=64, K=Tk, B = V in MN-major form (V rows are keys: the K
//              dimension of this product is the ROW index of the smem tile)                    -> TMEM columns [320, 384)
//   epilogue   O / l -> bf16 ctx, log-sum-exp -> lse (the backward recomputes P from it)
//
// The scores never leave the SM; the two GEMMs of a head run on the tensor cores at M=128 instead of the 16-row
// mma.sync tiles of vit_attention.cu.  EXPERIMENTAL (CG_ATTN_TC=1): validated against the oracle, not yet faster -- see
// cg_attention_fwd_tc below; vit_attention.cu remains the default path (and the only one for T > 272 and the backward).
#include <stdlib.h>
#include "common.cuh"
#include "tcgen05.cuh"

using namespace tc;

namespace {

constexpr int ATC_THREADS = 320;  // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..9 softmax/epilogue
constexpr int O_COL = 320;        // TMEM column of the O accumulator
constexpr float LOG2E_F = 1.4426950408889634f;

__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void bar_softmax() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// instruction descriptor, kind::f16: D fp32 (bit 4), A/B bf16 (bits 7, 10), b_major (bit 16), N>>3 at 17, M>>4 at 24
__device__ __forceinline__ uint32_t idesc_bf16(int m, int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(ATC_THREADS, 1) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                                                                    int T, int Tk, int kv_box, int heads, float scale, __nv_bfloat16* __restrict__ ctx,
                                                                    float* __restrict__ lse) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int D = heads * 64;
  const int h = blockIdx.x, n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t kv_bytes = (uint32_t)((Tk * 128 + 1023) & ~1023);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = sQ + 16384, sV = sK + kv_bytes, sP = sV + kv_bytes, sBar = sP + 5 * 16384;
  const uint32_t kv_full = sBar, q_full = sBar + 8, s_full = sBar + 16, p_ready = sBar + 24, o_full = sBar + 32, o_free = sBar + 40;
  const uint32_t tmem_slot = sBar + 48;
  float* pmax = reinterpret_cast<float*>(smem_raw + (sBar + 64 - smem_u32(smem_raw)));  // [2][128]
  float* psum = pmax + 256;                                                               // [2][128]
  const int mtiles = (T + 127) / 128;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
    mbar_init(kv_full, 1); mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(p_ready, 1); mbar_init(o_full, 1); mbar_init(o_free, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  const int row0 = n * T;  // first row of this image in the packed [Nimg*T, 3D] qkv matrix

  if (warp == 0) {
    // ================= TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2u * (uint32_t)Tk * 128u);
      for (int r = 0; r < Tk; r += kv_box) {
        tma_load_2d(sK + (uint32_t)r * 128u, &tmKV, kv_full, D + h * 64, row0 + r);
        tma_load_2d(sV + (uint32_t)r * 128u, &tmKV, kv_full, 2 * D + h * 64, row0 + r);
      }
      for (int mt = 0; mt < mtiles; ++mt) {
        if (mt > 0) mbar_wait(s_full, (uint32_t)((mt - 1) & 1));  // MMA 1 of the previous tile has consumed the Q buffer
        mbar_arrive_expect_tx(q_full, 16384u);
        tma_load_2d(sQ, &tmQ, q_full, h * 64, row0 + mt * 128);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer
    if (lane == 0) {
      const int n1 = Tk > 256 ? 256 : Tk, n2 = Tk - n1;
      const uint32_t id_s1 = idesc_bf16(128, n1, false), id_s2 = idesc_bf16(128, n2 > 0 ? n2 : 16, false), id_o = idesc_bf16(128, 64, true);
      mbar_wait(kv_full, 0);
      for (int mt = 0; mt < mtiles; ++mt) {
        const uint32_t ph = (uint32_t)(mt & 1);
        mbar_wait(q_full, ph);
        tcgen05_fence_after();
        const uint64_t qd = make_smem_desc(sQ), kd = make_smem_desc(sK);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16(tmem_base, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), id_s1, k != 0 ? 1u : 0u);
          if (n2 > 0) umma_bf16(tmem_base + 256u, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k) + (uint64_t)((256 * 128) >> 4), id_s2, k != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        mbar_wait(p_ready, ph);  // P is in shared memory (and S has been read)
        tcgen05_fence_after();
        if (mt > 0) mbar_wait(o_free, (uint32_t)((mt - 1) & 1));  // epilogue of the previous tile has read O
        tcgen05_fence_after();
        const int ksteps = Tk >> 4;
        for (int s = 0; s < ksteps; ++s) {
          const uint64_t pd = make_smem_desc(sP + (uint32_t)(s >> 2) * 16384u) + (uint64_t)(2 * (s & 3));
          const uint64_t vd = make_smem_desc(sV + (uint32_t)s * 2048u);  // 16 key rows further (MN-major: rows are K)
          umma_bf16(tmem_base + (uint32_t)O_COL, pd, vd, id_o, s != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
      }
    }
  } else {
    // ================= softmax + epilogue: warps 2..9
    const int q = warp & 3;            // TMEM lane quarter (warp id % 4)
    const int hh = (warp - 2) >> 2;    // which half of the key columns
    const int r = q * 32 + lane;       // row inside the 128-row tile
    const int ng = (Tk + 31) >> 5, ng0 = (ng + 1) >> 1;
    const int g_lo = hh == 0 ? 0 : ng0, g_hi = hh == 0 ? ng0 : ng;
    const float c = scale * LOG2E_F;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int mt = 0; mt < mtiles; ++mt) {
      const uint32_t ph = (uint32_t)(mt & 1);
      mbar_wait(s_full, ph);
      tcgen05_fence_after();
      // pass 1: row maximum over this thread's column groups
      float mx = -INFINITY;
      for (int gi = g_lo; gi < g_hi; ++gi) {
        uint32_t v[32];
        tmem_ld32(t_lane + (uint32_t)(gi * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (gi * 32 + j < T) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      pmax[hh * 128 + r] = mx;
      bar_softmax();
      const float m = fmaxf(pmax[r], pmax[128 + r]);
      const float mc = m * c;
      // pass 2: P = exp2(s*c - m*c) -> bf16 -> smem (A operand layout), partial row sum
      float sum = 0.f;
      for (int gi = g_lo; gi < g_hi; ++gi) {
        uint32_t v[32];
        tmem_ld32(t_lane + (uint32_t)(gi * 32), v);
        tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          p[j] = (gi * 32 + j < T) ? exp2f(fmaf(__uint_as_float(v[j]), c, -mc)) : 0.f;
          // the row sum must see the same bf16-rounded values the second MMA multiplies with V
          p[j] = __bfloat162float(__float2bfloat16(p[j]));
          sum += p[j];
        }
        const uint32_t rowb = sP + (uint32_t)(gi >> 1) * 16384u + (uint32_t)r * 128u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t chunk = (uint32_t)(((gi & 1) * 4 + j) ^ (r & 7));
          const uint32_t a = rowb + (chunk << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf2(p[8 * j], p[8 * j + 1])), "r"(pack_bf2(p[8 * j + 2], p[8 * j + 3])),
                       "r"(pack_bf2(p[8 * j + 4], p[8 * j + 5])), "r"(pack_bf2(p[8 * j + 6], p[8 * j + 7]))
                       : "memory");
        }
      }
      psum[hh * 128 + r] = sum;
      fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
      tcgen05_fence_before();
      bar_softmax();
      if (warp == 2 && lane == 0) mbar_arrive(p_ready);
      if (hh == 0) {
        mbar_wait(o_full, ph);
        tcgen05_fence_after();
        const float l = psum[r] + psum[128 + r];
        const float inv = 1.f / l;
        const int row = mt * 128 + r;
        uint32_t o0[32], o1[32];
        tmem_ld32(t_lane + (uint32_t)O_COL, o0);
        tmem_ld32(t_lane + (uint32_t)O_COL + 32u, o1);
        tmem_ld_wait();
        if (row < T) {
          uint4* dst = reinterpret_cast<uint4*>(ctx + ((long long)(row0 + row)) * D + h * 64);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_uint4(pack_bf2(__uint_as_float(o0[8 * j]) * inv, __uint_as_float(o0[8 * j + 1]) * inv),
                                pack_bf2(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv),
                                pack_bf2(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv),
                                pack_bf2(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv));
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[4 + j] = make_uint4(pack_bf2(__uint_as_float(o1[8 * j]) * inv, __uint_as_float(o1[8 * j + 1]) * inv),
                                    pack_bf2(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv),
                                    pack_bf2(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv),
                                    pack_bf2(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv));
          lse[((long long)n * heads + h) * T + row] = m * scale + logf(l);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

// Returns 0 when the tcgen05 path handled the call, 1 when the caller should use the mma.sync kernel, or an error.
int cg_attention_fwd_tc(const void* qkv, int Nimg, int T, int heads, void* ctx, float* lse, cudaStream_t s) {
  const int Tk = (T + 15) & ~15;
  if (Tk > 272) return 1;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("CG_ATTN_TC");
    // Off by default: parity-green on B200, but one CTA/SM with serialised TMA -> MMA -> softmax -> MMA phases measured
    // 151 us/layer (ViT-L/14 x 64) against 106 us for the 2-CTA/SM mma.sync kernel.  Needs a persistent CTA with K/V
    // prefetch and CUDA-core handling of the 1-row remainder tile (T = 257 = 2*128 + 1) before it pays off (round 2).
    enabled = e ? (atoi(e) != 0) : 0;
  }
  if (!enabled) return 1;
  const int D = heads * 64;
  const int kv_box = Tk <= 256 ? Tk : Tk / 2;
  CUtensorMap tq, tkv;
  int rc = cg_make_tensor_map_bf16(&tq, qkv, (long long)Nimg * T, 3LL * D, 3LL * D, 128);
  if (rc) return rc;
  rc = cg_make_tensor_map_bf16(&tkv, qkv, (long long)Nimg * T, 3LL * D, 3LL * D, kv_box);
  if (rc) return rc;
  const size_t kv_bytes = (size_t)((Tk * 128 + 1023) & ~1023);
  const size_t smem = 1024 + 16384 + 2 * kv_bytes + 5 * 16384 + 64 + 4 * 128 * sizeof(float) + 64;
  static size_t configured = 0;
  if (smem > configured) {
    CG_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  attn_fwd_tc_kernel<<<dim3(heads, Nimg), ATC_THREADS, smem, s>>>(tq, tkv, T, Tk, kv_box, heads, 0.125f, reinterpret_cast<__nv_bfloat16*>(ctx), lse);
  CG_LAUNCH_CHECK();
  return 0;
}
