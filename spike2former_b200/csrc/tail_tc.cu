// Semantic-inference tail on the tensor cores (decode_heads/maskformer_head.py:163-177):
//   logits[n, c, Y, X] = sum_q softmax(cls)[n, q, c] * sigmoid(bilinear_x2(mask_pred)[n, q, Y, X])
// One CTA owns a run of 128-pixel tiles of one image.  Per tile, eight warps build the A operand in shared
// memory (bilinear upsample + sigmoid of the Q mask logits of 128 output pixels, split into bf16 hi + lo so the
// product keeps ~16 mantissa bits), one thread issues tcgen05.mma.kind::f16 (M=128, N=classes, K=16) for
// hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM, and all warps drain TMEM to the NCHW logits (coalesced
// 128-byte rows per class) or reduce it to an argmax label (tail fusion).
// The B operand (class probabilities, bf16 hi/lo, already in the UMMA K-major SWIZZLE_128B image) is produced once
// per image by tail_prep_kernel and copied to shared memory once per CTA.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace s2f {

constexpr int TL_BM = 128;       // pixels per tile
constexpr int TL_KQ = 128;       // queries padded to two 64-element (128-byte) K atoms
constexpr int TL_THREADS = 256;

__device__ __forceinline__ uint32_t tl_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, k) in a K-major SWIZZLE_128B operand image with `rows` rows and 64-element atoms
__host__ __device__ __forceinline__ int tl_off(int rows, int r, int k) {
  const int atom = k >> 6, kk = k & 63;
  return atom * rows * 128 + r * 128 + ((((kk >> 3) ^ (r & 7)) << 4) | ((kk & 7) << 1));
}

__device__ __forceinline__ uint64_t tl_desc(uint32_t saddr) {      // SWIZZLE_128B, SBO = 1024 B, version 1
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void tl_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void tl_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// ---------------------------------------------------------------------------------------------------------------
// softmax over K+1 classes (last one dropped) -> B operand image: [hi|lo][atom][Np rows = class][64 q] bf16
__global__ void __launch_bounds__(256) tail_prep_kernel(const float* __restrict__ cls, uint8_t* __restrict__ bpack, int Q,
                                                        int K, int Np) {
  extern __shared__ float prob[];                 // [Q][K]
  const int img = blockIdx.x;
  const int K1 = K + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = warp; q < Q; q += blockDim.x / 32) {
    const float* c = cls + ((int64_t)img * Q + q) * K1;
    float mx = -INFINITY;
    for (int i = lane; i < K1; i += 32) mx = fmaxf(mx, c[i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = lane; i < K1; i += 32) sum += expf(c[i] - mx);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    for (int i = lane; i < K; i += 32) prob[q * K + i] = expf(c[i] - mx) / sum;
  }
  __syncthreads();
  const int plane = 2 * Np * 128;                 // bytes of one (hi or lo) image: 2 atoms
  uint8_t* dst = bpack + (int64_t)img * 2 * plane;
  for (int e = threadIdx.x; e < Np * TL_KQ; e += blockDim.x) {
    const int c = e / TL_KQ, q = e % TL_KQ;
    const float v = (c < K && q < Q) ? prob[q * K + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
    const int off = tl_off(Np, c, q);
    *reinterpret_cast<__nv_bfloat16*>(dst + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(dst + plane + off) = lo;
  }
}

struct TailP {
  const float* mask_pred; const uint8_t* bpack; float* logits; uint8_t* labels;
  int Q, K, Np, h, w, H, W, tiles_per_img, tiles_per_cta;
};

__global__ void __launch_bounds__(TL_THREADS, 1) tail_tc_kernel(const TailP p) {
  extern __shared__ __align__(1024) uint8_t tl_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tl_raw) + 1023) & ~uintptr_t(1023));
  const int b_plane = 2 * p.Np * 128;             // one of B_hi / B_lo (two atoms)
  const int a_plane = 2 * TL_BM * 128;            // one of A_hi / A_lo: 32 KB
  uint8_t* sB = smem;                             // [hi | lo]
  uint8_t* sA = smem + 2 * b_plane;               // [hi | lo]; 2*b_plane is a multiple of 1024 because Np % 8 == 0 ... Np*512
  uint64_t* bar = reinterpret_cast<uint64_t*>(sA + 2 * a_plane);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);          // [2][128] (value) + [2][128] (index) for labels

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.y;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tl_smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tl_smem_u32(tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // B operand: linear copy of the pre-swizzled image
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.bpack + (int64_t)img * 2 * b_plane);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = threadIdx.x; i < 2 * b_plane / 16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const float sh = (float)p.h / (float)p.H, sw = (float)p.w / (float)p.W;
  const float* mp = p.mask_pred + (int64_t)img * p.h * p.w * p.Q;
  const int64_t HW = (int64_t)p.H * p.W;
  // instruction descriptor: D=F32 (1<<4), A=B=BF16 (1<<7, 1<<10), K-major, N>>3 @17, M>>4 @24
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(TL_BM >> 4) << 24);
  uint32_t parity = 0;

  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(p.tiles_per_img, t_begin + p.tiles_per_cta);
  for (int tile = t_begin; tile < t_end; ++tile) {
    const int64_t pix0 = (int64_t)tile * TL_BM;
    // ---- 1. A operand: 16 pixels per warp, lanes over q
    for (int pr = 0; pr < TL_BM / 8; ++pr) {
      const int r = warp * (TL_BM / 8) + pr;
      const int64_t pix = pix0 + r;
      const bool ok = pix < HW;
      int o00 = 0, o01 = 0, o10 = 0, o11 = 0;
      float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
      if (ok) {
        const int Y = (int)(pix / p.W), X = (int)(pix % p.W);
        float sy = sh * ((float)Y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
        float sx = sw * ((float)X + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < p.h - 1 ? 1 : 0), x1 = x0 + (x0 < p.w - 1 ? 1 : 0);
        const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
        o00 = (y0 * p.w + x0) * p.Q; o01 = (y0 * p.w + x1) * p.Q; o10 = (y1 * p.w + x0) * p.Q; o11 = (y1 * p.w + x1) * p.Q;
        w00 = hy * hx; w01 = hy * lx; w10 = ly * hx; w11 = ly * lx;
      }
#pragma unroll
      for (int it = 0; it < TL_KQ / 32; ++it) {
        const int q = it * 32 + lane;
        float s = 0.f;
        if (ok && q < p.Q) {
          const float up = w00 * __ldg(mp + o00 + q) + w01 * __ldg(mp + o01 + q) + w10 * __ldg(mp + o10 + q) + w11 * __ldg(mp + o11 + q);
          s = __fdividef(1.f, 1.f + __expf(-up));
        }
        const __nv_bfloat16 hi = __float2bfloat16(s);
        const __nv_bfloat16 lo = __float2bfloat16(s - __bfloat162float(hi));
        const int off = tl_off(TL_BM, r, q);
        *reinterpret_cast<__nv_bfloat16*>(sA + off) = hi;
        *reinterpret_cast<__nv_bfloat16*>(sA + a_plane + off) = lo;
      }
    }
    // generic-proxy writes -> visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- 2. MMA: hi*hi + lo*hi + hi*lo
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = tl_smem_u32(sA), a_lo = a_hi + a_plane, b_hi = tl_smem_u32(sB), b_lo = b_hi + b_plane;
      uint32_t acc = 0;
#pragma unroll
      for (int combo = 0; combo < 3; ++combo) {
        const uint32_t abase = combo == 1 ? a_lo : a_hi, bbase = combo == 2 ? b_lo : b_hi;
#pragma unroll
        for (int atom = 0; atom < 2; ++atom) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = tl_desc(abase + atom * TL_BM * 128) + (uint64_t)(k * 2);
            const uint64_t db = tl_desc(bbase + atom * p.Np * 128) + (uint64_t)(k * 2);
            tl_mma_bf16(tmem_base, da, db, idesc, acc);
            acc = 1;
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tl_smem_u32(bar)) : "memory");
    }
    // ---- 3. wait for the accumulator
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "TL_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra TL_DONE;\n\t"
        "bra TL_WAIT;\n\t"
        "TL_DONE:\n\t"
        "}" ::"r"(tl_smem_u32(bar)), "r"(parity) : "memory");
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- 4. epilogue: warp w -> TMEM lanes 32*(w%4); warps 0-3 take the first half of the classes, 4-7 the second
    {
      const int quad = warp & 3, half = warp >> 2;
      const int r = quad * 32 + lane;
      const int64_t pix = pix0 + r;
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
      const int ncol16 = p.Np / 16;
      const int c_begin = half * ((ncol16 + 1) / 2), c_end = half ? ncol16 : (ncol16 + 1) / 2;
      float best = -INFINITY;
      int best_c = 0;
      for (int cb = c_begin; cb < c_end; ++cb) {
        uint32_t v[16];
        tl_ld16(trow + cb * 16, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (pix < HW) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = cb * 16 + j;
            if (c < p.K) {
              const float y = __uint_as_float(v[j]);
              if (p.logits) p.logits[((int64_t)img * p.K + c) * HW + pix] = y;
              if (y > best) { best = y; best_c = c; }
            }
          }
        }
      }
      if (p.labels) {
        s_red[half * TL_BM + r] = best;
        reinterpret_cast<int*>(s_red)[2 * TL_BM + half * TL_BM + r] = best_c;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (p.labels && threadIdx.x < TL_BM) {
      const int r = threadIdx.x;
      const int64_t pix = pix0 + r;
      if (pix < HW) {
        const float b0 = s_red[r], b1 = s_red[TL_BM + r];
        const int c0 = reinterpret_cast<int*>(s_red)[2 * TL_BM + r], c1 = reinterpret_cast<int*>(s_red)[3 * TL_BM + r];
        p.labels[(int64_t)img * HW + pix] = (uint8_t)((b1 > b0) ? c1 : c0);      // first maximum wins, as torch.argmax
      }
    }
    // s_red and sA are rewritten only after the next tile's first __syncthreads -> safe
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

}  // namespace s2f

using namespace s2f;

extern "C" int64_t s2f_semantic_tail_ws_bytes(int n, int K) {
  const int Np = (K + 15) / 16 * 16;
  return (int64_t)n * 2 * 2 * Np * 128;
}

extern "C" int s2f_semantic_tail_tc(const float* mask_pred, const float* cls, float* logits, uint8_t* labels, void* ws,
                                    int n, int Q, int K, int h, int w, int H, int W, void* stream) {
  S2F_REQUIRE(mask_pred && cls && ws && (logits || labels), "semantic_tail_tc: null pointer");
  S2F_REQUIRE(Q >= 1 && Q <= TL_KQ, "semantic_tail_tc: at most 128 queries");
  S2F_REQUIRE(K >= 1 && K <= 256, "semantic_tail_tc: at most 256 classes");
  cudaStream_t st = (cudaStream_t)stream;
  const int Np = (K + 15) / 16 * 16;
  const size_t prep_sm = sizeof(float) * (size_t)Q * K;
  S2F_REQUIRE(prep_sm <= 160 * 1024, "semantic_tail_tc: Q*K too large");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(tail_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(tail_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr = true;
  }
  tail_prep_kernel<<<n, 256, prep_sm, st>>>(cls, reinterpret_cast<uint8_t*>(ws), Q, K, Np);
  int rc = check_launch("tail_prep_kernel");
  if (rc) return rc;
  TailP p;
  p.mask_pred = mask_pred; p.bpack = reinterpret_cast<const uint8_t*>(ws); p.logits = logits; p.labels = labels;
  p.Q = Q; p.K = K; p.Np = Np; p.h = h; p.w = w; p.H = H; p.W = W;
  const int64_t HW = (int64_t)H * W;
  p.tiles_per_img = (int)ceil_div(HW, TL_BM);
  // enough CTAs for ~2 waves over 148 SMs, at least 4 tiles each so the B copy is amortised
  int ctas_per_img = (int)ceil_div(2 * 148, n);
  if (ctas_per_img > p.tiles_per_img / 4) ctas_per_img = p.tiles_per_img / 4;
  if (ctas_per_img < 1) ctas_per_img = 1;
  p.tiles_per_cta = (int)ceil_div(p.tiles_per_img, ctas_per_img);
  ctas_per_img = (int)ceil_div(p.tiles_per_img, p.tiles_per_cta);
  const size_t smem = (size_t)2 * 2 * Np * 128 + 2 * 2 * TL_BM * 128 + 64 + 4 * TL_BM * 4 + 1024;
  S2F_REQUIRE(smem <= 220 * 1024, "semantic_tail_tc: shared memory budget exceeded");
  tail_tc_kernel<<<dim3(ctas_per_img, n), TL_THREADS, smem, st>>>(p);
  return check_launch("tail_tc_kernel");
}
