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
#include <cuda_fp16.h>
#include <stdlib.h>

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

__device__ __forceinline__ void tl_mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
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
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const int off = tl_off(Np, c, q);
    *reinterpret_cast<__half*>(dst + off) = hi;
    *reinterpret_cast<__half*>(dst + plane + off) = lo;
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
  // instruction descriptor: D=F32 (1<<4), A=B=F16 (format 0), K-major, N>>3 @17, M>>4 @24
  const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(TL_BM >> 4) << 24);
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
        const __half hi = __float2half_rn(s);
        const __half lo = __float2half_rn(s - __half2float(hi));
        const int off = tl_off(TL_BM, r, q);
        *reinterpret_cast<__half*>(sA + off) = hi;
        *reinterpret_cast<__half*>(sA + a_plane + off) = lo;
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
            tl_mma_f16(tmem_base, da, db, idesc, acc);
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

// ---------------------------------------------------------------------------------------------------------------
// Exact x2 upsampling (H == 2h, W == 2w; every Spike2Former config): warp-specialised, pipelined version.
//
//   warps 0..13  producers: build the A operand (sigmoid of the bilinear x2 upsample, fp16 hi + lo) of a 4 x 32 pixel
//                tile.  One task = one low-resolution cell (its 4 corners are shared by a 2 x 2 block of output pixels)
//                x 8 queries: 8 x LDG.128, 32 sigmoids, 8 x STS.128 straight into the SWIZZLE_128B K-major image.
//   warp  14     one thread issues tcgen05.mma.kind::f16 (hi*hi + lo*hi + hi*lo) into one of two TMEM accumulators.
//   warps 15..18 epilogue (one per TMEM lane quadrant): tcgen05.ld the other accumulator, store the NCHW logits (128 B
//                per class per warp, class offsets folded into the STG immediates when H*W is the compile-time 512^2) or the
//                fused argmax label.
// Two A stages in shared memory and two accumulators in TMEM keep the three roles running on different tiles.
// Tile geometry: output rows Y = 4*ty - 1 + {0..3} (cell rows kc = 2*ty - 1, 2*ty; rows are offset by one so that the two
// rows of a cell never straddle tiles), columns X = 32*tx + {0..31} (cells jc = 16*tx - 1 .. 16*tx + 15; the first and
// last cell contribute one column each).  Clamped corner indices reproduce upsample_bilinear2d's border rule.
#ifndef S2F_T2_PROD
#define S2F_T2_PROD 14
#endif
#ifndef S2F_T2_EPI
#define S2F_T2_EPI 4
#endif
// 7 producer warps: 442 tasks per tile = two balanced rounds of 224 threads; 14: one round (the producers are a chain
// of dependent loads / SFU ops at ~0.14 IPC per warp, so the tile time is the per-warp task count).
constexpr int T2_PROD_WARPS = S2F_T2_PROD;
constexpr int T2_EPI_WARPS = S2F_T2_EPI;                    // 8: two per TMEM lane quadrant, half of the classes each; 4: one per quadrant
static_assert(T2_EPI_WARPS == 4 || T2_EPI_WARPS == 8, "one or two epilogue warps per TMEM lane quadrant");
constexpr int T2_THREADS = (T2_PROD_WARPS + 1 + T2_EPI_WARPS) * 32;    // 544
constexpr int T2_CELLS_X = 17;
constexpr int T2_ACC_COLS = 256;                            // TMEM columns per accumulator buffer

struct Tail2P {
  const float* mask_pred; const uint8_t* bpack; float* logits; uint8_t* labels;
  int Q, K, Np, h, w, H, W, tiles_x, tiles_y, ctas_per_img;
  int debug;      // S2F_TAIL_DEBUG (experiments): 1 epilogue idle, 2 no MMA, 4 producers idle
};

__device__ __forceinline__ void t2_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "T2_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra T2_DONE;\n\t"
      "bra T2_WAIT;\n\t"
      "T2_DONE:\n\t"
      "}" ::"r"(tl_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void t2_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tl_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void t2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tl_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool t2_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
// sigmoid(x) = 1 / (1 + 2^(-x log2 e)) with the two SFU approximations and no range fix-ups: ex2 saturates to 0 / +inf
// and rcp(+inf) = 0, which are the correct limits.  Branch-free on purpose: a per-element conditional around it makes
// the compiler serialise 32 dependent MUFU chains per task.
__device__ __forceinline__ float t2_sigmoid(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
// the same from z = -x * log2(e): the scaling is folded into the vertical interpolation weights by the caller
__device__ __forceinline__ float t2_sigmoid_l2(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}

__device__ __forceinline__ uint32_t t2_pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// 8 sigmoid values -> one 16-byte hi chunk and one 16-byte lo chunk of row r, query group g
__device__ __forceinline__ void t2_store8(uint8_t* sA, int a_plane, int r, int g, const float* s) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 hh = __floats2half2_rn(s[2 * j], s[2 * j + 1]);
    const float2 back = __half22float2(hh);
    hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
    lo[j] = t2_pack_half2(s[2 * j] - back.x, s[2 * j + 1] - back.y);
  }
  const int off = (g >> 3) * (TL_BM * 128) + r * 128 + ((((g & 7) ^ (r & 7))) << 4);
  // explicit shared-memory stores: through the generic pointer these were ST.E.128 (generic address resolution, and
  // the compiler has to order them against the global corner loads)
  const uint32_t a0 = tl_smem_u32(sA) + (uint32_t)off;
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + (uint32_t)a_plane), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
}

// HWC: compile-time H*W of the output (0 = run time).  With a constant plane size the per-class store offsets
// (c * H*W * 4 bytes) fold into the STG immediates; with a run-time one every store paid six integer instructions.
template <int HWC>
__global__ void __launch_bounds__(T2_THREADS, 1) tail_x2_kernel(const Tail2P p) {
  extern __shared__ __align__(1024) uint8_t t2_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(t2_raw) + 1023) & ~uintptr_t(1023));
  const int b_plane = 2 * p.Np * 128;             // one of B_hi / B_lo (two 64-query atoms)
  const int a_plane = 2 * TL_BM * 128;            // one of A_hi / A_lo: 32 KB
  const int a_stage = 2 * a_plane;                // 64 KB
  uint8_t* sB = smem;
  uint8_t* sA = smem + 2 * b_plane;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 2 * a_stage);
  uint64_t* a_full = bars, *a_empty = bars + 2, *t_full = bars + 4, *t_empty = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* s_best = reinterpret_cast<float*>(tmem_slot + 4);          // [2 stages][128] partial argmax of the upper class half
  int* s_bidx = reinterpret_cast<int*>(s_best + 2 * TL_BM);

  // warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the role branches on the uniform
  // datapath and the global-memory descriptors in uniform registers (no per-store R2UR pair)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int img = blockIdx.y;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tl_smem_u32(&a_full[s])), "r"(T2_PROD_WARPS));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tl_smem_u32(&a_empty[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tl_smem_u32(&t_full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tl_smem_u32(&t_empty[s])), "r"(T2_EPI_WARPS));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == T2_PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tl_smem_u32(tmem_slot)), "n"(2 * T2_ACC_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {  // B operand (pre-swizzled image of this picture's class probabilities) + zero the A stages (padding queries stay 0)
    const uint4* src = reinterpret_cast<const uint4*>(p.bpack + (int64_t)img * 2 * b_plane);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = threadIdx.x; i < 2 * b_plane / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    uint4* za = reinterpret_cast<uint4*>(sA);
    for (int i = threadIdx.x; i < 2 * a_stage / 16; i += blockDim.x) za[i] = make_uint4(0, 0, 0, 0);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_img = p.tiles_x * p.tiles_y;
  const int64_t HW = (int64_t)p.H * p.W;

  if (warp < T2_PROD_WARPS) {
    // ================================================================== producers
    const int ptid = threadIdx.x;                                  // 0 .. 32*T2_PROD_WARPS - 1
    constexpr int NPROD = T2_PROD_WARPS * 32;
    const int ngroups = (p.Q + 7) >> 3;
    const int ntasks = 2 * T2_CELLS_X * ngroups;
    const float* mp = p.mask_pred + (int64_t)img * p.h * p.w * p.Q;
    const bool vec = (p.Q & 3) == 0;
    float c[4][8];                                                 // the four corners of the current task, 8 queries each
    // corner loads of task t of `tile` (issued early: they are in flight while the previous task is being finished)
    auto load_task = [&](int tile, int t) {
      const int ty = tile / p.tiles_x, tx = tile % p.tiles_x;
      const int g = t % ngroups, cell = t / ngroups;
      const int cx = cell % T2_CELLS_X, kr = cell / T2_CELLS_X;
      const int jc = 16 * tx - 1 + cx, kc = 2 * ty - 1 + kr;
      const int x0 = min(max(jc, 0), p.w - 1), x1 = min(max(jc + 1, 0), p.w - 1);
      const int y0 = min(max(kc, 0), p.h - 1), y1 = min(max(kc + 1, 0), p.h - 1);
      const int q0 = g * 8;
      const float* src[4] = {mp + ((int64_t)y0 * p.w + x0) * p.Q + q0, mp + ((int64_t)y0 * p.w + x1) * p.Q + q0,
                             mp + ((int64_t)y1 * p.w + x0) * p.Q + q0, mp + ((int64_t)y1 * p.w + x1) * p.Q + q0};
      if (vec) {
        const bool second = q0 + 4 < p.Q;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(src[k]));
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (second) b = __ldg(reinterpret_cast<const float4*>(src[k]) + 1);
          c[k][0] = a.x; c[k][1] = a.y; c[k][2] = a.z; c[k][3] = a.w;
          c[k][4] = b.x; c[k][5] = b.y; c[k][6] = b.z; c[k][7] = b.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int j = 0; j < 8; ++j) c[k][j] = (q0 + j < p.Q) ? __ldg(src[k] + j) : 0.f;
      }
    };
    if (blockIdx.x < tiles_img && ptid < ntasks) load_task(blockIdx.x, ptid);
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles_img; tile += p.ctas_per_img, ++it) {
      const int s = it & 1;
      t2_wait(&a_empty[s], ((it >> 1) & 1) ^ 1);
      uint8_t* dstA = sA + s * a_stage;
      if (!(p.debug & 4))
      for (int t = ptid; t < ntasks; t += NPROD) {
        const int g = t % ngroups, cell = t / ngroups;
        const int cx = cell % T2_CELLS_X, kr = cell / T2_CELLS_X;
        const int q0 = g * 8;
        float qmask[8];                                        // 1 for real queries, 0 for the padding of the last group
#pragma unroll
        for (int j = 0; j < 8; ++j) qmask[j] = (q0 + j < p.Q) ? 1.f : 0.f;
        // horizontal interpolation of both corner rows for the two output columns of the cell
        float top[2][8], bot[2][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          top[0][j] = 0.75f * c[0][j] + 0.25f * c[1][j];
          top[1][j] = 0.25f * c[0][j] + 0.75f * c[1][j];
          bot[0][j] = 0.75f * c[2][j] + 0.25f * c[3][j];
          bot[1][j] = 0.25f * c[2][j] + 0.75f * c[3][j];
        }
        // c is dead: start the loads of this thread's next task (same tile or the CTA's next tile)
        if (t + NPROD < ntasks) load_task(tile, t + NPROD);
        else if (tile + p.ctas_per_img < tiles_img && ptid < ntasks) load_task(tile + p.ctas_per_img, ptid);
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int ox = 2 * cx - 1 + dx;
          if (ox < 0 || ox >= 32) continue;
#pragma unroll
          for (int dy = 0; dy < 2; ++dy) {
            // vertical weights pre-multiplied by -log2(e): the interpolation directly yields the ex2 argument
            const float wt = (dy ? 0.25f : 0.75f) * -1.4426950408889634f, wb = (dy ? 0.75f : 0.25f) * -1.4426950408889634f;
            float sv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) sv[j] = t2_sigmoid_l2(wt * top[dx][j] + wb * bot[dx][j]);
            if (q0 + 8 > p.Q) {                                    // padding queries of the last group contribute 0
#pragma unroll
              for (int j = 0; j < 8; ++j) sv[j] *= qmask[j];
            }
            t2_store8(dstA, a_plane, (2 * kr + dy) * 32 + ox, g, sv);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) t2_arrive(&a_full[s]);
    }
  } else if (warp == T2_PROD_WARPS) {
    // ================================================================== MMA issuer
    // the whole warp runs the loop and one elected lane issues: operands stay in uniform registers
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(TL_BM >> 4) << 24);   // F32 += F16 x F16
    const int ksteps = (p.Q + 15) >> 4;
    const uint32_t b_hi = tl_smem_u32(sB), b_lo = b_hi + b_plane;
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles_img; tile += p.ctas_per_img, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      t2_wait(&a_full[s], ph);
      t2_wait(&t_empty[s], ph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = tl_smem_u32(sA + s * a_stage), a_lo = a_hi + a_plane;
      const uint32_t d = tmem_base + (uint32_t)(s * T2_ACC_COLS);
      if (t2_elect_one()) {
        uint32_t acc = 0;
#pragma unroll 1
        for (int combo = (p.debug & 2) ? 3 : 0; combo < 3; ++combo) {
          const uint32_t abase = combo == 1 ? a_lo : a_hi, bbase = combo == 2 ? b_lo : b_hi;
          for (int ks = 0; ks < ksteps; ++ks) {
            const int atom = ks >> 2, k = ks & 3;
            const uint64_t da = tl_desc(abase + atom * TL_BM * 128) + (uint64_t)(k * 2);
            const uint64_t db = tl_desc(bbase + atom * p.Np * 128) + (uint64_t)(k * 2);
            tl_mma_f16(d, da, db, idesc, acc);
            acc = 1;
          }
        }
        t2_commit(&a_empty[s]);        // A stage reusable once these MMAs have read it
        t2_commit(&t_full[s]);         // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ================================================================== epilogue
    // warp -> (TMEM lane quadrant = output row of the tile, class half).  A single warp can only issue an instruction
    // every few cycles, so the 150 stores per pixel are spread over two warps per quadrant.
    const int quad = warp & 3;
    const int half = (warp - (T2_PROD_WARPS + 1)) >> 2;
    const int ncol16 = p.Np / 16;
    const int cb_begin = half ? (ncol16 + 1) / 2 : 0, cb_end = (half || T2_EPI_WARPS == 4) ? ncol16 : (ncol16 + 1) / 2;
    const uint32_t HW4 = (uint32_t)HW;
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles_img; tile += p.ctas_per_img, ++it) {
      const int s = it & 1;
      t2_wait(&t_full[s], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int ty = tile / p.tiles_x, tx = tile % p.tiles_x;
      const int Y = 4 * ty - 1 + quad, X = 32 * tx + lane;
      const bool ok = Y >= 0 && Y < p.H && X < p.W;
      const int64_t pix = (int64_t)Y * p.W + X;
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * T2_ACC_COLS);
      float best = -INFINITY;
      int best_c = 0;
      // warp-uniform null-ness (p.logits), lane predicate `ok` on every store: inside a lane-divergent branch the
      // compiler re-materialises the global-memory descriptor (2 x R2UR) for each of the 150 stores
      float* lg = p.logits ? p.logits + (int64_t)img * p.K * HW + (ok ? pix : 0) : nullptr;
      for (int cb = (p.debug & 1) ? cb_end : cb_begin; cb < cb_end; cb += 2) {
        uint32_t v[32];
        tl_ld16(trow + cb * 16, v);
        const bool two = cb + 1 < cb_end;                           // warp-uniform
        if (two) tl_ld16(trow + cb * 16 + 16, v + 16);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int c0 = cb * 16;
        const int nv = min(two ? 32 : 16, p.K - c0);               // valid classes among the loaded columns
        if (lg) {
          if (HWC > 0) {
            float* pj = lg + (size_t)c0 * HWC;                       // one 64-bit address per 32 classes
            if (nv == 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (ok) pj[(size_t)j * HWC] = __uint_as_float(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (ok && j < nv) pj[(size_t)j * HWC] = __uint_as_float(v[j]);
            }
          } else {
            char* pj = reinterpret_cast<char*>(lg) + (size_t)c0 * HW4 * 4u;
            const size_t stride = (size_t)HW4 * 4u;                  // running pointer: one 64-bit add per store
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (ok && j < nv) *reinterpret_cast<float*>(pj) = __uint_as_float(v[j]);
              pj += stride;
            }
          }
        }
        if (p.labels) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float y = __uint_as_float(v[j]);
            if (j < nv && y > best) { best = y; best_c = c0 + j; }
          }
        }
      }
      if (p.labels) {                                               // combine the two class halves: first maximum wins
        const int r = quad * 32 + lane;
        if (T2_EPI_WARPS == 4) {
          if (ok) p.labels[(int64_t)img * HW + pix] = (uint8_t)best_c;
        } else {
          if (half) { s_best[s * TL_BM + r] = best; s_bidx[s * TL_BM + r] = best_c; }
          asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
          if (!half && ok) {
            const float b1 = s_best[s * TL_BM + r];
            p.labels[(int64_t)img * HW + pix] = (uint8_t)((b1 > best) ? s_bidx[s * TL_BM + r] : best_c);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) t2_arrive(&t_empty[s]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == T2_PROD_WARPS) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * T2_ACC_COLS));
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
  static std::atomic<uint64_t> attr{0};
  if (first_use_on_this_device(attr)) {
    cudaFuncSetAttribute(tail_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(tail_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  }
  tail_prep_kernel<<<n, 256, prep_sm, st>>>(cls, reinterpret_cast<uint8_t*>(ws), Q, K, Np);
  int rc = check_launch("tail_prep_kernel");
  if (rc) return rc;
  const size_t smem2 = (size_t)2 * 2 * Np * 128 + 2 * 2 * 2 * TL_BM * 128 + 128 + 4 * TL_BM * 4 + 1024;
  if (H == 2 * h && W == 2 * w && smem2 <= 227 * 1024) {          // K <= 192 classes: both A stages + B fit
    Tail2P q;
    q.mask_pred = mask_pred; q.bpack = reinterpret_cast<const uint8_t*>(ws); q.logits = logits; q.labels = labels;
    q.Q = Q; q.K = K; q.Np = Np; q.h = h; q.w = w; q.H = H; q.W = W;
    q.tiles_x = (W + 31) / 32; q.tiles_y = (h + 1 + 1) / 2;          // cell rows -1 .. h-1, two per tile
    const int tiles_img = q.tiles_x * q.tiles_y;
    // CTAs per image: one CTA per SM (227 KB of shared memory), so the grid runs in k waves of all SMs; pick the k <= 8
    // that wastes the fewest SM-waves (batch 32: 4 CTAs per image fill only 128 of 148 SMs, 9 per image = 288 CTAs
    // fill 97 % of two waves; batch 64: 16 per image = 1024 CTAs fill 99 % of seven)
    const int sms = sm_count();
    int per_img = 1;
    double best_fill = 0.0;
    for (int k = 1; k <= 8; ++k) {
      int c = (sms * k) / n;
      if (c < 1) c = 1;
      if (c > tiles_img) c = tiles_img;
      const int waves = (c * n + sms - 1) / sms;
      const double fill = (double)(c * n) / ((double)sms * waves);
      if (fill > best_fill + 0.02) { best_fill = fill; per_img = c; }
    }
    q.ctas_per_img = per_img;
    { const char* dbg = getenv("S2F_TAIL_DEBUG"); q.debug = dbg ? atoi(dbg) : 0; }
    static std::atomic<uint64_t> attr2{0};
    if (first_use_on_this_device(attr2)) {
      cudaError_t e = cudaFuncSetAttribute(tail_x2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_x2_kernel<512 * 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return fail(S2F_ERR_CUDA, "semantic_tail_tc: smem attribute: %s", cudaGetErrorString(e));
    }
    if ((int64_t)H * W == 512 * 512) tail_x2_kernel<512 * 512><<<dim3(per_img, n), T2_THREADS, smem2, st>>>(q);
    else tail_x2_kernel<0><<<dim3(per_img, n), T2_THREADS, smem2, st>>>(q);
    return check_launch("tail_x2_kernel");
  }
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
