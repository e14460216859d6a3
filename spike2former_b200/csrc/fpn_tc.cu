// Top-down FPN merge in one launch (pixel_decoder.py:451-462):
//   y = BN(lateral_1x1(spikes)) + bilinear_x2(prev)        spikes_out = NI-LIF(y)
// for the two finest levels, where the lateral conv has a tiny K (32 / 64 input channels) and the work is the epilogue:
// 537 M outputs per 32 images at 256^2.  The int8 digit-plane GEMM (gemm_tc.cu, 128 x 64 tiles, three int32 planes per
// output) spent 37 instructions per output there, most of them per-tile bookkeeping and plane merging.  This kernel:
//   * tile = 16 x 8 pixels x ALL 256 channels: one fp32 accumulator of 256 TMEM columns (two of them, ping-pong);
//   * lateral conv on tcgen05.mma.kind::f16: the spike levels 0..8 are exact in fp16 (converted from int8 by the producer
//     warp on the way into shared memory), the weights are fp16 hi + lo with one power-of-two scale per output channel
//     (ops.pack_pw_f16): D = A W_hi^T + A W_lo^T, 2 * Cin / 16 MMAs per tile, 4 bytes of TMEM drain per output;
//   * the coarser level's 6 x 10-pixel patch is copied by cp.async into shared memory with a 16-byte skew per pixel, so
//     that an epilogue warp whose lanes are PIXELS (the layout tcgen05.ld delivers) reads its four bilinear corners
//     with conflict-free LDS.128;
//   * 16 epilogue warps (four per TMEM lane quadrant, 64 channels each): affine, ATen's bilinear expression
//     hy*(hx*p00 + lx*p01) + ly*(hx*p10 + lx*p11), NI-LIF, one STG.128 per 16 levels: ~12 instructions per output.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace s2f {

constexpr int FP_EW = 16;                          // epilogue warps
constexpr int FP_THREADS = (1 + FP_EW) * 32;       // warp 0: producer + MMA issue
constexpr int FP_TW = 16, FP_TH = 8;               // fine tile
constexpr int FP_PW = FP_TW / 2 + 2, FP_PH = FP_TH / 2 + 2;     // coarse patch: 10 x 6 pixels
constexpr int FP_COUT = 256;
constexpr int FP_PIX_BYTES = FP_COUT * 4 + 16;     // skewed pixel stride of the patch
constexpr int FP_PATCH_BYTES = FP_PW * FP_PH * FP_PIX_BYTES;    // 62400
constexpr int FP_A_BYTES = 128 * 128;              // one A stage: 128 pixels x 64 fp16 (K padded to 64)

struct FpnP {
  const int8_t* a; const uint8_t* bpack; const float* scale; const float* shift; const float* prev; int8_t* out_spike;
  int n, H, W, Cin, Hp, Wp, tiles_x, tiles_y, tiles;
};

__device__ __forceinline__ uint32_t fp_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t fp_desc(uint32_t saddr) {      // K-major SWIZZLE_128B, SBO = 1024 B, version 1
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void fp_mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void fp_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void fp_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "FP_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra FP_DONE;\n\t"
      "bra FP_WAIT;\n\t"
      "FP_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fp_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fp_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 fp_lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

__device__ __forceinline__ float fp_add_sat(float a, float b) {       // clamp(a + b, 0, 1) in one FADD.SAT
  float r;
  asm("add.rn.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
// four values u = clamp(y, 0, 8) / 8 in [0, 1] -> four int8 levels round_half_even(8 u): one FFMA each puts the level
// into the mantissa of 2^23 (common.cuh::level_bits8 without its saturating multiply), three PRMT pack them
__device__ __forceinline__ uint32_t pack_unit4(float a, float b, float c, float d) {
  const uint32_t la = __float_as_uint(fmaf(a, 8.f, 8388608.f)), lb = __float_as_uint(fmaf(b, 8.f, 8388608.f));
  const uint32_t lc = __float_as_uint(fmaf(c, 8.f, 8388608.f)), ld = __float_as_uint(fmaf(d, 8.f, 8388608.f));
  return __byte_perm(__byte_perm(la, lb, 0x0040), __byte_perm(lc, ld, 0x0040), 0x5410);
}

template <int CHUNKS>   // 16-byte int8 chunks per input pixel = Cin / 16
__global__ void __launch_bounds__(FP_THREADS, 1) fpn_merge_f16_kernel(const FpnP p) {
  extern __shared__ __align__(1024) uint8_t fp_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fp_raw) + 1023) & ~uintptr_t(1023));
  constexpr int b_plane = FP_COUT * 128;            // one of W_hi / W_lo: one 64-element K atom
  uint8_t* sB = smem;                               // [hi | lo]                       64 KB
  uint8_t* sA = sB + 2 * b_plane;                   // [2 stages]                      32 KB
  uint8_t* sP = sA + 2 * FP_A_BYTES;                // [2 stages] skewed patches      124.8 KB
  float* s_sc = reinterpret_cast<float*>(sP + 2 * FP_PATCH_BYTES);     // [256] scale, [256] shift
  float* s_sh = s_sc + FP_COUT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_sh + FP_COUT);        // acc_full[2], prev_full[2], slot_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const uint32_t bar0 = fp_u32(bars);
  auto acc_full = [&](int s) { return bar0 + 8u * s; };
  auto prev_full = [&](int s) { return bar0 + 16u + 8u * s; };
  auto slot_empty = [&](int s) { return bar0 + 32u + 8u * s; };

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(acc_full(s)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(prev_full(s)), "r"(FP_EW * 32));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(slot_empty(s)), "r"(FP_EW));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fp_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.bpack);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = tid; i < 2 * b_plane / 16; i += FP_THREADS) dst[i] = __ldg(src + i);
    for (int i = tid; i < FP_COUT; i += FP_THREADS) { s_sc[i] = __ldg(p.scale + i) * 0.125f; s_sh[i] = __ldg(p.shift + i) * 0.125f; }   // 1/8 scale: exact
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ A operand + MMA issue
    // int8 levels -> fp16, K-major SWIZZLE_128B rows (pixel m = lane + 32 j); the NEXT tile's levels are already in
    // registers while this tile's are converted, so the global latency is off the per-tile critical path.
    const uint32_t idesc = (1u << 4) | ((uint32_t)(FP_COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int ksteps = CHUNKS;                                   // K = 16 per MMA
    constexpr int chunks = CHUNKS;
    uint4 nxt[4][CHUNKS];
    auto load_a = [&](int tile) {
      const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = lane + 32 * j;
        const int y = min(ty * FP_TH + (m >> 4), p.H - 1), x = min(tx * FP_TW + (m & 15), p.W - 1);
        const uint4* src = reinterpret_cast<const uint4*>(p.a + (((int64_t)img * p.H + y) * p.W + x) * p.Cin);
#pragma unroll
        for (int c = 0; c < chunks; ++c) nxt[j][c] = __ldg(src + c);
      }
    };
    if ((int)blockIdx.x < p.tiles) load_a(blockIdx.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      if (it >= 2) fp_wait(slot_empty(s), (uint32_t)(((it >> 1) - 1) & 1));     // epilogue of tile it-2 is done with slot s
      {
        const uint32_t adst = fp_u32(sA) + (uint32_t)s * FP_A_BYTES;
        const __half2 bias = __floats2half2_rn(1024.f, 1024.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int m = lane + 32 * j;
          const uint32_t row = adst + (uint32_t)m * 128u;
#pragma unroll
          for (int c = 0; c < chunks; ++c) {
            {                                                          // 16 levels -> two 16-byte chunks of fp16
              const uint4 v = nxt[j][c];
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
              uint32_t h[8];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                // bytes (b0, b1) -> halves (1024 + b0, 1024 + b1) by PRMT into the mantissa of 0x6400, then - 1024: exact
                const uint32_t lo = __byte_perm(w[i], 0x64646464u, 0x4140), hi = __byte_perm(w[i], 0x64646464u, 0x4342);
                const __half2 hl = __hsub2(*reinterpret_cast<const __half2*>(&lo), bias);
                const __half2 hh = __hsub2(*reinterpret_cast<const __half2*>(&hi), bias);
                h[2 * i] = *reinterpret_cast<const uint32_t*>(&hl);
                h[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&hh);
              }
              const uint32_t c0 = (uint32_t)((2 * c) ^ (m & 7)) << 4, c1 = (uint32_t)((2 * c + 1) ^ (m & 7)) << 4;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + c0), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + c1), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
            }
          }
        }
      }
      if (tile + (int)gridDim.x < p.tiles) load_a(tile + gridDim.x);   // in flight behind the MMAs and the next slot wait
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = fp_u32(sA) + (uint32_t)s * FP_A_BYTES, b_hi = fp_u32(sB), b_lo = b_hi + b_plane;
        const uint32_t tacc = tmem_base + (uint32_t)(s * FP_COUT);
        for (int combo = 0; combo < 2; ++combo)
          for (int k = 0; k < ksteps; ++k)
            fp_mma_f16(tacc, fp_desc(a0) + (uint64_t)(k * 2), fp_desc(combo ? b_lo : b_hi) + (uint64_t)(k * 2), idesc,
                       (uint32_t)((combo | k) != 0));
        fp_commit(acc_full(s));
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue: lane = pixel, warp = 64 channels
    const int ew = warp - 1, q = warp & 3;              // TMEM lane quadrant is fixed by the warp index modulo 4
    // the four warps of a quadrant take the four 64-channel groups: rank of this warp among the warps with warp % 4 == q
    const int cg = (warp - (q == 0 ? 4 : q)) >> 2;      // warps q, q+4, q+8, q+12 (q = 0: 4, 8, 12, 16) -> 0..3
    const int m = q * 32 + lane;
    const float shy = (float)p.Hp / (float)p.H, swx = (float)p.Wp / (float)p.W;
    // The coarse patch of a tile (6 x 10 pixels x 1 KB) is fetched by these 512 threads themselves, one tile ahead:
    // 3840 16-byte cp.async, 7.5 per thread, all in flight at once; a warp's 32 chunks are half a pixel.
    auto issue_patch = [&](int tile, int s) {
      const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
      const int cy0 = ty * (FP_TH / 2) - 1, cx0 = tx * (FP_TW / 2) - 1;
      const float* pimg = p.prev + (int64_t)img * p.Hp * p.Wp * FP_COUT + ((ew & 1) * 32 + lane) * 4;
      const uint32_t pdst = fp_u32(sP) + (uint32_t)s * FP_PATCH_BYTES + (uint32_t)((ew & 1) * 32 + lane) * 16u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int lp = (ew >> 1) + 8 * i;                               // warp-uniform patch pixel
        if (lp < FP_PW * FP_PH) {
          const int gy = min(max(cy0 + lp / FP_PW, 0), p.Hp - 1), gx = min(max(cx0 + lp % FP_PW, 0), p.Wp - 1);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pdst + (uint32_t)lp * FP_PIX_BYTES),
                       "l"(pimg + ((int64_t)gy * p.Wp + gx) * FP_COUT) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(prev_full(s)) : "memory");
    };
    if ((int)blockIdx.x < p.tiles) issue_patch(blockIdx.x, 0);
    const uint32_t sc_s = fp_u32(s_sc) + (uint32_t)(cg * 64) * 4u, sh_s = fp_u32(s_sh) + (uint32_t)(cg * 64) * 4u;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      if (tile + (int)gridDim.x < p.tiles) {
        if (it >= 1) fp_wait(slot_empty(s ^ 1), (uint32_t)(((it - 1) >> 1) & 1));   // every warp is done with tile it-1
        issue_patch(tile + gridDim.x, s ^ 1);
      }
      const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
      const int y = ty * FP_TH + (m >> 4), x = tx * FP_TW + (m & 15);
      const bool ok = y < p.H && x < p.W;
      // upsample_bilinear2d, align_corners = False: src = (dst + 0.5) * scale - 0.5, clamped at 0
      float sy = shy * ((float)y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
      float sx = swx * ((float)x + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
      int y0 = (int)sy, x0 = (int)sx;
      y0 = min(y0, p.Hp - 1); x0 = min(x0, p.Wp - 1);                    // only for pixels outside the map (not stored)
      const int y1 = y0 + (y0 < p.Hp - 1 ? 1 : 0), x1 = x0 + (x0 < p.Wp - 1 ? 1 : 0);
      const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
      // everything below is carried at 1/8 scale (exact: a power of two), so that the final add saturates to [0, 1]
      const float2 hx2 = make_float2(hx, hx), lx2 = make_float2(lx, lx);
      const float2 hy8 = make_float2(hy * 0.125f, hy * 0.125f), ly8 = make_float2(ly * 0.125f, ly * 0.125f);
      const int cy0 = ty * (FP_TH / 2) - 1, cx0 = tx * (FP_TW / 2) - 1;
      const int r0 = min(max(y0 - cy0, 0), FP_PH - 1), r1 = min(max(y1 - cy0, 0), FP_PH - 1);
      const int c0 = min(max(x0 - cx0, 0), FP_PW - 1), c1 = min(max(x1 - cx0, 0), FP_PW - 1);
      const uint32_t pbase = fp_u32(sP) + (uint32_t)s * FP_PATCH_BYTES + (uint32_t)(cg * 64) * 4u;
      const uint32_t a00 = pbase + (uint32_t)(r0 * FP_PW + c0) * FP_PIX_BYTES, a01 = pbase + (uint32_t)(r0 * FP_PW + c1) * FP_PIX_BYTES;
      const uint32_t a10 = pbase + (uint32_t)(r1 * FP_PW + c0) * FP_PIX_BYTES, a11 = pbase + (uint32_t)(r1 * FP_PW + c1) * FP_PIX_BYTES;
      int8_t* dst = p.out_spike + (((int64_t)img * p.H + y) * p.W + x) * FP_COUT + cg * 64;
      const uint32_t par = (uint32_t)((it >> 1) & 1);
      fp_wait(acc_full(s), par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem_base + (uint32_t)(s * FP_COUT + cg * 64) + ((uint32_t)(q * 32) << 16);
      uint32_t v[2][16];
      fp_ld16(trow, v[0]);
      fp_wait(prev_full(s), par);
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cb < 3) fp_ld16(trow + (cb + 1) * 16, v[(cb + 1) & 1]);
        const uint32_t* vc = v[cb & 1];
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t co = (uint32_t)(cb * 16 + j * 4) * 4u;
          const float4 sc4 = fp_lds128(sc_s + co), sh4 = fp_lds128(sh_s + co);
          const float4 p00 = fp_lds128(a00 + co), p01 = fp_lds128(a01 + co), p10 = fp_lds128(a10 + co), p11 = fp_lds128(a11 + co);
          float2 ya = __ffma2_rn(make_float2(__uint_as_float(vc[j * 4 + 0]), __uint_as_float(vc[j * 4 + 1])),
                                 make_float2(sc4.x, sc4.y), make_float2(sh4.x, sh4.y));
          float2 yb = __ffma2_rn(make_float2(__uint_as_float(vc[j * 4 + 2]), __uint_as_float(vc[j * 4 + 3])),
                                 make_float2(sc4.z, sc4.w), make_float2(sh4.z, sh4.w));
          const float2 t0a = __ffma2_rn(lx2, make_float2(p01.x, p01.y), __fmul2_rn(hx2, make_float2(p00.x, p00.y)));
          const float2 t0b = __ffma2_rn(lx2, make_float2(p01.z, p01.w), __fmul2_rn(hx2, make_float2(p00.z, p00.w)));
          const float2 t1a = __ffma2_rn(lx2, make_float2(p11.x, p11.y), __fmul2_rn(hx2, make_float2(p10.x, p10.y)));
          const float2 t1b = __ffma2_rn(lx2, make_float2(p11.z, p11.w), __fmul2_rn(hx2, make_float2(p10.z, p10.w)));
          const float2 ua = __ffma2_rn(ly8, t1a, __fmul2_rn(hy8, t0a));
          const float2 ub = __ffma2_rn(ly8, t1b, __fmul2_rn(hy8, t0b));
          pk[j] = pack_unit4(fp_add_sat(ya.x, ua.x), fp_add_sat(ya.y, ua.y), fp_add_sat(yb.x, ub.x), fp_add_sat(yb.y, ub.y));
        }
        if (ok) *reinterpret_cast<uint4*>(dst + cb * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) fp_arrive(slot_empty(s));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Quad version: one epilogue lane = one 2 x 2 block of fine pixels = ONE coarse pixel's neighbourhood.
// The 2 x 2 fine pixels (2r, 2r+1) x (2k, 2k+1) interpolate between the 3 x 3 coarse pixels (r-1..r+1) x (k-1..k+1), so a
// lane that owns the whole block reads 9 shared-memory vectors for 4 outputs per channel instead of 16 -- the one-pixel
// version above was bound by exactly those LDS.128 wavefronts (63 % of the shared-memory pipe, ncu).  The four pixels of
// a block are four MMAs (sub-position (dy, dx) -> its own A tile and its own 64 TMEM columns), so a work unit is
//   16 x 8 blocks = 32 x 16 fine pixels  x  64 output channels        (4 units per pixel tile, 2 TMEM stages of 256 columns)
// and its coarse patch is 18 x 10 pixels x 64 channels (pixel stride 272 B: conflict-free LDS.128 across 8 lanes).
// Horizontal interpolation is done once per coarse row and shared by the two fine rows; weights follow
// upsample_bilinear2d(align_corners=False) on the fixed index pattern (k-1, k | k, k+1): (0.25, 0.75 | 0.75, 0.25), with
// (0, 1) for the first fine column / row, where ATen's clamped source coordinate has weight 1 on pixel 0.
// The 512 epilogue threads also fetch the next unit's patch (cp.async) and convert the next pixel tile's int8 levels to
// the fp16 A tiles (one A row per thread); warp 0 only waits on barriers and issues the MMAs.
constexpr int FQ_BW = 16, FQ_BH = 8;                              // blocks per tile
constexpr int FQ_PW = FQ_BW + 2, FQ_PH = FQ_BH + 2;               // patch: 18 x 10 coarse pixels
constexpr int FQ_PIX_BYTES = 64 * 4 + 16;                         // 272
constexpr int FQ_PATCH_BYTES = FQ_PW * FQ_PH * FQ_PIX_BYTES;      // 48960
constexpr int FQ_THREADS = FP_EW * 32;                          // 16 warps: 128 registers per thread
constexpr int FQ_A_SUB = 128 * 128;                               // one sub-position tile: 128 blocks x 64 fp16

__device__ __forceinline__ void fq_ld4(uint32_t taddr, float2& a, float2& b) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
  a = make_float2(__uint_as_float(r0), __uint_as_float(r1));
  b = make_float2(__uint_as_float(r2), __uint_as_float(r3));
}

template <int CHUNKS>
__global__ void __launch_bounds__(FQ_THREADS, 1) fpn_merge_quad_kernel(const FpnP p) {
  extern __shared__ __align__(1024) uint8_t fp_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fp_raw) + 1023) & ~uintptr_t(1023));
  constexpr int b_plane = FP_COUT * 128;
  uint8_t* sB = smem;                               // [hi | lo]                             64 KB
  uint8_t* sA = sB + 2 * b_plane;                   // [4 sub-positions]                     64 KB
  uint8_t* sP = sA + 4 * FQ_A_SUB;                  // [2 stages] skewed 64-channel patches  95.6 KB
  float* s_sc = reinterpret_cast<float*>(sP + 2 * FQ_PATCH_BYTES);
  float* s_sh = s_sc + FP_COUT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_sh + FP_COUT);   // acc_full[2], prev_full[2], slot_empty[2], a_full, a_free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t bar0 = fp_u32(bars);
  auto acc_full = [&](int s) { return bar0 + 8u * s; };
  auto prev_full = [&](int s) { return bar0 + 16u + 8u * s; };
  auto slot_empty = [&](int s) { return bar0 + 32u + 8u * s; };
  const uint32_t a_full = bar0 + 48u, a_free = bar0 + 56u;

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(acc_full(s)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(prev_full(s)), "r"(FP_EW * 32));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(slot_empty(s)), "r"(FP_EW));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_full), "r"(FP_EW));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_free), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fp_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.bpack);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = tid; i < 2 * b_plane / 16; i += FQ_THREADS) dst[i] = __ldg(src + i);
    for (int i = tid; i < FP_COUT; i += FQ_THREADS) { s_sc[i] = __ldg(p.scale + i) * 0.125f; s_sh[i] = __ldg(p.shift + i) * 0.125f; }
    uint4* za = reinterpret_cast<uint4*>(sA);       // K columns beyond Cin stay zero (never read by the MMAs, kept clean)
    for (int i = tid; i < 4 * FQ_A_SUB / 16; i += FQ_THREADS) za[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_img = p.tiles_x * p.tiles_y;

  {
    // ------------------------------------------------------------------ every warp: loaders + epilogue; warp 0 also issues
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // MMAs of unit (g) into TMEM stage s: four sub-position tiles x (W_hi, W_lo) x K steps; one elected lane of warp 0
    auto issue_mma = [&](int g, int s) {
      const uint32_t b_hi = fp_u32(sB) + (uint32_t)g * 64u * 128u, b_lo = b_hi + b_plane;
#pragma unroll
      for (int sub = 0; sub < 4; ++sub) {
        const uint32_t a0 = fp_u32(sA) + (uint32_t)sub * FQ_A_SUB;
        const uint32_t tacc = tmem_base + (uint32_t)(s * 256 + sub * 64);
#pragma unroll
        for (int combo = 0; combo < 2; ++combo)
#pragma unroll
          for (int k = 0; k < CHUNKS; ++k)
            fp_mma_f16(tacc, fp_desc(a0) + (uint64_t)(k * 2), fp_desc(combo ? b_lo : b_hi) + (uint64_t)(k * 2), idesc,
                       (uint32_t)((combo | k) != 0));
      }
      fp_commit(acc_full(s));
      if (g == 3) fp_commit(a_free);
    };
    const int ew = warp, q = warp & 3;                  // TMEM lane quadrant = warp index modulo 4
    const int cs = warp >> 2;                           // 16-channel slice of the unit's 64
    const int et = ew * 32 + lane;                      // 0..511
    const int m = q * 32 + lane, by = m >> 4, bx = m & 15;
    // ---- A rows: thread et converts row (et & 127) of sub-position (et >> 7)
    const int a_sub = et >> 7, a_m = et & 127, a_dy = a_sub >> 1, a_dx = a_sub & 1;
    uint4 nxt[CHUNKS];
    auto load_a = [&](int tile) {
      const int img = tile / tiles_img, tr = tile % tiles_img, ty = tr / p.tiles_x, tx = tr % p.tiles_x;
      const int y = min(ty * (2 * FQ_BH) + 2 * (a_m >> 4) + a_dy, p.H - 1), x = min(tx * (2 * FQ_BW) + 2 * (a_m & 15) + a_dx, p.W - 1);
      const uint4* src = reinterpret_cast<const uint4*>(p.a + (((int64_t)img * p.H + y) * p.W + x) * p.Cin);
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) nxt[c] = __ldg(src + c);
    };
    auto store_a = [&]() {
      const uint32_t row = fp_u32(sA) + (uint32_t)a_sub * FQ_A_SUB + (uint32_t)a_m * 128u;
      const __half2 bias = __floats2half2_rn(1024.f, 1024.f);
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        const uint32_t w[4] = {nxt[c].x, nxt[c].y, nxt[c].z, nxt[c].w};
        uint32_t h[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t lo = __byte_perm(w[i], 0x64646464u, 0x4140), hi = __byte_perm(w[i], 0x64646464u, 0x4342);
          const __half2 hl = __hsub2(*reinterpret_cast<const __half2*>(&lo), bias);
          const __half2 hh = __hsub2(*reinterpret_cast<const __half2*>(&hi), bias);
          h[2 * i] = *reinterpret_cast<const uint32_t*>(&hl);
          h[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        const uint32_t c0 = (uint32_t)((2 * c) ^ (a_m & 7)) << 4, c1 = (uint32_t)((2 * c + 1) ^ (a_m & 7)) << 4;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + c0), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + c1), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) fp_arrive(a_full);
    };
    // ---- coarse patch of unit (tile, g): 180 pixels x 16 chunks of 16 bytes over 512 threads
    auto issue_patch = [&](int tile, int g, int s) {
      const int img = tile / tiles_img, tr = tile % tiles_img, ty = tr / p.tiles_x, tx = tr % p.tiles_x;
      const int cy0 = ty * FQ_BH - 1, cx0 = tx * FQ_BW - 1;
      const float* pimg = p.prev + (int64_t)img * p.Hp * p.Wp * FP_COUT + g * 64 + (et & 15) * 4;
      const uint32_t pdst = fp_u32(sP) + (uint32_t)s * FQ_PATCH_BYTES + (uint32_t)(et & 15) * 16u;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int lp = (et >> 4) + 32 * i;
        if (lp < FQ_PW * FQ_PH) {
          const int gy = min(max(cy0 + lp / FQ_PW, 0), p.Hp - 1), gx = min(max(cx0 + lp % FQ_PW, 0), p.Wp - 1);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pdst + (uint32_t)lp * FQ_PIX_BYTES),
                       "l"(pimg + ((int64_t)gy * p.Wp + gx) * FP_COUT) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(prev_full(s)) : "memory");
    };
    if ((int)blockIdx.x < p.tiles) {
      load_a(blockIdx.x);
      store_a();
      if ((int)(blockIdx.x + gridDim.x) < p.tiles) load_a(blockIdx.x + gridDim.x);
      issue_patch(blockIdx.x, 0, 0);
      if (warp == 0) {
        fp_wait(a_full, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) issue_mma(0, 0);
        __syncwarp();
      }
    }
    int it = 0, u = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int img = tile / tiles_img, tr = tile % tiles_img, ty = tr / p.tiles_x, tx = tr % p.tiles_x;
      const int yT = ty * (2 * FQ_BH) + 2 * by, xL = tx * (2 * FQ_BW) + 2 * bx;
      // weights on the fixed pattern; everything vertical is carried at 1/8 scale (exact)
      const float wL0 = xL == 0 ? 0.f : 0.25f, wL1 = xL == 0 ? 1.f : 0.75f;
      const float wT0 = yT == 0 ? 0.f : 0.03125f, wT1 = yT == 0 ? 0.125f : 0.09375f;
      const float wR0 = 0.75f, wR1 = 0.25f, wB0 = 0.09375f, wB1 = 0.03125f;
      const bool okx0 = xL < p.W, okx1 = xL + 1 < p.W, oky0 = yT < p.H, oky1 = yT + 1 < p.H;
      int8_t* dst = p.out_spike + (((int64_t)img * p.H + yT) * p.W + xL) * FP_COUT + cs * 16;
      const int64_t row_b = (int64_t)p.W * FP_COUT;
      for (int g = 0; g < 4; ++g, ++u) {
        const int s = u & 1;
        // ---- one unit ahead: the next unit's patch, (at g == 3) the next pixel tile's A rows, and its MMAs
        {
          const int ng = (g + 1) & 3, ntile = g == 3 ? tile + (int)gridDim.x : tile;
          if (ntile < p.tiles) {
            if (u >= 1) {
              fp_wait(slot_empty(s ^ 1), (uint32_t)(((u - 1) >> 1) & 1));       // every warp is done with unit u - 1
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            issue_patch(ntile, ng, s ^ 1);
            if (g == 3) {
              // this tile's last MMAs were issued one unit ago: once they have completed the A tiles take the next
              // tile's levels (already in registers), and the tile after that is requested
              fp_wait(a_free, (uint32_t)(it & 1));
              store_a();
              if (tile + 2 * (int)gridDim.x < p.tiles) load_a(tile + 2 * gridDim.x);
            }
            if (warp == 0) {
              if (g == 3) {
                fp_wait(a_full, (uint32_t)((it + 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              }
              if (lane == 0) issue_mma(ng, s ^ 1);
              __syncwarp();
            }
          }
        }
        const uint32_t par = (uint32_t)((u >> 1) & 1);
        fp_wait(acc_full(s), par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        fp_wait(prev_full(s), par);
        // per-lane geometry re-derived from the thread index here (an opaque read): kept live across the loaders above
        // it was spilled, and the reload queued behind the cp.asyncs in the load pipe (11 % of all stall samples)
        uint32_t t_;
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t_));
        const uint32_t cs_ = t_ >> 7, q_ = (t_ >> 5) & 3u, by_ = (t_ & 127u) >> 4, bx_ = t_ & 15u;
        const uint32_t trow = tmem_base + (uint32_t)(s * 256) + cs_ * 16u + ((q_ * 32u) << 16);
        const uint32_t pb = fp_u32(sP) + (uint32_t)s * FQ_PATCH_BYTES + (by_ * FQ_PW + bx_) * FQ_PIX_BYTES + cs_ * 64u;
        const uint32_t sc_s = fp_u32(s_sc) + ((uint32_t)(g * 64) + cs_ * 16u) * 4u, sh_s = fp_u32(s_sh) + ((uint32_t)(g * 64) + cs_ * 16u) * 4u;
        uint32_t pk[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 acc[4][2];
#pragma unroll
          for (int sub = 0; sub < 4; ++sub) fq_ld4(trow + (uint32_t)(sub * 64 + j * 4), acc[sub][0], acc[sub][1]);
          const uint32_t co = (uint32_t)(j * 4) * 4u;
          const float4 sc4 = fp_lds128(sc_s + co), sh4 = fp_lds128(sh_s + co);
          // horizontal interpolation of one coarse row: left / right fine column, channel pairs (0,1) and (2,3)
          auto hrow = [&](int rr, float2 (&l)[2], float2 (&r)[2]) {
            const uint32_t ra = pb + (uint32_t)(rr * FQ_PW) * FQ_PIX_BYTES + co;
            const float4 c0 = fp_lds128(ra), c1 = fp_lds128(ra + FQ_PIX_BYTES), c2 = fp_lds128(ra + 2 * FQ_PIX_BYTES);
            l[0] = __ffma2_rn(make_float2(wL1, wL1), make_float2(c1.x, c1.y), __fmul2_rn(make_float2(wL0, wL0), make_float2(c0.x, c0.y)));
            l[1] = __ffma2_rn(make_float2(wL1, wL1), make_float2(c1.z, c1.w), __fmul2_rn(make_float2(wL0, wL0), make_float2(c0.z, c0.w)));
            r[0] = __ffma2_rn(make_float2(wR1, wR1), make_float2(c2.x, c2.y), __fmul2_rn(make_float2(wR0, wR0), make_float2(c1.x, c1.y)));
            r[1] = __ffma2_rn(make_float2(wR1, wR1), make_float2(c2.z, c2.w), __fmul2_rn(make_float2(wR0, wR0), make_float2(c1.z, c1.w)));
          };
          auto emit = [&](int sub, float w0, float w1, const float2 (&t0)[2], const float2 (&t1)[2]) {
            const float2 ua = __ffma2_rn(make_float2(w1, w1), t1[0], __fmul2_rn(make_float2(w0, w0), t0[0]));
            const float2 ub = __ffma2_rn(make_float2(w1, w1), t1[1], __fmul2_rn(make_float2(w0, w0), t0[1]));
            const float2 ya = __ffma2_rn(acc[sub][0], make_float2(sc4.x, sc4.y), make_float2(sh4.x, sh4.y));
            const float2 yb = __ffma2_rn(acc[sub][1], make_float2(sc4.z, sc4.w), make_float2(sh4.z, sh4.w));
            pk[sub][j] = pack_unit4(fp_add_sat(ya.x, ua.x), fp_add_sat(ya.y, ua.y), fp_add_sat(yb.x, ub.x), fp_add_sat(yb.y, ub.y));
          };
          float2 la[2], ra_[2], lb[2], rb[2];
          hrow(0, la, ra_);
          hrow(1, lb, rb);
          // the accumulator registers are operands of the wait, so that no use of them can be scheduled above it
          asm volatile("tcgen05.wait::ld.sync.aligned;"
                       : "+f"(acc[0][0].x), "+f"(acc[0][0].y), "+f"(acc[0][1].x), "+f"(acc[0][1].y), "+f"(acc[1][0].x), "+f"(acc[1][0].y),
                         "+f"(acc[1][1].x), "+f"(acc[1][1].y), "+f"(acc[2][0].x), "+f"(acc[2][0].y), "+f"(acc[2][1].x), "+f"(acc[2][1].y),
                         "+f"(acc[3][0].x), "+f"(acc[3][0].y), "+f"(acc[3][1].x), "+f"(acc[3][1].y)
                       :: "memory");
          emit(0, wT0, wT1, la, lb);
          emit(1, wT0, wT1, ra_, rb);
          hrow(2, la, ra_);                         // row 0's registers are free again
          emit(2, wB0, wB1, lb, la);
          emit(3, wB0, wB1, rb, ra_);
        }
        int8_t* d = dst + g * 64;
        if (oky0 && okx0) *reinterpret_cast<uint4*>(d) = make_uint4(pk[0][0], pk[0][1], pk[0][2], pk[0][3]);
        if (oky0 && okx1) *reinterpret_cast<uint4*>(d + FP_COUT) = make_uint4(pk[1][0], pk[1][1], pk[1][2], pk[1][3]);
        if (oky1 && okx0) *reinterpret_cast<uint4*>(d + row_b) = make_uint4(pk[2][0], pk[2][1], pk[2][2], pk[2][3]);
        if (oky1 && okx1) *reinterpret_cast<uint4*>(d + row_b + FP_COUT) = make_uint4(pk[3][0], pk[3][1], pk[3][2], pk[3][3]);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) fp_arrive(slot_empty(s));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

}  // namespace s2f

using namespace s2f;

extern "C" int s2f_fpn_merge_f16(const int8_t* a, const void* w_packed, const float* scale, const float* shift,
                                 const float* prev, int8_t* out_spike, int n, int H, int W, int Cin, int Cout, int Hp, int Wp,
                                 float d_max, void* stream) {
  S2F_REQUIRE(a && w_packed && scale && shift && prev && out_spike, "fpn_merge_f16: null pointer");
  S2F_REQUIRE(Cout == FP_COUT, "fpn_merge_f16: Cout must be 256");
  S2F_REQUIRE(Cin == 16 || Cin == 32 || Cin == 48 || Cin == 64, "fpn_merge_f16: Cin must be a multiple of 16, at most 64");
  S2F_REQUIRE(H == 2 * Hp && W == 2 * Wp, "fpn_merge_f16: exact x2 upsampling only");
  S2F_REQUIRE(d_max == 8.f, "fpn_merge_f16: d_max must be 8");
  S2F_REQUIRE((int64_t)n * H * W < (1ll << 31), "fpn_merge_f16: problem too large");
  FpnP p;
  p.a = a; p.bpack = reinterpret_cast<const uint8_t*>(w_packed); p.scale = scale; p.shift = shift; p.prev = prev;
  p.out_spike = out_spike; p.n = n; p.H = H; p.W = W; p.Cin = Cin; p.Hp = Hp; p.Wp = Wp;
  static const bool use_quad = []() { const char* e = getenv("S2F_FPN_QUAD"); return !(e && e[0] == '0'); }();
  if (use_quad) {
    p.tiles_x = (W + 2 * FQ_BW - 1) / (2 * FQ_BW); p.tiles_y = (H + 2 * FQ_BH - 1) / (2 * FQ_BH); p.tiles = n * p.tiles_x * p.tiles_y;
    const size_t smemq = 1024 + 2 * FP_COUT * 128 + 4 * FQ_A_SUB + 2 * FQ_PATCH_BYTES + 2 * FP_COUT * sizeof(float) + 8 * 8 + 16;
    static std::atomic<uint64_t> onceq{0};
    if (first_use_on_this_device(onceq)) {
      cudaFuncSetAttribute(fpn_merge_quad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemq);
      cudaFuncSetAttribute(fpn_merge_quad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemq);
      cudaFuncSetAttribute(fpn_merge_quad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemq);
      cudaFuncSetAttribute(fpn_merge_quad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemq);
    }
    const int gridq = p.tiles < sm_count() ? p.tiles : sm_count();
    switch (Cin >> 4) {
      case 1: fpn_merge_quad_kernel<1><<<gridq, FQ_THREADS, smemq, (cudaStream_t)stream>>>(p); break;
      case 2: fpn_merge_quad_kernel<2><<<gridq, FQ_THREADS, smemq, (cudaStream_t)stream>>>(p); break;
      case 3: fpn_merge_quad_kernel<3><<<gridq, FQ_THREADS, smemq, (cudaStream_t)stream>>>(p); break;
      default: fpn_merge_quad_kernel<4><<<gridq, FQ_THREADS, smemq, (cudaStream_t)stream>>>(p); break;
    }
    return check_launch("fpn_merge_quad_kernel");
  }
  p.tiles_x = (W + FP_TW - 1) / FP_TW; p.tiles_y = (H + FP_TH - 1) / FP_TH; p.tiles = n * p.tiles_x * p.tiles_y;
  const size_t smem = 1024 + 2 * FP_COUT * 128 + 2 * FP_A_BYTES + 2 * FP_PATCH_BYTES + 2 * FP_COUT * sizeof(float) + 6 * 8 + 16;
  static std::atomic<uint64_t> once{0};
  if (first_use_on_this_device(once)) {
    cudaFuncSetAttribute(fpn_merge_f16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(fpn_merge_f16_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(fpn_merge_f16_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(fpn_merge_f16_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  const int grid = p.tiles < sm_count() ? p.tiles : sm_count();
  switch (Cin >> 4) {
    case 1: fpn_merge_f16_kernel<1><<<grid, FP_THREADS, smem, (cudaStream_t)stream>>>(p); break;
    case 2: fpn_merge_f16_kernel<2><<<grid, FP_THREADS, smem, (cudaStream_t)stream>>>(p); break;
    case 3: fpn_merge_f16_kernel<3><<<grid, FP_THREADS, smem, (cudaStream_t)stream>>>(p); break;
    default: fpn_merge_f16_kernel<4><<<grid, FP_THREADS, smem, (cudaStream_t)stream>>>(p); break;
  }
  return check_launch("fpn_merge_f16_kernel");
}
