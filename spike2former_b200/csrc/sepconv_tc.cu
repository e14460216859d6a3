// SepConv tail in ONE launch (sdtv2.py:176-179): depthwise k x k over the int8 spike levels, then the 1x1 `pwconv2`
// over the REAL-valued stencil output + folded BatchNorm (+ residual) (+ NI-LIF of the result).
//
//   x2 = dwconv(spikes)            fp32, exact reference tap order (dw_tile.cuh)           -- CUDA cores, FFMA2
//   y  = W_pw x2 * scale + shift   tcgen05.mma.kind::f16, fp32 accumulate in TMEM         -- tensor cores
//
// The stencil output never reaches HBM (the two-kernel form wrote and re-read 1.07 GB of fp32 per 32 images) and the
// pointwise product no longer runs on legacy mma.sync TF32.  Both operands of the 1x1 are real numbers, so each is split
// into fp16 hi + lo (22 significant bits) and three products are accumulated: hi*hi + lo*hi + hi*lo (the dropped
// lo*lo term is < 2^-22 relative).  Weights are pre-split on the host with one power-of-two scale per output channel
// (ops.pack_pw_f16), the stencil output is split by the threads that produced it, straight into the K-major
// SWIZZLE_128B operand image.
//
// CTA = G stencil groups of 128 threads, persistent over 16 x 8 pixel tiles.  Each group owns a tile at a time: per
// 64-channel round its threads (8 x 2 pixels x 4 channels each) run the stencil in registers, wait until the tensor core
// has finished reading the group's A buffer, store hi / lo and arrive on the group's `full` barrier; warp 0 of the group
// then issues the 12 MMAs of the round (M = 128 pixels, N = Cout, K = 16) into one of the group's TWO TMEM accumulators
// and commits to `free` (and to `accfull` after the last round).  The accumulator of tile i is drained after the
// stencil of tile i + 1 (warp q <-> TMEM lanes 32q..32q+31: affine + residual + fp32 / spike stores), so neither the MMA
// latency nor the accumulator wait is ever exposed.  The tensor core idles most of the time by design: the stencil bounds.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "dw_tile.cuh"

namespace s2f {

constexpr int SC_GT = 128;                 // threads per stencil group
constexpr int SC_TW = 16, SC_TH = 8;       // tile: 16 x 8 pixels = the 128 rows of one MMA

struct SepP {
  const int8_t* a; const float* w_dw; const uint8_t* bpack; const float* scale; const float* shift; const float* residual;
  float* out_f32; int8_t* out_spike;
  int n, H, W, Cm, Cout, Np, tiles_x, tiles_y, tiles;
  float a_mul;          // stencil accumulator -> A operand (powers of two: 2^49 * a_scale * A_PRE)
  float d_max;
};

__device__ __forceinline__ uint32_t sc_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sc_desc(uint32_t saddr) {      // K-major SWIZZLE_128B, SBO = 1024 B, version 1
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void sc_mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void sc_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void sc_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "SC_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra SC_DONE;\n\t"
      "bra SC_WAIT;\n\t"
      "SC_DONE:\n\t"
      "}" ::"r"(sc_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool sc_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}" : "=r"(ok) : "r"(sc_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void sc_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sc_u32(bar)) : "memory");
}
__device__ __forceinline__ void sc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sc_u32(bar)) : "memory");
}

__device__ __forceinline__ float4 sc_lds128(uint32_t saddr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// The register file is split per scheduler (16 K registers each): with the MMA issued from a ninth warp one scheduler
// would hold three warps and ptxas caps every thread at 168 registers (the stencil then spills its prefetched row).
// So the CTA is exactly G * 4 warps and warp 0 of each group issues its group's MMAs.
template <int KS, int CT, int G>
__global__ void __launch_bounds__(G * SC_GT, 1) sepconv_dwpw_kernel(const SepP p) {
  extern __shared__ __align__(1024) uint8_t sc_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sc_raw) + 1023) & ~uintptr_t(1023));
  const int Cm = CT ? CT : p.Cm;
  const int KB = Cm >> 6;                           // 64-channel rounds
  const int c4n = Cm >> 2;
  const int b_plane = KB * p.Np * 128;              // one of B_hi / B_lo
  constexpr int a_plane = 128 * 128;                // one of A_hi / A_lo of a group: 16 KB
  uint8_t* sB = smem;
  uint8_t* sA = smem + 2 * b_plane;                 // [G][hi | lo]
  float4* wsm = reinterpret_cast<float4*>(sA + G * 2 * a_plane);         // [G][KS*KS][16]: the current round's 64 channels
  float* s_sc = reinterpret_cast<float*>(wsm + G * KS * KS * 16);        // [Np] scale, [Np] shift
  float* s_sh = s_sc + p.Np;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_sh + p.Np);         // [G] full, [G] free, [G][2] accfull
  uint64_t* bar_free = bar_full + G;
  uint64_t* bar_acc = bar_free + G;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 2 * G);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    for (int g = 0; g < G; ++g) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sc_u32(bar_full + g)), "r"(SC_GT));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sc_u32(bar_free + g)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sc_u32(bar_acc + 2 * g)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sc_u32(bar_acc + 2 * g + 1)), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int ACC_COLS = G <= 2 ? 128 : 64;       // columns per accumulator (>= Np, host-checked); two per group
  constexpr int TMEM_COLS = 512;
  static_assert(G * 2 * ACC_COLS <= 512, "two accumulators per group");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sc_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.bpack);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = tid; i < 2 * b_plane / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = tid; i < p.Np; i += blockDim.x) {
      s_sc[i] = i < p.Cout ? __ldg(p.scale + i) : 0.f;
      s_sh[i] = i < p.Cout ? __ldg(p.shift + i) : 0.f;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // B image: generic-proxy writes -> tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int stride = gridDim.x * G;

  const int g = warp >> 2, tg = tid - g * SC_GT, q = warp & 3;
  const int cq = tg & 15, pt = tg >> 4, sx = pt & 1, sy = pt >> 1;
  const uint32_t a_hi = sc_u32(sA + g * 2 * a_plane), a_lo = a_hi + a_plane;
  const uint32_t b_hi = sc_u32(sB), b_lo = b_hi + b_plane;
  float4* wg = wsm + g * KS * KS * 16;
  const uint32_t wg_s = sc_u32(wg), sc_s = sc_u32(s_sc), sh_s = sc_u32(s_sh);
  const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr int pad = (KS - 1) / 2;
  auto stage_round = [&](int r) {                                    // this round's k*k x 64 stencil weights, pre-scaled
    for (int i = tg; i < KS * KS * 16; i += SC_GT) {
      float4 v = __ldg(reinterpret_cast<const float4*>(p.w_dw) + (i >> 4) * c4n + r * 16 + (i & 15));
      v.x *= DwScale<int8_t>::w_pre; v.y *= DwScale<int8_t>::w_pre; v.z *= DwScale<int8_t>::w_pre; v.w *= DwScale<int8_t>::w_pre;
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(wg_s + (uint32_t)i * 16u), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(SC_GT) : "memory");
  };
  // drain accumulator `buf` of this group into the outputs of tile (img, ty, tx): lane <-> pixel row m = 32 q + lane
  auto drain = [&](int tile, uint32_t buf, uint32_t parity) {
    const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
    const int m = q * 32 + lane;
    const int y = ty * SC_TH + (m >> 4), x = tx * SC_TW + (m & 15);
    const bool ok = y < p.H && x < p.W;
    const int64_t o0 = (((int64_t)img * p.H + y) * p.W + x) * p.Cout;
    const int ncb = p.Np >> 4;
    float4 res[4];
    auto load_res = [&](int cb) {
      if (p.residual && ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (cb * 16 + j * 4 < p.Cout) res[j] = __ldg(reinterpret_cast<const float4*>(p.residual + o0 + cb * 16 + j * 4));
      }
    };
    load_res(0);
    sc_wait(bar_acc + 2 * g + buf, parity);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t trow = tmem_base + (uint32_t)((g * 2 + buf) * ACC_COLS) + ((uint32_t)(q * 32) << 16);
    for (int cb = 0; cb < ncb; ++cb) {
      uint32_t v[16];
      sc_ld16(trow + cb * 16, v);
      float4 rc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) rc[j] = res[j];
      if (cb + 1 < ncb) load_res(cb + 1);                              // next chunk's residual under this chunk's math
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (ok) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = cb * 16 + j * 4;
          const float4 sc4 = sc_lds128(sc_s + c * 4), sh4 = sc_lds128(sh_s + c * 4);
          float y0 = fmaf(__uint_as_float(v[j * 4 + 0]), sc4.x, sh4.x), y1 = fmaf(__uint_as_float(v[j * 4 + 1]), sc4.y, sh4.y);
          float y2 = fmaf(__uint_as_float(v[j * 4 + 2]), sc4.z, sh4.z), y3 = fmaf(__uint_as_float(v[j * 4 + 3]), sc4.w, sh4.w);
          if (p.residual && c < p.Cout) { y0 += rc[j].x; y1 += rc[j].y; y2 += rc[j].z; y3 += rc[j].w; }
          if (p.out_f32 && c < p.Cout) *reinterpret_cast<float4*>(p.out_f32 + o0 + c) = make_float4(y0, y1, y2, y3);
          pk[j] = pack_levels4(y0, y1, y2, y3, p.d_max);
        }
        if (p.out_spike) {
          if (cb * 16 + 16 <= p.Cout) {
            *reinterpret_cast<uint4*>(p.out_spike + o0 + cb * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (cb * 16 + j * 4 < p.Cout) *reinterpret_cast<uint32_t*>(p.out_spike + o0 + cb * 16 + j * 4) = pk[j];
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");    // a later round-0 MMA overwrites this accumulator
  };

  // the residual rows a drain will read, requested one stencil round ahead (the drain's loads then hit L1 / L2)
  auto prefetch_res = [&](int tile) {
    if (!p.residual) return;
    const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
    const int m = q * 32 + lane;
    const int y = ty * SC_TH + (m >> 4), x = tx * SC_TW + (m & 15);
    if (y < p.H && x < p.W) {
      const float* r0 = p.residual + (((int64_t)img * p.H + y) * p.W + x) * p.Cout;
      for (int c = 0; c < p.Cout; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + c));
    }
  };
  if (KB == 1) stage_round(0);
  uint32_t fills = 0, tiles_done = 0;
  int prev_tile = -1;
  for (int tile = blockIdx.x * G + g; tile < p.tiles; tile += stride) {
    const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
    const int ho0 = ty * SC_TH + sy * 2, wo0 = tx * SC_TW + sx * DW_TW;
    const int8_t* img_base = p.a + (int64_t)img * p.H * p.W * Cm;
    const uint32_t buf = tiles_done & 1;
    for (int r = 0; r < KB; ++r) {
      if (KB > 1) {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(SC_GT) : "memory");      // everyone is done with the old weights
        stage_round(r);
      }
      if (r == KB - 1 && prev_tile >= 0) prefetch_res(prev_tile);
      float2 acc[2][DW_TW][2];
      dw_tile_8x2<int8_t, KS, CT>(img_base + r * 64 + cq * 4, p.H, p.W, Cm, ho0, wo0, pad, wg_s + (uint32_t)cq * 16u, 256u, acc);
      if (fills > 0) sc_wait(bar_free + g, (fills - 1) & 1);       // the MMAs of the previous round have read A
#pragma unroll
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int px = 0; px < DW_TW; ++px) {
          const int m = (sy * 2 + t) * SC_TW + sx * DW_TW + px;
          const float v0 = acc[t][px][0].x * p.a_mul, v1 = acc[t][px][0].y * p.a_mul;
          const float v2 = acc[t][px][1].x * p.a_mul, v3 = acc[t][px][1].y * p.a_mul;
          const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y), l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
          const uint32_t off = (uint32_t)(m * 128 + ((((cq >> 1) ^ (m & 7)) << 4) | ((cq & 1) << 3)));
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_hi + off), "r"(*reinterpret_cast<const uint32_t*>(&h01)),
                       "r"(*reinterpret_cast<const uint32_t*>(&h23)) : "memory");
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_lo + off), "r"(*reinterpret_cast<const uint32_t*>(&l01)),
                       "r"(*reinterpret_cast<const uint32_t*>(&l23)) : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      sc_arrive(bar_full + g);
      if (q == 0) {
        // ---- this group's MMA issuer: the whole warp waits for the group's A image, one lane issues
        sc_wait(bar_full + g, fills & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t bk = (uint32_t)(r * p.Np * 128);
          const uint32_t tacc = tmem_base + (uint32_t)((g * 2 + buf) * ACC_COLS);
#pragma unroll
          for (int combo = 0; combo < 3; ++combo) {
            const uint32_t abase = combo == 1 ? a_lo : a_hi, bbase = (combo == 2 ? b_lo : b_hi) + bk;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              sc_mma_f16(tacc, sc_desc(abase) + (uint64_t)(k * 2), sc_desc(bbase) + (uint64_t)(k * 2), idesc,
                         (uint32_t)((r | combo | k) != 0));
          }
          sc_commit(bar_free + g);
          if (r == KB - 1) sc_commit(bar_acc + 2 * g + buf);
        }
        __syncwarp();
      }
      ++fills;
    }
    // the PREVIOUS tile's accumulator has been complete for a whole tile time: drain it now, no wait exposed
    if (prev_tile >= 0) drain(prev_tile, buf ^ 1, ((tiles_done - 1) >> 1) & 1);
    prev_tile = tile;
    ++tiles_done;
  }
  if (prev_tile >= 0) {
    prefetch_res(prev_tile);
    drain(prev_tile, (tiles_done - 1) & 1, ((tiles_done - 1) >> 1) & 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

}  // namespace s2f

using namespace s2f;

extern "C" int64_t s2f_sepconv_bpack_bytes(int Cm, int Cout) {
  const int Np = (Cout + 15) / 16 * 16;
  return (int64_t)2 * (Cm / 64) * Np * 128;
}

extern "C" int s2f_sepconv_dwpw(const int8_t* a, float a_scale, const float* w_dw, const void* w_pw_packed, float a_pre,
                                const float* scale, const float* shift, const float* residual, float* out_f32,
                                int8_t* out_spike, int n, int H, int W, int Cm, int Cout, int k, float d_max, void* stream) {
  S2F_REQUIRE(a && w_dw && w_pw_packed && scale && shift && (out_f32 || out_spike), "sepconv_dwpw: null pointer");
  S2F_REQUIRE(k == 7 || k == 3 || k == 5, "sepconv_dwpw: k must be 3, 5 or 7");
  S2F_REQUIRE(Cm % 64 == 0 && Cm >= 64, "sepconv_dwpw: the depthwise width must be a multiple of 64");
  S2F_REQUIRE(Cout % 4 == 0 && Cout >= 16 && Cout <= 128, "sepconv_dwpw: 16 <= Cout <= 128, Cout % 4 == 0");
  S2F_REQUIRE((int64_t)W * Cm < (1ll << 31) && (int64_t)n * H * W < (1ll << 31), "sepconv_dwpw: problem too large");
  SepP p;
  p.a = a; p.w_dw = w_dw; p.bpack = reinterpret_cast<const uint8_t*>(w_pw_packed); p.scale = scale; p.shift = shift;
  p.residual = residual; p.out_f32 = out_f32; p.out_spike = out_spike;
  p.n = n; p.H = H; p.W = W; p.Cm = Cm; p.Cout = Cout; p.Np = (Cout + 15) / 16 * 16;
  p.tiles_x = (W + SC_TW - 1) / SC_TW; p.tiles_y = (H + SC_TH - 1) / SC_TH; p.tiles = n * p.tiles_x * p.tiles_y;
  p.a_mul = a_scale * DwScale<int8_t>::post * a_pre;
  p.d_max = d_max;
  auto smem_for = [&](int G) {
    return 1024 + (size_t)s2f_sepconv_bpack_bytes(Cm, Cout) + (size_t)G * 2 * 128 * 128 + (size_t)G * k * k * 64 * sizeof(float) +
           (size_t)2 * p.Np * sizeof(float) + 4 * G * sizeof(uint64_t) + 16;
  };
  // Two stencil groups = 8 warps = two per scheduler, up to 255 registers each.  Three groups (12 warps, capped at 168
  // registers, the prefetched row spills) measured the same or slower: S2F_SEPCONV_GROUPS=3 keeps the experiment.
  static const char* genv = getenv("S2F_SEPCONV_GROUPS");
  int G = 2;
  if (genv && genv[0] == '3' && p.Np <= 64 && smem_for(3) <= 227 * 1024) G = 3;
  const size_t smem = smem_for(G);
  S2F_REQUIRE(smem <= 227 * 1024, "sepconv_dwpw: operands do not fit shared memory");
  const int ctas = (int)ceil_div(p.tiles, G);
  const int grid = ctas < sm_count() ? ctas : sm_count();
  cudaStream_t st = (cudaStream_t)stream;
#define S2F_SC_G(KS, CT, G_)                                                                                         \
  do {                                                                                                              \
    static std::atomic<uint64_t> once{0};                                                                           \
    if (first_use_on_this_device(once))                                                                             \
      cudaFuncSetAttribute(sepconv_dwpw_kernel<KS, CT, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
    sepconv_dwpw_kernel<KS, CT, G_><<<grid, G_ * SC_GT, smem, st>>>(p);                                             \
  } while (0)
#define S2F_SC(KS, CT)                                                  \
  do {                                                                  \
    if (G == 3) S2F_SC_G(KS, CT, 3); else S2F_SC_G(KS, CT, 2);          \
  } while (0)
#define S2F_SC_K(KS)                              \
  do {                                            \
    if (Cm == 64) S2F_SC(KS, 64);                 \
    else if (Cm == 128) S2F_SC(KS, 128);          \
    else if (Cm == 256) S2F_SC(KS, 256);          \
    else S2F_SC(KS, 0);                           \
  } while (0)
  if (k == 7) S2F_SC_K(7); else if (k == 5) S2F_SC_K(5); else S2F_SC_K(3);
#undef S2F_SC
#undef S2F_SC_G
#undef S2F_SC_K
  return check_launch("sepconv_dwpw_kernel");
}
