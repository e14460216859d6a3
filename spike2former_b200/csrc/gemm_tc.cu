// Spike GEMM / implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a).
//
//   A  int8 spike levels, channels-last, fetched by TMA (2D for 1x1, 4D boxes with zero-filled halos for
//      3x3 / strided convolutions) into 128B/64B/32B-swizzled K-major shared-memory tiles;
//   B  fp32 weights split on the host into `pieces` signed base-128 digit planes (int8) with one power-of-two
//      scale per output channel; the planes of a 64-channel tile are stacked along N, so ONE
//      tcgen05.mma.kind::i8 (M=128, N=64*pieces, K=32) feeds all planes and the int32 accumulation is exact;
//   D  int32 accumulators in TMEM, read back with tcgen05.ld by the epilogue warps, which recombine the digit planes,
//      apply the folded BatchNorm affine (+ residual / fused FPN merge) and emit fp32 and/or NI-LIF int8 levels.
//      The epilogue is compiled per output kind (EPI_GENERIC / EPI_SPIKE / EPI_STAGED / EPI_STAGED_UP, see below).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..9 = epilogue (TMEM lane quadrant = warp_id % 4, column half = (warp_id - 2) / 4).  The kernel is
// persistent (one CTA per SM) with two accumulators in TMEM, so the epilogue of a tile overlaps the next main loop.
// The warp index is taken through a shuffle so that the compiler treats the role branches as warp-uniform.
#include <cuda.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace s2f {

constexpr int TC_BM = 128;        // rows (pixels / tokens) per tile == UMMA M
constexpr int TC_BN = 64;         // output channels per tile
// Epilogue warps per CTA (template parameter EW): 8 -> 32 channels per warp, 16 -> 16 channels per warp.  Measured on
// B200: EW = 16 (4 warps per scheduler, 96 registers) changes nothing -- 14.3 us vs 14.4 us for 256->512 at 16 K rows.
// The small-K layers are bound by draining the accumulator: three int32 planes = 96 KB of tcgen05.ld per 128x64 tile at
// ~64 B/clk/SM = 1536 clk, during which the MMAs of the next tile make little progress; per-tile time fits
// 3*K + 1536 clk (K=256: 2300 measured, K=512: ~3000, K=2304: ~8400 = 82 % MMA).  EW = 8 is the default.
__host__ __device__ constexpr int tc_threads(int ew) { return 64 + 32 * ew; }       // warp 0 TMA, warp 1 MMA, the rest epilogue

__host__ __device__ inline int tc_bk(int cin) { return cin >= 128 ? 128 : (cin >= 64 ? 64 : 32); }
__host__ __device__ inline int tc_cin_pad(int cin) { const int bk = tc_bk(cin); return (cin + bk - 1) / bk * bk; }

struct TcParams {
  const float* scale; const float* shift; const float* residual;
  const float* up_prev; int up_H, up_W;     // fused FPN merge: y += bilinear_up(up_prev [n, up_H, up_W, Cout])
  float* out_f32; int8_t* out_spike;
  int M_total;          // n * Ho * Wo
  int M_img;            // Ho * Wo
  int Ho, Wo, Cout;
  int mode_conv;        // 0: 2D rows, 1: 4D boxes
  int taps_w, taps;     // KW, KH*KW
  int stride, pad;
  int TW, TH, tiles_w, tiles_h;
  int cin_chunks;       // ceil(Cin / BK)
  int bk;               // bytes of K per stage (32 / 64 / 128)
  int pieces, stages;
  int out_transposed;
  int tiles_n, tiles_m, ctas_per_n;
  int group;            // K chunks per pipeline stage (3 = one kernel row of a 3x3 layer with narrow Cin)
  int b_resident;       // all K chunks of the weight tile stay in shared memory for the CTA's lifetime
  int w_img_rows;       // packed weight rows per image (0: one weight matrix for all images)
  int ss_img_stride;    // scale/shift elements per image (0: shared)
  float d_max;
  int up_tma;           // the coarser FPN level arrives as TMA patches in shared memory (exact 2x upsample)
  int up_PW, up_PH;     // patch extent in source pixels: TW/2 + 2, TH/2 + 2
  int staged;           // epilogue through a per-warp shared-memory transpose: coalesced residual / up_prev / fp32 traffic
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// One lane of a fully converged warp; used *inside* warp-uniform control flow so that the compiler keeps descriptors,
// addresses and barriers in uniform registers (a divergent `if (lane == 0)` region makes it wrap every tcgen05 / TMA
// instruction in an ELECT + R2UR.BROADCAST loop, ~25 extra instructions per MMA).
__device__ __forceinline__ bool elect_one() {
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | SBO>>4 @32 | version=1 @46 | layout @61
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, int bk) {
  const uint64_t layout = bk == 128 ? 2 : (bk == 64 ? 4 : 6);       // SWIZZLE_128B / 64B / 32B
  const uint64_t sbo = (uint64_t)(8 * bk) >> 4;                       // 8 rows of one swizzle span
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// ------------------------------------------------------------------------------------------------ kernel
// Persistent, weight-stationary schedule.  CTA c owns output-channel tile n = c % tiles_n and walks the row tiles
// m = c / tiles_n, + ctas_per_n, ...  so that
//   * the tile's folded-BN scale / shift (and the digit-plane recombination factors) are staged in shared memory once;
//   * when all K chunks of the weight tile fit (b_resident), the weights are fetched once per CTA and only the spike
//     tile streams through the TMA pipeline -- 128*K bytes per tile instead of (128 + 192)*K;
//   * CTAs that run side by side work on the same row tile for different channel tiles, so the A tile is an L2 hit.
// Two accumulator buffers in TMEM (2 x 256 columns) let the epilogue of a tile overlap the main loop of the next.
constexpr int TC_ACC_COLS = 256;

struct TileOrigin { int img, ho0, wo0; };

__device__ __forceinline__ TileOrigin tile_origin(const TcParams& p, int tile_m) {
  TileOrigin o;
  o.img = 0; o.ho0 = 0; o.wo0 = 0;
  if (p.mode_conv) {
    const int per_img = p.tiles_w * p.tiles_h;
    o.img = tile_m / per_img;
    const int t = tile_m % per_img;
    o.ho0 = (t / p.tiles_w) * p.TH;
    o.wo0 = (t % p.tiles_w) * p.TW;
  } else if (p.w_img_rows) {
    o.img = (tile_m * TC_BM) / p.M_img;        // M_img % 128 == 0 is required for per-image weights
  }
  return o;
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// Stage the per-channel epilogue constants of one channel tile: ss[0][64] = scale, ss[1][64] = shift.
// (The digit planes are merged before the affine: y = (d0 * 128^(P-1) + merged low planes) * scale + shift.)
template <int EW>
__device__ __forceinline__ void stage_affine(const TcParams& p, float* ss, int co_base, int img, int et) {
  for (; et < 2 * TC_BN; et += 32 * EW) {
    const int which = et >> 6, ch = co_base + (et & (TC_BN - 1));
    float v = 0.f;
    if (ch < p.Cout) v = __ldg((which ? p.shift : p.scale) + (int64_t)img * p.ss_img_stride + ch);
    ss[et] = v;
  }
}

template <int PIECES, int NK>
__device__ __forceinline__ void mma_role(const TcParams& p, uint32_t ring_addr, uint32_t bres_addr, int stage_bytes,
                                         uint64_t* full, uint64_t* empty, uint64_t* tmem_full, uint64_t* tmem_empty,
                                         uint32_t tmem_base, int slot) {
  constexpr int nB = TC_BN * PIECES;
  constexpr int BKB = 32 * NK;                                   // bytes of K per chunk
  constexpr uint32_t A_STEP = (TC_BM * BKB) >> 4, B_STEP = (nB * BKB) >> 4;      // descriptor units (16 B)
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=S32, A=B=INT8, K-major, N>>3 @17, M>>4 @24
  constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nB >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  const uint64_t da_ring = smem_desc(ring_addr, BKB);
  const uint64_t db_res = smem_desc(bres_addr, BKB);
  const uint32_t stage_step = (uint32_t)stage_bytes >> 4;
  const int G = p.group, num_groups = (p.taps * p.cin_chunks) / G, stages = p.stages;
  const bool resident = p.b_resident != 0;
  int stage = 0, phase = 0, it = 0;
  for (int tile_m = slot; tile_m < p.tiles_m; tile_m += p.ctas_per_n, ++it) {
    const int acc = it & 1;
    mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_ACC_COLS);
    uint32_t c_units = 0;                                    // resident-B offset of the current chunk, descriptor units
    for (int g = 0; g < num_groups; ++g) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da0 = da_ring + (uint64_t)((uint32_t)stage * stage_step);
        const uint64_t db0 = resident ? db_res + (uint64_t)c_units : da0 + (uint64_t)((uint32_t)G * A_STEP);
        for (int gi = 0; gi < G; ++gi) {
          const uint64_t da = da0 + (uint64_t)((uint32_t)gi * A_STEP), db = db0 + (uint64_t)((uint32_t)gi * B_STEP);
#pragma unroll
          for (int k = 0; k < NK; ++k)
            umma_i8(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (g | gi | k) ? 1u : 0u);   // +32 B per K step
        }
        umma_commit(&empty[stage]);
        if (g == num_groups - 1) umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
      c_units += (uint32_t)G * B_STEP;
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  }
}



// ------------------------------------------------------------------------------------------------ epilogue variants
// The kernel is compiled once per epilogue kind: run-time flags in the (latency-bound, 2 warps per scheduler)
// epilogue cost branches, instruction-cache misses and dead registers -- 478 executed instructions per warp and tile
// in the flag-driven version against ~300 here.
constexpr int EPI_GENERIC = 0;     // every combination (transposed / ragged / per-image affine ...): flag driven
constexpr int EPI_SPIKE = 1;       // int8 levels only, Cout % 16 == 0, d_max = 8: the bulk of the network's launches
constexpr int EPI_STAGED = 2;      // fp32 / residual outputs through the shared-memory transpose
constexpr int EPI_STAGED_UP = 3;   // the same + fused FPN merge from TMA patches
#ifndef S2F_EW_SPIKE
#define S2F_EW_SPIKE 8
#endif
#ifndef S2F_EW_STAGED
#define S2F_EW_STAGED 8
#endif
#ifndef S2F_EW_UP
#define S2F_EW_UP 8
#endif
#ifndef S2F_LDTM_X32
#define S2F_LDTM_X32 1
#endif
constexpr int EW_SPIKE = S2F_EW_SPIKE, EW_STAGED = S2F_EW_STAGED, EW_UP = S2F_EW_UP;      // epilogue warps per kind

// exact plane merge: d0 * 128^(P-1) + (low planes merged in int32), one rounding
template <int PIECES>
__device__ __forceinline__ float merge_planes(uint32_t d0, uint32_t d1, uint32_t d2) {
  if (PIECES == 3) return fmaf((float)(int)d0, 16384.f, (float)((int)d1 * 128 + (int)d2));
  if (PIECES == 2) return (float)((int)d0 * 128 + (int)d1);
  return (float)(int)d0;
}

// Spike-only epilogue: lane = accumulator row, warp = 32 channels.  All six TMEM loads of the warp's slice are issued
// back to back and the accumulator is handed back to the MMA warp as soon as they have landed in registers.
// ss: [64] scale/8, [64] shift/8 (the power-of-two scaling is exact), so that the level is
//   rne(8 * sat(v * scale/8 + shift/8)):  FFMA.SAT + FFMA (2^23 trick) per output, then PRMT packing.
template <int PIECES, int EW>
__device__ __forceinline__ void epilogue_spike(const TcParams& p, float* ss, uint64_t* tmem_full, uint64_t* tmem_empty,
                                               uint32_t tmem_base, int slot, int tile_n, int warp, int lane) {
  constexpr int COLS = TC_BN / (EW / 4);                   // channels per warp: 32 or 16
  constexpr int NCH = COLS / 16;                           // 16-column TMEM loads per plane
  const int quad = warp & 3, slice = (warp - 2) >> 2;
  const int r = quad * 32 + lane;
  const int et = threadIdx.x - 64;
  const int co_base = tile_n * TC_BN;
  for (int e = et; e < 2 * TC_BN; e += 32 * EW) {
    const int ch = co_base + (e & (TC_BN - 1));
    ss[e] = ch < p.Cout ? 0.125f * __ldg(((e >> 6) ? p.shift : p.scale) + ch) : 0.f;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
  const int cw0 = co_base + slice * COLS;                  // this warp's first channel
  int nchunks = 0;                                         // Cout % 16 == 0 on this path
#pragma unroll
  for (int jj = 0; jj < NCH; ++jj) nchunks += (cw0 + 16 * jj < p.Cout) ? 1 : 0;
  const uint32_t ss_addr = smem_u32(ss) + (uint32_t)(slice * COLS) * 4u;
  int it = 0;
  for (int tile_m = slot; tile_m < p.tiles_m; tile_m += p.ctas_per_n, ++it) {
    const int acc = it & 1;
    int64_t m;
    if (p.mode_conv) {
      const TileOrigin o = tile_origin(p, tile_m);
      const int ho = o.ho0 + r / p.TW, wo = o.wo0 + r % p.TW;
      m = (ho < p.Ho && wo < p.Wo) ? ((int64_t)o.img * p.Ho + ho) * p.Wo + wo : -1;
    } else {
      m = (int64_t)tile_m * TC_BM + r;
      if (m >= p.M_total) m = -1;
    }
    int8_t* dst = p.out_spike + m * p.Cout + cw0;
    mbar_wait(&tmem_full[acc], (it >> 1) & 1);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * TC_ACC_COLS) + (uint32_t)(slice * COLS);
    uint32_t d[3][COLS];                                      // [plane][channel of this warp's slice]
    if (nchunks > 0) {
      if (NCH == 2 && S2F_LDTM_X32) {                         // one 32-column load per plane
        tmem_ld32(trow, d[0]);
        if (PIECES > 1) tmem_ld32(trow + TC_BN, d[1]);
        if (PIECES > 2) tmem_ld32(trow + 2 * TC_BN, d[2]);
      } else {
#pragma unroll
        for (int jj = 0; jj < NCH; ++jj) {
          tmem_ld16(trow + jj * 16, d[0] + 16 * jj);
          if (PIECES > 1) tmem_ld16(trow + TC_BN + jj * 16, d[1] + 16 * jj);
          if (PIECES > 2) tmem_ld16(trow + 2 * TC_BN + jj * 16, d[2] + 16 * jj);
        }
      }
      tmem_ld_wait();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tmem_empty[acc]);            // accumulator drained: the next tile's MMAs may start
    if (m < 0) continue;
#pragma unroll
    for (int jj = 0; jj < NCH; ++jj) {
      if (jj < nchunks) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 sc = lds128(ss_addr + (uint32_t)(jj * 16 + 4 * q) * 4u);
          const float4 sh = lds128(ss_addr + (uint32_t)(TC_BN + jj * 16 + 4 * q) * 4u);
          const float s4[4] = {sc.x, sc.y, sc.z, sc.w}, h4[4] = {sh.x, sh.y, sh.z, sh.w};
          uint32_t b[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;
            const float v = merge_planes<PIECES>(d[0][16 * jj + j], d[1][16 * jj + j], d[2][16 * jj + j]);
            b[e] = __float_as_uint(fmaf(__saturatef(fmaf(v, s4[e], h4[e])), 8.f, 8388608.f));
          }
          w[q] = __byte_perm(__byte_perm(b[0], b[1], 0x0040), __byte_perm(b[2], b[3], 0x0040), 0x5410);
        }
        *reinterpret_cast<uint4*>(dst + jj * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ staged epilogue
// tcgen05.ld hands every lane one accumulator ROW, so an epilogue that goes to global memory straight from those
// registers touches 32 different rows per instruction (32 L1 wavefronts for 512 bytes).  Layers that read a residual /
// the coarser FPN level or write fp32 are bound by exactly that, so they take this path instead: the warp parks its
// 32 rows x 32 channels of merged accumulators in a private shared-memory tile (row stride 36 floats: conflict-free for
// both the row-per-lane STS.128 and the 4-rows-x-128-byte LDS.128), releases the TMEM accumulator, and then walks the
// tile with lane = (row % 4, channel quad): every global access is 4 rows x 128 contiguous bytes.
__host__ __device__ constexpr int tc_stg_ld(int ew) { return TC_BN / (ew / 4) + 4; }          // floats per staged row: 36 / 20
__host__ __device__ constexpr int tc_stg_floats(int ew) { return 32 * tc_stg_ld(ew); }        // per epilogue warp
constexpr int TC_UP_SLOT = 60 * TC_BN * 4;                  // one patch: <= 60 source pixels x 64 channels fp32 (TW x TH = 16x8 or 8x16)
__host__ __device__ constexpr int tc_aux_off(int ew) { return 512 + 2 * 4 * TC_BN * 4 + ew * tc_stg_floats(ew) * 4; }   // barriers | affine | staging -> patches

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int PIECES, bool UP_TMA, int EW>
__device__ __forceinline__ void epilogue_staged(const TcParams& p, float* stg, uint64_t* tmem_full, uint64_t* tmem_empty,
                                                uint32_t tmem_base, int slot, int tile_n, int warp, int lane,
                                                const uint8_t* up_patch, uint64_t* up_full, uint64_t* up_empty) {
  constexpr int COLS = TC_BN / (EW / 4);                      // channels per warp: 32 or 16
  constexpr int NCH = COLS / 16;
  constexpr int LD = tc_stg_ld(EW);
  constexpr int LPR = COLS / 4;                               // lanes per row in phase 2: 8 or 4
  constexpr int RPS = 32 / LPR;                               // rows per phase-2 step: 4 or 8
  constexpr int NSTEP = 32 / RPS;                             // 8 or 4
  const int quad = warp & 3, slice = (warp - 2) >> 2;
  const int r = quad * 32 + lane;
  const int cq = lane % LPR, rsub = lane / LPR;
  const int cw = tile_n * TC_BN + slice * COLS + 4 * cq;      // this lane's four channels in phase 2
  const bool cvalid = cw < p.Cout;                            // Cout % 4 == 0 on this path
  const bool warp_has_cols = tile_n * TC_BN + slice * COLS < p.Cout;
  const uint32_t stg_addr = smem_u32(stg);
  float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
  int loaded_img = -1;
  int it = 0;
  for (int tile_m = slot; tile_m < p.tiles_m; tile_m += p.ctas_per_n, ++it) {
    const TileOrigin o = tile_origin(p, tile_m);
    const int acc = it & 1;
    const int aff_img = p.ss_img_stride ? o.img : 0;
    if (aff_img != loaded_img) {                              // warp-uniform; once per CTA unless the weights are per image
      if (cvalid) {
        sc = __ldg(reinterpret_cast<const float4*>(p.scale + (int64_t)aff_img * p.ss_img_stride + cw));
        sh = __ldg(reinterpret_cast<const float4*>(p.shift + (int64_t)aff_img * p.ss_img_stride + cw));
      }
      loaded_img = aff_img;
    }
    int m;                // flat output row of this lane's TMEM row (n*Ho*Wo index < 2^31), -1 if outside
    if (p.mode_conv) {
      const int ho = o.ho0 + r / p.TW, wo = o.wo0 + r % p.TW;
      m = (ho < p.Ho && wo < p.Wo) ? (o.img * p.Ho + ho) * p.Wo + wo : -1;
    } else {
      m = tile_m * TC_BM + r;
      if (m >= p.M_total) m = -1;
    }
    // fused FPN merge: the four source pixels (indices into up_prev's [n*up_H*up_W] pixel axis) and the two fractions
    int u00 = 0, u01 = 0, u10 = 0, u11 = 0;
    float up_lx = 0.f, up_ly = 0.f;
    if (p.up_prev && m >= 0) {
      const int im = m / p.M_img, rr = m % p.M_img;
      const int yo = rr / p.Wo, xo = rr % p.Wo;
      float sy = ((float)p.up_H / (float)p.Ho) * ((float)yo + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
      float sx = ((float)p.up_W / (float)p.Wo) * ((float)xo + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
      const int y0 = (int)sy, x0 = (int)sx;
      const int y1 = y0 + (y0 < p.up_H - 1 ? 1 : 0), x1 = x0 + (x0 < p.up_W - 1 ? 1 : 0);
      up_ly = sy - (float)y0; up_lx = sx - (float)x0;
      if (UP_TMA) {          // pixel indices inside the tile's patch (origin: ho0/2 - 1, wo0/2 - 1), one byte each
        const int py = o.ho0 / 2 - 1, px = o.wo0 / 2 - 1;
        const int l0 = (y0 - py) * p.up_PW - px, l1 = (y1 - py) * p.up_PW - px;
        u00 = (l0 + x0) | ((l0 + x1) << 8) | ((l1 + x0) << 16) | ((l1 + x1) << 24);
      } else {
        const int pb = im * p.up_H * p.up_W;
        u00 = pb + y0 * p.up_W + x0; u01 = pb + y0 * p.up_W + x1; u10 = pb + y1 * p.up_W + x0; u11 = pb + y1 * p.up_W + x1;
      }
    }
    // the residual rows of phase 2 do not depend on the accumulator: issue their (coalesced) loads now, so that the
    // DRAM / L2 latency is covered by the wait for the MMAs and by phase 1
    // element offsets of this lane's eight phase-2 rows relative to the tile's first row (32-bit; negative = no row):
    // the 64-bit part of every address is a per-tile, warp-uniform base
    const int64_t mt0 = p.mode_conv ? ((int64_t)o.img * p.Ho + o.ho0) * p.Wo + o.wo0 : (int64_t)tile_m * TC_BM;
    const int rel_c = m >= 0 ? (m - (int)mt0) * p.Cout : INT_MIN / 2;
    int orow[NSTEP];
#pragma unroll
    for (int i = 0; i < NSTEP; ++i) {
      const int t = __shfl_sync(0xffffffffu, rel_c, RPS * i + rsub);
      orow[i] = cvalid ? t + cw : -1;                           // cvalid: this lane's channel quad exists
    }
    const float* res_t = p.residual ? p.residual + mt0 * p.Cout : nullptr;
    float* of_t = p.out_f32 ? p.out_f32 + mt0 * p.Cout : nullptr;
    int8_t* os_t = p.out_spike ? p.out_spike + mt0 * p.Cout : nullptr;
    float4 res[NSTEP];
    if (p.residual) {
#pragma unroll
      for (int i = 0; i < NSTEP; ++i) {
        res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (orow[i] >= 0) res[i] = __ldg(reinterpret_cast<const float4*>(res_t + orow[i]));
      }
    }
    mbar_wait(&tmem_full[acc], (it >> 1) & 1);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * TC_ACC_COLS);
    // ---- phase 1: accumulators -> merged fp32 -> this lane's row of the staging tile (16 channels at a time: the
    // residual rows already occupy 32 registers)
    if (warp_has_cols) {
#pragma unroll
      for (int jj = 0; jj < NCH; ++jj) {
        const int j0 = slice * COLS + jj * 16;
        uint32_t d0[16], d1[16], d2[16];
        tmem_ld16(trow + j0, d0);
        if (PIECES > 1) tmem_ld16(trow + TC_BN + j0, d1);
        if (PIECES > 2) tmem_ld16(trow + 2 * TC_BN + j0, d2);
        tmem_ld_wait();
        const uint32_t wa = stg_addr + (uint32_t)(lane * LD + jj * 16) * 4u;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128(wa + 16u * q, merge_planes<PIECES>(d0[4 * q], d1[4 * q], d2[4 * q]),
                 merge_planes<PIECES>(d0[4 * q + 1], d1[4 * q + 1], d2[4 * q + 1]),
                 merge_planes<PIECES>(d0[4 * q + 2], d1[4 * q + 2], d2[4 * q + 2]),
                 merge_planes<PIECES>(d0[4 * q + 3], d1[4 * q + 3], d2[4 * q + 3]));
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tmem_empty[acc]);              // the MMA warp may refill this accumulator now
    // ---- phase 2: lane = (row % 4, channel quad); 8 steps of 4 rows x 128 bytes
    if (warp_has_cols) {
      const uint32_t ra = stg_addr + (uint32_t)(rsub * LD + 4 * cq) * 4u;
      if (!UP_TMA && !p.up_prev) {
#pragma unroll
        for (int i = 0; i < NSTEP; ++i) {
          const float4 v = lds128(ra + (uint32_t)(RPS * i * LD) * 4u);
          if (orow[i] >= 0) {
            float y0 = fmaf(v.x, sc.x, sh.x), y1 = fmaf(v.y, sc.y, sh.y), y2 = fmaf(v.z, sc.z, sh.z), y3 = fmaf(v.w, sc.w, sh.w);
            if (p.residual) { y0 += res[i].x; y1 += res[i].y; y2 += res[i].z; y3 += res[i].w; }
            if (of_t) *reinterpret_cast<float4*>(of_t + orow[i]) = make_float4(y0, y1, y2, y3);
            if (os_t) *reinterpret_cast<uint32_t*>(os_t + orow[i]) = pack_levels4_d8(y0, y1, y2, y3);
          }
        }
      } else if (UP_TMA) {
        const int us = it % 3;
        mbar_wait(&up_full[us], (uint32_t)((it / 3) & 1));
        const uint32_t pa = smem_u32(up_patch) + (uint32_t)us * (uint32_t)TC_UP_SLOT + (uint32_t)(slice * COLS + 4 * cq) * 4u;
#pragma unroll
        for (int i = 0; i < NSTEP; ++i) {
          const int src = RPS * i + rsub;
          const uint32_t pk = (uint32_t)__shfl_sync(0xffffffffu, u00, src);
          const float lx = __shfl_sync(0xffffffffu, up_lx, src), ly = __shfl_sync(0xffffffffu, up_ly, src);
          const float4 v = lds128(ra + (uint32_t)(RPS * i * LD) * 4u);
          if (orow[i] >= 0) {
            const float hx = 1.f - lx, hy = 1.f - ly;
            const float4 p00 = lds128(pa + (pk & 0xffu) * (TC_BN * 4u)), p01 = lds128(pa + ((pk >> 8) & 0xffu) * (TC_BN * 4u));
            const float4 p10 = lds128(pa + ((pk >> 16) & 0xffu) * (TC_BN * 4u)), p11 = lds128(pa + (pk >> 24) * (TC_BN * 4u));
            float y0 = fmaf(v.x, sc.x, sh.x), y1 = fmaf(v.y, sc.y, sh.y), y2 = fmaf(v.z, sc.z, sh.z), y3 = fmaf(v.w, sc.w, sh.w);
            if (p.residual) { y0 += res[i].x; y1 += res[i].y; y2 += res[i].z; y3 += res[i].w; }
            y0 = y0 + (hy * (hx * p00.x + lx * p01.x) + ly * (hx * p10.x + lx * p11.x));
            y1 = y1 + (hy * (hx * p00.y + lx * p01.y) + ly * (hx * p10.y + lx * p11.y));
            y2 = y2 + (hy * (hx * p00.z + lx * p01.z) + ly * (hx * p10.z + lx * p11.z));
            y3 = y3 + (hy * (hx * p00.w + lx * p01.w) + ly * (hx * p10.w + lx * p11.w));
            if (of_t) *reinterpret_cast<float4*>(of_t + orow[i]) = make_float4(y0, y1, y2, y3);
            if (os_t) *reinterpret_cast<uint32_t*>(os_t + orow[i]) = pack_levels4_d8(y0, y1, y2, y3);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < NSTEP; ++i) {
          const int src = RPS * i + rsub;
          const int s00 = __shfl_sync(0xffffffffu, u00, src), s01 = __shfl_sync(0xffffffffu, u01, src);
          const int s10 = __shfl_sync(0xffffffffu, u10, src), s11 = __shfl_sync(0xffffffffu, u11, src);
          const float lx = __shfl_sync(0xffffffffu, up_lx, src), ly = __shfl_sync(0xffffffffu, up_ly, src);
          const float4 v = lds128(ra + (uint32_t)(RPS * i * LD) * 4u);
          if (orow[i] >= 0) {
            const float hx = 1.f - lx, hy = 1.f - ly;
            const float4 p00 = __ldg(reinterpret_cast<const float4*>(p.up_prev + (int64_t)s00 * p.Cout + cw));
            const float4 p01 = __ldg(reinterpret_cast<const float4*>(p.up_prev + (int64_t)s01 * p.Cout + cw));
            const float4 p10 = __ldg(reinterpret_cast<const float4*>(p.up_prev + (int64_t)s10 * p.Cout + cw));
            const float4 p11 = __ldg(reinterpret_cast<const float4*>(p.up_prev + (int64_t)s11 * p.Cout + cw));
            float y0 = fmaf(v.x, sc.x, sh.x), y1 = fmaf(v.y, sc.y, sh.y), y2 = fmaf(v.z, sc.z, sh.z), y3 = fmaf(v.w, sc.w, sh.w);
            if (p.residual) { y0 += res[i].x; y1 += res[i].y; y2 += res[i].z; y3 += res[i].w; }
            // ATen upsample_bilinear2d: hy*(hx*p00 + lx*p01) + ly*(hx*p10 + lx*p11), then cur + up
            y0 = y0 + (hy * (hx * p00.x + lx * p01.x) + ly * (hx * p10.x + lx * p11.x));
            y1 = y1 + (hy * (hx * p00.y + lx * p01.y) + ly * (hx * p10.y + lx * p11.y));
            y2 = y2 + (hy * (hx * p00.z + lx * p01.z) + ly * (hx * p10.z + lx * p11.z));
            y3 = y3 + (hy * (hx * p00.w + lx * p01.w) + ly * (hx * p10.w + lx * p11.w));
            if (of_t) *reinterpret_cast<float4*>(of_t + orow[i]) = make_float4(y0, y1, y2, y3);
            if (os_t) *reinterpret_cast<uint32_t*>(os_t + orow[i]) = pack_levels4_d8(y0, y1, y2, y3);
          }
        }
      }
    }
    __syncwarp();                                              // the tile is free for the next phase 1
    if (UP_TMA && lane == 0) mbar_arrive(&up_empty[it % 3]);
  }
}

template <int PIECES, int EPI, int EW>
__global__ void __launch_bounds__(tc_threads(EW), 1)
gemm_i8_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_up, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int nB = TC_BN * PIECES;                    // MMA N
  const int a_bytes = TC_BM * p.bk, b_bytes = nB * p.bk;
  const int num_chunks = p.taps * p.cin_chunks;
  const int stage_bytes = p.group * (p.b_resident ? a_bytes : a_bytes + b_bytes);      // multiples of 1024
  const int num_groups = num_chunks / p.group;
  uint8_t* b_res = smem;                                                   // [num_chunks][nB][bk] when resident
  uint8_t* ring = smem + (p.b_resident ? (size_t)num_chunks * b_bytes : 0);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;             // [2]
  uint64_t* tmem_empty = tmem_full + 2;               // [2]
  uint64_t* b_full = tmem_empty + 2;                  // [1]
  uint64_t* up_full = b_full + 1;                     // [3] patches of the coarser FPN level (p.up_tma)
  uint64_t* up_empty = up_full + 3;                   // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(up_empty + 3);
  uint8_t* up_patch = reinterpret_cast<uint8_t*>(full) + tc_aux_off(EW);      // [3][TC_UP_SLOT], 512-byte aligned
  float* ss_stage = reinterpret_cast<float*>(tmem_slot + 2);      // [2][4][64] floats, 16-byte aligned
  float* stg_all = ss_stage + 2 * 4 * TC_BN;                      // [EW warps][32][COLS + 4] floats when p.staged

  // warp index through a shuffle: the compiler then knows it is warp-uniform (role branches on the uniform datapath,
  // descriptors and barrier addresses in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x % p.tiles_n;
  const int slot = blockIdx.x / p.tiles_n;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EW); }
    mbar_init(b_full, 1);
    for (int s = 0; s < 3; ++s) { mbar_init(&up_full[s], 1); mbar_init(&up_empty[s], EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * TC_ACC_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (whole warp runs the loop; one elected lane issues)
    if (p.b_resident) {
      if (elect_one()) {
        mbar_expect_tx(b_full, (uint32_t)(num_chunks * b_bytes));
        for (int c = 0; c < num_chunks; ++c) tma_load_2d(b_res + (size_t)c * b_bytes, &map_b, b_full, c * p.bk, tile_n * nB);
      }
      __syncwarp();
    }
    int stage = 0, phase = 0, upi = 0;
    for (int tile_m = slot; tile_m < p.tiles_m; tile_m += p.ctas_per_n, ++upi) {
      const TileOrigin o = tile_origin(p, tile_m);
      if (EPI == EPI_STAGED_UP) {
        const int us = upi % 3;
        mbar_wait(&up_empty[us], (uint32_t)(((upi / 3) & 1) ^ 1));
        if (elect_one()) {
          mbar_expect_tx(&up_full[us], (uint32_t)(p.up_PW * p.up_PH * TC_BN * 4));
          tma_load_4d(up_patch + (size_t)us * TC_UP_SLOT, &map_up, &up_full[us], tile_n * TC_BN, o.wo0 / 2 - 1, o.ho0 / 2 - 1, o.img);
        }
        __syncwarp();
      }
      const int w_row0 = (p.w_img_rows ? o.img * p.w_img_rows : 0) + tile_n * nB;
      const int x0 = o.wo0 * p.stride - p.pad, y0 = o.ho0 * p.stride - p.pad, row0 = tile_m * TC_BM;
      int c = 0, cc = 0, kh = 0, kw = 0;                   // running chunk index / channel chunk / tap coordinates
      for (int g = 0; g < num_groups; ++g) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = ring + (size_t)stage * stage_bytes;
        if (elect_one()) {
          mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
          for (int gi = 0; gi < p.group; ++gi) {
            uint8_t* sa = st + gi * a_bytes;
            if (p.mode_conv) tma_load_4d(sa, &map_a, &full[stage], cc * p.bk, x0 + kw, y0 + kh, o.img);
            else tma_load_2d(sa, &map_a, &full[stage], cc * p.bk, row0);
            if (!p.b_resident) tma_load_2d(st + p.group * a_bytes + gi * b_bytes, &map_b, &full[stage], c * p.bk, w_row0);
            ++c;
            if (++cc == p.cin_chunks) { cc = 0; if (++kw == p.taps_w) { kw = 0; ++kh; } }
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp runs the loop; one elected lane issues).  Everything per chunk is a handful of
    // uniform-datapath instructions: descriptors advance by constant strides (address >> 4 fits the 14-bit field).
    if (p.b_resident) { mbar_wait(b_full, 0); tc_fence_after(); }
    if (p.bk == 128) mma_role<PIECES, 4>(p, smem_u32(ring), smem_u32(b_res), stage_bytes, full, empty, tmem_full, tmem_empty, tmem_base, slot);
    else if (p.bk == 64) mma_role<PIECES, 2>(p, smem_u32(ring), smem_u32(b_res), stage_bytes, full, empty, tmem_full, tmem_empty, tmem_base, slot);
    else mma_role<PIECES, 1>(p, smem_u32(ring), smem_u32(b_res), stage_bytes, full, empty, tmem_full, tmem_empty, tmem_base, slot);
  } else if (EPI == EPI_SPIKE) {
    epilogue_spike<PIECES, EW>(p, ss_stage, tmem_full, tmem_empty, tmem_base, slot, tile_n, warp, lane);
  } else if (EPI == EPI_STAGED || EPI == EPI_STAGED_UP) {
    epilogue_staged<PIECES, EPI == EPI_STAGED_UP, EW>(p, stg_all + (warp - 2) * tc_stg_floats(EW), tmem_full, tmem_empty, tmem_base, slot,
                                                  tile_n, warp, lane, up_patch, up_full, up_empty);
  } else {
    // ===== epilogue (8 warps): warp w reads TMEM lanes [32*(w%4), +32) -- thread = one output row -- and one half
    // (32 channels) of the tile's columns.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                   // column slice: 64 / (EW / 4) channels each
    constexpr int COLS_PER_WARP = TC_BN / (EW / 4);
    const int r = quad * 32 + lane;
    const int et = threadIdx.x - 64;                    // 0..255 within the epilogue group
    const int co_base = tile_n * TC_BN;
    const bool per_tile_affine = p.ss_img_stride != 0;
    if (!per_tile_affine) {
      stage_affine<EW>(p, ss_stage, co_base, 0, et);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");     // epilogue warps only
    }
    int it = 0, staged_img = -1, staged_buf = 0;
    for (int tile_m = slot; tile_m < p.tiles_m; tile_m += p.ctas_per_n, ++it) {
      const TileOrigin o = tile_origin(p, tile_m);
      const int acc = it & 1;
      if (per_tile_affine && o.img != staged_img) {      // per-image weights: constants change with the image only
        staged_buf ^= 1;                                 // the other buffer may still be read by a slower warp
        stage_affine<EW>(p, ss_stage + staged_buf * (4 * TC_BN), co_base, o.img, et);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
        staged_img = o.img;
      }
      const uint32_t ss_addr = smem_u32(ss_stage + staged_buf * (4 * TC_BN));
      int64_t m;            // flat output row (n*Ho*Wo index), -1 if outside
      if (p.mode_conv) {
        const int ho = o.ho0 + r / p.TW, wo = o.wo0 + r % p.TW;
        m = (ho < p.Ho && wo < p.Wo) ? ((int64_t)o.img * p.Ho + ho) * p.Wo + wo : -1;
      } else {
        m = (int64_t)tile_m * TC_BM + r;
        if (m >= p.M_total) m = -1;
      }
      // fused FPN merge: source coordinates of this row in the coarser map (upsample_bilinear2d, align_corners=False)
      if (p.residual && m >= 0) {                      // this warp's 128-byte residual slice -> L1 before the accumulator wait
        const int cpre = co_base + half * COLS_PER_WARP;
        if (cpre < p.Cout) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.residual + m * p.Cout + cpre));
      }
      const float* up00 = nullptr; const float* up01 = nullptr; const float* up10 = nullptr; const float* up11 = nullptr;
      float up_hx = 0.f, up_lx = 0.f, up_hy = 0.f, up_ly = 0.f;
      if (p.up_prev && m >= 0) {
        const int im = (int)(m / p.M_img), rr = (int)(m % p.M_img);
        const int yo = rr / p.Wo, xo = rr % p.Wo;
        float sy = ((float)p.up_H / (float)p.Ho) * ((float)yo + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
        float sx = ((float)p.up_W / (float)p.Wo) * ((float)xo + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < p.up_H - 1 ? 1 : 0), x1 = x0 + (x0 < p.up_W - 1 ? 1 : 0);
        up_ly = sy - (float)y0; up_lx = sx - (float)x0; up_hy = 1.f - up_ly; up_hx = 1.f - up_lx;
        const float* pb = p.up_prev + (int64_t)im * p.up_H * p.up_W * p.Cout;
        up00 = pb + ((int64_t)y0 * p.up_W + x0) * p.Cout; up01 = pb + ((int64_t)y0 * p.up_W + x1) * p.Cout;
        up10 = pb + ((int64_t)y1 * p.up_W + x0) * p.Cout; up11 = pb + ((int64_t)y1 * p.up_W + x1) * p.Cout;
        // pull this warp's 128-byte slice of the four corners into L1 now: the loads in the chunk loop below then hit
        // L1 instead of exposing an L2 round trip per chunk on the (latency-bound) epilogue path
        const int cpre = co_base + half * COLS_PER_WARP;
        if (cpre < p.Cout) {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(up00 + cpre));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(up01 + cpre));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(up10 + cpre));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(up11 + cpre));
        }
      }
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * TC_ACC_COLS);
#pragma unroll 1
      for (int jj = 0; jj < COLS_PER_WARP / 16; ++jj) {
        const int j0 = half * COLS_PER_WARP + jj * 16;
        if (co_base + j0 >= p.Cout) break;               // warp-uniform
        uint32_t d0[16], d1[16], d2[16];
        tmem_ld16(trow + j0, d0);
        if (PIECES > 1) tmem_ld16(trow + TC_BN + j0, d1);
        if (PIECES > 2) tmem_ld16(trow + 2 * TC_BN + j0, d2);
        tmem_ld_wait();
        if (m < 0) continue;
        // v = d0 * 128^(P-1) + (low planes merged exactly in int32: |d1 * 128 + d2| < 2^31 for K < 32768), one rounding;
        // y = v * scale + shift.  Per output: one IMAD, two I2FP, two FFMA.
        float y[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 sc = lds128(ss_addr + (uint32_t)(j0 + 4 * q) * 4u);
          const float4 sh = lds128(ss_addr + (uint32_t)(TC_BN + j0 + 4 * q) * 4u);
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;
            if (PIECES == 3) v[e] = fmaf((float)(int)d0[j], 16384.f, (float)((int)d1[j] * 128 + (int)d2[j]));
            else if (PIECES == 2) v[e] = (float)((int)d0[j] * 128 + (int)d1[j]);
            else v[e] = (float)(int)d0[j];
          }
          y[4 * q] = fmaf(v[0], sc.x, sh.x); y[4 * q + 1] = fmaf(v[1], sc.y, sh.y);
          y[4 * q + 2] = fmaf(v[2], sc.z, sh.z); y[4 * q + 3] = fmaf(v[3], sc.w, sh.w);
        }
        const int co0 = co_base + j0;
        const int nvalid = min(16, p.Cout - co0);
        const bool full16 = nvalid == 16 && (p.Cout & 3) == 0;
        const int64_t row_off = m * p.Cout + co0;
        if (p.residual) {
          if (full16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 rv = *reinterpret_cast<const float4*>(p.residual + row_off + 4 * q);
              y[4 * q] += rv.x; y[4 * q + 1] += rv.y; y[4 * q + 2] += rv.z; y[4 * q + 3] += rv.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nvalid) y[j] += p.residual[row_off + j];
          }
        }
        if (p.up_prev) {                                   // host guarantees Cout % 16 == 0 here
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 p00 = __ldg(reinterpret_cast<const float4*>(up00 + co0) + q), p01 = __ldg(reinterpret_cast<const float4*>(up01 + co0) + q);
            const float4 p10 = __ldg(reinterpret_cast<const float4*>(up10 + co0) + q), p11 = __ldg(reinterpret_cast<const float4*>(up11 + co0) + q);
            // ATen: hy*(hx*p00 + lx*p01) + ly*(hx*p10 + lx*p11), then cur + up
            y[4 * q] = y[4 * q] + (up_hy * (up_hx * p00.x + up_lx * p01.x) + up_ly * (up_hx * p10.x + up_lx * p11.x));
            y[4 * q + 1] = y[4 * q + 1] + (up_hy * (up_hx * p00.y + up_lx * p01.y) + up_ly * (up_hx * p10.y + up_lx * p11.y));
            y[4 * q + 2] = y[4 * q + 2] + (up_hy * (up_hx * p00.z + up_lx * p01.z) + up_ly * (up_hx * p10.z + up_lx * p11.z));
            y[4 * q + 3] = y[4 * q + 3] + (up_hy * (up_hx * p00.w + up_lx * p01.w) + up_ly * (up_hx * p10.w + up_lx * p11.w));
          }
        }
        if (!p.out_transposed) {
          if (p.out_f32) {
            if (full16) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(p.out_f32 + row_off + 4 * q) = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < nvalid) p.out_f32[row_off + j] = y[j];
            }
          }
          if (p.out_spike) {
            if (nvalid == 16 && (p.Cout & 15) == 0) {
              uint32_t w[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) w[q] = pack_levels4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3], p.d_max);
              *reinterpret_cast<uint4*>(p.out_spike + row_off) = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < nvalid) p.out_spike[row_off + j] = (int8_t)(level_bits(y[j], p.d_max) & 0xffu);
            }
          }
        } else {
          const int64_t im = m / p.M_img, pm = m % p.M_img;
          const int64_t tbase = im * (int64_t)p.M_img * p.Cout + pm;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < nvalid) {
              const int64_t oo = tbase + (int64_t)(co0 + j) * p.M_img;
              if (p.out_f32) p.out_f32[oo] = y[j];
              if (p.out_spike) p.out_spike[oo] = (int8_t)(level_bits(y[j], p.d_max) & 0xffu);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TC_ACC_COLS));
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static CUtensorMapSwizzle swz_for(int bk) {
  return bk == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

}  // namespace s2f

using namespace s2f;

extern "C" int s2f_gemm_i8_tc(const s2f_gemm_tc_args* a, void* stream) {
  S2F_REQUIRE(a && a->a && a->w_packed && a->scale && a->shift, "gemm_i8_tc: a, w_packed, scale, shift are required");
  S2F_REQUIRE(a->out_f32 || a->out_spike, "gemm_i8_tc: no output requested");
  S2F_REQUIRE(a->pieces >= 1 && a->pieces <= 3, "gemm_i8_tc: pieces must be 1..3");
  S2F_REQUIRE(a->Cin >= 32 && a->Cin % 16 == 0, "gemm_i8_tc: Cin must be >= 32 and a multiple of 16");
  S2F_REQUIRE((a->KH == 1 && a->KW == 1) || (a->KH == 3 && a->KW == 3), "gemm_i8_tc: 1x1 or 3x3 only");
  S2F_REQUIRE(a->stride == 1 || a->stride == 2, "gemm_i8_tc: stride 1 or 2");
  S2F_REQUIRE((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w_packed) & 15) == 0,
              "gemm_i8_tc: operands must be 16-byte aligned");
  S2F_REQUIRE(a->a_ld == 0 || (a->KH == 1 && a->stride == 1 && a->pad == 0 && a->a_ld >= a->Cin && a->a_ld % 16 == 0 && !a->up_prev),
              "gemm_i8_tc: a_ld needs a plain 1x1 layer and a multiple of 16 >= Cin");
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(S2F_ERR_CUDA, "gemm_i8_tc: %s", "cuTensorMapEncodeTiled entry point not found");

  TcParams p{};
  const int Ho = (a->H + 2 * a->pad - a->KH) / a->stride + 1, Wo = (a->W + 2 * a->pad - a->KW) / a->stride + 1;
  S2F_REQUIRE(Ho > 0 && Wo > 0, "gemm_i8_tc: empty output");
  p.scale = a->scale; p.shift = a->shift; p.residual = a->residual; p.out_f32 = a->out_f32; p.out_spike = a->out_spike;
  p.up_prev = a->up_prev; p.up_H = a->up_H; p.up_W = a->up_W;
  if (a->up_prev) {
    S2F_REQUIRE(a->up_H > 0 && a->up_W > 0 && a->Cout % 16 == 0 && !a->out_transposed && (reinterpret_cast<uintptr_t>(a->up_prev) & 15) == 0,
                "gemm_i8_tc: fused upsample needs Cout % 16 == 0, 16-byte aligned up_prev and a non-transposed output");
  }
  p.Ho = Ho; p.Wo = Wo; p.Cout = a->Cout; p.M_img = Ho * Wo; p.M_total = a->n * Ho * Wo;
  p.taps_w = a->KW; p.taps = a->KH * a->KW; p.stride = a->stride; p.pad = a->pad;
  p.bk = tc_bk(a->Cin); p.cin_chunks = (a->Cin + p.bk - 1) / p.bk;
  p.pieces = a->pieces; p.out_transposed = a->out_transposed; p.d_max = a->d_max > 0.f ? a->d_max : 8.f;
  const int per_img_w = a->per_image_weights;     // 1: weights + affine per image, 2: weights per image, shared affine
  p.mode_conv = (a->KH == 1 && a->stride == 1 && a->pad == 0) ? 0 : 1;
  p.staged = (p.d_max == 8.f && !a->out_transposed && a->Cout % 4 == 0 && (a->out_f32 || a->residual || a->up_prev) &&
              (!a->residual || (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0) &&
              (!a->out_f32 || (reinterpret_cast<uintptr_t>(a->out_f32) & 15) == 0) &&
              (!a->out_spike || (reinterpret_cast<uintptr_t>(a->out_spike) & 3) == 0) &&
              (reinterpret_cast<uintptr_t>(a->scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->shift) & 15) == 0 &&
              (int64_t)p.M_total < (1ll << 31) && (!a->up_prev || (int64_t)a->n * a->up_H * a->up_W < (1ll << 31))) ? 1 : 0;
  // fused FPN merge of an exact 2x coarser level: spatial 16x8 tiles, so that the tile's source pixels form one small
  // box (TW/2 + 2 by TH/2 + 2) that the producer warp fetches with TMA next to the spike tile
  const bool up2x = p.staged && a->up_prev && !per_img_w && Ho == 2 * a->up_H && Wo == 2 * a->up_W && Wo >= 8 &&
                    a->Cout % 4 == 0 && (int64_t)a->n * Ho * Wo < (1ll << 31);
  if (up2x) p.mode_conv = 1;
  int epi = EPI_GENERIC;
  if (up2x) epi = EPI_STAGED_UP;
  else if (p.staged) epi = EPI_STAGED;
  else if (a->out_spike && !a->out_f32 && !a->residual && !a->up_prev && !a->out_transposed && per_img_w != 1 && a->Cout % 16 == 0 &&
           p.d_max == 8.f && (reinterpret_cast<uintptr_t>(a->out_spike) & 15) == 0)
    epi = EPI_SPIKE;
  const int ew = epi == EPI_SPIKE ? EW_SPIKE : (epi == EPI_STAGED ? EW_STAGED : (epi == EPI_STAGED_UP ? EW_UP : 8));
  const int nB = TC_BN * p.pieces;
  const int tiles_n = (a->Cout + TC_BN - 1) / TC_BN;
  if (per_img_w) {
    S2F_REQUIRE(!p.mode_conv && p.M_img % TC_BM == 0, "gemm_i8_tc: per-image weights need a 1x1 layer with Ho*Wo % 128 == 0");
    p.w_img_rows = tiles_n * nB;
    p.ss_img_stride = per_img_w == 1 ? a->Cout : 0;
  }
  const int kpad = p.taps * tc_cin_pad(a->Cin);

  CUtensorMap map_a, map_b;
  int tiles_m;
  if (p.mode_conv) {
    p.TW = 16; while (p.TW / 2 >= Wo && p.TW > 1) p.TW /= 2;
    p.TH = TC_BM / p.TW;
    p.tiles_w = (Wo + p.TW - 1) / p.TW; p.tiles_h = (Ho + p.TH - 1) / p.TH;
    tiles_m = a->n * p.tiles_w * p.tiles_h;
    cuuint64_t dims[4] = {(cuuint64_t)a->Cin, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->n};
    cuuint64_t strides[3] = {(cuuint64_t)a->Cin, (cuuint64_t)a->W * a->Cin, (cuuint64_t)a->H * a->W * a->Cin};
    cuuint32_t box[4] = {(cuuint32_t)p.bk, (cuuint32_t)(p.TW * a->stride), (cuuint32_t)(p.TH * a->stride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)a->stride, (cuuint32_t)a->stride, 1};
    S2F_REQUIRE(box[1] <= 256 && box[2] <= 256, "gemm_i8_tc: box too large");
    CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<int8_t*>(a->a), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz_for(p.bk), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(S2F_ERR_CUDA, "gemm_i8_tc: cuTensorMapEncodeTiled(A,4D) failed (%s) code %lld", "", (long long)r);
  } else {
    tiles_m = (p.M_total + TC_BM - 1) / TC_BM;
    cuuint64_t dims[2] = {(cuuint64_t)a->Cin, (cuuint64_t)p.M_total};
    cuuint64_t strides[1] = {(cuuint64_t)(a->a_ld ? a->a_ld : a->Cin)};
    cuuint32_t box[2] = {(cuuint32_t)p.bk, (cuuint32_t)TC_BM};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(a->a), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz_for(p.bk), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(S2F_ERR_CUDA, "gemm_i8_tc: cuTensorMapEncodeTiled(A,2D) failed (%s) code %lld", "", (long long)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)tiles_n * nB * (per_img_w ? a->n : 1)};
    cuuint64_t strides[1] = {(cuuint64_t)kpad};
    cuuint32_t box[2] = {(cuuint32_t)p.bk, (cuuint32_t)nB};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(a->w_packed), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz_for(p.bk), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(S2F_ERR_CUDA, "gemm_i8_tc: cuTensorMapEncodeTiled(B) failed (%s) code %lld", "", (long long)r);
  }
  p.tiles_n = tiles_n;
  p.tiles_m = tiles_m;
  CUtensorMap map_up = map_b;
  if (up2x) {
    p.up_tma = 1; p.up_PW = p.TW / 2 + 2; p.up_PH = p.TH / 2 + 2;
    S2F_REQUIRE(p.up_PW * p.up_PH * TC_BN * 4 <= TC_UP_SLOT && p.up_PW * p.up_PH <= 255, "gemm_i8_tc: FPN patch too large");
    cuuint64_t dims[4] = {(cuuint64_t)a->Cout, (cuuint64_t)a->up_W, (cuuint64_t)a->up_H, (cuuint64_t)a->n};
    cuuint64_t strides[3] = {(cuuint64_t)a->Cout * 4, (cuuint64_t)a->up_W * a->Cout * 4, (cuuint64_t)a->up_H * a->up_W * a->Cout * 4};
    cuuint32_t box[4] = {(cuuint32_t)TC_BN, (cuuint32_t)p.up_PW, (cuuint32_t)p.up_PH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&map_up, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a->up_prev), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(S2F_ERR_CUDA, "gemm_i8_tc: cuTensorMapEncodeTiled(up_prev) failed (%s) code %lld", "", (long long)r);
  }

  S2F_REQUIRE((int64_t)kpad * 8 * 64 * 129 < (1ll << 31), "gemm_i8_tc: K too large for the int32 plane merge");
  const int num_chunks = p.taps * p.cin_chunks;
  const size_t fixed = 1024 /*alignment*/ + 512 /*barriers*/ + 2 * 4 * TC_BN * sizeof(float) +
                       (p.staged ? (size_t)ew * tc_stg_floats(ew) * sizeof(float) : 0) + (p.up_tma ? 3 * (size_t)TC_UP_SLOT : 0);
  const size_t budget = 226 * 1024 - fixed;
  const size_t b_all = (size_t)num_chunks * nB * p.bk;
  // weight-stationary when the whole K extent of the weight tile fits and still leaves >= 4 A stages
  p.b_resident = (!per_img_w && b_all + 6 * (size_t)TC_BM * p.bk <= budget) ? 1 : 0;
  p.group = (p.mode_conv && p.taps == 9 && p.cin_chunks == 1 && p.bk <= 64) ? 3 : 1;
  if (p.group == 1 && p.b_resident && num_chunks % 2 == 0) p.group = 2;      // 32 KB spike stages: half the barrier rounds
  const int stage_bytes = p.group * (p.b_resident ? TC_BM * p.bk : (TC_BM + nB) * p.bk);
  p.stages = (int)((budget - (p.b_resident ? b_all : 0)) / stage_bytes);
  if (p.stages > 12) p.stages = 12;
  if (p.stages < 2) p.stages = 2;
  size_t smem = (p.b_resident ? b_all : 0) + (size_t)p.stages * stage_bytes + fixed;
  if (smem < 120 * 1024) smem = 120 * 1024;          // never two CTAs on one SM: each allocates all 512 TMEM columns
  const int num_sms = sm_count();
  // CTAs per channel tile: fill the SMs, never more than there are row tiles
  int per_n = num_sms / tiles_n;
  if (per_n < 1) per_n = 1;
  if (per_n > tiles_m) per_n = tiles_m;
  p.ctas_per_n = per_n;
  const unsigned grid = (unsigned)(per_n * tiles_n);
  cudaError_t attr_err = cudaSuccess;
#define S2F_TC_LAUNCH(P, E, W)                                                                                            \
  do {                                                                                                                  \
    static std::atomic<uint64_t> attr_set{0};                                                                           \
    if (first_use_on_this_device(attr_set))                                                                             \
      attr_err = cudaFuncSetAttribute(gemm_i8_tc_kernel<P, E, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
    if (attr_err == cudaSuccess) gemm_i8_tc_kernel<P, E, W><<<grid, tc_threads(W), smem, (cudaStream_t)stream>>>(map_a, map_b, map_up, p); \
  } while (0)
#define S2F_TC_EPI(P)                                               \
  do {                                                              \
    if (epi == EPI_SPIKE) S2F_TC_LAUNCH(P, EPI_SPIKE, EW_SPIKE);    \
    else if (epi == EPI_STAGED) S2F_TC_LAUNCH(P, EPI_STAGED, EW_STAGED); \
    else if (epi == EPI_STAGED_UP) S2F_TC_LAUNCH(P, EPI_STAGED_UP, EW_UP); \
    else S2F_TC_LAUNCH(P, EPI_GENERIC, 8);                          \
  } while (0)
  if (p.pieces == 3) S2F_TC_EPI(3);
  else if (p.pieces == 2) S2F_TC_EPI(2);
  else S2F_TC_EPI(1);
#undef S2F_TC_EPI
#undef S2F_TC_LAUNCH
  if (attr_err != cudaSuccess) return fail(S2F_ERR_CUDA, "gemm_i8_tc: smem attribute: %s", cudaGetErrorString(attr_err));
  return check_launch("gemm_i8_tc_kernel");
}

// Device-side packer for weights that are produced on the GPU (one matrix per image): one block per row.
namespace s2f {
__global__ void __launch_bounds__(128) pack_rows_i8_kernel(const float* __restrict__ w, int ld, int rows_per_img, int K,
                                                           int kpad, int pieces, int8_t* __restrict__ packed,
                                                           float* __restrict__ scale_out, float* __restrict__ shift_out,
                                                           const float* __restrict__ bias_col, float post_scale) {
  const int row = blockIdx.x;                       // global row = img * rows_per_img + r
  const int img = row / rows_per_img, r = row % rows_per_img;
  const float* src = w + (int64_t)row * ld;
  __shared__ float red[128];
  float mx = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) mx = fmaxf(mx, fabsf(src[k]));
  red[threadIdx.x] = mx; __syncthreads();
  for (int s = 64; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
  mx = red[0];
  int e = 0;
  if (mx > 0.f) { frexpf(mx, &e); if (ldexpf(1.f, e - 1) == mx) e -= 1; }
  const int F = 7 * pieces - 1;
  const float inv = ldexpf(1.f, F - e);
  const int tiles_n = (rows_per_img + TC_BN - 1) / TC_BN;
  const int tile = r / TC_BN, rr = r % TC_BN;
  int8_t* base = packed + (int64_t)img * tiles_n * pieces * TC_BN * kpad;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    int I = __float2int_rn(src[k] * inv);
    int digit[3] = {0, 0, 0};
    for (int pz = pieces - 1; pz >= 1; --pz) {
      const int lo = (((I + 64) % 128 + 128) % 128) - 64;
      digit[pz] = lo;
      I = (I - lo) / 128;
    }
    digit[0] = I;
    for (int pz = 0; pz < pieces; ++pz) base[((int64_t)(tile * pieces + pz) * TC_BN + rr) * kpad + k] = (int8_t)digit[pz];
  }
  if (threadIdx.x == 0) {
    scale_out[row] = ldexpf(1.f, e - F) * post_scale;
    shift_out[row] = bias_col ? bias_col[(int64_t)row * ld] : 0.f;
  }
}
}  // namespace s2f

extern "C" int s2f_pack_rows_i8_device(const float* w, int ld, int n_img, int rows_per_img, int K, int pieces,
                                       int8_t* packed, float* scale_out, float* shift_out, const float* bias_col,
                                       float post_scale, void* stream) {
  S2F_REQUIRE(w && packed && scale_out && shift_out, "pack_rows_i8_device: null pointer");
  S2F_REQUIRE(pieces >= 1 && pieces <= 3 && K >= 32 && K % tc_bk(K) == 0, "pack_rows_i8_device: K must be a multiple of its K block");
  pack_rows_i8_kernel<<<n_img * rows_per_img, 128, 0, (cudaStream_t)stream>>>(w, ld, rows_per_img, K, K, pieces, packed,
                                                                              scale_out, shift_out, bias_col, post_scale);
  return check_launch("pack_rows_i8_kernel");
}

// Split fp32 weights [Cout, taps*Cin] (Cin fastest) into `pieces` signed base-128 digit planes.
// Packed layout: rows = tiles_n * pieces * 64 (tile-major, then plane, then channel-in-tile), row length
// taps * cin_pad bytes, zero padded.  rowscale[co] = 2^(e - (7*pieces - 1)) so that w ~= rowscale * sum_p digit_p * 128^(pieces-1-p).
extern "C" int64_t s2f_pack_weights_i8(const float* w, int Cout, int taps, int Cin, int pieces, int8_t* w_packed,
                                       float* w_rowscale) {
  if (Cout <= 0 || taps <= 0 || Cin <= 0 || pieces < 1 || pieces > 3) return -1;
  const int cin_pad = tc_cin_pad(Cin);
  const int64_t kpad = (int64_t)taps * cin_pad;
  const int tiles_n = (Cout + TC_BN - 1) / TC_BN;
  const int64_t bytes = (int64_t)tiles_n * pieces * TC_BN * kpad;
  if (!w_packed) return bytes;
  if (!w || !w_rowscale) return -1;
  memset(w_packed, 0, (size_t)bytes);
  const int F = 7 * pieces - 1;
  for (int co = 0; co < Cout; ++co) {
    const float* row = w + (int64_t)co * taps * Cin;
    float mx = 0.f;
    for (int64_t k = 0; k < (int64_t)taps * Cin; ++k) mx = fmaxf(mx, fabsf(row[k]));
    int e = 0;
    if (mx > 0.f) { frexpf(mx, &e); if (ldexpf(1.f, e - 1) == mx) e -= 1; }   // smallest e with mx <= 2^e
    const double inv = ldexp(1.0, F - e);
    w_rowscale[co] = (float)ldexp(1.0, e - F);
    const int tile = co / TC_BN, r = co % TC_BN;
    for (int t = 0; t < taps; ++t)
      for (int c = 0; c < Cin; ++c) {
        long long I = llrint((double)row[(int64_t)t * Cin + c] * inv);
        int digit[3] = {0, 0, 0};
        for (int pz = pieces - 1; pz >= 1; --pz) {
          int lo = (int)(((I + 64) % 128 + 128) % 128) - 64;
          digit[pz] = lo;
          I = (I - lo) / 128;
        }
        digit[0] = (int)I;     // |I| <= 64
        for (int pz = 0; pz < pieces; ++pz)
          w_packed[((int64_t)(tile * pieces + pz) * TC_BN + r) * kpad + (int64_t)t * cin_pad + c] = (int8_t)digit[pz];
      }
  }
  return bytes;
}
