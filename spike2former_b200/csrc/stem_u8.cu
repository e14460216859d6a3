// Fused data-preprocessor + stem on the int8 tensor cores.
//
//   reference:  x = (float(img[[2,1,0]]) - mean) / std                      (mmseg/models/data_preprocessor.py:121-126)
//               y = BN(conv7x7_s2_p3(x) + b) ; s = LIF(y)                    (sdtv2.py:412-421, first MS_DownSampling)
//
// The pixels are exact integers (0..255), so the convolution is an exact int32 GEMM on tcgen05.mma.kind::i8 with an
// UNSIGNED A operand: sum_inb (W/std)_q * p, (W/std)_q = three signed base-128 digit planes as in the spike GEMM.
// The mean subtraction is a bias: sum_inb (W/std)_q * mean -- it depends on which taps fall inside the image (the
// reference pads the *normalised* image with zeros, i.e. raw pixels with `mean`), so the host tabulates the folded
// shift for every (rows cut at the top, bottom, columns cut left, right) combination: 4^4 x Cout floats in shared memory.
//
// K layout: 7 kernel rows x 24 bytes (21 pixel bytes in the image's own memory order + 3 zeros) = 168 -> 192 = six
// K=32 MMAs per tile; M = 128 output pixels (8 rows x 16 columns), N = 192 (three planes of one 64-channel tile).
// Warp roles: 8 producer warps gather the 7 x 21-byte runs of each output pixel straight from the uint8 image into the
// SWIZZLE_32B K-major A image (the fp32 image is never materialised), 1 warp issues the MMAs, 4 warps drain TMEM,
// apply scale / tabulated shift and write the fp32 stream and its int8 spike twin.  Two A stages, two accumulators.
#include "common.cuh"

namespace s2f {
namespace stem {

constexpr int BM = 128, TW = 16, TH = 8;
constexpr int KROW = 24, KTOT = 192, NCHUNK = KTOT / 32;        // bytes of K per kernel row / per tile row / K=32 chunks
constexpr int NB = 192;                                         // MMA N: 3 planes x 64 channel slots
constexpr int PROD_WARPS = 8, EPI_WARPS = 4;
constexpr int THREADS = (PROD_WARPS + 1 + EPI_WARPS) * 32;
constexpr int ACC_COLS = 256;
constexpr int A_STAGE = NCHUNK * BM * 32;                       // 24 KB
constexpr int B_BYTES = NCHUNK * NB * 32;                       // 36 KB
constexpr int PATCH_ROWS = 2 * TH + 5, PATCH_LD = 132;          // input rows of a tile; 111 (HWC) / 3 x 40 (CHW) bytes used; 132: rows shift by one bank
constexpr int PATCH_BYTES = PATCH_ROWS * PATCH_LD;

struct Params {
  const uint8_t* img; const int8_t* w_packed; int w_ld;         // packed weights: [192 rows][w_ld bytes]
  const float* scale; const float* shift_tab;                   // [Cout], [256][Cout]
  float* out_f32; int8_t* out_spike;
  int n, H, W, Ho, Wo, Cout;
  int64_t img_stride, row_stride; int px_stride, ch_stride;     // in bytes: HWC (W*3, 3, 1) or CHW (W, 1, H*W)
  int tiles_w, tiles_h, tiles, ctas;
  float d_max;
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void bar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "STEM_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra STEM_DONE;\n\t"
      "bra STEM_WAIT;\n\t"
      "STEM_DONE:\n\t"
      "}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mma_u8s8(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// K-major SWIZZLE_32B descriptor (same encoding as gemm_tc.cu::smem_desc for bk = 32): SBO = 8 rows x 32 B
__device__ __forceinline__ uint64_t desc32(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// byte offset of (row r, k) inside an operand image of `rows` rows: chunk-major, 32-byte rows, 16-byte units XOR-ed with
// address bit 7 (what TMA's SWIZZLE_32B writes for a 1024-byte aligned tile)
__device__ __forceinline__ uint32_t swz(int rows, int r, int k) {
  return (uint32_t)((k >> 5) * rows * 32 + r * 32 + ((((k >> 4) & 1) ^ ((r >> 2) & 1)) << 4) + (k & 15));
}

// One 32-byte K chunk of output pixel r: bytes k = 32*C .. 32*C+31 of the row (k = kh*24 + j, j < 21 pixel bytes of kernel
// row kh in the image's memory order, zeros otherwise), gathered from the staged patch and written as two 16-byte units.
// With lanes = consecutive pixels the two STS.128 of a quarter warp cover all 32 banks (that is what the swizzle is for).
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
template <int C, bool CHW>
__device__ __forceinline__ void build_chunk(uint32_t patch, uint32_t dst, int r) {
  const uint32_t base = patch + (uint32_t)((2 * (r >> 4)) * PATCH_LD + (CHW ? 2 : 6) * (r & 15));
  uint32_t w[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint32_t word = 0u;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      constexpr int dummy = 0; (void)dummy;
      const int k = 32 * C + 4 * q + b, kh = k / KROW, j = k % KROW;       // compile-time after unrolling
      if (kh < 7 && j < 21) word |= lds_u8(base + (uint32_t)(kh * PATCH_LD + (CHW ? (j / 7) * 40 + (j % 7) : j))) << (8 * b);
    }
    w[q] = word;
  }
  const uint32_t sw = (uint32_t)((r >> 2) & 1);
  const uint32_t row = dst + (uint32_t)(C * BM * 32 + r * 32);
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + ((0u ^ sw) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + ((1u ^ sw) << 4)), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
template <bool CHW>
__device__ __forceinline__ void build_tile(uint32_t patch, uint32_t dst, int ptid) {
  const int r = ptid & (BM - 1);
  if (ptid < BM) { build_chunk<0, CHW>(patch, dst, r); build_chunk<2, CHW>(patch, dst, r); build_chunk<4, CHW>(patch, dst, r); }
  else { build_chunk<1, CHW>(patch, dst, r); build_chunk<3, CHW>(patch, dst, r); build_chunk<5, CHW>(patch, dst, r); }
}

__global__ void __launch_bounds__(THREADS, 1) stem_u8_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                           // 36 KB
  uint8_t* sA = smem + B_BYTES;                                 // 2 x 24 KB (B_BYTES is a multiple of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 2 * A_STAGE);
  uint64_t *a_full = bars, *a_empty = bars + 2, *t_full = bars + 4, *t_empty = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);     // [64]
  float* s_tab = s_scale + 64;                                  // [256][Cout]
  uint8_t* s_patch = reinterpret_cast<uint8_t*>(s_tab + 256 * p.Cout);    // [2][PATCH_BYTES]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { bar_init(&a_full[s], PROD_WARPS); bar_init(&a_empty[s], 1); bar_init(&t_full[s], 1); bar_init(&t_empty[s], EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(2 * ACC_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // weights -> swizzled B image; epilogue constants
  for (int e = threadIdx.x; e < NB * (KTOT / 16); e += THREADS) {
    const int r = e / (KTOT / 16), u = e % (KTOT / 16);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.w_packed + (int64_t)r * p.w_ld + u * 16));
    *reinterpret_cast<uint4*>(sB + swz(NB, r, u * 16)) = v;
  }
  for (int e = threadIdx.x; e < 64; e += THREADS) s_scale[e] = e < p.Cout ? __ldg(p.scale + e) : 0.f;
  for (int e = threadIdx.x; e < 256 * p.Cout; e += THREADS) s_tab[e] = __ldg(p.shift_tab + e);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int per_img = p.tiles_w * p.tiles_h;

  if (warp < PROD_WARPS) {
    // ================================================================== producers
    // (1) the tile's input patch (21 rows x 37 pixels x 3 channels, zeros outside the image) -> shared memory with
    //     byte loads that run along the image rows: ~9 independent loads per thread, one latency per tile;
    // (2) producer-only barrier; (3) 128 pixels x 8 K-rows of 24 bytes are assembled from shared memory into the
    //     swizzled A image.  The patch buffer is double buffered like the A stage.
    const bool chw = p.ch_stride != 1;
    const int ptid = threadIdx.x;
    constexpr int NIT = (PATCH_ROWS * 111 + PROD_WARPS * 32 - 1) / (PROD_WARPS * 32);
    // per-thread patch bytes: position inside the patch (tile independent) and the registers of the NEXT tile's values
    int po[NIT], pry[NIT], pdx[NIT], pc[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int e = ptid + i * PROD_WARPS * 32;
      const int ee = e < PATCH_ROWS * 111 ? e : 0;
      const int ry = ee / 111, j = ee % 111;
      pc[i] = chw ? j / 37 : j % 3; pdx[i] = chw ? j % 37 : j / 3; pry[i] = ry;
      po[i] = e < PATCH_ROWS * 111 ? ry * PATCH_LD + (chw ? pc[i] * 40 + pdx[i] : j) : -1;
    }
    uint32_t pv[NIT];
    // all loads of a tile are issued back to back (unconditional, from clamped addresses): interleaved with the stores,
    // every store waited for its own load -- ten DRAM / L2 latencies per tile
    auto load_patch = [&](int tile_) {
      const int img_ = tile_ / per_img, tt_ = tile_ % per_img;
      const int ys_ = 2 * ((tt_ / p.tiles_w) * TH) - 3, xs_ = 2 * ((tt_ % p.tiles_w) * TW) - 3;
      const uint8_t* ib_ = p.img + (int64_t)img_ * p.img_stride;
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const int y = ys_ + pry[i], x = xs_ + pdx[i];
        const bool inb = y >= 0 && y < p.H && x >= 0 && x < p.W;
        const uint32_t v = __ldg(ib_ + (int64_t)(inb ? y : 0) * p.row_stride + (int64_t)(inb ? x : 0) * p.px_stride + (int64_t)pc[i] * p.ch_stride);
        pv[i] = inb ? v : 0u;
      }
    };
    if (blockIdx.x < p.tiles) load_patch(blockIdx.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += p.ctas, ++it) {
      const int s = it & 1;
      const uint32_t patch_s = s32(s_patch + s * PATCH_BYTES);
#pragma unroll
      for (int i = 0; i < NIT; ++i)
        if (po[i] >= 0) asm volatile("st.shared.u8 [%0], %1;" ::"r"(patch_s + (uint32_t)po[i]), "r"(pv[i]) : "memory");
      if (tile + p.ctas < p.tiles) load_patch(tile + p.ctas);    // in flight under the barrier and the A build below
      bar_wait(&a_empty[s], ((it >> 1) & 1) ^ 1);
      asm volatile("bar.sync 2, %0;" ::"n"(PROD_WARPS * 32) : "memory");
      const uint32_t dst = s32(sA + s * A_STAGE);
      if (chw) build_tile<true>(patch_s, dst, ptid);
      else build_tile<false>(patch_s, dst, ptid);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(&a_full[s]);
    }
  } else if (warp == PROD_WARPS) {
    // ================================================================== MMA issuer
    // idesc: D = S32 (2 << 4), A = U8 (0 << 7), B = S8 (1 << 10), K-major both, N >> 3 @17, M >> 4 @24
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += p.ctas, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      bar_wait(&a_full[s], ph);
      bar_wait(&t_empty[s], ph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = s32(sA + s * A_STAGE), b0 = s32(sB);
      const uint32_t d = tmem_base + (uint32_t)(s * ACC_COLS);
      if (elect()) {
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) mma_u8s8(d, desc32(a0 + c * BM * 32), desc32(b0 + c * NB * 32), idesc, c ? 1u : 0u);
        commit(&a_empty[s]);
        commit(&t_full[s]);
      }
      __syncwarp();
    }
  } else {
    // ================================================================== epilogue: lane = pixel, all Cout channels
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += p.ctas, ++it) {
      const int s = it & 1;
      const int img = tile / per_img, tt = tile % per_img;
      const int ho = (tt / p.tiles_w) * TH + (r >> 4), wo = (tt % p.tiles_w) * TW + (r & 15);
      const bool ok = ho < p.Ho && wo < p.Wo;
      // taps cut off by the image border: rows above / below, columns left / right (each 0..3)
      const int top = min(max(3 - 2 * ho, 0), 3), bot = min(max(2 * ho + 4 - p.H, 0), 3);
      const int lef = min(max(3 - 2 * wo, 0), 3), rig = min(max(2 * wo + 4 - p.W, 0), 3);
      const uint32_t tab_s = s32(s_tab) + (uint32_t)((((top * 4 + bot) * 4 + lef) * 4 + rig) * p.Cout) * 4u;
      const uint32_t scale_s = s32(s_scale);
      const int64_t o = (((int64_t)img * p.Ho + (ok ? ho : 0)) * p.Wo + (ok ? wo : 0)) * p.Cout;
      bar_wait(&t_full[s], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * ACC_COLS);
      for (int j0 = 0; j0 < p.Cout; j0 += 16) {                  // Cout % 16 == 0
        uint32_t d0[16], d1[16], d2[16];
        ld16(trow + j0, d0); ld16(trow + 64 + j0, d1); ld16(trow + 128 + j0, d2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float y[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {                            // explicit LDS.128 (generic loads otherwise)
          float4 sc, sh;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sc.x), "=f"(sc.y), "=f"(sc.z), "=f"(sc.w) : "r"(scale_s + (uint32_t)(j0 + 4 * q) * 4u));
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sh.x), "=f"(sh.y), "=f"(sh.z), "=f"(sh.w) : "r"(tab_s + (uint32_t)(j0 + 4 * q) * 4u));
          const float s4[4] = {sc.x, sc.y, sc.z, sc.w}, h4[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;
            const float v = fmaf((float)(int)d0[j], 16384.f, (float)((int)d1[j] * 128 + (int)d2[j]));
            y[j] = fmaf(v, s4[e], h4[e]);
          }
        }
        if (ok) {
          if (p.out_f32) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(p.out_f32 + o + j0 + 4 * q) = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
          }
          if (p.out_spike) {
            uint32_t w4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w4[q] = pack_levels4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3], p.d_max);
            *reinterpret_cast<uint4*>(p.out_spike + o + j0) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(&t_empty[s]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == PROD_WARPS) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * ACC_COLS));
  }
}

}  // namespace stem
}  // namespace s2f

using namespace s2f;

extern "C" int s2f_stem_u8(const uint8_t* img, int chw, const int8_t* w_packed, int w_ld, const float* scale,
                           const float* shift_tab, float* out_f32, int8_t* out_spike, int n, int H, int W, int Cout,
                           float d_max, void* stream) {
  S2F_REQUIRE(img && w_packed && scale && shift_tab && (out_f32 || out_spike), "stem_u8: null pointer");
  S2F_REQUIRE(n > 0 && H >= 8 && W >= 8, "stem_u8: image must be at least 8 x 8");
  S2F_REQUIRE(Cout >= 16 && Cout <= 64 && Cout % 16 == 0, "stem_u8: Cout must be 16, 32, 48 or 64");
  S2F_REQUIRE(w_ld >= stem::KTOT && w_ld % 16 == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
              "stem_u8: packed weights must be 16-byte aligned rows of >= 192 bytes");
  S2F_REQUIRE((!out_f32 || (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0) && (!out_spike || (reinterpret_cast<uintptr_t>(out_spike) & 15) == 0),
              "stem_u8: outputs must be 16-byte aligned");
  stem::Params p{};
  p.img = img; p.w_packed = w_packed; p.w_ld = w_ld; p.scale = scale; p.shift_tab = shift_tab;
  p.out_f32 = out_f32; p.out_spike = out_spike;
  p.n = n; p.H = H; p.W = W; p.Cout = Cout;
  p.Ho = (H + 6 - 7) / 2 + 1; p.Wo = (W + 6 - 7) / 2 + 1;
  p.img_stride = (int64_t)H * W * 3;
  if (chw) { p.row_stride = W; p.px_stride = 1; p.ch_stride = H * W; }
  else { p.row_stride = (int64_t)W * 3; p.px_stride = 3; p.ch_stride = 1; }
  S2F_REQUIRE((int64_t)H * W < (1ll << 31), "stem_u8: image too large");
  p.tiles_w = (p.Wo + stem::TW - 1) / stem::TW; p.tiles_h = (p.Ho + stem::TH - 1) / stem::TH;
  const int64_t tiles = (int64_t)n * p.tiles_w * p.tiles_h;
  S2F_REQUIRE(tiles < (1ll << 31), "stem_u8: too many tiles");
  p.tiles = (int)tiles;
  p.ctas = (int)(tiles < sm_count() ? tiles : sm_count());
  p.d_max = d_max > 0.f ? d_max : 8.f;
  const size_t smem = 1024 + stem::B_BYTES + 2 * stem::A_STAGE + 128 + 64 * 4 + (size_t)256 * Cout * 4 + 64 + 2 * stem::PATCH_BYTES;
  static std::atomic<uint64_t> attr{0};
  if (first_use_on_this_device(attr)) {
    cudaError_t e = cudaFuncSetAttribute(stem::stem_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(S2F_ERR_CUDA, "stem_u8: smem attribute: %s", cudaGetErrorString(e));
  }
  // one CTA per SM: each allocates all 512 TMEM columns
  stem::stem_u8_kernel<<<p.ctas, stem::THREADS, smem < 120 * 1024 ? 120 * 1024 : smem, (cudaStream_t)stream>>>(p);
  return check_launch("stem_u8_kernel");
}
