// Channels-last spatial kernels: depthwise stencil, DCNv3 bilinear gather, FPN upsample-add-LIF.
// All HBM/L2-bound; threads run over channels fastest so every warp access is a coalesced row.
#include <stdlib.h>

#include "common.cuh"
#include "dw_tile.cuh"

// kernel sizes >= S2F_DW_ROLL_MIN keep the loop over kernel rows rolled: the fully unrolled 7x7 body (~60 KB of SASS)
// thrashes the instruction cache (stall_no_instruction was the top stall; 0.62 -> 0.49 ms at 256^2 x 64 ch, batch 32)
#ifndef S2F_DCN_SPLIT8
#define S2F_DCN_SPLIT8 0
#endif
#ifndef S2F_DCN_MINB
#define S2F_DCN_MINB 2        // with 256-bit corner loads: all 9 points in flight at 91 registers beat 3 blocks of 80 (1.34 vs 1.41 ms)
#endif
#ifndef S2F_DCN_UNROLL
#define S2F_DCN_UNROLL 9
#endif
#ifndef S2F_DW_ROLL_MIN
#define S2F_DW_ROLL_MIN 7
#endif

namespace s2f {

// ------------------------------------------------------------------------------------------------
// Depthwise k x k, stride 1, 'same' padding (or no padding when no_pad, output shrinks by k-1).
// One thread = a strip of TW consecutive output pixels of one row x 4 channels.  Per kernel row it loads the
// TW + k - 1 input pixels once (one 32-bit word = 4 int8 channels, one PRMT per byte, see unpack4) and the
// k weight vectors once, so the inner loop is ~65 % FFMA.  Threads run over channels fastest: every warp access is a
// contiguous row segment.  Weights are tap-major [k*k, C].  fp32 accumulation in the reference's tap order (kh, kw).
// One input pixel x 4 channels as raw bits (int8: one 32-bit word; fp32: a float4), loaded unconditionally from a clamped
// address so that the compiler can issue all loads of a row back to back.
// int8 levels enter the FFMAs as fp32 *denormals*: the isolated byte b, reinterpreted as a float, is b * 2^-149, and
// FFMA takes denormal operands at full rate (no -ftz in this build).  With the weights pre-multiplied by 2^DW_WEXP every
// product and partial sum is the reference's value times 2^(DW_WEXP-149) -- a power of two, so each rounding is the
// same as in the unscaled fp32 sum -- and the final affine multiplies by 2^(149-DW_WEXP).  One PRMT per byte instead of
// PRMT + FADD (or I2F).
constexpr int DW_WEXP = 100;
__device__ __forceinline__ void unpack4(Raw8 r, bool ok, float& x0, float& x1, float& x2, float& x3) {
  const uint32_t raw = ok ? r.w : 0u;
  x0 = __uint_as_float(__byte_perm(raw, 0u, 0x4440));
  x1 = __uint_as_float(__byte_perm(raw, 0u, 0x4441));
  x2 = __uint_as_float(__byte_perm(raw, 0u, 0x4442));
  x3 = __uint_as_float(__byte_perm(raw, 0u, 0x4443));
}
__device__ __forceinline__ void unpack4(Raw32 r, bool ok, float& x0, float& x1, float& x2, float& x3) {
  x0 = ok ? r.v.x : 0.f; x1 = ok ? r.v.y : 0.f; x2 = ok ? r.v.z : 0.f; x3 = ok ? r.v.w : 0.f;
}

template <typename AT, int KS, int TW, int CT>      // CT: compile-time channel count (0 = use the runtime C)
__global__ void __launch_bounds__(256, 2) dwconv_kernel(const AT* __restrict__ a, float a_scale,
                                                     const float* __restrict__ w_tap, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, float* __restrict__ out_f32,
                                                     int8_t* __restrict__ out_spike, int n, int H, int W, int C_rt, int Ho,
                                                     int Wo, int pad, float d_max) {
  using RT = typename RawOf<AT>::type;
  const int C = CT ? CT : C_rt;                      // a constant C turns the interior loads into [base + immediate]
  const uint32_t c4n = (uint32_t)C >> 2;
  const uint32_t strips = (uint32_t)(Wo + TW - 1) / TW;
  const uint32_t total = (uint32_t)n * Ho * strips * c4n;          // host guarantees < 2^31
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = (int)(idx % c4n) * 4;
    uint32_t r = idx / c4n;
    const int wo0 = (int)(r % strips) * TW; r /= strips;
    const int ho = (int)(r % (uint32_t)Ho);
    const int img = (int)(r / (uint32_t)Ho);
    float acc[TW][4];
#pragma unroll
    for (int t = 0; t < TW; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
    const AT* base = a + (int64_t)img * H * W * C + c;
    const int wi0 = wo0 - pad;
    int off[TW + KS - 1];                             // clamped column offsets, shared by all kernel rows
    uint32_t valid = 0;
    const bool interior = wi0 >= 0 && wi0 + TW + KS - 1 <= W;
    if (interior) {
      valid = (1u << (TW + KS - 1)) - 1u;
#pragma unroll
      for (int i = 0; i < TW + KS - 1; ++i) off[i] = (wi0 + i) * C;
    } else {
#pragma unroll
      for (int i = 0; i < TW + KS - 1; ++i) {
        const int wi = wi0 + i;
        off[i] = min(max(wi, 0), W - 1) * C;
        valid |= (wi >= 0 && wi < W) ? (1u << i) : 0u;
      }
    }
    constexpr int KH_UNROLL = KS >= S2F_DW_ROLL_MIN ? 1 : KS;
    constexpr bool PREFETCH = (KS >= S2F_DW_ROLL_MIN) && sizeof(AT) == 1;     // rolled row loop: fetch row kh+1 under row kh
    auto load_row = [&](int kh_, RT (&dst)[TW + KS - 1]) {
      const int hi_ = ho - pad + kh_;
      const AT* row_ = base + (int64_t)min(max(hi_, 0), H - 1) * W * C;
      if (CT && interior) {
        const AT* row0 = row_ + wi0 * CT;
#pragma unroll
        for (int i = 0; i < TW + KS - 1; ++i) dst[i] = load_raw(row0 + i * CT);
      } else {
#pragma unroll
        for (int i = 0; i < TW + KS - 1; ++i) dst[i] = load_raw(row_ + off[i]);
      }
    };
    RT nxt[TW + KS - 1];
    if (PREFETCH) load_row(0, nxt);
#pragma unroll KH_UNROLL
    for (int kh = 0; kh < KS; ++kh) {
      // rows outside the map are loaded from the clamped row and masked: no branch, so the loads of all kernel rows can
      // be scheduled ahead of the FFMAs of the first one
      const int hi = ho - pad + kh;
      const bool row_ok = hi >= 0 && hi < H;
      RT raw[TW + KS - 1];
      if (PREFETCH) {
#pragma unroll
        for (int i = 0; i < TW + KS - 1; ++i) raw[i] = nxt[i];
        if (kh + 1 < KS) load_row(kh + 1, nxt);
      } else {
        load_row(kh, raw);
      }
      float4 wv[KS];
#pragma unroll
      for (int kw = 0; kw < KS; ++kw) {
        wv[kw] = __ldg(reinterpret_cast<const float4*>(w_tap + (kh * KS + kw) * C + c));
        if (sizeof(AT) == 1) {                          // exact: a power of two, |w| * 2^100 stays far below FLT_MAX
          wv[kw].x *= DwScale<AT>::w_pre; wv[kw].y *= DwScale<AT>::w_pre; wv[kw].z *= DwScale<AT>::w_pre; wv[kw].w *= DwScale<AT>::w_pre;
        }
      }
#pragma unroll
      for (int i = 0; i < TW + KS - 1; ++i) {
        float x0, x1, x2, x3;
        unpack4(raw[i], row_ok && ((valid >> i) & 1u), x0, x1, x2, x3);
#pragma unroll
        for (int kw = 0; kw < KS; ++kw) {
          const int t = i - kw;                       // compile-time after unrolling
          if (t >= 0 && t < TW) {
            acc[t][0] = fmaf(x0, wv[kw].x, acc[t][0]); acc[t][1] = fmaf(x1, wv[kw].y, acc[t][1]);
            acc[t][2] = fmaf(x2, wv[kw].z, acc[t][2]); acc[t][3] = fmaf(x3, wv[kw].w, acc[t][3]);
          }
        }
      }
    }
    // y = acc * (a_scale * scale) + shift: one FFMA per output (a_scale is a power of two, so the product is exact)
    const float asc = a_scale * DwScale<AT>::post;       // a_scale is a power of two as well
    float4 sc = make_float4(asc, asc, asc, asc), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale) {
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + c));
      sc = make_float4(s4.x * asc, s4.y * asc, s4.z * asc, s4.w * asc);
      sh = __ldg(reinterpret_cast<const float4*>(shift + c));
    }
    const int64_t o0 = (((int64_t)img * Ho + ho) * Wo + wo0) * C + c;
#pragma unroll
    for (int t = 0; t < TW; ++t) {
      if (wo0 + t >= Wo) break;
      const float y[4] = {fmaf(acc[t][0], sc.x, sh.x), fmaf(acc[t][1], sc.y, sh.y), fmaf(acc[t][2], sc.z, sh.z),
                          fmaf(acc[t][3], sc.w, sh.w)};
      const int64_t o = o0 + (int64_t)t * C;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(y[0], y[1], y[2], y[3]);
      if (out_spike) {
        const uint32_t pk = pack_levels4(y[0], y[1], y[2], y[3], d_max);
        *reinterpret_cast<uint32_t*>(out_spike + o) = pk;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Depthwise k x k, register-tiled over TWO output rows with the kernel in shared memory (the production stencil; the
// arithmetic core is dw_tile.cuh).  A block of 128 threads = 128 consecutive (column strip, channel quad) pairs, walking
// down a chunk of RH output rows so that six of the eight input rows of a k = 7 step are L1 hits.  Same bits as
// dwconv_kernel (tests/test_kernels_gpu.py).
constexpr int DWT_THREADS = 128;

template <typename AT, int KS, int CT>
__global__ void __launch_bounds__(DWT_THREADS, 3) dwconv_tile_kernel(const AT* __restrict__ a, float a_scale,
                                                                  const float* __restrict__ w_tap,
                                                                  const float* __restrict__ scale,
                                                                  const float* __restrict__ shift,
                                                                  float* __restrict__ out_f32, int8_t* __restrict__ out_spike,
                                                                  int n, int H, int W, int C_rt, int Ho, int Wo, int pad,
                                                                  float d_max, int RH) {
  constexpr int TW = DW_TW;
  extern __shared__ float4 dw_wsm[];                 // [KS*KS][C/4], pre-scaled
  const int C = CT ? CT : C_rt;
  const int c4n = C >> 2;
  dw_stage_weights<AT>(dw_wsm, w_tap, KS * KS * c4n, threadIdx.x, DWT_THREADS);
  __syncthreads();
  const int strips = (Wo + TW - 1) / TW;
  const int row_tiles = strips * c4n;                // (strip, channel quad) pairs of one output row pair
  const int groups = (row_tiles + DWT_THREADS - 1) / DWT_THREADS;
  const int hchunks = (Ho + RH - 1) / RH;
  const int tasks = n * hchunks * groups;
  const float asc = a_scale * DwScale<AT>::post;     // powers of two: exact
  for (int task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int flat = (task % groups) * DWT_THREADS + (int)threadIdx.x;
    if (flat >= row_tiles) continue;
    const int hc = (task / groups) % hchunks, img = task / (groups * hchunks);
    const int cq = flat % c4n, wo0 = (flat / c4n) * TW, c = cq * 4;
    const AT* base = a + (int64_t)img * H * W * C + c;
    float4 sc = make_float4(asc, asc, asc, asc), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale) {
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + c));
      sc = make_float4(s4.x * asc, s4.y * asc, s4.z * asc, s4.w * asc);
      sh = __ldg(reinterpret_cast<const float4*>(shift + c));
    }
    const int h_end = min(Ho, (hc + 1) * RH);
    for (int ho0 = hc * RH; ho0 < h_end; ho0 += 2) {
      float2 acc[2][TW][2];
      dw_tile_8x2<AT, KS, CT>(base, H, W, C, ho0, wo0, pad, (uint32_t)__cvta_generic_to_shared(dw_wsm + cq), (uint32_t)c4n * 16u, acc);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (ho0 + t >= Ho) break;
        const int64_t o0 = (((int64_t)img * Ho + ho0 + t) * Wo + wo0) * C + c;
#pragma unroll
        for (int p = 0; p < TW; ++p) {
          if (wo0 + p >= Wo) break;
          const float y0 = fmaf(acc[t][p][0].x, sc.x, sh.x), y1 = fmaf(acc[t][p][0].y, sc.y, sh.y);
          const float y2 = fmaf(acc[t][p][1].x, sc.z, sh.z), y3 = fmaf(acc[t][p][1].y, sc.w, sh.w);
          const int64_t o = o0 + (int64_t)p * C;
          if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(y0, y1, y2, y3);
          if (out_spike) *reinterpret_cast<uint32_t*>(out_spike + o) = pack_levels4(y0, y1, y2, y3, d_max);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// DCNv3 core.  Follows dcnv3_core_pytorch's float op order (normalised location -> grid_sample
// unnormalise) so sampling coordinates agree with the reference to the last bits:
//   loc = ref + grid*os + off*os/size ; g = 2*loc - 1 ; ix = ((g + 1)*size - 1)/2   (padded image coords)
// One thread = one (pixel, group, 4 channels).  x is read through the 1-pixel zero border analytically.
struct DcnGrid { float gx[25], gy[25]; };       // per sampling point: grid offset * offset_scale (K <= 5)

// SPLIT threads share one (pixel, group): each takes CQ of the group's channel quads (shorter dependent gather chains,
// more warps in flight; the coordinate arithmetic is repeated, the kernel is latency-bound).
__device__ __forceinline__ void ldg256(const float4* p, float (&v)[8]) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

template <int CQ, int SPLIT = 1, bool WIDE = false>      // float4 channel quads per thread (Cg = 4 * CQ * SPLIT)
__global__ void __launch_bounds__(256, S2F_DCN_MINB) dcnv3_kernel(const float* __restrict__ x, const float* __restrict__ offset,
                                                    const int8_t* __restrict__ mask, float mask_scale,
                                                    float* __restrict__ out, int n, int H, int W, int G, int K,
                                                    float os, const DcnGrid grid) {
  constexpr int Cg = 4 * CQ * SPLIT;
  const int C = G * Cg, P = K * K, pad = (K - 1) / 2;
  const float Hin = (float)(H + 2 * pad), Win = (float)(W + 2 * pad);
  const int64_t total = (int64_t)n * H * W * G * SPLIT;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx;
    const int part = (int)(r % SPLIT); r /= SPLIT;
    const int g = (int)(r % G); r /= G;
    const int wo = (int)(r % W); r /= W;
    const int ho = (int)(r % H);
    const int img = (int)(r / H);
    const int64_t pix = ((int64_t)img * H + ho) * W + wo;
    const float* off = offset + pix * (int64_t)(G * P * 2) + (int64_t)g * P * 2;
    const int8_t* mk = mask + pix * (int64_t)(G * P) + (int64_t)g * P;
    const float* xb = x + (int64_t)img * H * W * C + g * Cg + part * 4 * CQ;
    const float half = (float)pad;   // dilation 1: (K-1)/2
    const float ref_x = __fdiv_rn((float)wo + half + 0.5f, Win);
    const float ref_y = __fdiv_rn((float)ho + half + 0.5f, Hin);
    float acc[CQ][4];
#pragma unroll
    for (int c = 0; c < CQ; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
    constexpr int PT_UNROLL = S2F_DCN_UNROLL;
#pragma unroll PT_UNROLL
    for (int pt = 0; pt < P; ++pt) {
      const float2 o2 = __ldg(reinterpret_cast<const float2*>(off) + pt);     // (x, y) offset of this point: 8-byte aligned
      const float m = (float)mk[pt] * mask_scale;
      // point order of _generate_dilation_grids: x index is the slow one (dcnv3_func.py:125-137); the per-point grid
      // terms (gx * os, gy * os) depend on the point only and come precomputed (same fp32 operations, done once)
      const float lx = __fadd_rn(__fadd_rn(ref_x, grid.gx[pt]), __fdiv_rn(__fmul_rn(o2.x, os), Win));
      const float ly = __fadd_rn(__fadd_rn(ref_y, grid.gy[pt]), __fdiv_rn(__fmul_rn(o2.y, os), Hin));
      const float sgx = __fadd_rn(__fmul_rn(2.f, lx), -1.f), sgy = __fadd_rn(__fmul_rn(2.f, ly), -1.f);
      // grid_sample's un-normalisation, align_corners=False, exactly as the oracle's kernel rounds it: ATen's
      // vectorised CPU kernel (cpu/GridSamplerKernel.cpp, ComputeLocationBase::unnormalize) evaluates
      // (g + 1) * (size / 2) - 0.5 with ONE rounding (the compiler contracts it to an FMA; verified bit for bit
      // against F.grid_sample on 4e5 random points, tools/check_grid_sample_order.py).  Rounding the product
      // separately moves ix by an ulp (~2e-6 px at x ~ 20), which the subtraction of floor(ix) turns into a
      // 1e-5 relative error of the bilinear weight -- the largest arithmetic difference of the whole path.
      const float ix = __fmaf_rn(__fadd_rn(sgx, 1.f), Win * 0.5f, -0.5f);
      const float iy = __fmaf_rn(__fadd_rn(sgy, 1.f), Hin * 0.5f, -0.5f);
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = (int)fx - pad, y0 = (int)fy - pad;   // back to unpadded coordinates
      const float tx = ix - fx, ty = iy - fy;
      // corner weights (0 outside the map) and clamped addresses: all loads are unconditional and independent
      const bool inx0 = x0 >= 0 && x0 < W, inx1 = x0 + 1 >= 0 && x0 + 1 < W;
      const bool iny0 = y0 >= 0 && y0 < H, iny1 = y0 + 1 >= 0 && y0 + 1 < H;
      const float q00 = (inx0 && iny0) ? (1.f - tx) * (1.f - ty) : 0.f, q01 = (inx1 && iny0) ? tx * (1.f - ty) : 0.f;
      const float q10 = (inx0 && iny1) ? (1.f - tx) * ty : 0.f, q11 = (inx1 && iny1) ? tx * ty : 0.f;
      const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
      const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
      const float4* p00 = reinterpret_cast<const float4*>(xb + ((int64_t)yc0 * W + xc0) * C);
      const float4* p01 = reinterpret_cast<const float4*>(xb + ((int64_t)yc0 * W + xc1) * C);
      const float4* p10 = reinterpret_cast<const float4*>(xb + ((int64_t)yc1 * W + xc0) * C);
      const float4* p11 = reinterpret_cast<const float4*>(xb + ((int64_t)yc1 * W + xc1) * C);
      if (CQ == 2 && WIDE) {
        // 8 channels per corner in ONE 256-bit load (sm_100: LDG.E.256): half the load instructions and L1 sector
        // look-ups of two LDG.128 that each take half of the same 32-byte sector
        float v[4][8];
        ldg256(p00, v[0]); ldg256(p01, v[1]); ldg256(p10, v[2]); ldg256(p11, v[3]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          s0 = fmaf(v[0][4 * c], q00, s0); s1 = fmaf(v[0][4 * c + 1], q00, s1); s2 = fmaf(v[0][4 * c + 2], q00, s2); s3 = fmaf(v[0][4 * c + 3], q00, s3);
          s0 = fmaf(v[1][4 * c], q01, s0); s1 = fmaf(v[1][4 * c + 1], q01, s1); s2 = fmaf(v[1][4 * c + 2], q01, s2); s3 = fmaf(v[1][4 * c + 3], q01, s3);
          s0 = fmaf(v[2][4 * c], q10, s0); s1 = fmaf(v[2][4 * c + 1], q10, s1); s2 = fmaf(v[2][4 * c + 2], q10, s2); s3 = fmaf(v[2][4 * c + 3], q10, s3);
          s0 = fmaf(v[3][4 * c], q11, s0); s1 = fmaf(v[3][4 * c + 1], q11, s1); s2 = fmaf(v[3][4 * c + 2], q11, s2); s3 = fmaf(v[3][4 * c + 3], q11, s3);
          acc[c][0] = fmaf(s0, m, acc[c][0]); acc[c][1] = fmaf(s1, m, acc[c][1]);
          acc[c][2] = fmaf(s2, m, acc[c][2]); acc[c][3] = fmaf(s3, m, acc[c][3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < CQ; ++c) {
          const float4 v00 = __ldg(p00 + c), v01 = __ldg(p01 + c), v10 = __ldg(p10 + c), v11 = __ldg(p11 + c);
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          s0 = fmaf(v00.x, q00, s0); s1 = fmaf(v00.y, q00, s1); s2 = fmaf(v00.z, q00, s2); s3 = fmaf(v00.w, q00, s3);
          s0 = fmaf(v01.x, q01, s0); s1 = fmaf(v01.y, q01, s1); s2 = fmaf(v01.z, q01, s2); s3 = fmaf(v01.w, q01, s3);
          s0 = fmaf(v10.x, q10, s0); s1 = fmaf(v10.y, q10, s1); s2 = fmaf(v10.z, q10, s2); s3 = fmaf(v10.w, q10, s3);
          s0 = fmaf(v11.x, q11, s0); s1 = fmaf(v11.y, q11, s1); s2 = fmaf(v11.z, q11, s2); s3 = fmaf(v11.w, q11, s3);
          acc[c][0] = fmaf(s0, m, acc[c][0]); acc[c][1] = fmaf(s1, m, acc[c][1]);
          acc[c][2] = fmaf(s2, m, acc[c][2]); acc[c][3] = fmaf(s3, m, acc[c][3]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CQ; ++c)
      *reinterpret_cast<float4*>(out + pix * C + g * Cg + part * 4 * CQ + c * 4) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
  }
}

// ------------------------------------------------------------------------------------------------
// y = cur + bilinear(prev -> HxW), align_corners=False (upsample_bilinear2d: src = (dst+0.5)*scale-0.5,
// clamped at 0); spikes = NI-LIF(y).  One thread = one pixel x 4 channels.
__global__ void __launch_bounds__(256) upsample_add_lif_kernel(const float* __restrict__ cur,
                                                               const float* __restrict__ prev,
                                                               int8_t* __restrict__ out_spike,
                                                               float* __restrict__ out_f32, int n, int H, int W, int Hp,
                                                               int Wp, int C, float d_max) {
  const int c4n = C >> 2;
  const float sh = (float)Hp / (float)H, sw = (float)Wp / (float)W;
  const int64_t total = (int64_t)n * H * W * c4n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c4n) * 4;
    int64_t r = idx / c4n;
    const int xo = (int)(r % W); r /= W;
    const int yo = (int)(r % H);
    const int img = (int)(r / H);
    float sy = sh * ((float)yo + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
    float sx = sw * ((float)xo + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < Hp - 1 ? 1 : 0), x1 = x0 + (x0 < Wp - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* pb = prev + (int64_t)img * Hp * Wp * C + c;
    const float4 p00 = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)y0 * Wp + x0) * C));
    const float4 p01 = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)y0 * Wp + x1) * C));
    const float4 p10 = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)y1 * Wp + x0) * C));
    const float4 p11 = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)y1 * Wp + x1) * C));
    const int64_t o = (((int64_t)img * H + yo) * W + xo) * C + c;
    const float4 cv = *reinterpret_cast<const float4*>(cur + o);
    // ATen: hy*(hx*p00 + lx*p01) + ly*(hx*p10 + lx*p11)
    float y[4];
    y[0] = cv.x + (hy * (hx * p00.x + lx * p01.x) + ly * (hx * p10.x + lx * p11.x));
    y[1] = cv.y + (hy * (hx * p00.y + lx * p01.y) + ly * (hx * p10.y + lx * p11.y));
    y[2] = cv.z + (hy * (hx * p00.z + lx * p01.z) + ly * (hx * p10.z + lx * p11.z));
    y[3] = cv.w + (hy * (hx * p00.w + lx * p01.w) + ly * (hx * p10.w + lx * p11.w));
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(y[0], y[1], y[2], y[3]);
    if (out_spike) {
      const uint32_t pk = pack_levels4(y[0], y[1], y[2], y[3], d_max);
      *reinterpret_cast<uint32_t*>(out_spike + o) = pk;
    }
  }
}

__global__ void __launch_bounds__(256) affine_add_lif_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                             const float* __restrict__ residual,
                                                             float* __restrict__ out_f32, int8_t* __restrict__ out_spike,
                                                             int64_t N4, int C, float d_max) {
  for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < N4; i4 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i4 * 4;
    float4 v = *reinterpret_cast<const float4*>(x + i);
    if (scale) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + (int)(i % C)));
      v.x = __fmul_rn(v.x, sc.x); v.y = __fmul_rn(v.y, sc.y); v.z = __fmul_rn(v.z, sc.z); v.w = __fmul_rn(v.w, sc.w);
    }
    if (residual) {
      const float4 r = *reinterpret_cast<const float4*>(residual + i);
      v.x = __fadd_rn(r.x, v.x); v.y = __fadd_rn(r.y, v.y); v.z = __fadd_rn(r.z, v.z); v.w = __fadd_rn(r.w, v.w);
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + i) = v;
    if (out_spike) {
      const uint32_t pk = pack_levels4(v.x, v.y, v.z, v.w, d_max);
      *reinterpret_cast<uint32_t*>(out_spike + i) = pk;
    }
  }
}

static inline int grid_for(int64_t total, int threads) {
  const int64_t want = ceil_div(total, threads);
  return (int)(want < 148 * 16 ? (want > 0 ? want : 1) : 148 * 16);
}

}  // namespace s2f

using namespace s2f;

extern "C" int s2f_dwconv(const void* a, int a_is_spike, float a_scale, const float* w, const float* scale,
                          const float* shift, const float* pad_value, float* out_f32, int8_t* out_spike, int n, int H,
                          int W, int C, int k, int no_pad, float d_max, void* stream) {
  S2F_REQUIRE(a && w && (out_f32 || out_spike), "dwconv: a, w and an output are required");
  S2F_REQUIRE(pad_value == nullptr, "dwconv: pad_value is not supported (RepConv is re-parameterised on the host)");
  S2F_REQUIRE(C % 4 == 0, "dwconv: C must be a multiple of 4");
  S2F_REQUIRE(k == 3 || k == 5 || k == 7, "dwconv: k must be 3, 5 or 7");
  S2F_REQUIRE((scale == nullptr) == (shift == nullptr), "dwconv: scale and shift come together");
  const int pad = no_pad ? 0 : (k - 1) / 2;
  const int Ho = H + 2 * pad - k + 1, Wo = W + 2 * pad - k + 1;
  S2F_REQUIRE(Ho > 0 && Wo > 0, "dwconv: empty output");
  cudaStream_t st = (cudaStream_t)stream;
  const float asc = a_is_spike ? a_scale : 1.f;
  S2F_REQUIRE((int64_t)W * C < (1ll << 31), "dwconv: row too large for 32-bit offsets");
  // Production path: two-row register tiles with the kernel in shared memory.  3 resident blocks of 128 threads per SM.
  const char* lg = getenv("S2F_DW_LEGACY");            // A/B switch for tools/bench_kernels.py (read per call)
  const bool legacy = lg && lg[0] == '1';
  const size_t wsm = (size_t)k * k * C * sizeof(float);
  const int64_t row_tiles = (int64_t)((Wo + 7) / 8) * (C / 4), groups = ceil_div(row_tiles, DWT_THREADS);
  if (!legacy && wsm <= 64 * 1024 && (int64_t)n * Ho * groups < (1ll << 30)) {
    const int resident = sm_count() * 3;
    int RH = 32;                                       // rows a block walks down: as long as >= 8 waves of tasks remain
    while (RH > 2 && (int64_t)n * ceil_div(Ho, RH) * groups < 8ll * resident) RH >>= 1;
    const int64_t tasks = (int64_t)n * ceil_div(Ho, RH) * groups;
    const int g = (int)(tasks < resident ? tasks : resident);
#define S2F_DWT_C(AT, KS, CT)                                                                                    \
  do {                                                                                                           \
    static std::atomic<uint64_t> once{0};                                                                        \
    if (first_use_on_this_device(once))                                                                          \
      cudaFuncSetAttribute(dwconv_tile_kernel<AT, KS, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); \
    dwconv_tile_kernel<AT, KS, CT><<<g, DWT_THREADS, wsm, st>>>(reinterpret_cast<const AT*>(a), asc, w, scale, shift, \
                                                                 out_f32, out_spike, n, H, W, C, Ho, Wo, pad, d_max, RH); \
  } while (0)
#define S2F_DWT(AT, KS)                                                       \
  do {                                                                        \
    if (C == 64) S2F_DWT_C(AT, KS, 64);                                       \
    else if (C == 128) S2F_DWT_C(AT, KS, 128);                                \
    else if (C == 256) S2F_DWT_C(AT, KS, 256);                                \
    else if (C == 512) S2F_DWT_C(AT, KS, 512);                                \
    else S2F_DWT_C(AT, KS, 0);                                                \
  } while (0)
    if (a_is_spike) {
      if (k == 3) S2F_DWT(int8_t, 3); else if (k == 5) S2F_DWT(int8_t, 5); else S2F_DWT(int8_t, 7);
    } else {
      if (k == 3) S2F_DWT(float, 3); else if (k == 5) S2F_DWT(float, 5); else S2F_DWT(float, 7);
    }
#undef S2F_DWT_C
#undef S2F_DWT
    return check_launch("dwconv_tile_kernel");
  }
  constexpr int TW = 8;
  const int64_t total = (int64_t)n * Ho * ((Wo + TW - 1) / TW) * (C / 4);
  S2F_REQUIRE(total < (1ll << 31), "dwconv: problem too large for 32-bit task indices");
  const int g = grid_for(total, 256);
#define S2F_DW_C(AT, KS, CT)                                                                                               \
  dwconv_kernel<AT, KS, TW, CT><<<g, 256, 0, st>>>(reinterpret_cast<const AT*>(a), asc, w, scale, shift, out_f32, out_spike, \
                                                   n, H, W, C, Ho, Wo, pad, d_max)
#define S2F_DW(AT, KS)                                                                            \
  do {                                                                                            \
    if (sizeof(AT) == 1 && C == 64) S2F_DW_C(AT, KS, 64);                                         \
    else if (sizeof(AT) == 1 && C == 128) S2F_DW_C(AT, KS, 128);                                  \
    else if (sizeof(AT) == 1 && C == 256) S2F_DW_C(AT, KS, 256);                                  \
    else if (sizeof(AT) == 1 && C == 512) S2F_DW_C(AT, KS, 512);                                  \
    else S2F_DW_C(AT, KS, 0);                                                                     \
  } while (0)
  if (a_is_spike) {
    if (k == 3) S2F_DW(int8_t, 3); else if (k == 5) S2F_DW(int8_t, 5); else S2F_DW(int8_t, 7);
  } else {
    if (k == 3) S2F_DW(float, 3); else if (k == 5) S2F_DW(float, 5); else S2F_DW(float, 7);
  }
#undef S2F_DW_C
#undef S2F_DW
  return check_launch("dwconv_kernel");
}

extern "C" int s2f_dcnv3_gather(const float* x, const float* offset, const int8_t* mask, float mask_scale, float* out,
                                int n, int H, int W, int G, int Cg, int K, float offset_scale, void* stream) {
  S2F_REQUIRE(x && offset && mask && out, "dcnv3_gather: null pointer");
  S2F_REQUIRE(Cg % 4 == 0, "dcnv3_gather: group channels must be a multiple of 4");
  S2F_REQUIRE(K % 2 == 1 && K >= 1, "dcnv3_gather: K must be odd");
  S2F_REQUIRE(Cg <= 16 && ((G * K * K * 2) % 2 == 0), "dcnv3_gather: at most 16 channels per group");
  const int64_t total = (int64_t)n * H * W * G;
  const int grid = grid_for(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  S2F_REQUIRE(K <= 5, "dcnv3_gather: K must be <= 5");
  DcnGrid dg;
  {
    const int pad = (K - 1) / 2;
    const float Hin = (float)(H + 2 * pad), Win = (float)(W + 2 * pad), half = (float)pad;
    for (int pt = 0; pt < K * K; ++pt) {      // the same fp32 operations the kernel used to repeat per thread
      const volatile float gx = ((float)(pt / K) - half) / Win, gy = ((float)(pt % K) - half) / Hin;
      const volatile float gxs = gx * offset_scale, gys = gy * offset_scale;
      dg.gx[pt] = gxs; dg.gy[pt] = gys;
    }
  }
  switch (Cg / 4) {
    case 1: dcnv3_kernel<1><<<grid, 256, 0, st>>>(x, offset, mask, mask_scale, out, n, H, W, G, K, offset_scale, dg); break;
    case 2:
      if (S2F_DCN_SPLIT8) dcnv3_kernel<1, 2><<<grid_for(total * 2, 256), 256, 0, st>>>(x, offset, mask, mask_scale, out, n, H, W, G, K, offset_scale, dg);
      else if ((reinterpret_cast<uintptr_t>(x) & 31) == 0) dcnv3_kernel<2, 1, true><<<grid, 256, 0, st>>>(x, offset, mask, mask_scale, out, n, H, W, G, K, offset_scale, dg);
      else dcnv3_kernel<2><<<grid, 256, 0, st>>>(x, offset, mask, mask_scale, out, n, H, W, G, K, offset_scale, dg);
      break;
    case 3: dcnv3_kernel<3><<<grid, 256, 0, st>>>(x, offset, mask, mask_scale, out, n, H, W, G, K, offset_scale, dg); break;
    default: dcnv3_kernel<4><<<grid, 256, 0, st>>>(x, offset, mask, mask_scale, out, n, H, W, G, K, offset_scale, dg); break;
  }
  return check_launch("dcnv3_kernel");
}

extern "C" int s2f_upsample_add_lif(const float* cur, const float* prev, int8_t* out_spike, float* out_f32, int n,
                                    int H, int W, int Hp, int Wp, int C, float d_max, void* stream) {
  S2F_REQUIRE(cur && prev && (out_spike || out_f32), "upsample_add_lif: null pointer");
  S2F_REQUIRE(C % 4 == 0, "upsample_add_lif: C must be a multiple of 4");
  const int64_t total = (int64_t)n * H * W * (C / 4);
  upsample_add_lif_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(cur, prev, out_spike, out_f32, n, H, W,
                                                                                  Hp, Wp, C, d_max);
  return check_launch("upsample_add_lif_kernel");
}

extern "C" int s2f_affine_add_lif(const float* x, const float* scale, const float* residual, float* out_f32,
                                  int8_t* out_spike, int64_t N, int C, float d_max, void* stream) {
  S2F_REQUIRE(x && (out_f32 || out_spike), "affine_add_lif: null pointer");
  S2F_REQUIRE(N % 4 == 0 && C % 4 == 0, "affine_add_lif: N and C must be multiples of 4");
  if (N == 0) return S2F_OK;
  affine_add_lif_kernel<<<grid_for(N / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, scale, residual, out_f32, out_spike,
                                                                                N / 4, C, d_max);
  return check_launch("affine_add_lif_kernel");
}
