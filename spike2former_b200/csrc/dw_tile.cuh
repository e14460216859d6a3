// Register-tiled depthwise stencil core shared by dwconv_tile_kernel (spatial.cu) and the fused SepConv kernel
// (sepconv_tc.cu): one thread = 8 x 2 output pixels x 4 channels of a channels-last map.
//  * the k*k weight vectors live in shared memory, already multiplied by 2^100 for int8 operands (an int8 level enters
//    the FMA as the fp32 *denormal* level * 2^-149: one PRMT per byte, every rounding identical to the plain fp32 sum;
//    FFMA / FFMA2 take denormal operands at full rate on sm_100 -- converting to normal floats first measured no faster);
//  * the multiply-adds are fma.rn.f32x2 (FFMA2, sm_100): two channels per instruction, one LDS.128 per 16 of them;
//  * every input row is loaded and unpacked ONCE for both output rows (rolled loop over the k + 1 input rows, the next
//    row's words prefetched under the current row's arithmetic); rows outside the map are skipped, columns outside are
//    loaded from clamped addresses and zero-masked (no branch around a load).
// Per output the taps are accumulated in the reference's (kh, kw) order: same bits as dwconv_kernel.
#pragma once
#include "common.cuh"

namespace s2f {

struct Raw8 { uint32_t w; };
struct Raw32 { float4 v; };
__device__ __forceinline__ Raw8 load_raw(const int8_t* p) { return Raw8{__ldg(reinterpret_cast<const uint32_t*>(p))}; }
__device__ __forceinline__ Raw32 load_raw(const float* p) { return Raw32{__ldg(reinterpret_cast<const float4*>(p))}; }
template <typename AT> struct RawOf;
template <> struct RawOf<int8_t> { using type = Raw8; };
template <> struct RawOf<float> { using type = Raw32; };
template <typename AT> struct DwScale;
template <> struct DwScale<int8_t> { static constexpr float w_pre = 0x1p100f, post = 0x1p49f; };   // 2^100, 2^(149-100)
template <> struct DwScale<float> { static constexpr float w_pre = 1.f, post = 1.f; };

__device__ __forceinline__ void unpack2x2(Raw8 r, bool ok, float2& lo, float2& hi) {
  const uint32_t raw = ok ? r.w : 0u;
  lo = make_float2(__uint_as_float(__byte_perm(raw, 0u, 0x4440)), __uint_as_float(__byte_perm(raw, 0u, 0x4441)));
  hi = make_float2(__uint_as_float(__byte_perm(raw, 0u, 0x4442)), __uint_as_float(__byte_perm(raw, 0u, 0x4443)));
}
__device__ __forceinline__ void unpack2x2(Raw32 r, bool ok, float2& lo, float2& hi) {
  lo = ok ? make_float2(r.v.x, r.v.y) : make_float2(0.f, 0.f);
  hi = ok ? make_float2(r.v.z, r.v.w) : make_float2(0.f, 0.f);
}

// stage the [k*k][C] tap-major weights, pre-scaled, as float4 per channel quad
template <typename AT>
__device__ __forceinline__ void dw_stage_weights(float4* wsm, const float* __restrict__ w_tap, int n4, int tid, int nthreads) {
  for (int i = tid; i < n4; i += nthreads) {
    float4 v = __ldg(reinterpret_cast<const float4*>(w_tap) + i);
    v.x *= DwScale<AT>::w_pre; v.y *= DwScale<AT>::w_pre; v.z *= DwScale<AT>::w_pre; v.w *= DwScale<AT>::w_pre;
    wsm[i] = v;
  }
}

constexpr int DW_TW = 8;      // output pixels per thread along x (two rows of them)

// Explicit shared-space load: a pointer into a re-aligned dynamic shared buffer is a GENERIC pointer to the compiler
// (LD.E.128 with address-space resolution, tracked on the long scoreboard like a global load).
__device__ __forceinline__ float4 dw_lds128(uint32_t saddr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// acc[t][p][0/1] = sum over taps for output pixel (ho0 + t, wo0 + p), channels (c, c+1) / (c+2, c+3), in units of
// DwScale<AT>::post.  `base` points at channel c of pixel (0, 0) of the image; w_saddr is the SHARED-space byte address of
// this thread's channel quad of tap 0, w_tap_bytes the distance between taps.
template <typename AT, int KS, int CT>
__device__ __forceinline__ void dw_tile_8x2(const AT* __restrict__ base, int H, int W, int C, int ho0, int wo0, int pad,
                                            uint32_t w_saddr, uint32_t w_tap_bytes, float2 (&acc)[2][DW_TW][2]) {
  using RT = typename RawOf<AT>::type;
  constexpr int NI = DW_TW + KS - 1;
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int p = 0; p < DW_TW; ++p) acc[t][p][0] = acc[t][p][1] = make_float2(0.f, 0.f);
  const int wi0 = wo0 - pad;
  int off[NI];
  uint32_t valid = 0;
  const bool interior = wi0 >= 0 && wi0 + NI <= W;
  if (interior) {
    valid = (1u << NI) - 1u;
#pragma unroll
    for (int i = 0; i < NI; ++i) off[i] = (wi0 + i) * C;
  } else {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int wi = wi0 + i;
      off[i] = min(max(wi, 0), W - 1) * C;
      valid |= (wi >= 0 && wi < W) ? (1u << i) : 0u;
    }
  }
  auto load_row = [&](int hi_, RT (&dst)[NI]) {
    const AT* row_ = base + (int64_t)hi_ * W * C;
    if (CT && interior) {
      const AT* row0 = row_ + wi0 * CT;
#pragma unroll
      for (int i = 0; i < NI; ++i) dst[i] = load_raw(row0 + i * CT);
    } else {
#pragma unroll
      for (int i = 0; i < NI; ++i) dst[i] = load_raw(row_ + off[i]);
    }
  };
  // input row ho0 - pad + ih feeds output row t with kernel row kh = ih - t.  Rows outside the map are SKIPPED (the
  // range is uniform over the warp): their products are +0 and an accumulator that starts at +0 never becomes -0 under
  // round-to-nearest, so leaving them out changes no bit.
  const int ih_lo = max(0, pad - ho0), ih_hi = min(KS, H - 1 - (ho0 - pad));
  RT nxt[NI];
  if (ih_lo <= ih_hi) load_row(ho0 - pad + ih_lo, nxt);
#pragma unroll 1
  for (int ih = ih_lo; ih <= ih_hi; ++ih) {
    float2 xl[NI], xh[NI];
    if (interior) {
#pragma unroll
      for (int i = 0; i < NI; ++i) unpack2x2(nxt[i], true, xl[i], xh[i]);
    } else {
#pragma unroll
      for (int i = 0; i < NI; ++i) unpack2x2(nxt[i], (valid >> i) & 1u, xl[i], xh[i]);
    }
    if (ih < ih_hi) load_row(ho0 - pad + ih + 1, nxt);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int kh = ih - t;
      if (kh >= 0 && kh < KS) {                     // uniform over the warp
        const uint32_t wrow = w_saddr + (uint32_t)(kh * KS) * w_tap_bytes;
#pragma unroll
        for (int kw = 0; kw < KS; ++kw) {
          const float4 wv = dw_lds128(wrow + (uint32_t)kw * w_tap_bytes);
          const float2 wl = make_float2(wv.x, wv.y), wh = make_float2(wv.z, wv.w);
#pragma unroll
          for (int p = 0; p < DW_TW; ++p) {
            acc[t][p][0] = __ffma2_rn(xl[p + kw], wl, acc[t][p][0]);
            acc[t][p][1] = __ffma2_rn(xh[p + kw], wh, acc[t][p][1]);
          }
        }
      }
    }
  }
}

}  // namespace s2f
