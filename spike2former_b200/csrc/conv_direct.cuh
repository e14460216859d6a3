// Direct convolution for real-valued (fp32) activations on CUDA cores: one thread = PX output pixels x NC output
// channels, weights of the channel block broadcast from shared memory.
//
// Used for the two layer kinds whose activation operand is not a spike tensor:
//   * the stem, MS_DownSampling 3 -> C/2, 7x7 stride 2 on the fp32 image (sdtv2.py:412-421), K = 147;
//   * SepConv.pwconv2, a 1x1 convolution of the fp32 depthwise output (sdtv2.py:176-178).
// Per k step a thread issues NC/4 broadcast LDS.128 for PX*NC FFMA (1:8 at PX=2), so the kernel is FMA-bound; the
// accumulation runs over k in ascending order exactly like the tiled kernel in conv_simt.cu (bit-identical results).
#pragma once
#include "common.cuh"

namespace s2f {

struct ConvP {
  const void* a; const float* w; const float* scale; const float* shift; const float* residual;
  float* out_f32; int8_t* out_spike;
  int n, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo, K, ldw;
  int64_t a_img_stride, a_stride_m, a_stride_k, w_img_stride;
  float a_scale, d_max;
  int out_transposed, generic;
};

template <int PX, int NC, bool IS1X1, int CIN_CT = 0, int KW_CT = 0>      // CIN_CT / KW_CT: compile-time Cin and KW of the stem (0 = run time)
__global__ void __launch_bounds__(256, CIN_CT > 0 ? 2 : 1) conv_direct_kernel(const ConvP p) {
  extern __shared__ __align__(16) float Ws[];                    // [K][NC], zero beyond Cout
  const int co0 = blockIdx.y * NC;
  for (int e = threadIdx.x; e < p.K * NC; e += blockDim.x) {
    const int k = e / NC, c = e % NC;
    Ws[e] = (co0 + c < p.Cout) ? __ldg(p.w + (int64_t)(co0 + c) * p.ldw + k) : 0.f;
  }
  __syncthreads();
  const float* A = reinterpret_cast<const float*>(p.a);
  const int M_img = p.Ho * p.Wo;
  const int64_t M_total = (int64_t)p.n * M_img;
  const int64_t tiles = ceil_div(M_total, 256 * PX);
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t m_first = (tile * 256 + threadIdx.x) * PX;
    float acc[PX][NC];
#pragma unroll
    for (int i = 0; i < PX; ++i)
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[i][c] = 0.f;

    if (IS1X1) {
      const float* row[PX];
      bool ok[PX];
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        ok[i] = m_first + i < M_total;
        row[i] = A + (ok[i] ? (m_first + i) : 0) * (int64_t)p.Cin;
      }
      for (int k = 0; k < p.K; k += 4) {
        float4 a[PX];
#pragma unroll
        for (int i = 0; i < PX; ++i) a[i] = __ldg(reinterpret_cast<const float4*>(row[i] + k));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4* wr = reinterpret_cast<const float4*>(Ws + (k + kk) * NC);
#pragma unroll
          for (int c4 = 0; c4 < NC / 4; ++c4) {
            const float4 w = wr[c4];
#pragma unroll
            for (int i = 0; i < PX; ++i) {
              const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
              acc[i][4 * c4] = fmaf(av, w.x, acc[i][4 * c4]); acc[i][4 * c4 + 1] = fmaf(av, w.y, acc[i][4 * c4 + 1]);
              acc[i][4 * c4 + 2] = fmaf(av, w.z, acc[i][4 * c4 + 2]); acc[i][4 * c4 + 3] = fmaf(av, w.w, acc[i][4 * c4 + 3]);
            }
          }
        }
      }
    } else {
      int img[PX], hb[PX], wb[PX];
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        const int64_t m = m_first + i < M_total ? m_first + i : 0;
        img[i] = (int)(m / M_img);
        const int r = (int)(m % M_img);
        hb[i] = (r / p.Wo) * p.stride - p.pad;
        wb[i] = (r % p.Wo) * p.stride - p.pad;
      }
      if constexpr (CIN_CT > 0) {
        // the stem (Cin = 3, 7x7): a ROLLED loop over the 49 taps, two taps per iteration (ping-pong registers), the
        // loads of the next tap in flight under the 3 * 64 FFMAs of the current one.  The body is ~7 KB of SASS; the
        // version that unrolled a whole kernel row (24 KB) stalled on instruction fetch as often as it issued
        // (ncu: no_instruction 0.93 per issue, FMA pipe 42 %).
        const int ntaps = p.KH * KW_CT;
        const int64_t sm = p.a_stride_m ? p.a_stride_m : CIN_CT, sk = p.a_stride_k ? p.a_stride_k : 1;
        auto load_tap = [&](int kh, int kw, float (&dst)[PX][CIN_CT]) {
#pragma unroll
          for (int i = 0; i < PX; ++i) {
            const int hi = hb[i] + kh, wi = wb[i] + kw;
            const bool ok = hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            // pixel / channel strides: channels-last (Cin, 1) or the planar NCHW image as the caller holds it (1, H*W)
            const float* src = A + (int64_t)img[i] * p.a_img_stride + ((int64_t)(ok ? hi : 0) * p.W + (ok ? wi : 0)) * sm;
#pragma unroll
            for (int ci = 0; ci < CIN_CT; ++ci) { const float v = __ldg(src + ci * sk); dst[i][ci] = ok ? v : 0.f; }
          }
        };
        auto fma_tap = [&](int tap, const float (&av)[PX][CIN_CT]) {
          const float* wrow = Ws + tap * CIN_CT * NC;
#pragma unroll
          for (int ci = 0; ci < CIN_CT; ++ci) {
            const float4* wr = reinterpret_cast<const float4*>(wrow + ci * NC);
#pragma unroll
            for (int c4 = 0; c4 < NC / 4; ++c4) {
              const float4 w = wr[c4];
#pragma unroll
              for (int i = 0; i < PX; ++i) {
                acc[i][4 * c4] = fmaf(av[i][ci], w.x, acc[i][4 * c4]); acc[i][4 * c4 + 1] = fmaf(av[i][ci], w.y, acc[i][4 * c4 + 1]);
                acc[i][4 * c4 + 2] = fmaf(av[i][ci], w.z, acc[i][4 * c4 + 2]); acc[i][4 * c4 + 3] = fmaf(av[i][ci], w.w, acc[i][4 * c4 + 3]);
              }
            }
          }
        };
        float va[PX][CIN_CT], vb[PX][CIN_CT];
        int kh = 0, kw = 0;
        auto advance = [&]() { if (++kw == KW_CT) { kw = 0; ++kh; } };
        load_tap(0, 0, va);
        advance();
#pragma unroll 1
        for (int tap = 0; tap < ntaps; tap += 2) {
          if (tap + 1 < ntaps) { load_tap(kh, kw, vb); advance(); }
          fma_tap(tap, va);
          if (tap + 1 < ntaps) {
            if (tap + 2 < ntaps) { load_tap(kh, kw, va); advance(); }
            fma_tap(tap + 1, vb);
          }
        }
      } else {
      int k = 0;
      for (int kh = 0; kh < p.KH; ++kh)
        for (int kw = 0; kw < p.KW; ++kw) {
          const float* src[PX];
          bool ok[PX];
#pragma unroll
          for (int i = 0; i < PX; ++i) {
            const int hi = hb[i] + kh, wi = wb[i] + kw;
            ok[i] = hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            src[i] = A + (((int64_t)img[i] * p.H + (ok[i] ? hi : 0)) * p.W + (ok[i] ? wi : 0)) * p.Cin;
          }
          for (int ci = 0; ci < p.Cin; ++ci, ++k) {
            float av[PX];
#pragma unroll
            for (int i = 0; i < PX; ++i) { const float v = __ldg(src[i] + ci); av[i] = ok[i] ? v : 0.f; }
            const float4* wr = reinterpret_cast<const float4*>(Ws + k * NC);
#pragma unroll
            for (int c4 = 0; c4 < NC / 4; ++c4) {
              const float4 w = wr[c4];
#pragma unroll
              for (int i = 0; i < PX; ++i) {
                acc[i][4 * c4] = fmaf(av[i], w.x, acc[i][4 * c4]); acc[i][4 * c4 + 1] = fmaf(av[i], w.y, acc[i][4 * c4 + 1]);
                acc[i][4 * c4 + 2] = fmaf(av[i], w.z, acc[i][4 * c4 + 2]); acc[i][4 * c4 + 3] = fmaf(av[i], w.w, acc[i][4 * c4 + 3]);
              }
            }
          }
        }
      }
    }
    // ---- epilogue: same op order as conv_simt_kernel
#pragma unroll
    for (int i = 0; i < PX; ++i) {
      const int64_t m = m_first + i;
      if (m >= M_total) continue;
#pragma unroll
      for (int c4 = 0; c4 < NC / 4; ++c4) {
        const int co = co0 + 4 * c4;
        if (co >= p.Cout) break;
        float y[4] = {acc[i][4 * c4] * p.a_scale, acc[i][4 * c4 + 1] * p.a_scale, acc[i][4 * c4 + 2] * p.a_scale, acc[i][4 * c4 + 3] * p.a_scale};
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.scale) sc = __ldg(reinterpret_cast<const float4*>(p.scale + co));
        if (p.shift) sh = __ldg(reinterpret_cast<const float4*>(p.shift + co));
        y[0] = __fadd_rn(__fmul_rn(y[0], sc.x), sh.x); y[1] = __fadd_rn(__fmul_rn(y[1], sc.y), sh.y);
        y[2] = __fadd_rn(__fmul_rn(y[2], sc.z), sh.z); y[3] = __fadd_rn(__fmul_rn(y[3], sc.w), sh.w);
        const int64_t o = m * p.Cout + co;
        if (p.residual) {
          const float4 rv = *reinterpret_cast<const float4*>(p.residual + o);
          y[0] += rv.x; y[1] += rv.y; y[2] += rv.z; y[3] += rv.w;
        }
        if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + o) = make_float4(y[0], y[1], y[2], y[3]);
        if (p.out_spike)
          *reinterpret_cast<uint32_t*>(p.out_spike + o) =
              pack_levels4(y[0], y[1], y[2], y[3], p.d_max);
      }
    }
  }
}

// Launch the direct kernel when the layer fits it; returns false (nothing launched) otherwise.
inline bool launch_conv_direct(const ConvP& p, bool a_is_spike, cudaStream_t st) {
  constexpr int PX = 2, NC = 32;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (a_is_spike || p.generic || p.w_img_stride != 0 || p.out_transposed || (p.Cout & 3)) return false;
  const bool stem = !(p.KH == 1 && p.KW == 1) && p.Cin == 3 && p.KW == 7;
  if (p.a_img_stride != (int64_t)p.H * p.W * p.Cin) return false;
  if (!stem && (p.a_stride_m || p.a_stride_k)) return false;
  if (!al16(p.scale) || !al16(p.shift) || !al16(p.residual) || !al16(p.out_f32) || (reinterpret_cast<uintptr_t>(p.out_spike) & 3)) return false;
  const bool is1x1 = p.KH == 1 && p.KW == 1 && p.stride == 1 && p.pad == 0;
  if (is1x1 && ((p.Cin & 3) || !al16(p.a))) return false;
  if (!is1x1 && p.Cin > 8) return false;                          // gather path: narrow inputs only (the stem)
  const size_t smem = (size_t)p.K * NC * sizeof(float);
  if (smem > 96 * 1024) return false;
  const int64_t M_total = (int64_t)p.n * p.Ho * p.Wo;
  const int64_t tiles = ceil_div(M_total, 256 * PX);
  const int ny = (int)ceil_div(p.Cout, NC);
  int64_t gx = (148 * 2 + ny - 1) / ny;                           // ~2 resident blocks per SM in total
  if (gx > tiles) gx = tiles;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)ny);
  if (is1x1) {
    cudaFuncSetAttribute(conv_direct_kernel<PX, NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    conv_direct_kernel<PX, NC, true><<<grid, 256, smem, st>>>(p);
  } else if (p.Cin == 3 && p.KW == 7) {
    cudaFuncSetAttribute(conv_direct_kernel<PX, NC, false, 3, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    conv_direct_kernel<PX, NC, false, 3, 7><<<grid, 256, smem, st>>>(p);
  } else {
    cudaFuncSetAttribute(conv_direct_kernel<PX, NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    conv_direct_kernel<PX, NC, false><<<grid, 256, smem, st>>>(p);
  }
  return true;
}

}  // namespace s2f
