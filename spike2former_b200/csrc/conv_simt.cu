// General implicit-GEMM convolution on CUDA cores (fp32 FMA), channels-last.
//
// This is the shape-agnostic kernel: real-valued activations (stem image, SepConv.pwconv2 whose
// input is the depthwise output, mask einsum), ragged / strided operands (the reference's
// reinterpreting reshapes), tiny token counts.  Spike-operand layers of regular shape go to the
// tcgen05 kernel in gemm_tc.cu instead.
//
//   out[img, m, co] = epi( sum_{kh,kw,ci} A[img, ho*s-p+kh, wo*s-p+kw, ci] * W[co, kh, kw, ci] )
//   epi(acc) = acc*scale[co] + shift[co] (+ residual) -> fp32 and/or NI-LIF int8 level.
#include "common.cuh"
#include "conv_direct.cuh"
#include "conv_tf32.cuh"

namespace s2f {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4, NTHREADS = 256;

template <typename AT>
__device__ __forceinline__ float a_val(const AT* p, int64_t i);
template <>
__device__ __forceinline__ float a_val<int8_t>(const int8_t* p, int64_t i) { return (float)p[i]; }
template <>
__device__ __forceinline__ float a_val<float>(const float* p, int64_t i) { return p[i]; }

// VEC: Cin % 16 == 0 and plain channels-last addressing, so one K-step of 16 sits inside one tap.
template <typename AT, bool VEC>
__global__ void __launch_bounds__(NTHREADS) conv_simt_kernel(const ConvP p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int img = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int M = p.Ho * p.Wo;
  const AT* A = reinterpret_cast<const AT*>(p.a) + (int64_t)img * p.a_img_stride;
  const float* Wt = p.w + (int64_t)img * p.w_img_stride;
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN);   // 16 column groups
  const int ty = tid / (BN / TN);   // 16 row groups
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // A-loader assignment
  const int a_row = tid % BM;          // 128 rows, two threads per row (k halves of 8)
  const int a_kh8 = tid / BM;          // 0 or 1
  const int am = m0 + a_row;
  const bool a_row_ok = am < M;
  const int a_ho = a_row_ok ? am / p.Wo : 0, a_wo = a_row_ok ? am % p.Wo : 0;
  // B-loader: 64 rows x 16 k, one float4 (4 k) per thread
  const int b_row = tid / 4, b_k4 = (tid % 4) * 4;
  const int bn = n0 + b_row;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    // ---- A tile
    if (VEC) {
      const int tap = k0 / p.Cin, ci0 = k0 % p.Cin;
      const int kh = tap / p.KW, kw = tap % p.KW;
      const int hi = a_ho * p.stride - p.pad + kh, wi = a_wo * p.stride - p.pad + kw;
      const bool ok = a_row_ok && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
      float vals[8];
      if (ok) {
        const int64_t off = ((int64_t)hi * p.W + wi) * p.Cin + ci0 + a_kh8 * 8;
        if (sizeof(AT) == 1) {
          const int2 raw = *reinterpret_cast<const int2*>(reinterpret_cast<const int8_t*>(A) + off);
          const int8_t* b = reinterpret_cast<const int8_t*>(&raw);
#pragma unroll
          for (int j = 0; j < 8; ++j) vals[j] = (float)b[j];
        } else {
          const float4 r0 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(A) + off);
          const float4 r1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(A) + off + 4);
          vals[0] = r0.x; vals[1] = r0.y; vals[2] = r0.z; vals[3] = r0.w;
          vals[4] = r1.x; vals[5] = r1.y; vals[6] = r1.z; vals[7] = r1.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) vals[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) As[a_kh8 * 8 + j][a_row] = vals[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + a_kh8 * 8 + j;
        float v = 0.f;
        if (a_row_ok && k < p.K) {
          if (p.generic) {
            v = a_val<AT>(A, (int64_t)am * p.a_stride_m + (int64_t)k * p.a_stride_k);
          } else {
            const int tap = k / p.Cin, ci = k % p.Cin;
            const int kh = tap / p.KW, kw = tap % p.KW;
            const int hi = a_ho * p.stride - p.pad + kh, wi = a_wo * p.stride - p.pad + kw;
            if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) v = a_val<AT>(A, ((int64_t)hi * p.W + wi) * p.Cin + ci);
          }
        }
        As[a_kh8 * 8 + j][a_row] = v;
      }
    }
    // ---- B tile (weights [Cout, ldw], ldw % 4 == 0, zero padded beyond K)
    {
      float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bn < p.Cout && k0 + b_k4 < p.ldw) wv = *reinterpret_cast<const float4*>(Wt + (int64_t)bn * p.ldw + k0 + b_k4);
      Bs[b_k4 + 0][b_row] = wv.x; Bs[b_k4 + 1][b_row] = wv.y; Bs[b_k4 + 2][b_row] = wv.z; Bs[b_k4 + 3][b_row] = wv.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
      bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  const int64_t out_img = (int64_t)img * M * p.Cout;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tx * TN + j;
      if (co >= p.Cout) continue;
      float y = acc[i][j] * p.a_scale;
      const float sc = p.scale ? p.scale[co] : 1.f, sh = p.shift ? p.shift[co] : 0.f;
      y = __fadd_rn(__fmul_rn(y, sc), sh);
      if (p.residual) y += p.residual[out_img + (int64_t)m * p.Cout + co];
      const int64_t o = out_img + (p.out_transposed ? (int64_t)co * M + m : (int64_t)m * p.Cout + co);
      if (p.out_f32) p.out_f32[o] = y;
      if (p.out_spike) p.out_spike[o] = (int8_t)(int)spike_level(y, p.d_max);
    }
  }
}

}  // namespace s2f

using namespace s2f;

extern "C" int s2f_conv_simt(const s2f_conv_args* a, void* stream) {
  S2F_REQUIRE(a && a->a && a->w, "conv_simt: a and w are required");
  S2F_REQUIRE(a->out_f32 || a->out_spike, "conv_simt: no output requested");
  S2F_REQUIRE(a->n > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0 && a->KH > 0 && a->KW > 0 && a->stride > 0,
              "conv_simt: bad dims");
  ConvP p;
  p.a = a->a; p.w = a->w; p.scale = a->scale; p.shift = a->shift; p.residual = a->residual;
  p.out_f32 = a->out_f32; p.out_spike = a->out_spike;
  p.n = a->n; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout; p.KH = a->KH; p.KW = a->KW;
  p.stride = a->stride; p.pad = a->pad;
  p.Ho = (a->H + 2 * a->pad - a->KH) / a->stride + 1;
  p.Wo = (a->W + 2 * a->pad - a->KW) / a->stride + 1;
  p.K = a->KH * a->KW * a->Cin;
  p.ldw = (p.K + 3) / 4 * 4;
  p.generic = (a->KH == 1 && a->KW == 1 && (a->a_stride_m != 0 || a->a_stride_k != 0)) ? 1 : 0;
  p.a_stride_m = a->a_stride_m; p.a_stride_k = a->a_stride_k;
  p.a_img_stride = a->a_img_stride ? a->a_img_stride : (int64_t)a->H * a->W * a->Cin;
  p.w_img_stride = a->w_img_stride;
  p.a_scale = a->a_is_spike ? a->a_scale : 1.f;
  p.d_max = a->d_max > 0.f ? a->d_max : 8.f;
  p.out_transposed = a->out_transposed;
  S2F_REQUIRE(p.Ho > 0 && p.Wo > 0, "conv_simt: empty output");
  if (launch_pw_tf32(p, a->a_is_spike != 0, (cudaStream_t)stream)) return check_launch("pw_tf32_kernel");
  if (launch_conv_direct(p, a->a_is_spike != 0, (cudaStream_t)stream)) return check_launch("conv_direct_kernel");
  S2F_REQUIRE(p.generic || (a->a_stride_m == 0 && a->a_stride_k == 0),
              "conv_simt: A strides are supported for 1x1 layers and for the fp32 7x7 stem (Cin = 3) only");
  const int M = p.Ho * p.Wo;
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(p.Cout, BN), (unsigned)p.n);
  S2F_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv_simt: grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  const bool base_ok = (reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0;
  S2F_REQUIRE((reinterpret_cast<uintptr_t>(a->w) & 15) == 0, "conv_simt: w must be 16-byte aligned");
  const bool vec = !p.generic && (p.Cin % 16 == 0) && base_ok && (p.a_img_stride % 16 == 0);
  if (a->a_is_spike) {
    if (vec) conv_simt_kernel<int8_t, true><<<grid, NTHREADS, 0, st>>>(p);
    else conv_simt_kernel<int8_t, false><<<grid, NTHREADS, 0, st>>>(p);
  } else {
    if (vec) conv_simt_kernel<float, true><<<grid, NTHREADS, 0, st>>>(p);
    else conv_simt_kernel<float, false><<<grid, NTHREADS, 0, st>>>(p);
  }
  return check_launch("conv_simt_kernel");
}
