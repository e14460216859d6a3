// 1x1 convolution of fp32 activations on the tensor cores with 3xTF32 error compensation.
//
// SepConv.pwconv2 (sdtv2.py:176-178) multiplies the *real-valued* depthwise output by fp32 weights, so neither operand
// fits the int8 spike GEMM.  Each fp32 value is split into big = tf32(x) and small = tf32(x - big); the product is
// accumulated as  A_big W_big + A_small W_big + A_big W_small  (mma.sync.m16n8k8 tf32, fp32 accumulate), which keeps
// ~21 mantissa bits per product -- fp32-grade results at tensor-core speed.  The layer is memory-bound after that:
// one CTA = 128 pixels x all Cout (<= 128) channels, A and W chunks of 32 input channels staged (and split) in shared
// memory, epilogue staged through shared memory so the residual read and both stores run along channels.
#pragma once
#include "common.cuh"
#include "conv_direct.cuh"

namespace s2f {

constexpr int TF_BM = 128, TF_BK = 32, TF_LD = TF_BK + 4;

// fp32 -> tf32 (10-bit mantissa) by round-to-nearest on the bit pattern: two full-rate integer instructions instead
// of the conversion-pipe cvt.rna.tf32.f32 (ties away from zero, like cvt.rna; finite inputs only)
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NT>       // n-tiles of 8 output channels per warp: Cout <= 8 * NT
__global__ void __launch_bounds__(256, (NT <= 4 ? 4 : (NT <= 8 ? 3 : 2))) pw_tf32_kernel(const ConvP p) {
  extern __shared__ __align__(16) uint32_t tf_smem[];
  constexpr int NP = 8 * NT;                                  // padded Cout
  uint32_t* Ab = tf_smem;                                     // [128][TF_LD] big parts of the A chunk
  uint32_t* As = Ab + TF_BM * TF_LD;                          // small parts
  uint32_t* Wb = As + TF_BM * TF_LD;                          // [NP][TF_LD]
  uint32_t* Ws = Wb + NP * TF_LD;
  float* Cs = reinterpret_cast<float*>(tf_smem);              // epilogue tile [128][NP + 4], reuses the operand space
  const float* A = reinterpret_cast<const float*>(p.a);
  const int64_t M_total = (int64_t)p.n * p.Ho * p.Wo;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t tiles = ceil_div(M_total, TF_BM);
  // Register-staged software pipeline: the global loads of the next chunk (possibly the next tile's first chunk) are
  // in flight while the tensor cores work on the current one and while the epilogue runs.
  constexpr int WV = NT / 4;                                  // float4 weight loads per thread and chunk (NP * 8 / 256)
  float4 ra[4], rw[WV];
  auto issue_loads = [&](int64_t tile_, int k0_) {
    const int64_t mb = tile_ * TF_BM;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + 256 * i, row = e >> 3, c4 = e & 7;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mb + row < M_total) ra[i] = __ldg(reinterpret_cast<const float4*>(A + (mb + row) * p.Cin + k0_) + c4);
    }
#pragma unroll
    for (int i = 0; i < WV; ++i) {
      const int e = threadIdx.x + 256 * i, row = e >> 3, c4 = e & 7;
      rw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < p.Cout) rw[i] = __ldg(reinterpret_cast<const float4*>(p.w + (int64_t)row * p.ldw + k0_) + c4);
    }
  };
  auto split_store = [&](const float4& v, uint32_t* big, uint32_t* small, int row, int c4) {
    const uint32_t bb[4] = {to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w)};
    const uint32_t ss[4] = {to_tf32(v.x - __uint_as_float(bb[0])), to_tf32(v.y - __uint_as_float(bb[1])),
                            to_tf32(v.z - __uint_as_float(bb[2])), to_tf32(v.w - __uint_as_float(bb[3]))};
    *reinterpret_cast<uint4*>(big + row * TF_LD + 4 * c4) = make_uint4(bb[0], bb[1], bb[2], bb[3]);
    *reinterpret_cast<uint4*>(small + row * TF_LD + 4 * c4) = make_uint4(ss[0], ss[1], ss[2], ss[3]);
  };
  if (blockIdx.x < tiles) issue_loads(blockIdx.x, 0);
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t m0 = tile * TF_BM;
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    for (int k0 = 0; k0 < p.K; k0 += TF_BK) {
      __syncthreads();                                        // previous chunk (or the previous tile's epilogue) consumed
#pragma unroll
      for (int i = 0; i < 4; ++i) { const int e = threadIdx.x + 256 * i; split_store(ra[i], Ab, As, e >> 3, e & 7); }
#pragma unroll
      for (int i = 0; i < WV; ++i) { const int e = threadIdx.x + 256 * i; split_store(rw[i], Wb, Ws, e >> 3, e & 7); }
      __syncthreads();
      if (k0 + TF_BK < p.K) issue_loads(tile, k0 + TF_BK);
      else if (tile + gridDim.x < tiles) issue_loads(tile + gridDim.x, 0);
#pragma unroll
      for (int k8 = 0; k8 < TF_BK / 8; ++k8) {
        const int ao = (warp * 16 + g) * TF_LD + k8 * 8 + t;
        const uint32_t ab[4] = {Ab[ao], Ab[ao + 8 * TF_LD], Ab[ao + 4], Ab[ao + 8 * TF_LD + 4]};
        const uint32_t as[4] = {As[ao], As[ao + 8 * TF_LD], As[ao + 4], As[ao + 8 * TF_LD + 4]};
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int bo = (j * 8 + g) * TF_LD + k8 * 8 + t;
          const uint32_t wb0 = Wb[bo], wb1 = Wb[bo + 4], ws0 = Ws[bo], ws1 = Ws[bo + 4];
          mma_tf32(acc[j], as, wb0, wb1);                     // small terms first
          mma_tf32(acc[j], ab, ws0, ws1);
          mma_tf32(acc[j], ab, wb0, wb1);
        }
      }
    }
    __syncthreads();                                          // operands dead -> reuse as the epilogue tile
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      float* c0 = Cs + (warp * 16 + g) * (NP + 4) + j * 8 + 2 * t;
      *reinterpret_cast<float2*>(c0) = make_float2(acc[j][0], acc[j][1]);
      *reinterpret_cast<float2*>(c0 + 8 * (NP + 4)) = make_float2(acc[j][2], acc[j][3]);
    }
    __syncthreads();
    const int c4n = p.Cout >> 2;
    for (int e = threadIdx.x; e < TF_BM * c4n; e += 256) {
      const int row = e / c4n, co = (e % c4n) * 4;
      const int64_t m = m0 + row;
      if (m >= M_total) continue;
      const float4 a4 = *reinterpret_cast<const float4*>(Cs + row * (NP + 4) + co);
      float y[4] = {a4.x * p.a_scale, a4.y * p.a_scale, a4.z * p.a_scale, a4.w * p.a_scale};
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
      if (p.out_spike) *reinterpret_cast<uint32_t*>(p.out_spike + o) = pack_levels4(y[0], y[1], y[2], y[3], p.d_max);
    }
  }
}

// Launch when the layer fits (fp32 1x1, Cin % 32 == 0, Cout % 8 == 0 and <= 128); returns false otherwise.
inline bool launch_pw_tf32(const ConvP& p, bool a_is_spike, cudaStream_t st) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (a_is_spike || p.generic || p.w_img_stride != 0 || p.out_transposed) return false;
  if (!(p.KH == 1 && p.KW == 1 && p.stride == 1 && p.pad == 0)) return false;
  if ((p.Cin % TF_BK) || (p.Cout % 8) || p.Cout > 128 || p.Cout < 8 || (p.ldw & 3)) return false;
  if (p.a_img_stride != (int64_t)p.H * p.W * p.Cin) return false;
  if (!al16(p.a) || !al16(p.w) || !al16(p.scale) || !al16(p.shift) || !al16(p.residual) || !al16(p.out_f32) ||
      (reinterpret_cast<uintptr_t>(p.out_spike) & 3)) return false;
  const int nt = (p.Cout + 7) / 8;
  const int NT = nt <= 4 ? 4 : (nt <= 8 ? 8 : 16);
  const size_t smem = (size_t)(2 * TF_BM * TF_LD + 2 * 8 * NT * TF_LD) * 4;
  const size_t smem_c = (size_t)TF_BM * (8 * NT + 4) * 4;
  const size_t bytes = smem > smem_c ? smem : smem_c;
  const int64_t tiles = ceil_div((int64_t)p.n * p.Ho * p.Wo, TF_BM);
  const int per_sm = NT == 4 ? 4 : (NT == 8 ? 3 : 2);
  const int grid = (int)(tiles < 148 * per_sm ? tiles : 148 * per_sm);
#define S2F_TF(N_)                                                                                         \
  do {                                                                                                     \
    cudaFuncSetAttribute(pw_tf32_kernel<N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);    \
    cudaFuncSetAttribute(pw_tf32_kernel<N_>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);         \
    pw_tf32_kernel<N_><<<grid, 256, bytes, st>>>(p);                                                      \
  } while (0)
  if (NT == 4) S2F_TF(4); else if (NT == 8) S2F_TF(8); else S2F_TF(16);
#undef S2F_TF
  return true;
}

}  // namespace s2f
