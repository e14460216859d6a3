// Training-mode neuron: Q_IFNode forward + `quant.backward`
// (neuron.py:166-197, surrogate.py:522-538) for T = 1, v0 = 0.
//
// forward : y = rint(clamp(x, 0, D)) / norm   (fp32, what the next torch op consumes)
//           tag = level | 0x80 if x is outside [0, D]   (one byte per neuron: the surrogate gradient needs to know
//           whether the input was clamped, which the level alone cannot tell -- level 0 / D occur inside the range too)
// backward: gx = gy / norm where the tag's range bit is clear, else 0
// The backward pass therefore reads 1 B + 4 B and writes 4 B per neuron and never re-reads the fp32 pre-activation
// (autograd would otherwise keep 4 B per neuron alive for every one of the 270 neurons of a step).
// 16 neurons per thread iteration: 4 x LDG.128 in, 4 x STG.128 + 1 x STG.128 out; scalar tail.
#include "common.cuh"

namespace s2f {
namespace {

__device__ __forceinline__ uint32_t tag_of(float x, float d_max, float inv_norm, float& y) {
  const float s = rintf(fminf(fmaxf(x, 0.f), d_max));
  y = s * inv_norm;
  return (uint32_t)s | ((x < 0.f || x > d_max) ? 0x80u : 0u);        // NaN compares false: passes, like the reference's mask
}

__global__ void __launch_bounds__(256) nilif_train_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              uint8_t* __restrict__ tag, int64_t N, float d_max,
                                                              float inv_norm) {
  const int64_t n16 = N >> 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const float4* xp = reinterpret_cast<const float4*>(x) + i * 4;
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldg(xp + j);
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 o;
      const uint32_t a = tag_of(v[j].x, d_max, inv_norm, o.x), b = tag_of(v[j].y, d_max, inv_norm, o.y);
      const uint32_t c = tag_of(v[j].z, d_max, inv_norm, o.z), d = tag_of(v[j].w, d_max, inv_norm, o.w);
      w[j] = a | (b << 8) | (c << 16) | (d << 24);
      reinterpret_cast<float4*>(y)[i * 4 + j] = o;
    }
    reinterpret_cast<uint4*>(tag)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (int64_t i = (n16 << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float o;
    tag[i] = (uint8_t)tag_of(x[i], d_max, inv_norm, o);
    y[i] = o;
  }
}

__global__ void __launch_bounds__(256) nilif_train_bwd_kernel(const uint8_t* __restrict__ tag, const float* __restrict__ gy,
                                                              float* __restrict__ gx, int64_t N, float inv_norm) {
  const int64_t n16 = N >> 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(tag) + i);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gy) + i * 4 + j);
      float4 o;
      o.x = (w[j] & 0x80u) ? 0.f : g.x * inv_norm;
      o.y = (w[j] & 0x8000u) ? 0.f : g.y * inv_norm;
      o.z = (w[j] & 0x800000u) ? 0.f : g.z * inv_norm;
      o.w = (w[j] & 0x80000000u) ? 0.f : g.w * inv_norm;
      reinterpret_cast<float4*>(gx)[i * 4 + j] = o;
    }
  }
  for (int64_t i = (n16 << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    gx[i] = (tag[i] & 0x80u) ? 0.f : gy[i] * inv_norm;
}

inline int grid_of(int64_t N) {
  const int64_t want = ceil_div(ceil_div(N, 16), 256);
  return (int)(want < 1 ? 1 : (want < 148 * 8 ? want : 148 * 8));
}

}  // namespace
}  // namespace s2f

using namespace s2f;

extern "C" int s2f_nilif_train_fwd(const float* x, float* y_norm, uint8_t* tag, int64_t N, float d_max, float norm,
                                   void* stream) {
  S2F_REQUIRE(x && y_norm && tag, "nilif_train_fwd: null pointer");
  S2F_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y_norm) | reinterpret_cast<uintptr_t>(tag)) & 15) == 0,
              "nilif_train_fwd: pointers must be 16-byte aligned");
  S2F_REQUIRE(d_max >= 1.f && d_max <= 127.f && norm > 0.f, "nilif_train_fwd: 1 <= d_max <= 127");
  if (N == 0) return S2F_OK;
  nilif_train_fwd_kernel<<<grid_of(N), 256, 0, (cudaStream_t)stream>>>(x, y_norm, tag, N, d_max, 1.f / norm);
  return check_launch("nilif_train_fwd_kernel");
}

extern "C" int s2f_nilif_train_bwd(const uint8_t* tag, const float* gy, float* gx, int64_t N, float norm, void* stream) {
  S2F_REQUIRE(tag && gy && gx, "nilif_train_bwd: null pointer");
  S2F_REQUIRE(((reinterpret_cast<uintptr_t>(tag) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(gx)) & 15) == 0,
              "nilif_train_bwd: pointers must be 16-byte aligned");
  if (N == 0) return S2F_OK;
  nilif_train_bwd_kernel<<<grid_of(N), 256, 0, (cudaStream_t)stream>>>(tag, gy, gx, N, 1.f / norm);
  return check_launch("nilif_train_bwd_kernel");
}
