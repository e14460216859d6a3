// Callers either side of the hot path:
//   * s2f_preprocess_u8: SegDataPreProcessor.forward (mmseg/models/data_preprocessor.py:109-152) + stack_batch
//     (mmseg/utils/misc.py:30-110) for a batch of equally sized uint8 images: optional BGR<->RGB flip, fp32 cast,
//     (x - mean) / std with the reference's two roundings, right/bottom padding with pad_val -- written straight in the
//     channels-last layout the stem reads, so neither the fp32 NCHW image nor its NHWC copy is ever materialised.
//   * s2f_level_hist: per-tensor histogram of spike levels (the firing-rate / energy census of
//     tools/cal_firing_num.py:140-171) read from the int8 levels the kernels already emit.
// Both are HBM-bound streaming kernels: 128-bit accesses, grid sized to the SM count.
#include "common.cuh"

namespace s2f {

struct PreP {
  const uint8_t* img; float* out;
  int n, H, W, Hp, Wp, chw, swap_rb;
  float mean[3], std[3], pad_val;
  int normalize;
};

// one thread = 4 consecutive output pixels of one row (12 floats = three float4 stores)
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const PreP p) {
  const int wq = (p.Wp + 3) >> 2;
  const int64_t total = (int64_t)p.n * p.Hp * wq;
  const int64_t plane = (int64_t)p.H * p.W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int xq = (int)(idx % wq);
    int64_t r = idx / wq;
    const int y = (int)(r % p.Hp);
    const int img = (int)(r / p.Hp);
    float v[12];
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const int x = 4 * xq + px;
      const bool in = y < p.H && x < p.W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int cs = p.swap_rb ? 2 - c : c;                    // inputs[[2, 1, 0], ...] (data_preprocessor.py:122)
        float f = p.pad_val;
        if (in) {
          const uint8_t b = p.chw ? __ldg(p.img + ((int64_t)img * 3 + cs) * plane + (int64_t)y * p.W + x)
                                  : __ldg(p.img + (((int64_t)img * p.H + y) * p.W + x) * 3 + cs);
          f = (float)b;                                          // .float() (:124)
          if (p.normalize) f = __fdiv_rn(__fsub_rn(f, p.mean[c]), p.std[c]);      // (x - mean) / std (:126)
        }
        v[3 * px + c] = f;
      }
    }
    float* o = p.out + (((int64_t)img * p.Hp + y) * p.Wp + 4 * xq) * 3;
    if (4 * xq + 3 < p.Wp && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
      float4* o4 = reinterpret_cast<float4*>(o);
      o4[0] = make_float4(v[0], v[1], v[2], v[3]);
      o4[1] = make_float4(v[4], v[5], v[6], v[7]);
      o4[2] = make_float4(v[8], v[9], v[10], v[11]);
    } else {
      for (int e = 0; e < 12; ++e)
        if (4 * xq + e / 3 < p.Wp) o[e] = v[e];
    }
  }
}

// levels int8 [N] -> hist[0..15] (uint64 counts), ties: nothing to do with pre-activations here.
// One thread = 16 levels (one LDG.128); per-warp counts are combined with shuffles, per-block in shared memory.
__global__ void __launch_bounds__(256) level_hist_kernel(const int8_t* __restrict__ lv, int64_t N,
                                                         unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[16];
  if (threadIdx.x < 16) sh[threadIdx.x] = 0;
  __syncthreads();
  unsigned int cnt[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) cnt[i] = 0;
  const int64_t n16 = N >> 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(lv) + i);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const unsigned l = (ws[k] >> (8 * b)) & 15u;
#pragma unroll
        for (int v = 0; v < 16; ++v) cnt[v] += (l == (unsigned)v);
      }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n16 << 4; i < N; ++i) cnt[lv[i] & 15] += 1;
#pragma unroll
  for (int v = 0; v < 16; ++v) {
    unsigned c = cnt[v];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sh[v], c);
  }
  __syncthreads();
  if (threadIdx.x < 16 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

}  // namespace s2f

using namespace s2f;

extern "C" int s2f_preprocess_u8(const uint8_t* img, int chw, float* out, int n, int H, int W, int Hp, int Wp,
                                 const float* mean, const float* std, int swap_rb, float pad_val, void* stream) {
  S2F_REQUIRE(img && out, "preprocess_u8: null pointer");
  S2F_REQUIRE(n > 0 && H > 0 && W > 0 && Hp >= H && Wp >= W, "preprocess_u8: padded size must cover the image");
  S2F_REQUIRE((mean == nullptr) == (std == nullptr), "preprocess_u8: mean and std come together");
  PreP p{};
  p.img = img; p.out = out; p.n = n; p.H = H; p.W = W; p.Hp = Hp; p.Wp = Wp; p.chw = chw; p.swap_rb = swap_rb;
  p.pad_val = pad_val; p.normalize = mean != nullptr;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean ? mean[c] : 0.f; p.std[c] = std ? std[c] : 1.f; }
  const int64_t total = (int64_t)n * Hp * ((Wp + 3) / 4);
  int64_t blocks = ceil_div(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  preprocess_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("preprocess_u8_kernel");
}

extern "C" int s2f_level_hist(const int8_t* levels, int64_t N, unsigned long long* hist16, void* stream) {
  if (N == 0) return S2F_OK;
  S2F_REQUIRE(levels && hist16 && N > 0, "level_hist: null pointer");
  S2F_REQUIRE((reinterpret_cast<uintptr_t>(levels) & 15) == 0, "level_hist: levels must be 16-byte aligned");
  int64_t blocks = ceil_div(N >> 4, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  level_hist_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(levels, N, hist16);
  return check_launch("level_hist_kernel");
}
