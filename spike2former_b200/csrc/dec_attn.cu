// Masked spike attention in the reference's own association order (mmcv_spike/transformer.py:262-270, 345-353):
//   scores[q, n] = (Q[q] . K[n]) / sqrt(embed_dim);  scores.masked_fill(mask, 0);  out[q] = sum_n scores[q, n] V[n]
// followed by the attn_spike neuron.  A mask breaks the Q (K^T V) re-association s2f_linear_attn relies on, so this
// kernel walks the keys: all operands are integer spike levels, scores fit int32 (<= 64 d), the output sum is kept
// in int64 and rounded once -- exact where the reference's fp32 sum over up to 2^18 keys is order dependent.
// One warp = one query of one head; a block stages 32-key chunks of K and V of that head in shared memory for its
// 8 queries; lane j scores key j with dp4a, the scores travel by shuffle, lane e accumulates channels e and e + 32.
#include "common.cuh"

namespace s2f {
namespace {

constexpr int DA_Q = 8;          // queries (warps) per block
constexpr int DA_KEYS = 32;      // keys per chunk

__global__ void __launch_bounds__(DA_Q * 32) dec_attn_kernel(const int8_t* __restrict__ q, const int8_t* __restrict__ k,
                                                              const int8_t* __restrict__ v, const uint8_t* __restrict__ mask,
                                                              int8_t* __restrict__ out_spike, float* __restrict__ out_f32,
                                                              int Nq, int Nk, int heads, int d, int q_ld, int kv_ld,
                                                              int out_ld, float out_scale, float d_max) {
  extern __shared__ int32_t sm[];
  const int dw = (d + 3) / 4;                 // 32-bit words per head row
  const int rw = dw + 1;                      // padded row (bank spread)
  int32_t* ks = sm;                           // [DA_KEYS][rw]
  int32_t* vs = ks + DA_KEYS * rw;            // [DA_KEYS][rw]
  int32_t* qs = vs + DA_KEYS * rw;            // [DA_Q][rw]
  const int img = blockIdx.z, head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * DA_Q + warp;
  const bool q_ok = qi < Nq;
  // stage the block's query rows (zero padded to whole words)
  for (int t = threadIdx.x; t < DA_Q * dw; t += blockDim.x) {
    const int r = t / dw, w = t % dw, qq = blockIdx.x * DA_Q + r;
    int32_t word = 0;
    if (qq < Nq) {
      const int8_t* src = q + ((int64_t)img * Nq + qq) * q_ld + head * d + w * 4;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (w * 4 + b < d) word |= (int32_t)(uint8_t)src[b] << (8 * b);
    }
    qs[r * rw + w] = word;
  }
  long long acc0 = 0, acc1 = 0;               // channels lane and lane + 32
  const uint8_t* mrow = mask ? mask + (((int64_t)img * heads + head) * Nq + (q_ok ? qi : 0)) * Nk : nullptr;
  for (int n0 = 0; n0 < Nk; n0 += DA_KEYS) {
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * DA_KEYS * dw; t += blockDim.x) {
      const int which = t / (DA_KEYS * dw), rem = t % (DA_KEYS * dw), r = rem / dw, w = rem % dw;
      int32_t word = 0;
      if (n0 + r < Nk) {
        const int8_t* src = (which ? v : k) + ((int64_t)img * Nk + n0 + r) * kv_ld + head * d + w * 4;
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (w * 4 + b < d) word |= (int32_t)(uint8_t)src[b] << (8 * b);
      }
      (which ? vs : ks)[r * rw + w] = word;
    }
    __syncthreads();
    int score = 0;
    if (q_ok && n0 + lane < Nk) {
      for (int w = 0; w < dw; ++w) score = __dp4a(qs[warp * rw + w], ks[lane * rw + w], score);
      if (mrow && mrow[n0 + lane]) score = 0;                   // masked_fill(mask, 0)
    }
    const int8_t* vb = reinterpret_cast<const int8_t*>(vs);
    for (int j = 0; j < DA_KEYS; ++j) {
      const int s = __shfl_sync(0xffffffffu, score, j);
      if (s != 0) {                                             // warp-uniform
        if (lane < d) acc0 += (long long)s * vb[j * rw * 4 + lane];
        if (lane + 32 < d) acc1 += (long long)s * vb[j * rw * 4 + lane + 32];
      }
    }
  }
  if (!q_ok) return;
  const int64_t o = ((int64_t)img * Nq + qi) * out_ld + head * d;
  if (lane < d) {
    const float y = (float)acc0 * out_scale;
    if (out_f32) out_f32[o + lane] = y;
    if (out_spike) out_spike[o + lane] = (int8_t)spike_level(y, d_max);
  }
  if (lane + 32 < d) {
    const float y = (float)acc1 * out_scale;
    if (out_f32) out_f32[o + lane + 32] = y;
    if (out_spike) out_spike[o + lane + 32] = (int8_t)spike_level(y, d_max);
  }
}

}  // namespace
}  // namespace s2f

using namespace s2f;

extern "C" int s2f_dec_attn(const int8_t* q, const int8_t* k, const int8_t* v, const uint8_t* mask, int8_t* out_spike,
                            float* out_f32, int n, int Nq, int Nk, int heads, int d, int q_ld, int kv_ld, int out_ld,
                            float out_scale, float d_max, void* stream) {
  S2F_REQUIRE(q && k && v && (out_spike || out_f32), "dec_attn: null pointer");
  S2F_REQUIRE(d >= 1 && d <= 64 && heads >= 1 && heads <= 65535 && n >= 0 && n <= 65535, "dec_attn: head dim must be <= 64");
  S2F_REQUIRE(q_ld >= heads * d && kv_ld >= heads * d && out_ld >= heads * d, "dec_attn: row strides smaller than heads*d");
  if (n == 0 || Nq == 0) return S2F_OK;
  const int rw = (d + 3) / 4 + 1;
  const size_t smem = sizeof(int32_t) * (size_t)(2 * DA_KEYS + DA_Q) * rw;
  const dim3 grid((unsigned)ceil_div(Nq, DA_Q), (unsigned)heads, (unsigned)n);
  dec_attn_kernel<<<grid, DA_Q * 32, smem, (cudaStream_t)stream>>>(q, k, v, mask, out_spike, out_f32, Nq, Nk, heads, d, q_ld,
                                                                   kv_ld, out_ld, out_scale, d_max);
  return check_launch("dec_attn_kernel");
}
