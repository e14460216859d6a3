// Spike-driven linear attention and the head's tail kernels.
#include "common.cuh"

namespace s2f {

// attn_tc.cu: K^T V on tcgen05.mma.kind::i8 (TMA-fed, MN-major operands); the mma.sync kernel below covers unaligned operands
bool kv_tc_eligible(const int8_t* k, const int8_t* v, int heads, int d, int kv_ld);
int kv_tc_launch(const int8_t* k, const int8_t* v, int32_t* kv_ws, int n, int Nk, int heads, int d, int kv_ld, cudaStream_t st);
// Q (K^T V) as a per-image spike GEMM on the tcgen05 kernel of gemm_tc.cu (the workspace holds the digit planes)
int64_t attn_ws_bytes(int n, int heads, int d);
bool qkv_tc_eligible(const int8_t* q, int8_t* out_spike, float* out_f32, int Nq, int Nk, int heads, int d, int q_ld, int out_ld, float d_max);
int qkv_tc_launch(const int8_t* q, int32_t* kv_ws, int8_t* out_spike, float* out_f32, int n, int Nq, int heads, int d, int q_ld,
                  float out_scale, float d_max, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// kv[img, h, i, j] = sum_tok K[img, tok, h*d+i] * V[img, tok, h*d+j]      (exact int32)
// grid (n*heads, splits); each block reduces a token slice and atomically adds its d x d partial.
constexpr int KV_TOK = 64;   // tokens staged per smem tile

__global__ void __launch_bounds__(256) kv_kernel(const int8_t* __restrict__ k, const int8_t* __restrict__ v,
                                                 int32_t* __restrict__ kv, int Nk, int heads, int d, int ld, int tok_per_block) {
  extern __shared__ int8_t smem[];
  int8_t* ks = smem;                     // [KV_TOK][d]
  int8_t* vs = smem + KV_TOK * d;        // [KV_TOK][d]
  const int img = blockIdx.x / heads, h = blockIdx.x % heads;
  const int C = ld;
  const int t_begin = blockIdx.y * tok_per_block;
  const int t_end = min(Nk, t_begin + tok_per_block);
  const int dd = d * d;
  // each thread owns up to 16 (i,j) pairs: pair p = tid + r*256
  int acc[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) acc[r] = 0;
  const int8_t* kb = k + ((int64_t)img * Nk) * C + h * d;
  const int8_t* vb = v + ((int64_t)img * Nk) * C + h * d;
  for (int t0 = t_begin; t0 < t_end; t0 += KV_TOK) {
    const int nt = min(KV_TOK, t_end - t0);
    for (int e = threadIdx.x; e < nt * d; e += blockDim.x) {
      const int tt = e / d, c = e % d;
      ks[e] = kb[(int64_t)(t0 + tt) * C + c];
      vs[e] = vb[(int64_t)(t0 + tt) * C + c];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int p = threadIdx.x + r * 256;
      if (p < dd) {
        const int i = p / d, j = p % d;
        int s = 0;
        for (int tt = 0; tt < nt; ++tt) s += (int)ks[tt * d + i] * (int)vs[tt * d + j];
        acc[r] += s;
      }
    }
    __syncthreads();
  }
  int32_t* out = kv + ((int64_t)img * heads + h) * dd;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int p = threadIdx.x + r * 256;
    if (p < dd && acc[r] != 0) atomicAdd(out + p, acc[r]);
  }
}

// out[img, tok, h*d+j] = (sum_i Q[img,tok,h*d+i] * kv[img,h,i,j]) * out_scale -> NI-LIF
// block = (img, 32-token tile); kv of all heads is staged in smem.
__global__ void __launch_bounds__(256) qkv_kernel(const int8_t* __restrict__ q, const int32_t* __restrict__ kv,
                                                  int8_t* __restrict__ out_spike, float* __restrict__ out_f32, int Nq,
                                                  int heads, int d, int q_ld, int out_ld, float out_scale, float d_max) {
  extern __shared__ int32_t kvs[];       // [heads][d][d]
  const int img = blockIdx.y;
  const int C = heads * d;
  const int tok0 = blockIdx.x * 32;
  const int32_t* kvb = kv + (int64_t)img * heads * d * d;
  for (int e = threadIdx.x; e < heads * d * d; e += blockDim.x) kvs[e] = kvb[e];
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * out_ld; e += blockDim.x) {
    const int tt = e / out_ld, c = e % out_ld;
    const int tok = tok0 + tt;
    if (tok >= Nq) continue;
    if (c >= C) {                                    // channel padding of the output row (TMA needs 16-byte rows)
      const int64_t oz = ((int64_t)img * Nq + tok) * out_ld + c;
      if (out_f32) out_f32[oz] = 0.f;
      if (out_spike) out_spike[oz] = 0;
      continue;
    }
    const int h = c / d, j = c % d;
    const int8_t* qrow = q + ((int64_t)img * Nq + tok) * q_ld + h * d;
    const int32_t* kvh = kvs + h * d * d + j;
    long long s = 0;
    for (int i = 0; i < d; ++i) s += (long long)qrow[i] * (long long)kvh[i * d];
    const float y = (float)s * out_scale;
    const int64_t o = ((int64_t)img * Nq + tok) * out_ld + c;
    if (out_f32) out_f32[o] = y;
    if (out_spike) out_spike[o] = (int8_t)(int)spike_level(y, d_max);
  }
}

// ------------------------------------------------------------------------------------------------
// Fast path (d % 4 == 0, 4-byte aligned rows): K^T V on the int8 tensor-core path.
//   kv[i][j] = sum_tok K[tok][i] V[tok][j] is a GEMM whose reduction axis (tokens) is the slow axis of both operands in
//   memory, so a block first transposes 128-token tiles of K and V into shared memory (4 tokens x 4 channels per
//   thread, a 4x4 byte transpose with PRMT), after which every mma.sync.m16n8k32.s8 fragment is one 32-bit LDS.
//   Each of the 4 warps reduces its own 32 tokens of the tile into the full d x d accumulator; warps are summed
//   through shared memory and one atomicAdd per entry goes to the global workspace.
constexpr int KT_TOK = 128;                    // tokens per staged tile
constexpr int KT_LD = KT_TOK + 16;             // bytes per transposed row (36 words: conflict-free fragment loads)

__device__ __forceinline__ void mma_s8(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MT>      // MT = ceil(d / 16) row tiles; column tiles NT = 2 * MT
__global__ void __launch_bounds__(128) kv_mma_kernel(const int8_t* __restrict__ k, const int8_t* __restrict__ v,
                                                     int32_t* __restrict__ kv, int Nk, int heads, int d, int ld,
                                                     int tok_per_block) {
  constexpr int NT = 2 * MT;
  constexpr int DP = 16 * MT;                  // padded head dim
  __shared__ __align__(16) uint8_t tiles[2 * DP * KT_LD];
  uint8_t* kt = tiles;
  uint8_t* vt = tiles + DP * KT_LD;
  const int img = blockIdx.x / heads, h = blockIdx.x % heads;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int t_begin = blockIdx.y * tok_per_block;
  const int t_end = min(Nk, t_begin + tok_per_block);
  const int8_t* kb = k + ((int64_t)img * Nk) * ld + h * d;
  const int8_t* vb = v + ((int64_t)img * Nk) * ld + h * d;
  // rows d..DP-1 of the transposed tiles are never written by the loader: clear them once
  for (int e = threadIdx.x; e < (DP - d) * KT_LD / 4; e += blockDim.x) {
    reinterpret_cast<uint32_t*>(kt + d * KT_LD)[e] = 0u;
    reinterpret_cast<uint32_t*>(vt + d * KT_LD)[e] = 0u;
  }
  int acc[MT][NT][4];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = acc[mi][ni][2] = acc[mi][ni][3] = 0;
  const int cgn = d >> 2;                      // channel groups of 4
  const int gid = lane >> 2, tig = lane & 3;
  for (int t0 = t_begin; t0 < t_end; t0 += KT_TOK) {
    __syncthreads();                           // previous tile fully consumed
    for (int e = threadIdx.x; e < (KT_TOK / 4) * cgn; e += blockDim.x) {
      const int cg = e % cgn, tg = e / cgn;
      const int tok = t0 + tg * 4;
      uint32_t wk[4], wv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const bool ok = tok + r < t_end;
        wk[r] = ok ? __ldg(reinterpret_cast<const uint32_t*>(kb + (int64_t)(tok + r) * ld + cg * 4)) : 0u;
        wv[r] = ok ? __ldg(reinterpret_cast<const uint32_t*>(vb + (int64_t)(tok + r) * ld + cg * 4)) : 0u;
      }
      auto tr = [&](const uint32_t (&w)[4], uint8_t* dst) {
        const uint32_t a_lo = __byte_perm(w[0], w[1], 0x5140), a_hi = __byte_perm(w[0], w[1], 0x7362);
        const uint32_t b_lo = __byte_perm(w[2], w[3], 0x5140), b_hi = __byte_perm(w[2], w[3], 0x7362);
        uint8_t* o = dst + (cg * 4) * KT_LD + tg * 4;
        *reinterpret_cast<uint32_t*>(o) = __byte_perm(a_lo, b_lo, 0x5410);
        *reinterpret_cast<uint32_t*>(o + KT_LD) = __byte_perm(a_lo, b_lo, 0x7632);
        *reinterpret_cast<uint32_t*>(o + 2 * KT_LD) = __byte_perm(a_hi, b_hi, 0x5410);
        *reinterpret_cast<uint32_t*>(o + 3 * KT_LD) = __byte_perm(a_hi, b_hi, 0x7632);
      };
      tr(wk, kt);
      tr(wv, vt);
    }
    __syncthreads();
    const int tb = warp * 32;                  // this warp's 32 tokens of the tile
    uint32_t bf[NT][2];
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
      const uint8_t* pv = vt + (ni * 8 + gid) * KT_LD + tb + tig * 4;
      bf[ni][0] = *reinterpret_cast<const uint32_t*>(pv);
      bf[ni][1] = *reinterpret_cast<const uint32_t*>(pv + 16);
    }
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
      uint32_t af[4];
      const uint8_t* pk = kt + (mi * 16 + gid) * KT_LD + tb + tig * 4;
      af[0] = *reinterpret_cast<const uint32_t*>(pk);
      af[1] = *reinterpret_cast<const uint32_t*>(pk + 8 * KT_LD);
      af[2] = *reinterpret_cast<const uint32_t*>(pk + 16);
      af[3] = *reinterpret_cast<const uint32_t*>(pk + 8 * KT_LD + 16);
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) mma_s8(acc[mi][ni], af, bf[ni]);
    }
  }
  // cross-warp sum through shared memory (the transposed tiles are dead by now)
  __syncthreads();
  int32_t* red = reinterpret_cast<int32_t*>(tiles);       // [DP][DP] int32: DP*DP*4 bytes <= 2*DP*KT_LD
  static_assert(DP * 4 <= 2 * KT_LD, "reduction buffer must fit in the two transposed tiles");
  for (int e = threadIdx.x; e < DP * DP; e += blockDim.x) red[e] = 0;
  __syncthreads();
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
      const int r0 = mi * 16 + gid, c0 = ni * 8 + tig * 2;
      atomicAdd(&red[r0 * DP + c0], acc[mi][ni][0]);
      atomicAdd(&red[r0 * DP + c0 + 1], acc[mi][ni][1]);
      atomicAdd(&red[(r0 + 8) * DP + c0], acc[mi][ni][2]);
      atomicAdd(&red[(r0 + 8) * DP + c0 + 1], acc[mi][ni][3]);
    }
  __syncthreads();
  int32_t* out = kv + ((int64_t)img * heads + h) * d * d;
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, j = e % d;
    const int val = red[i * DP + j];
    if (val != 0) atomicAdd(out + e, val);
  }
}

// out[img, tok, h*d + j..j+3] for one (token, 4 channels) per thread: q as 32-bit words, kv rows as int4 from smem.
// WIDE = 64-bit accumulation (needed only when 8 * 64 * Nk * d can exceed 2^31).
template <bool WIDE>
__global__ void __launch_bounds__(256) qkv_fast_kernel(const int8_t* __restrict__ q, const int32_t* __restrict__ kv,
                                                       int8_t* __restrict__ out_spike, float* __restrict__ out_f32,
                                                       int Nq, int heads, int d, int q_ld, int out_ld, float out_scale,
                                                       float d_max, int tok_per_block) {
  extern __shared__ __align__(16) int32_t kvs4[];     // [heads][d][d]
  const int img = blockIdx.y;
  const int C = heads * d;
  const int32_t* kvb = kv + (int64_t)img * heads * d * d;
  for (int e = threadIdx.x; e < heads * d * d / 4; e += blockDim.x)
    reinterpret_cast<int4*>(kvs4)[e] = __ldg(reinterpret_cast<const int4*>(kvb) + e);
  __syncthreads();
  const int g4n = out_ld >> 2;                         // 4-channel groups per output row (incl. padding columns)
  const int tok0 = blockIdx.x * tok_per_block;
  const int ntok = min(tok_per_block, Nq - tok0);
  for (int e = threadIdx.x; e < ntok * g4n; e += blockDim.x) {
    const int tok = tok0 + e / g4n, c = (e % g4n) * 4;
    const int64_t o = ((int64_t)img * Nq + tok) * out_ld + c;
    if (c >= C) {
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (out_spike) *reinterpret_cast<uint32_t*>(out_spike + o) = 0u;
      continue;
    }
    const int h = c / d, j = c % d;
    const int8_t* qrow = q + ((int64_t)img * Nq + tok) * q_ld + h * d;
    const int32_t* kvh = kvs4 + h * d * d + j;
    float y[4];
    if (WIDE) {
      long long s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      for (int i = 0; i < d; i += 4) {
        const uint32_t qw = __ldg(reinterpret_cast<const uint32_t*>(qrow + i));
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const long long qi = (int8_t)((qw >> (8 * b)) & 0xff);
          const int4 kr = *reinterpret_cast<const int4*>(kvh + (i + b) * d);
          s0 += qi * kr.x; s1 += qi * kr.y; s2 += qi * kr.z; s3 += qi * kr.w;
        }
      }
      y[0] = (float)s0 * out_scale; y[1] = (float)s1 * out_scale; y[2] = (float)s2 * out_scale; y[3] = (float)s3 * out_scale;
    } else {
      int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      for (int i = 0; i < d; i += 4) {
        const uint32_t qw = __ldg(reinterpret_cast<const uint32_t*>(qrow + i));
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int qi = (int8_t)((qw >> (8 * b)) & 0xff);
          const int4 kr = *reinterpret_cast<const int4*>(kvh + (i + b) * d);
          s0 += qi * kr.x; s1 += qi * kr.y; s2 += qi * kr.z; s3 += qi * kr.w;
        }
      }
      y[0] = (float)s0 * out_scale; y[1] = (float)s1 * out_scale; y[2] = (float)s2 * out_scale; y[3] = (float)s3 * out_scale;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(y[0], y[1], y[2], y[3]);
    if (out_spike)
      *reinterpret_cast<uint32_t*>(out_spike + o) =
          pack_levels4(y[0], y[1], y[2], y[3], d_max);
  }
}

// ------------------------------------------------------------------------------------------------
// Q (K^T V) on the int8 tensor-core path as well.  kv = K^T V is an int32 matrix (up to Nk * 64 per entry), so it is
// split -- exactly like a weight matrix of the spike GEMM -- into three signed base-128 digit planes in shared memory
// (laid out [head][plane][j][i], i contiguous: the "col" B operand of mma.sync.m16n8k32), after which
//   out[tok][h*d + j] = sum_p 128^(2-p) * (Q_h digits_p)[tok][j]
// is three s8 MMAs per 16 tokens x 8 channels and an exact int32 merge.  A warp owns 16 tokens and walks the heads;
// its A fragments are four 32-bit loads straight from the q rows.  ~1300 instructions per 16 tokens x 256 channels
// against ~6400 for the IMAD loop of qkv_fast_kernel.
template <int KB>      // 32-wide K blocks: ceil(d / 32)
__global__ void __launch_bounds__(256) qkv_mma_kernel(const int8_t* __restrict__ q, const int32_t* __restrict__ kv,
                                                      int8_t* __restrict__ out_spike, float* __restrict__ out_f32, int Nq,
                                                      int heads, int d, int q_ld, int out_ld, float out_scale, float d_max,
                                                      int hpb) {
  extern __shared__ __align__(16) uint8_t planes[];          // [hpb heads of this block][3][d][KPAD]
  constexpr int KPAD = 32 * KB + 16;                         // conflict-free fragment loads (12 / 20 words per row)
  const int img = blockIdx.y;
  const int C = heads * d;
  const int h_first = blockIdx.z * hpb;                      // blockIdx.z: group of hpb heads (small Nq: more blocks)
  const int32_t* kvb = kv + ((int64_t)img * heads + h_first) * d * d;
  for (int e = threadIdx.x; e < hpb * d * d; e += blockDim.x) {
    const int h = e / (d * d), r = e % (d * d), i = r / d, j = r % d;
    int v = __ldg(kvb + e);
    const int d2 = ((v + 64) & 127) - 64; v = (v - d2) >> 7;
    const int d1 = ((v + 64) & 127) - 64; v = (v - d1) >> 7;
    uint8_t* dst = planes + ((size_t)(h * 3) * d + j) * KPAD + i;
    dst[0] = (uint8_t)(int8_t)v; dst[(size_t)d * KPAD] = (uint8_t)(int8_t)d1; dst[(size_t)2 * d * KPAD] = (uint8_t)(int8_t)d2;
  }
  __syncthreads();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int tok0 = blockIdx.x * 128 + warp * 16;
  if (tok0 >= Nq) return;
  const int r0 = tok0 + g, r1 = tok0 + g + 8;
  const bool ok0 = r0 < Nq, ok1 = r1 < Nq;
  const int8_t* q0 = q + ((int64_t)img * Nq + (ok0 ? r0 : 0)) * q_ld;
  const int8_t* q1 = q + ((int64_t)img * Nq + (ok1 ? r1 : 0)) * q_ld;
  const int64_t o0 = ((int64_t)img * Nq + r0) * out_ld, o1 = ((int64_t)img * Nq + r1) * out_ld;
  const int ntiles = d >> 3;
  for (int hl = 0; hl < hpb; ++hl) {
    const int h = h_first + hl;
    uint32_t a[KB][4];
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      const int k0 = kb * 32 + 4 * t, k1 = k0 + 16;
      a[kb][0] = (ok0 && k0 < d) ? __ldg(reinterpret_cast<const uint32_t*>(q0 + h * d + k0)) : 0u;
      a[kb][1] = (ok1 && k0 < d) ? __ldg(reinterpret_cast<const uint32_t*>(q1 + h * d + k0)) : 0u;
      a[kb][2] = (ok0 && k1 < d) ? __ldg(reinterpret_cast<const uint32_t*>(q0 + h * d + k1)) : 0u;
      a[kb][3] = (ok1 && k1 < d) ? __ldg(reinterpret_cast<const uint32_t*>(q1 + h * d + k1)) : 0u;
    }
    const uint8_t* ph = planes + (size_t)(hl * 3) * d * KPAD;
    for (int nt = 0; nt < ntiles; ++nt) {
      int acc[3][4];
#pragma unroll
      for (int pz = 0; pz < 3; ++pz) {
        acc[pz][0] = acc[pz][1] = acc[pz][2] = acc[pz][3] = 0;
        const uint8_t* pb = ph + ((size_t)pz * d + nt * 8 + g) * KPAD + 4 * t;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t b[2] = {*reinterpret_cast<const uint32_t*>(pb + kb * 32), *reinterpret_cast<const uint32_t*>(pb + kb * 32 + 16)};
          mma_s8(acc[pz], a[kb], b);
        }
      }
      float y[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) y[r] = (float)((acc[0][r] * 128 + acc[1][r]) * 128 + acc[2][r]) * out_scale;
      const int c = h * d + nt * 8 + 2 * t;
      if (ok0) {
        if (out_f32) *reinterpret_cast<float2*>(out_f32 + o0 + c) = make_float2(y[0], y[1]);
        if (out_spike) *reinterpret_cast<uint16_t*>(out_spike + o0 + c) =
            (uint16_t)((level_bits(y[0], d_max) & 0xffu) | ((level_bits(y[1], d_max) & 0xffu) << 8));
      }
      if (ok1) {
        if (out_f32) *reinterpret_cast<float2*>(out_f32 + o1 + c) = make_float2(y[2], y[3]);
        if (out_spike) *reinterpret_cast<uint16_t*>(out_spike + o1 + c) =
            (uint16_t)((level_bits(y[2], d_max) & 0xffu) | ((level_bits(y[3], d_max) & 0xffu) << 8));
      }
    }
  }
  // channel padding of the output rows (TMA needs 16-byte rows): zeros
  if (blockIdx.z == 0)
  for (int e = lane; e < 16 * (out_ld - C); e += 32) {
    const int rr = tok0 + e / (out_ld - C), c = C + e % (out_ld - C);
    if (rr < Nq) {
      const int64_t o = ((int64_t)img * Nq + rr) * out_ld + c;
      if (out_f32) out_f32[o] = 0.f;
      if (out_spike) out_spike[o] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sigmoid_lif_kernel(const float* __restrict__ x, int8_t* __restrict__ levels,
                                                          int64_t N, float d_max) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    // sigmoid(x) in (0,1): the level is 1 iff the fp32 sigmoid exceeds 0.5 (sigmoid(0) = 0.5 ties to 0)
    const float s = 1.f / (1.f + expf(-x[i]));
    levels[i] = (int8_t)(int)spike_level(s, d_max);
  }
}

// softmax over K+1 classes, drop the last ("no object") column: prob [n*Q, K]
__global__ void __launch_bounds__(128) softmax_drop_kernel(const float* __restrict__ cls, float* __restrict__ prob,
                                                           int rows, int K1) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  const float* c = cls + (int64_t)row * K1;
  __shared__ float red[128];
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < K1; i += blockDim.x) mx = fmaxf(mx, c[i]);
  red[threadIdx.x] = mx; __syncthreads();
  for (int s = 64; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
  mx = red[0]; __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < K1; i += blockDim.x) sum += expf(c[i] - mx);
  red[threadIdx.x] = sum; __syncthreads();
  for (int s = 64; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
  sum = red[0];
  for (int i = threadIdx.x; i < K1 - 1; i += blockDim.x) prob[(int64_t)row * (K1 - 1) + i] = expf(c[i] - mx) / sum;
}

// logits[n, c, Y, X] = sum_q prob[n,q,c] * sigmoid(bilinear(mask_pred[n, :, :, q])(Y, X))
// block = 64 consecutive output pixels of one row segment; stage sigmoid(up) [64][Q] and prob [Q][K] in smem.
constexpr int TAIL_PIX = 64;
__global__ void __launch_bounds__(256) semantic_tail_kernel(const float* __restrict__ mask_pred,
                                                            const float* __restrict__ prob, float* __restrict__ logits,
                                                            int Q, int K, int h, int w, int H, int W) {
  extern __shared__ float sm[];
  float* S = sm;                          // [Q][TAIL_PIX + 1]
  float* Pm = sm + Q * (TAIL_PIX + 1);    // [Q][K]
  const int img = blockIdx.y;
  const int64_t pix0 = (int64_t)blockIdx.x * TAIL_PIX;
  const int64_t HW = (int64_t)H * W;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  const float* mp = mask_pred + (int64_t)img * h * w * Q;
  for (int e = threadIdx.x; e < Q * K; e += blockDim.x) Pm[e] = prob[(int64_t)img * Q * K + e];
  for (int e = threadIdx.x; e < TAIL_PIX * Q; e += blockDim.x) {
    const int pp = e / Q, q = e % Q;
    const int64_t pix = pix0 + pp;
    float val = 0.f;
    if (pix < HW) {
      const int Y = (int)(pix / W), X = (int)(pix % W);
      float sy = sh * ((float)Y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
      float sx = sw * ((float)X + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
      const int y0 = (int)sy, x0 = (int)sx;
      const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
      const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
      const float p00 = mp[((int64_t)y0 * w + x0) * Q + q], p01 = mp[((int64_t)y0 * w + x1) * Q + q];
      const float p10 = mp[((int64_t)y1 * w + x0) * Q + q], p11 = mp[((int64_t)y1 * w + x1) * Q + q];
      const float up = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
      val = 1.f / (1.f + expf(-up));
    }
    S[q * (TAIL_PIX + 1) + pp] = val;
  }
  __syncthreads();
  // thread -> pixel (fastest) x class strided
  const int pp = threadIdx.x % TAIL_PIX;
  const int cg = threadIdx.x / TAIL_PIX;          // 0..3
  const int64_t pix = pix0 + pp;
  if (pix >= HW) return;
  for (int c = cg; c < K; c += 256 / TAIL_PIX) {
    float acc = 0.f;
    for (int q = 0; q < Q; ++q) acc = fmaf(Pm[q * K + c], S[q * (TAIL_PIX + 1) + pp], acc);
    logits[((int64_t)img * K + c) * HW + pix] = acc;
  }
}

}  // namespace s2f

using namespace s2f;

extern "C" int64_t s2f_linear_attn_ws_bytes(int n, int heads, int d) {
  if (n < 1 || heads < 1 || d < 1) return -1;
  return attn_ws_bytes(n, heads, d);
}

extern "C" int s2f_linear_attn(const int8_t* q, const int8_t* k, const int8_t* v, int32_t* kv_ws, int8_t* out_spike,
                               float* out_f32, int n, int Nq, int Nk, int heads, int d, int q_ld, int kv_ld,
                               int out_ld, float out_scale, float d_max, void* stream) {
  S2F_REQUIRE(q && k && v && kv_ws && (out_spike || out_f32), "linear_attn: null pointer");
  S2F_REQUIRE(d >= 1 && d <= 64 && heads >= 1, "linear_attn: head dim must be <= 64");
  S2F_REQUIRE(q_ld >= heads * d && kv_ld >= heads * d && out_ld >= heads * d, "linear_attn: row strides smaller than heads*d");
  S2F_REQUIRE((int64_t)64 * Nk < (1ll << 31), "linear_attn: Nk too large for int32 K^T V");
  cudaStream_t st = (cudaStream_t)stream;
  const int dd = d * d;
  cudaError_t e = cudaMemsetAsync(kv_ws, 0, sizeof(int32_t) * (size_t)n * heads * dd, st);
  if (e != cudaSuccess) return fail(S2F_ERR_CUDA, "linear_attn memset: %s", cudaGetErrorString(e));
  auto al4 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3) == 0; };
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool fast_kv = d % 4 == 0 && kv_ld % 4 == 0 && al4(k) && al4(v);
  int rc;
  if (kv_tc_eligible(k, v, heads, d, kv_ld)) {
    rc = kv_tc_launch(k, v, kv_ws, n, Nk, heads, d, kv_ld, st);
  } else if (fast_kv) {
    // token slices of >= 256 tokens; ~6 of these 4-warp blocks per SM keep enough loads in flight
    int splits = (int)ceil_div(148 * 6, (int64_t)n * heads);
    const int max_splits = (int)ceil_div(Nk, 256);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int tok_per_block = (int)ceil_div(ceil_div(Nk, splits), KT_TOK) * KT_TOK;
    splits = (int)ceil_div(Nk, tok_per_block);
    const dim3 grid(n * heads, splits);
    if (d <= 16) kv_mma_kernel<1><<<grid, 128, 0, st>>>(k, v, kv_ws, Nk, heads, d, kv_ld, tok_per_block);
    else if (d <= 32) kv_mma_kernel<2><<<grid, 128, 0, st>>>(k, v, kv_ws, Nk, heads, d, kv_ld, tok_per_block);
    else if (d <= 48) kv_mma_kernel<3><<<grid, 128, 0, st>>>(k, v, kv_ws, Nk, heads, d, kv_ld, tok_per_block);
    else kv_mma_kernel<4><<<grid, 128, 0, st>>>(k, v, kv_ws, Nk, heads, d, kv_ld, tok_per_block);
    rc = check_launch("kv_mma_kernel");
  } else {
    // enough token slices to fill the GPU, at least 256 tokens each
    int splits = (int)ceil_div(148 * 4, (int64_t)n * heads);
    const int max_splits = (int)ceil_div(Nk, 256);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int tok_per_block = (int)ceil_div(ceil_div(Nk, splits), KV_TOK) * KV_TOK;
    splits = (int)ceil_div(Nk, tok_per_block);
    kv_kernel<<<dim3(n * heads, splits), 256, 2 * KV_TOK * d, st>>>(k, v, kv_ws, Nk, heads, d, kv_ld, tok_per_block);
    rc = check_launch("kv_kernel");
  }
  if (rc) return rc;
  if (qkv_tc_eligible(q, out_spike, out_f32, Nq, Nk, heads, d, q_ld, out_ld, d_max))
    return qkv_tc_launch(q, kv_ws, out_spike, out_f32, n, Nq, heads, d, q_ld, out_scale, d_max, st);
  const size_t sm = sizeof(int32_t) * (size_t)heads * dd;
  S2F_REQUIRE(sm <= 200 * 1024, "linear_attn: heads*d*d too large for shared memory");
  const bool fast_q = d % 4 == 0 && q_ld % 4 == 0 && out_ld % 4 == 0 && al4(q) && al16(kv_ws) && (!out_f32 || al16(out_f32)) &&
                      (!out_spike || al4(out_spike));
  const bool narrow = (double)8 * 64 * (double)Nk * d < 2147483648.0 && (double)Nk * 64 < 127.0 * 16384.0;
  if (fast_q && narrow && d % 8 == 0 && (out_ld & 1) == 0 && (!out_spike || (reinterpret_cast<uintptr_t>(out_spike) & 1) == 0) &&
      (!out_f32 || (reinterpret_cast<uintptr_t>(out_f32) & 7) == 0)) {
    const int KB = (d + 31) / 32;
    // heads per block: all of them when the token tiles alone fill the GPU, fewer (a divisor of heads) otherwise
    const int64_t tok_blocks = ceil_div(Nq, 128) * n;
    int hpb = heads;
    while (hpb > 1 && hpb % 2 == 0 && tok_blocks * (heads / hpb) < 148 * 2) hpb /= 2;
    const size_t smq = (size_t)hpb * 3 * d * (32 * KB + 16);
    if (smq <= 200 * 1024) {
      const dim3 grid((unsigned)ceil_div(Nq, 128), n, (unsigned)(heads / hpb));
      if (KB == 1) {
        if (smq > 48 * 1024) cudaFuncSetAttribute(qkv_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smq);
        qkv_mma_kernel<1><<<grid, 256, smq, st>>>(q, kv_ws, out_spike, out_f32, Nq, heads, d, q_ld, out_ld, out_scale, d_max, hpb);
      } else {
        if (smq > 48 * 1024) cudaFuncSetAttribute(qkv_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smq);
        qkv_mma_kernel<2><<<grid, 256, smq, st>>>(q, kv_ws, out_spike, out_f32, Nq, heads, d, q_ld, out_ld, out_scale, d_max, hpb);
      }
      return check_launch("qkv_mma_kernel");
    }
  }
  if (fast_q) {
    const bool wide = (double)8 * 64 * (double)Nk * d >= 2147483648.0;
    // tokens per block: amortise the kv staging, keep >= ~2 blocks per SM
    int tpb = (int)ceil_div((int64_t)Nq * n, 148 * 2);
    if (tpb < 8) tpb = 8;
    if (tpb > 64) tpb = 64;
    const dim3 grid((unsigned)ceil_div(Nq, tpb), n);
    if (wide) {
      if (sm > 48 * 1024) cudaFuncSetAttribute(qkv_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      qkv_fast_kernel<true><<<grid, 256, sm, st>>>(q, kv_ws, out_spike, out_f32, Nq, heads, d, q_ld, out_ld, out_scale, d_max, tpb);
    } else {
      if (sm > 48 * 1024) cudaFuncSetAttribute(qkv_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      qkv_fast_kernel<false><<<grid, 256, sm, st>>>(q, kv_ws, out_spike, out_f32, Nq, heads, d, q_ld, out_ld, out_scale, d_max, tpb);
    }
    return check_launch("qkv_fast_kernel");
  }
  if (sm > 48 * 1024) cudaFuncSetAttribute(qkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  qkv_kernel<<<dim3((unsigned)ceil_div(Nq, 32), n), 256, sm, st>>>(q, kv_ws, out_spike, out_f32, Nq, heads, d, q_ld,
                                                                   out_ld, out_scale, d_max);
  return check_launch("qkv_kernel");
}

extern "C" int s2f_sigmoid_lif(const float* x, int8_t* levels, int64_t N, float d_max, void* stream) {
  S2F_REQUIRE(x && levels, "sigmoid_lif: null pointer");
  if (N == 0) return S2F_OK;
  const int64_t want = ceil_div(N, 256);
  sigmoid_lif_kernel<<<(int)(want < 2368 ? want : 2368), 256, 0, (cudaStream_t)stream>>>(x, levels, N, d_max);
  return check_launch("sigmoid_lif_kernel");
}

extern "C" int s2f_semantic_tail(const float* mask_pred, const float* cls, float* logits, float* prob_ws, int n, int Q,
                                 int K, int h, int w, int H, int W, void* stream) {
  S2F_REQUIRE(mask_pred && cls && logits && prob_ws, "semantic_tail: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  softmax_drop_kernel<<<n * Q, 128, 0, st>>>(cls, prob_ws, n * Q, K + 1);
  int rc = check_launch("softmax_drop_kernel");
  if (rc) return rc;
  const size_t sm = sizeof(float) * ((size_t)Q * (TAIL_PIX + 1) + (size_t)Q * K);
  S2F_REQUIRE(sm <= 200 * 1024, "semantic_tail: Q*K too large for shared memory");
  if (sm > 48 * 1024) cudaFuncSetAttribute(semantic_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  const int64_t HW = (int64_t)H * W;
  semantic_tail_kernel<<<dim3((unsigned)ceil_div(HW, TAIL_PIX), n), 256, sm, st>>>(mask_pred, prob_ws, logits, Q, K, h, w,
                                                                                   H, W);
  return check_launch("semantic_tail_kernel");
}
