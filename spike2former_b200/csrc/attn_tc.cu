// K^T V of the spike-driven (linear) attention on tcgen05.mma.kind::i8 -- the one contraction of the path whose two
// operands are both spikes (sdtv2.py:335-336, mmcv_spike/transformer.py:262-270, 345-353):
//   kv[img, h, i, j] = sum_tok K[img, tok, h*d + i] * V[img, tok, h*d + j]            exact int32
// The contraction runs over TOKENS, the outer dimension of the channels-last operands, so both MMA operands are
// MN-major: a TMA box of 128 tokens x 128 channels (SWIZZLE_128B) is, as it lands in shared memory, the canonical
// MN-major image ((16,8,m),(8,k)) of a 128-wide operand -- no transpose pass (the mma.sync version transposed every tile
// through shared memory with CUDA cores), and one tcgen05.mma (M = 128 value channels, N = 128 or 256 key channels,
// K = 32 tokens) does the work of 4 heads at once; the products between different heads are computed and dropped (the
// tensor core has 8x headroom here, the kernel is bound by streaming K and V once from HBM).
//   CTA = (128-channel slab of V, token split, image): warp 0 issues TMA, warp 1 the MMAs (4 per 128-token stage, 4-stage
//   mbarrier ring), warps 2..5 read the int32 accumulator (lane = value channel j) and add each head's d x d block to
//   kv_ws with coalesced red.add (consecutive lanes = consecutive j).
// Head widths that do not divide 128 (d = 48: stage 4 of the backbone) use a 256-wide key box that starts at the first
// head the slab touches.
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace s2f {

constexpr int KT_STAGES = 4;
constexpr int KT_TOK = 128;                 // tokens per stage
constexpr int KT_THREADS = 192;             // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

struct KvTcP {
  int32_t* kv;
  int Nk, heads, d, C, nw;                  // nw: key-channel columns of the accumulator (128 or 256)
  int tiles, tiles_per_split;
};

__device__ __forceinline__ uint32_t kt_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kt_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "KT_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra KT_DONE;\n\t"
      "bra KT_WAIT;\n\t"
      "KT_DONE:\n\t"
      "}" ::"r"(kt_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool kt_elect() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
// MN-major SWIZZLE_128B operand: 128 B of MN per row, 8 K rows per 1024-byte group (SBO), 128-wide atoms LBO apart
__device__ __forceinline__ uint64_t kt_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

__global__ void __launch_bounds__(KT_THREADS, 1)
kv_tc_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, const KvTcP p) {
  extern __shared__ __align__(1024) uint8_t kt_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(kt_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = KT_TOK * 128, b_bytes = KT_TOK * p.nw, stage_bytes = a_bytes + b_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + KT_STAGES * stage_bytes);
  uint64_t* empty = full + KT_STAGES;
  uint64_t* acc_full = empty + KT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int slab = blockIdx.x, split = blockIdx.y, img = blockIdx.z;
  const int t_begin = split * p.tiles_per_split, t_end = min(p.tiles, t_begin + p.tiles_per_split);
  const int j0 = slab * 128;                                     // first value channel of the slab
  const int i0 = p.nw == 128 ? j0 : (j0 / p.d) * p.d;            // first key channel of the accumulator columns
  if (threadIdx.x == 0) {
    for (int s = 0; s < KT_STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kt_u32(full + s)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kt_u32(empty + s)), "r"(1));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kt_u32(acc_full)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (p.nw == 128) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(kt_u32(tmem_slot)), "n"(128));
    else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(kt_u32(tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int s = it % KT_STAGES;
      if (it >= KT_STAGES) kt_wait(empty + s, (uint32_t)((it / KT_STAGES - 1) & 1));
      if (kt_elect()) {
        uint8_t* sa = smem + s * stage_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kt_u32(full + s)), "r"((uint32_t)stage_bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(kt_u32(sa)), "l"(&map_v), "r"(kt_u32(full + s)), "r"(j0), "r"(t * KT_TOK), "r"(img) : "memory");
        for (int b = 0; b < p.nw / 128; ++b)
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
              ::"r"(kt_u32(sa + a_bytes + b * a_bytes)), "l"(&map_k), "r"(kt_u32(full + s)), "r"(i0 + b * 128), "r"(t * KT_TOK), "r"(img)
              : "memory");
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issue
    // D[j, i] (+)= sum_tok V[tok, j] K[tok, i]: A = V tile, B = K tile, both MN-major; S32 accumulate, S8 x S8
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.nw >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int s = it % KT_STAGES;
      kt_wait(full + s, (uint32_t)((it / KT_STAGES) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (kt_elect()) {
        const uint32_t sa = kt_u32(smem + s * stage_bytes), sb = sa + a_bytes;
#pragma unroll
        for (int k = 0; k < KT_TOK / 32; ++k) {
          const uint64_t da = kt_desc(sa + k * 4096, 0), db = kt_desc(sb + k * 4096, (uint32_t)a_bytes);
          const uint32_t accum = (it | k) ? 1u : 0u;
          asm volatile(
              "{\n\t"
              ".reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
              "}" ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(kt_u32(empty + s)) : "memory");
        if (t == t_end - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(kt_u32(acc_full)) : "memory");
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue: lane = value channel j
    const int quad = warp & 3;
    const int jc = j0 + quad * 32 + lane;                        // this lane's value channel
    const bool j_ok = jc < p.C;
    const int h = min(jc, p.C - 1) / p.d, jj = jc - h * p.d;
    // columns that hold this warp's heads: from the first lane's head to the last lane's head
    const int jw0 = j0 + quad * 32, jw1 = min(jw0 + 31, p.C - 1);
    const int c_lo = (jw0 / p.d) * p.d - i0, c_hi = min((jw1 / p.d + 1) * p.d, p.C) - i0;     // accumulator columns [c_lo, c_hi)
    if (t_end > t_begin && jw0 < p.C) {
      kt_wait(acc_full, 0u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      int32_t* out = p.kv + ((int64_t)img * p.heads + h) * p.d * p.d + jj;
      const int ih0 = h * p.d - i0;                              // this lane's head starts at accumulator column ih0
      for (int c = c_lo & ~15; c < c_hi; c += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int ii = c + e - ih0;                            // key channel within this lane's head
          const int val = (int)r[e];
          if (j_ok && ii >= 0 && ii < p.d && val != 0) atomicAdd(out + (int64_t)ii * p.d, val);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (p.nw == 128) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

typedef CUresult (*KtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static KtEncodeFn kt_encode_fn() {
  static KtEncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<KtEncodeFn>(ptr);
  }
  return fn;
}

// Can the tensor-core K^T V take these operands?  (TMA: 16-byte aligned base and row stride; head width <= 64.)
bool kv_tc_eligible(const int8_t* k, const int8_t* v, int heads, int d, int kv_ld) {
  static const bool off = []() { const char* e = getenv("S2F_ATTN_TC"); return e && e[0] == '0'; }();
  if (off) return false;
  return d >= 8 && d <= 64 && heads * d >= 16 && kv_ld % 16 == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(v) & 15) == 0;
}

// kv_ws must be zero on entry (the token splits and the heads of a slab add into it).
int kv_tc_launch(const int8_t* k, const int8_t* v, int32_t* kv_ws, int n, int Nk, int heads, int d, int kv_ld, cudaStream_t st) {
  KtEncodeFn enc = kt_encode_fn();
  if (!enc) return fail(S2F_ERR_CUDA, "linear_attn: %s", "cuTensorMapEncodeTiled entry point not found");
  const int C = heads * d;
  KvTcP p;
  p.kv = kv_ws; p.Nk = Nk; p.heads = heads; p.d = d; p.C = C;
  p.nw = (128 % d == 0) ? 128 : 256;
  p.tiles = (Nk + KT_TOK - 1) / KT_TOK;
  const int slabs = (C + 127) / 128;
  int splits = (2 * sm_count()) / (slabs * n);
  if (splits > p.tiles / 4) splits = p.tiles / 4;              // at least four 128-token stages per CTA
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.tiles + splits - 1) / splits;
  splits = (p.tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  CUtensorMap mk, mv;
  const int8_t* ptrs[2] = {k, v};
  CUtensorMap* maps[2] = {&mk, &mv};
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Nk, (cuuint64_t)n};
    cuuint64_t strides[2] = {(cuuint64_t)kv_ld, (cuuint64_t)Nk * kv_ld};
    cuuint32_t box[3] = {128, (cuuint32_t)KT_TOK, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(maps[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(ptrs[i]), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(S2F_ERR_CUDA, "linear_attn: cuTensorMapEncodeTiled(k/v) failed (%s) code %lld", "", (long long)r);
  }
  const size_t smem = 1024 + (size_t)KT_STAGES * (KT_TOK * 128 + KT_TOK * p.nw) + (2 * KT_STAGES + 1) * 8 + 16;
  static std::atomic<uint64_t> once{0};
  if (first_use_on_this_device(once)) cudaFuncSetAttribute(kv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  kv_tc_kernel<<<dim3(slabs, splits, n), KT_THREADS, smem, st>>>(mk, mv, p);
  return check_launch("kv_tc_kernel");
}


// ---------------------------------------------------------------------------------------------------------------------
// Q (K^T V) as a spike GEMM with one weight matrix per image (gemm_tc.cu, per_image_weights = 2): the matrix is the
// block-diagonal K^T V,  W[j][i] = kv[h][i - h d][j - h d]  for i, j in the same head h, 0 elsewhere.  kv is a non-negative
// integer below 2^21 (Nk * 64), so three digits 0..127 represent it exactly; they are written straight in the kernel's
// packed tile layout (rows (tile, plane, 64 channels), K-major, plane 0 most significant).  One thread = 16 K bytes.
__host__ __device__ inline int kt_kpad(int C) { const int bk = C >= 128 ? 128 : (C >= 64 ? 64 : 32); return (C + bk - 1) / bk * bk; }

__global__ void __launch_bounds__(256) kv_pack_kernel(const int32_t* __restrict__ kv, int8_t* __restrict__ packed, float* __restrict__ scale,
                                                      float* __restrict__ shift, int n, int heads, int d, int kpad, int tiles_n,
                                                      float out_scale) {
  const int C = heads * d;
  const int chunks = kpad >> 4;
  const int64_t total = (int64_t)n * tiles_n * 3 * 64 * chunks;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid < C) { scale[gid] = out_scale; shift[gid] = 0.f; }
  if (gid >= total) return;
  const int ck = (int)(gid % chunks);
  int64_t row = gid / chunks;
  const int r = (int)(row & 63); row >>= 6;
  const int pl = (int)(row % 3); row /= 3;
  const int t = (int)(row % tiles_n), img = (int)(row / tiles_n);
  const int j = t * 64 + r;
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  if (j < C) {
    const int h = j / d, jj = j - h * d;
    const int32_t* src = kv + ((int64_t)img * heads + h) * d * d + jj;
    const int sh_ = pl == 0 ? 14 : (pl == 1 ? 7 : 0);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int ii = ck * 16 + e - h * d;
      if (ii >= 0 && ii < d) w[e >> 2] |= (uint32_t)((__ldg(src + (int64_t)ii * d) >> sh_) & 127) << (8 * (e & 3));
    }
  }
  reinterpret_cast<uint4*>(packed)[gid] = make_uint4(w[0], w[1], w[2], w[3]);
}

// workspace: [kv int32 n*heads*d*d][pad to 256][packed planes][scale C][shift C]
static inline int64_t kt_align(int64_t x) { return (x + 255) & ~int64_t(255); }
int64_t attn_ws_bytes(int n, int heads, int d) {
  const int C = heads * d;
  const int64_t kvb = kt_align((int64_t)n * heads * d * d * 4);
  const int64_t pk = kt_align((int64_t)n * ((C + 63) / 64) * 192 * kt_kpad(C));
  return kvb + pk + kt_align(2 * (int64_t)C * 4);
}

bool qkv_tc_eligible(const int8_t* q, int8_t* out_spike, float* out_f32, int Nq, int Nk, int heads, int d, int q_ld, int out_ld,
                     float d_max) {
  static const bool off = []() { const char* e = getenv("S2F_ATTN_TC"); return e && (e[0] == '0' || e[0] == '1'); }();   // 1: K^T V only
  if (off) return false;
  const int C = heads * d;
  return Nq % 128 == 0 && C >= 32 && C % 16 == 0 && out_ld == C && q_ld % 16 == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0 &&
         (int64_t)Nk * 64 < (1ll << 21) && d_max == 8.f && (!out_spike || (reinterpret_cast<uintptr_t>(out_spike) & 15) == 0) &&
         (!out_f32 || (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0);
}

int qkv_tc_launch(const int8_t* q, int32_t* kv_ws, int8_t* out_spike, float* out_f32, int n, int Nq, int heads, int d, int q_ld,
                  float out_scale, float d_max, cudaStream_t st) {
  const int C = heads * d, kpad = kt_kpad(C), tiles_n = (C + 63) / 64;
  uint8_t* base = reinterpret_cast<uint8_t*>(kv_ws);
  int8_t* packed = reinterpret_cast<int8_t*>(base + kt_align((int64_t)n * heads * d * d * 4));
  float* scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(packed) + kt_align((int64_t)n * tiles_n * 192 * kpad));
  float* shift = scale + C;
  const int64_t total = (int64_t)n * tiles_n * 192 * (kpad >> 4);
  kv_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(kv_ws, packed, scale, shift, n, heads, d, kpad, tiles_n, out_scale);
  int rc = check_launch("kv_pack_kernel");
  if (rc) return rc;
  s2f_gemm_tc_args a;
  memset(&a, 0, sizeof(a));
  a.a = q; a.w_packed = packed; a.scale = scale; a.shift = shift; a.out_f32 = out_f32; a.out_spike = out_spike;
  a.n = n; a.H = Nq; a.W = 1; a.Cin = C; a.Cout = C; a.KH = a.KW = 1; a.stride = 1; a.pad = 0; a.pieces = 3; a.d_max = d_max;
  a.per_image_weights = 2; a.a_ld = q_ld;
  return s2f_gemm_i8_tc(&a, st);
}

}  // namespace s2f
