// Tensor-pipe peak of this GPU for the MMA kinds the library uses ("int8 peak not measured yet ->
// builder measures it before quoting utilisation").  One CTA per SM, operands resident in shared memory (never
// reloaded), one elected thread issues back-to-back tcgen05.mma M=128 N=256 into two alternating TMEM accumulators:
// nothing but the tensor pipe can be the limit, so OPS = 2*128*256*K_per_mma * mmas / time is the ceiling any kernel
// of that kind can reach.  kind 0: kind::i8 (K = 32 per instruction), kind 1: kind::f16 with bf16 operands (K = 16).
#include "common.cuh"

namespace s2f {
namespace {

__device__ __forceinline__ uint32_t pk_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool pk_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int KIND>
__device__ __forceinline__ void pk_mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// K-major SWIZZLE_128B operand descriptor: 128-byte rows, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t pk_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) peak_mma_kernel(int iters) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (pk_smem(raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + 128 * 128;             // A: 128 rows x 128 B, B: 256 rows x 128 B
  for (uint32_t i = threadIdx.x * 16; i < (128 + 256) * 128; i += blockDim.x * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i), "r"(0x01010101u));
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pk_smem(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pk_smem(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    // D = S32 / F32, A = B = INT8 / BF16, K-major, N = 256, M = 128
    const uint32_t idesc = KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24))
                                     : ((1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24));
    const uint64_t da = pk_desc(a_addr), db = pk_desc(b_addr);
    if (pk_elect()) {
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + (uint32_t)((it & 1) * 256);
#pragma unroll
        for (int k = 0; k < 4; ++k) pk_mma<KIND>(d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it > 1 || k) ? 1u : 0u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pk_smem(&bar)) : "memory");
    }
    __syncwarp();
    asm volatile(
        "{\n\t.reg .pred P1;\n\tPK_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra PK_DONE;\n\tbra PK_WAIT;\n\tPK_DONE:\n\t}"
        ::"r"(pk_smem(&bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace
}  // namespace s2f

// Launches the issue-rate kernel on every SM; returns the number of operations (2 * MACs) one launch performs.
extern "C" int64_t s2f_peak_mma(int kind, int iters, void* stream) {
  using namespace s2f;
  if (kind < 0 || kind > 1 || iters < 2) { fail(S2F_ERR_ARG, "%s", "peak_mma: kind must be 0 (i8) or 1 (bf16), iters >= 2"); return -1; }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = (128 + 256) * 128 + 1024;
  auto fn = kind == 0 ? peak_mma_kernel<0> : peak_mma_kernel<1>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  fn<<<sms, 128, smem, (cudaStream_t)stream>>>(iters);
  if (check_launch("peak_mma_kernel") != S2F_OK) return -1;
  const int64_t k_per_mma = kind == 0 ? 32 : 16;
  return (int64_t)sms * iters * 4 * 2 * 128 * 256 * k_per_mma;
}
