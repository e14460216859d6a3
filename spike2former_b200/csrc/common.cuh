// Shared helpers for libs2f.so (sm_100a).  No torch headers: the library is a plain C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/s2f.h"

namespace s2f {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b, c);
  return code;
}

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return S2F_ERR_CUDA;
  }
  return S2F_OK;
}

#define S2F_REQUIRE(cond, msg)                                          \
  do {                                                                  \
    if (!(cond)) return s2f::fail(S2F_ERR_ARG, "%s (%s)", msg, 0, 0);   \
  } while (0)

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute and the SM count are PER DEVICE: a process that runs the model on a second GPU must set the
// kernels' shared-memory attribute there too.  `once` is a per-call-site bit mask indexed by the ordinal.
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
inline bool first_use_on_this_device(std::atomic<uint64_t>& once) {
  const uint64_t bit = 1ull << (current_device() & 63);
  return (once.fetch_or(bit, std::memory_order_relaxed) & bit) == 0;
}
inline int sm_count() {
  static std::atomic<int> cached[64];
  const int dev = current_device();
  int n = cached[dev & 63].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

// round-half-to-even of clamp(v, 0, d): torch.round(torch.clamp(v, 0, d)) -- surrogate.py:529
__device__ __forceinline__ float spike_level(float v, float d_max) { return rintf(fminf(fmaxf(v, 0.f), d_max)); }

// The same level as the low byte of a word, without the XU-pipe FRND / F2I instructions: for 0 <= v <= d_max < 2^22,
// v + 2^23 rounds (to nearest, ties to even -- the default fp32 add) to an integer whose value sits in the low mantissa
// bits of 0x4B000000 | level.  pack_levels4 assembles four of them into one 32-bit word with three PRMT.
__device__ __forceinline__ uint32_t level_bits(float v, float d_max) {
  return __float_as_uint(fminf(fmaxf(v, 0.f), d_max) + 8388608.f);
}
// Same for d_max = 8 (every Spike2Former config): clamp by the saturating multiply (FMUL.SAT clamps to [0, 1]; the
// power-of-two scalings are exact), then one FFMA puts the rounded level into the mantissa: 2 instructions, not 3.
__device__ __forceinline__ uint32_t level_bits8(float v) {
  return __float_as_uint(fmaf(__saturatef(v * 0.125f), 8.f, 8388608.f));
}
__device__ __forceinline__ uint32_t pack_levels4_d8(float a, float b, float c, float d) {
  const uint32_t lo = __byte_perm(level_bits8(a), level_bits8(b), 0x0040);
  const uint32_t hi = __byte_perm(level_bits8(c), level_bits8(d), 0x0040);
  return __byte_perm(lo, hi, 0x5410);
}
__device__ __forceinline__ uint32_t pack_levels4(float a, float b, float c, float d, float d_max) {
  if (d_max == 8.f) return pack_levels4_d8(a, b, c, d);          // uniform branch
  const uint32_t lo = __byte_perm(level_bits(a, d_max), level_bits(b, d_max), 0x0040);
  const uint32_t hi = __byte_perm(level_bits(c, d_max), level_bits(d, d_max), 0x0040);
  return __byte_perm(lo, hi, 0x5410);
}

__device__ __forceinline__ bool is_tie(float v, float d_max) {
  return (v > 0.f) && (v < d_max) && ((v - floorf(v)) == 0.5f);
}

}  // namespace s2f
