// tcgen05 / TMEM / TMA int8 spike GEMM -- placeholder until the kernel lands (returns UNSUPPORTED, never a fallback).
#include "common.cuh"

using namespace s2f;

extern "C" int s2f_gemm_i8_tc(const s2f_gemm_tc_args* args, void* stream) {
  (void)args; (void)stream;
  return fail(S2F_ERR_UNSUPPORTED, "gemm_i8_tc: %s", "not built in this revision");
}

extern "C" int64_t s2f_pack_weights_i8(const float* w, int Cout, int K, int pieces, int8_t* w_packed, float* w_rowscale) {
  (void)w; (void)Cout; (void)K; (void)pieces; (void)w_packed; (void)w_rowscale;
  return -1;
}
