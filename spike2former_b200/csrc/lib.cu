// Library-wide state and the small ABI entry points of libs2f.so.
#include "common.cuh"

namespace s2f {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace s2f

extern "C" const char* s2f_last_error(void) { return s2f::g_err; }
extern "C" int s2f_abi_version(void) { return 10; }
extern "C" uint64_t s2f_launch_count(void) { return s2f::g_launches.load(); }
