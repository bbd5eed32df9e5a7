// Library-level entry points of libptb_b200.so (error string, ABI version, launch counter).
#include "ptb_common.cuh"

namespace ptb {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace ptb

extern "C" int ptb_abi_version(void) { return PTB_ABI_VERSION; }
extern "C" const char* ptb_last_error(void) { return ptb::g_err; }
extern "C" uint64_t ptb_launch_count(void) { return ptb::g_launches.load(); }
