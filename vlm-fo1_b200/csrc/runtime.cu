// runtime.cu -- error reporting and launch accounting behind the C ABI.
#include "common.cuh"

namespace fo1 {

static thread_local char t_err[1024] = {0};
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

}  // namespace fo1

extern "C" int fo1_abi_version(void) { return FO1_ABI_VERSION; }
extern "C" const char* fo1_last_error(void) { return fo1::t_err; }
extern "C" uint64_t fo1_launch_count(void) { return fo1::g_launches.load(); }
extern "C" void fo1_launch_count_reset(void) { fo1::g_launches.store(0); }
