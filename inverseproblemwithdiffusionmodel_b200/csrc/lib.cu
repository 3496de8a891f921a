// Library-wide state of libipdm_b200.so: error text, launch counter, ABI version.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace ipdm {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace ipdm

extern "C" int ipdm_abi_version(void) { return 3; }
extern "C" const char* ipdm_last_error(void) { return ipdm::g_err; }
extern "C" unsigned long long ipdm_launch_count(void) { return ipdm::g_launches.load(); }
