// mk_api.cu -- version, error reporting and launch accounting of libmosaic_b200.
#include "mk_common.cuh"

#include <stdarg.h>

namespace mk {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}

}  // namespace mk

extern "C" int mk_abi_version(void) { return MK_ABI_VERSION; }
extern "C" const char *mk_version_string(void) { return "mosaic_b200 0.1 (sm_100a)"; }
extern "C" const char *mk_last_error(void) { return mk::t_err; }
extern "C" uint64_t mk_launch_count(void) { return mk::g_launches.load(); }
