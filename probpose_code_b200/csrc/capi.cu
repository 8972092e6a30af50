// C-ABI glue: error string, version, launch counter.
#include "common.cuh"

#include <string.h>

namespace pp {

static thread_local char g_error[1024] = "";
thread_local int64_t g_launch_count = 0;

bool pdl_enabled() {
  static const bool on = getenv("PP_NO_PDL") == nullptr;
  return on;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

}  // namespace pp

extern "C" const char* pp_last_error(void) { return pp::g_error; }
extern "C" const char* pp_version(void) { return "probpose_b200 0.1.0 sm_100a"; }
