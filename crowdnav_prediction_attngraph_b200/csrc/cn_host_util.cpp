#include "cn_host_util.h"
#include "../../include/crowdnav_b200.h"

static thread_local char g_err[1024] = "";

int cn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

extern "C" const char* cn_last_error(void) { return g_err; }
