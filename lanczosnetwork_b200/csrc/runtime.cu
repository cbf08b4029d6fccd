// ABI version, thread-local error text, launch counter.
#include "common.cuh"

namespace lnb {

static thread_local char g_err[512] = {0};
static thread_local int64_t g_launches = 0;

char* err_buf() { return g_err; }

void set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches += n; }

static unsigned long long* g_prof_buf = nullptr;
unsigned long long* prof_buffer() { return g_prof_buf; }
void set_prof_buffer(unsigned long long* p) { g_prof_buf = p; }

}  // namespace lnb

extern "C" {

int lnb_abi_version(void) { return 1; }

const char* lnb_last_error(void) { return lnb::err_buf(); }

int64_t lnb_launch_count(void) { return lnb::g_launches; }

}  // extern "C"
