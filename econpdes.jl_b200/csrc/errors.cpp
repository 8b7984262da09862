#include "errors.h"

#include "../../include/econpdes_cuda.h"

namespace econpdes {
namespace {
thread_local std::string g_last_error;
}
void set_error(const std::string& msg) { g_last_error = msg; }
const char* last_error() { return g_last_error.c_str(); }
}  // namespace econpdes

extern "C" {
int econpdes_abi_version(void) { return ECONPDES_ABI_VERSION; }
const char* econpdes_last_error(void) { return econpdes::last_error(); }
}
