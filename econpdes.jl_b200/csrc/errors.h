// Thread-local last-error string behind econpdes_last_error().
#pragma once
#include <string>

namespace econpdes {
void set_error(const std::string& msg);
const char* last_error();
}  // namespace econpdes
