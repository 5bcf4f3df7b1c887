// Internal helpers shared by the translation units of libb200infer.so (not part of the ABI).
#pragma once
#include <string>

namespace b2i {
extern thread_local std::string g_err;
// records a thread-local error message and returns `code`
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
}  // namespace b2i
