// emu_support.cpp -- TEST INFRASTRUCTURE ONLY: the two symbols the operator files expect from c_abi.cu.
#include <cstdio>

namespace fgs {
static thread_local char g_emu_err[512] = "";
int report_error(int code, const char* where, const char* detail) {
    std::snprintf(g_emu_err, sizeof(g_emu_err), "%s: %s", where, detail);
    return code;
}
}  // namespace fgs

extern "C" const char* fgs_last_error(void) { return fgs::g_emu_err; }
