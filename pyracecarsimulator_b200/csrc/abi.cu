// abi.cu -- version, error reporting and device enumeration of the C ABI.
#include "common.h"

namespace rl {
static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int32_t fail(int32_t code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}
}  // namespace rl

extern "C" {

int32_t rl_abi_version(void) { return RL_ABI_VERSION; }

const char *rl_last_error(void) { return rl::g_last_error.c_str(); }

int32_t rl_device_count(int32_t *count)
{
    if (!count) return rl::fail(RL_ERR_BAD_ARG, "rl_device_count: null count");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return rl::fail(RL_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = n;
    return RL_OK;
}

}  // extern "C"
