// abi.cu -- version, error reporting and device enumeration of the C ABI.
#include <map>
#include <mutex>

#include "common.h"

namespace rl {
static thread_local std::string g_last_error;

// Host ranges page-locked through rl_host_register (start -> end).  The host pipeline asks whether an
// output buffer is one of them: kernel stores into cudaHostRegister-ed pageable memory measured
// markedly slower than into cudaHostAlloc-ed memory (604 vs 374 us for 17.7 MB), the copy engine does
// not care (411 us), so registered buffers take the DMA pipeline and CUDA-allocated ones the
// zero-copy stores.
static std::mutex g_reg_mutex;
static std::map<uintptr_t, uintptr_t> g_registered;

int host_registered_range(const void *p, size_t bytes)
{
    std::lock_guard<std::mutex> lock(g_reg_mutex);
    const uintptr_t a = (uintptr_t)p, b = a + bytes;
    auto it = g_registered.upper_bound(a);
    if (it != g_registered.begin()) {
        auto prev = it;
        --prev;
        if (a < prev->second) return b <= prev->second ? 1 : -1;
    }
    // starts outside every registration: does it run into the next one?
    if (it != g_registered.end() && it->first < b) return -1;
    return 0;
}

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

// Page-lock a caller-owned host range so the *_host entry points can use it in place (ranges stored
// straight into it by the kernel, inputs copied from it without staging).  Nothing is registered when
// the range is page-locked already (*was_pinned = 1: the caller must not unregister it).
int32_t rl_host_register(int32_t device, void *ptr, int64_t bytes, int32_t *was_pinned)
{
    if (!ptr || bytes <= 0) return rl::fail(RL_ERR_BAD_ARG, "rl_host_register: bad argument");
    rl::DeviceGuard guard(device);
    if (!guard.ok) return rl::fail(RL_ERR_NO_DEVICE, "rl_host_register: no such device");
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) == cudaSuccess && a.type != cudaMemoryTypeUnregistered) {
        if (was_pinned) *was_pinned = 1;
        return a.type == cudaMemoryTypeHost ? RL_OK : rl::fail(RL_ERR_BAD_ARG, "rl_host_register: not a host pointer");
    }
    cudaGetLastError();
    if (was_pinned) *was_pinned = 0;
    RL_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    {
        std::lock_guard<std::mutex> lock(rl::g_reg_mutex);
        rl::g_registered[(uintptr_t)ptr] = (uintptr_t)ptr + (uintptr_t)bytes;
    }
    return RL_OK;
}

int32_t rl_host_unregister(int32_t device, void *ptr)
{
    if (!ptr) return rl::fail(RL_ERR_BAD_ARG, "rl_host_unregister: null pointer");
    rl::DeviceGuard guard(device);
    if (!guard.ok) return rl::fail(RL_ERR_NO_DEVICE, "rl_host_unregister: no such device");
    {
        std::lock_guard<std::mutex> lock(rl::g_reg_mutex);
        rl::g_registered.erase((uintptr_t)ptr);
    }
    RL_CUDA(cudaHostUnregister(ptr));
    return RL_OK;
}

}  // extern "C"
