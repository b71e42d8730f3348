// api.cu -- C-ABI plumbing shared by all ops: error text, device properties, log level.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_error[1024] = "";
static int g_log_level = 1;      // same default as the reference (nvdiffrast/torch/ops.py:27-41)
#include <atomic>
static std::atomic<unsigned long long> g_launches{0};   // kernels launched by this library, any thread (bench.py reports it as gpu_launches)

void b200dr_count_launches(int n) { g_launches.fetch_add((unsigned long long)(long long)n, std::memory_order_relaxed); }

void b200dr_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int b200dr_num_sms(void)
{
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0)
    {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

extern "C" int b200dr_version(void) { return B200DR_VERSION; }
extern "C" const char* b200dr_last_error(void) { return g_error; }
extern "C" int b200dr_get_log_level(void) { return g_log_level; }
extern "C" void b200dr_set_log_level(int level) { g_log_level = level; }
extern "C" unsigned long long b200dr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
