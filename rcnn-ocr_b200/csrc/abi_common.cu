// Error plumbing and device probing behind the C ABI (include/rcnn_ocr_b200.h).
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include "common.cuh"

namespace rcnn {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return RCNN_ERR_CUDA_BASE + (int)e;
}

// SMs the persistent GEMM kernels leave free (rcnn_reserve_sms): under data parallelism NCCL's all-reduce kernels need
// a few SMs of their own to run beside a GEMM that would otherwise hold one CTA on every SM until it ends
static int g_reserved_sms = 0;
int gemm_sms() {
    const int n = num_sms() - g_reserved_sms;
    return n < 2 ? 2 : n;
}

// rcnn_chain_launches: kernels of a dependent chain (the attention decoder's step loop) are launched with programmatic stream
// serialization, so that a kernel's CTAs are dispatched -- and run their prologue -- while its predecessor drains
static int g_chain = 0;
bool chain_launches() { return g_chain != 0; }

int num_sms() {
    static thread_local int cached_dev = -1, cached = 0;   // keyed by the device id: re-read when the thread switches device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

static long long *g_timeline = nullptr;
long long *debug_timeline() { return g_timeline; }
static unsigned int *g_refetch = nullptr;
unsigned int *debug_refetch() { return g_refetch; }
static unsigned long long g_launches = 0;
void count_launch() { ++g_launches; }

// ---- group-barrier counters for the recurrent kernels -------------------------------------------
// Regions [0, kEagerRegions) are a ring for eager launches (a region is reused after kEagerRegions later calls,
// in stream order behind its memset); a launch that is being CAPTURED into a CUDA graph keeps its region for
// every replay, so it takes one of the remaining regions for good (never handed out again): an eager launch on
// another stream can no longer memset a region a graph replay is spinning on.  The indices are atomic.
static const int kCounterRegion = 512, kEagerRegions = 64, kCaptureRegions = 448, kMaxDevices = 32;
static unsigned int *g_counters[kMaxDevices] = {};
static std::atomic<unsigned int> g_counter_next[kMaxDevices], g_capture_next[kMaxDevices];
static std::mutex g_counter_mutex;

unsigned int *group_counters(int n, cudaStream_t s) {
    int dev = 0;
    if (n > kCounterRegion || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        set_error("group_counters: n=%d device=%d unsupported", n, dev);
        return nullptr;
    }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    if (!g_counters[dev]) {   // one-time 512 KB per device (like a library handle's workspace)
        std::lock_guard<std::mutex> lock(g_counter_mutex);
        if (!g_counters[dev]) {
            if (cap != cudaStreamCaptureStatusNone) {
                set_error("group_counters: first use inside a stream capture (run one eager step first)");
                return nullptr;
            }
            unsigned int *q = nullptr;
            if (cudaMalloc(&q, sizeof(unsigned int) * kCounterRegion * (kEagerRegions + kCaptureRegions)) != cudaSuccess) {
                set_error("group_counters: cudaMalloc failed");
                return nullptr;
            }
            g_counters[dev] = q;
        }
    }
    unsigned int region;
    if (cap != cudaStreamCaptureStatusNone) {
        region = g_capture_next[dev].fetch_add(1u);
        if (region >= (unsigned)kCaptureRegions) {
            set_error("group_counters: more than %d captured recurrent launches on device %d", kCaptureRegions, dev);
            return nullptr;
        }
        region += kEagerRegions;
    } else {
        region = g_counter_next[dev].fetch_add(1u) % kEagerRegions;
    }
    unsigned int *p = g_counters[dev] + (size_t)region * kCounterRegion;
    if (cudaMemsetAsync(p, 0, sizeof(unsigned int) * kCounterRegion, s) != cudaSuccess) {
        set_error("group_counters: cudaMemsetAsync failed");
        return nullptr;
    }
    return p;
}

// ---- per-kernel event timing -----------------------------------------------------------
static const int kProfSlots = 16384;
static bool g_prof_on = false;
static cudaEvent_t *g_ev0 = nullptr, *g_ev1 = nullptr;
static int *g_slot_kernel = nullptr;
static int g_prof_used = 0;

int prof_begin(int kernel, cudaStream_t s) {
    if (!g_prof_on || g_prof_used >= kProfSlots) return -1;
    const int slot = g_prof_used++;
    g_slot_kernel[slot] = kernel;
    cudaEventRecord(g_ev0[slot], s);
    return slot;
}

void prof_end(int slot, cudaStream_t s) {
    if (slot >= 0) cudaEventRecord(g_ev1[slot], s);
}

}  // namespace rcnn

extern "C" {

int rcnn_reserve_sms(int n) {
    if (n < 0 || n > 64) {
        rcnn::set_error("reserve_sms: %d out of range (0 .. 64)", n);
        return RCNN_ERR_ARG;
    }
    rcnn::g_reserved_sms = n;
    return RCNN_OK;
}

int rcnn_chain_launches(int on) {
    rcnn::g_chain = on ? 1 : 0;
    return RCNN_OK;
}

int rcnn_debug_timeline(void *buf) { rcnn::g_timeline = (long long *)buf; return RCNN_OK; }
int rcnn_debug_refetch_counter(void *counter) { rcnn::g_refetch = (unsigned int *)counter; return RCNN_OK; }

unsigned long long rcnn_launch_count(void) { return rcnn::g_launches; }

int rcnn_prof_enable(int on) {
    using namespace rcnn;
    if (on && !g_ev0) {
        g_ev0 = new cudaEvent_t[kProfSlots];
        g_ev1 = new cudaEvent_t[kProfSlots];
        g_slot_kernel = new int[kProfSlots];
        for (int i = 0; i < kProfSlots; ++i) {
            RCNN_CUDA(cudaEventCreate(&g_ev0[i]));
            RCNN_CUDA(cudaEventCreate(&g_ev1[i]));
        }
    }
    g_prof_on = on != 0;
    return RCNN_OK;
}

int rcnn_prof_reset(void) {
    rcnn::g_prof_used = 0;
    return RCNN_OK;
}

int rcnn_prof_read(int kernel, double *total_ms, int *launches) {
    using namespace rcnn;
    double tot = 0.0;
    int cnt = 0;
    for (int i = 0; i < g_prof_used; ++i) {
        if (g_slot_kernel[i] != kernel) continue;
        RCNN_CUDA(cudaEventSynchronize(g_ev1[i]));
        float ms = 0.f;
        RCNN_CUDA(cudaEventElapsedTime(&ms, g_ev0[i], g_ev1[i]));
        tot += ms;
        ++cnt;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = cnt;
    return RCNN_OK;
}

int rcnn_version(void) { return 100; }

const char *rcnn_last_error(void) { return rcnn::g_err; }

int rcnn_device_check(void) {
    int dev = 0;
    RCNN_CUDA(cudaGetDevice(&dev));
    int major = 0;
    RCNN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        rcnn::set_error("device %d has compute capability %d.x; these kernels are sm_100a only", dev, major);
        return RCNN_ERR_DEVICE;
    }
    return RCNN_OK;
}

}  // extern "C"
