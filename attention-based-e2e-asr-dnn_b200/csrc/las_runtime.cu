// Runtime plumbing of the C ABI: error string, device info cache, launch counter.
#include "las_common.cuh"
#include <stdlib.h>
#include "las_b200.h"
#include <atomic>
#include <mutex>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void las_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void las_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static LasDeviceInfo g_info[64];
static bool g_info_ok[64];
static std::mutex g_mu;

static int fill_info(int dev) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_info_ok[dev]) return LAS_OK;
    cudaDeviceProp p;
    LAS_CUDA(cudaGetDeviceProperties(&p, dev));
    if (p.major != 10) {
        las_set_error("las_b200 targets sm_100a (B200); device %d is sm_%d%d -- no fallback path exists", dev, p.major, p.minor);
        return LAS_ERR_UNSUPPORTED;
    }
    g_info[dev].device = dev;
    g_info[dev].num_sms = p.multiProcessorCount;
    g_info[dev].max_smem_optin = (int)p.sharedMemPerBlockOptin;
    g_info[dev].coop_launch = p.cooperativeLaunch;
    g_info_ok[dev] = true;
    return LAS_OK;
}

const LasDeviceInfo* las_device_info() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!g_info_ok[dev]) {
        if (fill_info(dev) != LAS_OK) {
            // keep callers safe: conservative defaults (entry points call las_init / las_set_device_of first)
            static LasDeviceInfo fallback{0, 148, 232448, 1};
            return &fallback;
        }
    }
    return &g_info[dev];
}

int las_set_device_of(const void* dev_ptr) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, dev_ptr);
    if (e != cudaSuccess) {
        las_set_error("cudaPointerGetAttributes failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return LAS_ERR_CUDA;
    }
    if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) {
        las_set_error("pointer %p is not device memory (type %d): las_b200 has no CPU path", dev_ptr, (int)at.type);
        return LAS_ERR_ARG;
    }
    int cur = -1;
    LAS_CUDA(cudaGetDevice(&cur));
    if (cur != at.device) LAS_CUDA(cudaSetDevice(at.device));
    if (at.device < 0 || at.device >= 64) return LAS_ERR_ARG;
    if (!g_info_ok[at.device]) return fill_info(at.device);
    return LAS_OK;
}

extern "C" int las_abi_version(void) { return LAS_B200_ABI_VERSION; }

extern "C" int las_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        las_set_error("las_init: no CUDA device (%s); las_b200 has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
        cudaGetLastError();
        return LAS_ERR_CUDA;
    }
    LAS_CHECK_ARG(device >= 0 && device < n && device < 64, "las_init: bad device %d", device);
    LAS_CUDA(cudaSetDevice(device));
    return fill_info(device);
}

extern "C" const char* las_last_error(void) { return g_err; }
extern "C" long long las_launch_count(void) { return g_launches.load(); }
extern "C" void las_launch_count_reset(void) { g_launches.store(0); }

// ---- per-kernel-kind device timing (CUDA events recorded around launches on the launching stream) ----------------
// bench.py enables one or more kinds, runs its timed steps, synchronises, then collects (sum of event-pair durations,
// launch count, algorithmic work).  Disabled (the default) this costs one relaxed atomic load per launch site.
#include <vector>
namespace {
struct ProfRec { int kind; cudaEvent_t e0, e1; double work; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::atomic<unsigned> g_prof_mask{0};
}  // namespace

LasProfScope::LasProfScope(int kind, void* stream, double work) : active_(false), idx_(0), stream_(stream) {
    if (!(g_prof_mask.load(std::memory_order_relaxed) & (1u << kind))) return;
    ProfRec r{kind, nullptr, nullptr, work};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, (cudaStream_t)stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    idx_ = g_prof.size() - 1;
    active_ = true;
}
LasProfScope::~LasProfScope() {
    if (!active_) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_prof[idx_].e1, (cudaStream_t)stream_);
}

extern "C" void las_prof_enable(unsigned kind_mask) { g_prof_mask.store(kind_mask); }
unsigned las_prof_mask_get() { return g_prof_mask.load(std::memory_order_relaxed); }

extern "C" void las_prof_reset(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
}

// caller must have synchronised the device
extern "C" int las_prof_collect(int kind, double* total_ms, long long* count, double* total_work) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double ms = 0, work = 0;
    long long n = 0;
    for (auto& r : g_prof) {
        if (r.kind != kind) continue;
        float t = 0.f;
        cudaError_t e = cudaEventElapsedTime(&t, r.e0, r.e1);
        if (e != cudaSuccess) {
            las_set_error("las_prof_collect: %s (synchronise before collecting)", cudaGetErrorString(e));
            cudaGetLastError();
            return LAS_ERR_CUDA;
        }
        ms += t; work += r.work; ++n;
    }
    if (total_ms) *total_ms = ms;
    if (count) *count = n;
    if (total_work) *total_work = work;
    return LAS_OK;
}


// ---- programmatic dependent launch scope (las_common.cuh) ----
namespace { thread_local bool t_pdl = false; }
bool las_pdl_active() { return t_pdl; }
LasPdlScope::LasPdlScope() : prev_(t_pdl) {
    const char* e = getenv("LAS_PDL");
    t_pdl = !(e && atoi(e) == 0);
}
LasPdlScope::~LasPdlScope() { t_pdl = prev_; }
