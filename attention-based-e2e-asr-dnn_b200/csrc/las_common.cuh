// Common helpers for the las_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#define LAS_OK 0
#define LAS_ERR_ARG -1
#define LAS_ERR_CUDA -2
#define LAS_ERR_WORKSPACE -3
#define LAS_ERR_UNSUPPORTED -4

// thread-local error string readable through las_last_error() (C ABI: no exceptions cross the boundary)
void las_set_error(const char* fmt, ...);

#define LAS_CHECK_ARG(cond, ...)                      \
    do {                                              \
        if (!(cond)) {                                \
            las_set_error(__VA_ARGS__);               \
            return LAS_ERR_ARG;                       \
        }                                             \
    } while (0)

#define LAS_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess) {                                                                \
            las_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return LAS_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

void las_count_launch(int n);

#define LAS_LAUNCH_CHECK()                                                                      \
    do {                                                                                        \
        las_count_launch(1);                                                                    \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) {                                                                \
            las_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return LAS_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

// every entry point may be called from the autograd engine thread: make the pointer's device current
int las_set_device_of(const void* dev_ptr);

struct LasDeviceInfo {
    int device;
    int num_sms;
    int max_smem_optin;
    int coop_launch;
};
const LasDeviceInfo* las_device_info();   // for the current device (cached)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- device helpers ----
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------------
// Inside the decoder loops (thousands of tiny dependent kernels) every launch carries the programmatic-stream-serialization
// attribute: kernel N+1 is scheduled while kernel N still runs and sits in griddepcontrol.wait until N has completed and
// flushed, so launch latency, CTA scheduling and each kernel's prologue (barrier init, TMEM allocation, descriptor
// prefetch) overlap the predecessor's tail.  Every participating kernel executes pdl_wait() BEFORE its first global-memory
// access (reads AND writes: the predecessor may still be reading what we overwrite) and pdl_trigger() right after.
// Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool las_pdl_active();                 // true inside a LasPdlScope on this thread (and LAS_PDL != 0)
struct LasPdlScope {
    LasPdlScope();
    ~LasPdlScope();
    bool prev_;
};

// kernel launch that adds the PDL attribute when a LasPdlScope is active
template <typename... KArgs, typename... Args>
inline cudaError_t las_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = las_pdl_active() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- profiling scope (see las_runtime.cu) ----
enum LasProfKind { LAS_PROF_GEMM_GATES = 0, LAS_PROF_GEMM_OTHER = 1, LAS_PROF_REC_FWD = 2, LAS_PROF_REC_BWD = 3,
                   LAS_PROF_ATTN_FWD = 4, LAS_PROF_ATTN_BWD = 5, LAS_PROF_ADAM = 6, LAS_PROF_SPELLER_FWD = 7,
                   LAS_PROF_SPELLER_BWD = 8,
                   LAS_PROF_GEMM_GATES_SIDE = 9 /* gate GEMMs launched with max_ctas > 0: beside a recurrence kernel, on the SMs it leaves free */ };
class LasProfScope {
public:
    LasProfScope(int kind, void* stream, double work);
    ~LasProfScope();
private:
    bool active_;
    size_t idx_;
    void* stream_;
};

// ---- fused LSTM-cell epilogue of the decoder-step tensor-core GEMM (gemm_tc.cu <-> decoder.cu) ----
// The GEMM's B operand rows are permuted so that every 64-column tile holds 16 units x 4 gates ([i|f|g|o] x 16); each epilogue
// thread owns one batch row and finishes those 16 units: + table/bias, nonlinearities, cell update, dropout, all outputs.
struct LasLstmEpi {
    int H;                                        // units of this cell (GEMM N = 4H)
    const float* tab;                             // optional (V, 4H) table gathered by token (gate-major layout)
    const float* bias1; const float* bias2;       // optional (4H) biases (gate-major layout)
    const int* y; long long ld_y;                 // gold tokens (B, >= steps) or null
    const int* chars_prev;                        // (B) greedy argmax of the previous step or null
    int* tok_out;                                 // (B) token fed this step (saved for backward) or null
    int t, use_gold, sos_idx;
    const float* c_prev; long long ld_cp;
    float* c_out; long long ld_co;
    const float* mask;                            // (B, H) dropout mask or null
    float* G;                                     // (B, 4H) activated gates, gate-major layout (saved for backward)
    float* h1; long long ld_h1; float* h2; long long ld_h2;                       // fp32 destinations (h2 nullable)
    __nv_bfloat16* h1b; long long ld_h1b; __nv_bfloat16* h2b; long long ld_h2b;   // bf16 destinations (nullable)
};
