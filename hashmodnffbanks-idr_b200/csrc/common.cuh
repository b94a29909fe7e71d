// Shared helpers for libidrk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/idrk.h"

#define IDRK_CUDA_TRY(expr)                         \
    do {                                            \
        cudaError_t _e = (expr);                    \
        if (_e != cudaSuccess) return (int)_e;      \
    } while (0)

#define IDRK_LAUNCH_CHECK()                         \
    do {                                            \
        cudaError_t _e = cudaGetLastError();        \
        if (_e != cudaSuccess) return (int)_e;      \
    } while (0)

namespace idrk {

inline int sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

// Programmatic dependent launch.  Every kernel of the library is launched with the programmatic-stream-serialization
// attribute and starts with pdl_wait() (griddepcontrol.wait: blocks until the preceding grid in the stream has
// completed and its writes are visible) followed by pdl_trigger() (lets the NEXT grid be scheduled early; it then
// parks in its own pdl_wait()).  Semantics are those of plain stream order; what is saved is the grid-launch latency
// between the ~1100 dependent launches of a step, also inside captured CUDA graphs.  IDRK_PDL=0 disables it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline int pdl_enabled() {
    static const int on = [] { const char* e = getenv("IDRK_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace idrk
