// Shared host/device helpers for libmaggie_b200.so (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/maggie_b200.h"

namespace mg {

extern std::atomic<unsigned long long> g_launches;
void set_error(const char* fmt, ...);

// Programmatic dependent launch (PDL): every kernel is launched with the "programmatic stream serialization" attribute,
// signals `griddepcontrol.launch_dependents` as its first instruction and executes `griddepcontrol.wait` before it
// touches global memory.  The next kernel of the stream (or of the captured CUDA graph) is then scheduled onto the SMs
// while the tail of the current one drains, and only its memory accesses wait for the full completion + flush of its
// predecessor: the launch gap between the ~600 dependent kernels of a step disappears.  MAGGIE_B200_NO_PDL=1 launches
// plainly (the device-side instructions are no-ops then).
bool pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
    pdl_launch();
    pdl_wait();
}
#endif

// Every kernel launch goes through this so that mg_launch_count() is exact.
#define MG_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                    \
        cudaLaunchConfig_t cfg__ = {};                                                      \
        cfg__.gridDim = dim3(grid);                                                         \
        cfg__.blockDim = dim3(block);                                                       \
        cfg__.dynamicSmemBytes = (smem);                                                    \
        cfg__.stream = (cudaStream_t)(stream);                                              \
        cudaLaunchAttribute attr__[1];                                                      \
        attr__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                  \
        attr__[0].val.programmaticStreamSerializationAllowed = 1;                           \
        cfg__.attrs = attr__;                                                               \
        cfg__.numAttrs = mg::pdl_enabled() ? 1 : 0;                                         \
        cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__);                                    \
        mg::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
    } while (0)

#define MG_CHECK_LAUNCH(name)                                                               \
    do {                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) {                                                           \
            mg::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));          \
            return MG_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

#define MG_REQUIRE(cond, ...)                                                               \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            mg::set_error(__VA_ARGS__);                                                     \
            return MG_ERR_ARG;                                                              \
        }                                                                                   \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

constexpr int kNumSMs = 148;  // B200

}  // namespace mg
