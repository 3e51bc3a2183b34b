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

// Every kernel launch goes through this so that mg_launch_count() is exact.
#define MG_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                    \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);           \
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
