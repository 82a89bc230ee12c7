// Shared definitions for libglomar_b200.so (error codes, launch helpers).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>

// Return-code convention (include/glomar_b200.h): 0 ok; negative = index of the offending
// argument (1-based); positive = runtime failure class below.
enum GmgStatus : int {
    GMG_OK = 0,
    GMG_ERR_CUDA = 1,          // a CUDA runtime call or launch failed (see gmg_last_cuda_error)
    GMG_ERR_NOT_SPD = 2,       // Cholesky met a non-positive pivot (info out-param has the order)
    GMG_ERR_WORKSPACE = 3,     // workspace too small / misaligned
    GMG_ERR_UNSUPPORTED = 4,   // argument combination not supported
    GMG_ERR_SINGULAR = 5,      // LU met an exactly zero pivot
    GMG_ERR_COMM = 6,          // NCCL is not available in the process or a collective call failed
};

extern thread_local int g_gmg_last_cuda_error;

#define GMG_CUDA_CHECK(expr)                                   \
    do {                                                       \
        cudaError_t _e = (expr);                               \
        if (_e != cudaSuccess) {                               \
            g_gmg_last_cuda_error = (int)_e;                   \
            return GMG_ERR_CUDA;                               \
        }                                                      \
    } while (0)

#define GMG_LAUNCH_CHECK() GMG_CUDA_CHECK(cudaGetLastError())

static inline int64_t gmg_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t gmg_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
