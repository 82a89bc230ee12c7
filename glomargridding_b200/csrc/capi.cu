// Library-level entry points of libglomar_b200.so (version, error strings).
#include "gmg_common.cuh"
#include "../../include/glomar_b200.h"

thread_local int g_gmg_last_cuda_error = 0;

extern "C" {

int gmg_version(void) { return 100; }   // 0.1.0

int gmg_last_cuda_error(void) { return g_gmg_last_cuda_error; }

const char* gmg_strerror(int code) {
    if (code < 0) return "invalid argument (the negated code is the 1-based argument index)";
    switch (code) {
        case GMG_OK: return "ok";
        case GMG_ERR_CUDA: return cudaGetErrorString((cudaError_t)g_gmg_last_cuda_error);
        case GMG_ERR_NOT_SPD: return "matrix is not positive definite (non-positive Cholesky pivot)";
        case GMG_ERR_WORKSPACE: return "workspace too small or misaligned";
        case GMG_ERR_UNSUPPORTED: return "unsupported argument combination";
        case GMG_ERR_SINGULAR: return "matrix is singular (zero LU pivot)";
        case GMG_ERR_COMM: return "NCCL is not loaded in this process or a collective call failed";
        default: return "unknown error";
    }
}

}  // extern "C"
