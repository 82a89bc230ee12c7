// FP64 roofline denominators measured in the run (bench.py): the plain DFMA pipe and the DMMA (mma.sync m8n8k4 f64)
// path of THIS GPU at ITS clocks.  MEASURED_PEAKS.json (driver-written) carries the HBM copy bandwidth and the bf16
// GEMM rate only; the FP64 kernels of this library (K1 general nu: FP64 pipe; K4 / K5: DMMA) need these two.
// Not on any product path: a measurement utility behind the C ABI so that bench.py times it with the library that
// is being benchmarked, on the same box, next to the timed kernels ("FP64 peak is unmeasured").
#include "gmg_common.cuh"
#include "../../include/glomar_b200.h"

namespace {

constexpr int kIlp = 8;

__global__ void __launch_bounds__(256) mb_dfma_kernel(double* out, int iters, double a, double b) {
    double acc[kIlp];
#pragma unroll
    for (int k = 0; k < kIlp; ++k) acc[k] = threadIdx.x * 1e-3 + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kIlp; ++k) acc[k] = fma(acc[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < kIlp; ++k) s += acc[k];
    if (s == 123.456) out[0] = s;                  // keeps the chain alive, never true
}

__global__ void __launch_bounds__(256) mb_dmma_kernel(double* out, int iters, double a, double b) {
    double c0[kIlp], c1[kIlp];
#pragma unroll
    for (int k = 0; k < kIlp; ++k) { c0[k] = threadIdx.x * 1e-3; c1[k] = k; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kIlp; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < kIlp; ++k) s += c0[k] + c1[k];
    if (s == 123.456) out[0] = s;
}

}  // namespace

extern "C" int gmg_microbench_fp64(int what, int iters, int reps, double* d_scratch, double* tflops_best,
                                   double* tflops_sustained, gmg_stream_t stream) {
    if (what != GMG_MICROBENCH_DFMA && what != GMG_MICROBENCH_DMMA) return -1;
    if (iters <= 0) return -2;
    if (reps <= 0 || reps > 10000) return -3;
    if (!d_scratch) return -4;
    if (!tflops_best) return -5;
    cudaStream_t s = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    GMG_CUDA_CHECK(cudaGetDevice(&dev));
    GMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8;                    // 8 CTAs x 256 threads = 2048 threads per SM
    // flops of one launch: DFMA 2 per lane; DMMA m8n8k4 2 * 8 * 8 * 4 per WARP instruction
    const double flops = what == GMG_MICROBENCH_DFMA ? 2.0 * blocks * 256.0 * iters * kIlp
                                                     : 2.0 * 256.0 * blocks * 8.0 * iters * kIlp;
    auto launch = [&]() {
        if (what == GMG_MICROBENCH_DFMA) mb_dfma_kernel<<<blocks, 256, 0, s>>>(d_scratch, iters, 1.0000001, 1e-9);
        else mb_dmma_kernel<<<blocks, 256, 0, s>>>(d_scratch, iters, 1.0000001, 1e-9);
    };
    cudaEvent_t e0, e1, e2;
    GMG_CUDA_CHECK(cudaEventCreate(&e0));
    GMG_CUDA_CHECK(cudaEventCreate(&e1));
    GMG_CUDA_CHECK(cudaEventCreate(&e2));
    launch();
    launch();
    GMG_CUDA_CHECK(cudaStreamSynchronize(s));
    float best = 1e30f, total = 0.f;
    GMG_CUDA_CHECK(cudaEventRecord(e2, s));        // start of the back-to-back (sustained) window
    for (int r = 0; r < reps; ++r) {
        GMG_CUDA_CHECK(cudaEventRecord(e0, s));
        launch();
        GMG_CUDA_CHECK(cudaEventRecord(e1, s));
    }
    GMG_CUDA_CHECK(cudaEventSynchronize(e1));
    GMG_CUDA_CHECK(cudaEventElapsedTime(&total, e2, e1));
    float last = 0.f;
    GMG_CUDA_CHECK(cudaEventElapsedTime(&last, e0, e1));
    best = last;                                   // the last launch, at the clocks the window settled on ...
    for (int r = 0; r < 3; ++r) {                  // ... and three isolated ones (burst)
        GMG_CUDA_CHECK(cudaEventRecord(e0, s));
        launch();
        GMG_CUDA_CHECK(cudaEventRecord(e1, s));
        GMG_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms;
        GMG_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    GMG_LAUNCH_CHECK();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    *tflops_best = flops / (best * 1e-3) / 1e12;
    if (tflops_sustained) *tflops_sustained = flops * reps / (total * 1e-3) / 1e12;
    return GMG_OK;
}
