// FP64 roofline denominators for B200 (sm_100a): DFMA pipe, DMMA (mma.sync f64) and
// store-only HBM bandwidth.  MEASURED_PEAKS.json only has the HBM copy and bf16 GEMM
// figures, so the FP64 kernels' fractions are taken against what this prints.
//
//   mkdir -p build && nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench_fp64 microbench_fp64.cu
//   build/microbench_fp64        -> one JSON object on stdout
// (-cudart shared: a statically linked runtime drags every runtime symbol name into the binary; build/ is git-ignored.
//  The DFMA / DMMA figures are also available in the run through gmg_microbench_fp64 of the library, which is what
//  bench.py uses.)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x * 1e-3 + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k];
    if (s == 123.456) out[0] = s;
}

// mma.sync m8n8k4 f64: 8*8*4 = 256 FMA per warp instruction.
template <int ILP>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { c0[k] = threadIdx.x * 1e-3; c1[k] = k; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += c0[k] + c1[k];
    if (s == 123.456) out[0] = s;
}

// mma.sync m16n8k16 f64: 16*8*16 = 2048 FMA per warp instruction.
template <int ILP>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { c[k][0] = threadIdx.x * 1e-3; c[k][1] = k; c[k][2] = 1; c[k][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                "{%4,%4,%4,%4,%4,%4,%4,%4}, {%5,%5,%5,%5}, {%0,%1,%2,%3};\n"
                : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) store_kernel(double2* out, size_t n2, double v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) out[i] = make_double2(v, v + 1.0);
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n2) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) out[i] = in[i];
}

template <typename F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* d; CK(cudaMalloc(&d, 1024));
    const int iters = 1 << 14;
    const int blocks = sms * 8;   // 8 CTAs x 256 thr = 2048 threads/SM
    float t_fma = time_ms([&] { dfma_kernel<8><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); }, 5);
    double fma_tf = 2.0 * blocks * 256.0 * iters * 8 / (t_fma * 1e-3) / 1e12;
    float t_884 = time_ms([&] { dmma884_kernel<8><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); }, 5);
    double d884_tf = 2.0 * 256.0 * blocks * 8.0 /*warps*/ * iters * 8 / (t_884 * 1e-3) / 1e12;
    float t_16816 = time_ms([&] { dmma16816_kernel<4><<<blocks, 256>>>(d, iters / 4, 1.0000001, 1e-9); }, 5);
    double d16816_tf = 2.0 * 2048.0 * blocks * 8.0 * (iters / 4) * 4 / (t_16816 * 1e-3) / 1e12;

    // DFMA issue-to-issue latency of a dependent chain and how many independent chains per SM sub-partition
    // saturate the pipe: one CTA per SM, W warps per sub-partition, ILP independent chains per thread.
    // cycles per DFMA per sub-partition = clock * time / (iters * ILP * W)
    auto sweep = [&](int warps_per_smsp, int ilp) {
        const int threads = 128 * warps_per_smsp;
        float t;
        if (ilp == 1) t = time_ms([&] { dfma_kernel<1><<<sms, threads>>>(d, iters, 1.0000001, 1e-9); }, 3);
        else if (ilp == 2) t = time_ms([&] { dfma_kernel<2><<<sms, threads>>>(d, iters, 1.0000001, 1e-9); }, 3);
        else t = time_ms([&] { dfma_kernel<4><<<sms, threads>>>(d, iters, 1.0000001, 1e-9); }, 3);
        return (double)p.clockRate * 1e3 * t * 1e-3 / ((double)iters * ilp * warps_per_smsp);
    };
    fprintf(stderr, "DFMA cycles per warp-instruction per sub-partition (2.0 = pipe saturated):\n");
    for (int w : {1, 2, 3, 4, 6, 8})
        fprintf(stderr, "  warps/SMSP=%d: ILP1 %.2f  ILP2 %.2f  ILP4 %.2f\n", w, sweep(w, 1), sweep(w, 2), sweep(w, 4));
    const double dfma_latency = sweep(1, 1);

    size_t bytes = (size_t)8 << 30;   // 8 GiB, far larger than the 126 MB L2
    double2 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    size_t n2 = bytes / sizeof(double2);
    float t_st = time_ms([&] { store_kernel<<<sms * 16, 256>>>(a, n2, 1.5); }, 5);
    float t_cp = time_ms([&] { copy_kernel<<<sms * 16, 256>>>(a, b, n2); }, 5);
    float t_ms = time_ms([&] { CK(cudaMemsetAsync(a, 0, bytes)); }, 5);
    size_t free_b, total_b; CK(cudaMemGetInfo(&free_b, &total_b));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, "
           "\"dmma_m16n8k16_tflops\": %.2f, \"store_only_gbs\": %.1f, \"copy_rw_gbs\": %.1f, \"memset_gbs\": %.1f, "
           "\"hbm_total_gb\": %.1f, \"dfma_dependent_latency_cycles\": %.1f}\n",
           p.name, sms, p.clockRate, fma_tf, d884_tf, d16816_tf, bytes / (t_st * 1e-3) / 1e9,
           2.0 * bytes / (t_cp * 1e-3) / 1e9, bytes / (t_ms * 1e-3) / 1e9, total_b / 1e9, dfma_latency);
    return 0;
}
