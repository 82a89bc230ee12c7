// Throughput of the instructions that could seed 1/sqrt(x) in the ellipse kernel on B200:
// cycles per warp-instruction per SM sub-partition (8 warps per sub-partition, 4 independent chains each).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; } } while (0)

// DFMA with three register operands mixed with independent integer instructions: does a non-FP64 instruction
// issue "for free" between two FP64 instructions (FP64 pipe accepts one warp instruction per 2 cycles)?
template <int NINT>
__global__ void __launch_bounds__(1024) mix(double* out, int iters, const double* in) {
    double a[4], b = in[threadIdx.x & 7], c = in[8 + (threadIdx.x & 7)];
    int x[4];
    for (int i = 0; i < 4; ++i) { a[i] = in[16 + i] + threadIdx.x * 1e-3; x[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b), "d"(c));
#pragma unroll
            for (int j = 0; j < NINT; ++j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(it), "r"(j + 1));
        }
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) s += a[i] + x[i];
    if (s == 123.456) out[0] = s;
}

template <int OP>
__global__ void __launch_bounds__(1024) k(double* out, int iters, double seed) {
    double a[4];
    float f[4];
    for (int i = 0; i < 4; ++i) { a[i] = seed + threadIdx.x * 1e-3 + i; f[i] = (float)a[i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (OP == 0) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(a[i]));
            if (OP == 1) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            if (OP == 2) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(a[i])); a[i] += 1.0; }
            if (OP == 3) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(a[i]) : "f"(f[i])); f[i] += 1.0f; }
            if (OP == 4) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(a[i]));
            if (OP == 5) a[i] = fma(a[i], 1.0000001, 1e-9);
            if (OP == 6) a[i] += 1.0;     // baseline for the DADD in OP 2
            if (OP == 7) f[i] += 1.0f;    // baseline for the FADD in OP 3
        }
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) s += a[i] + f[i];
    if (s == 123.456) out[0] = s;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double* d; CK(cudaMalloc(&d, 64));
    const int iters = 1 << 12;
    const char* names[] = {"rsqrt.approx.f64 (MUFU.RSQ64H)", "rsqrt.approx.f32 (MUFU.RSQ)", "cvt.f32.f64 + DADD",
                           "cvt.f64.f32 + FADD", "rcp.approx.f64 (MUFU.RCP64H)", "DFMA", "DADD", "FADD"};
    for (int op = 0; op < 8; ++op) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            switch (op) {
                case 0: k<0><<<p.multiProcessorCount, 1024>>>(d, iters, 2.0); break;
                case 1: k<1><<<p.multiProcessorCount, 1024>>>(d, iters, 2.0); break;
                case 2: k<2><<<p.multiProcessorCount, 1024>>>(d, iters, 2.0); break;
                case 3: k<3><<<p.multiProcessorCount, 1024>>>(d, iters, 2.0); break;
                case 4: k<4><<<p.multiProcessorCount, 1024>>>(d, iters, 2.0); break;
                case 5: k<5><<<p.multiProcessorCount, 1024>>>(d, iters, 2.0); break;
                case 6: k<6><<<p.multiProcessorCount, 1024>>>(d, iters, 2.0); break;
                case 7: k<7><<<p.multiProcessorCount, 1024>>>(d, iters, 2.0); break;
            }
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        // 8 warps per sub-partition, 4 instructions per iteration each
        double cyc = (double)p.clockRate * 1e3 * best * 1e-3 / ((double)iters * 4 * 8);
        printf("%-34s %.2f cycles per warp-instruction per sub-partition\n", names[op], cyc);
    }
    double h[32]; for (int i = 0; i < 32; ++i) h[i] = 1.0 + 1e-9 * i;
    double* din; CK(cudaMalloc(&din, sizeof(h))); CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
    for (int nint = 0; nint <= 3; ++nint) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (nint == 0) mix<0><<<p.multiProcessorCount, 1024>>>(d, iters, din);
            if (nint == 1) mix<1><<<p.multiProcessorCount, 1024>>>(d, iters, din);
            if (nint == 2) mix<2><<<p.multiProcessorCount, 1024>>>(d, iters, din);
            if (nint == 3) mix<3><<<p.multiProcessorCount, 1024>>>(d, iters, din);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        double cyc = (double)p.clockRate * 1e3 * best * 1e-3 / ((double)iters * 4 * 8);
        printf("DFMA(3 reg operands) + %d LOP3: %.2f cycles per DFMA per sub-partition\n", nint, cyc);
    }
    return 0;
}
