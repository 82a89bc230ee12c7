// Bandwidth-bound N x N bookkeeping of the covariance path on the device:
//   gmg_cov2cor      reference utils.cov_2_cor, utils.py:637-671 (called by EllipseCovarianceBuilder.calculate_cor,
//                    ellipse/covariance.py:470-473, and by covariance_tools.laloux_clip):  cor = cov / outer(sd, sd),
//                    sd = sqrt(diag(cov)), exact zeros of cov kept as zeros; the host decides from two scalars
//                    (all diagonal entries == 1 ?  largest |diag - 1|) whether the diagonal is reset or an error raised
//   gmg_cor2cov      reference utils.cor_2_cov, utils.py:611-634:  cov = cor * outer(sd, sd), zeros kept
//   gmg_uncompress   EllipseCovarianceBuilder.uncompress_cov, ellipse/covariance.py:475-509: scatter the matrix of the
//                    unmasked points into full-grid size with fill values
// Arithmetic is done in the matrix' own precision (float32 or float64), like numpy does.  One thread per pair of
// neighbouring columns (16 B / 8 B vector accesses), one block row per blockIdx.y, grid-stride over rows.
#include "gmg_common.cuh"
#include "../../include/glomar_b200.h"

namespace {

template <typename T> struct V2;
template <> struct V2<double> { using type = double2; };
template <> struct V2<float> { using type = float2; };

template <typename T>
__global__ void __launch_bounds__(256) diag_sqrt_kernel(const T* __restrict__ C, int64_t ld, int64_t n,
                                                        T* __restrict__ sd) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) sd[i] = sqrt(C[i * ld + i]);
}

// stat[0] = number of diagonal entries != 1, stat[1] = bits of max |diag - 1| (as double; NaN counts as +inf)
template <typename T, bool TO_COR>
__global__ void __launch_bounds__(256) scale_kernel(const T* __restrict__ A, int64_t lda, int64_t n,
                                                    const T* __restrict__ sd, T* __restrict__ out, int64_t ldo,
                                                    unsigned long long* __restrict__ stat, int vec2) {
    const int64_t j = 2 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
    if (j >= n) return;
    const bool two = j + 1 < n;
    const T sj0 = sd[j], sj1 = two ? sd[j + 1] : (T)1;
    for (int64_t i = blockIdx.y; i < n; i += gridDim.y) {
        const T si = sd[i];
        T a0, a1 = (T)0;
        const T* src = A + i * lda + j;
        if (two && vec2) {
            const typename V2<T>::type v = *reinterpret_cast<const typename V2<T>::type*>(src);
            a0 = v.x; a1 = v.y;
        } else {
            a0 = src[0];
            if (two) a1 = src[1];
        }
        T r0, r1;
        if (TO_COR) {
            r0 = a0 / (si * sj0);                       // cov / np.outer(stdevs, stdevs)
            r1 = a1 / (si * sj1);
        } else {
            r0 = a0 * (si * sj0);                       // cor * np.outer(stdevs, stdevs)
            r1 = a1 * (si * sj1);
        }
        if (a0 == (T)0) r0 = (T)0;                      // cor[cov == 0] = 0  /  cov[cor == 0] = 0
        if (a1 == (T)0) r1 = (T)0;
        if (TO_COR && stat && (i == j || (two && i == j + 1))) {
            const T dg = (i == j) ? r0 : r1;
            if (!(dg == (T)1)) {
                atomicAdd(&stat[0], 1ull);
                const double dev = fabs((double)dg - 1.0);
                const double key = (dev == dev) ? dev : __longlong_as_double(0x7ff0000000000000ll);
                atomicMax(&stat[1], (unsigned long long)__double_as_longlong(key));
            }
        }
        T* dst = out + i * ldo + j;
        if (two && vec2) {
            typename V2<T>::type v;
            v.x = r0; v.y = r1;
            *reinterpret_cast<typename V2<T>::type*>(dst) = v;
        } else {
            dst[0] = r0;
            if (two) dst[1] = r1;
        }
    }
}

// inv[r] = index of full-grid point r in the compressed ordering, -1 if masked
template <typename T>
__global__ void __launch_bounds__(256) uncompress_kernel(const T* __restrict__ C, int64_t ldc,
                                                         const int* __restrict__ inv, int64_t M, T fill, T diag_fill,
                                                         T* __restrict__ out, int64_t ldo) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= M) return;
    const int ic = inv[c];
    for (int64_t r = blockIdx.y; r < M; r += gridDim.y) {
        const int ir = inv[r];
        T v = (r == c) ? diag_fill : fill;
        if (ir >= 0 && ic >= 0) v = C[(int64_t)ir * ldc + ic];
        out[r * ldo + c] = v;
    }
}

template <typename T>
int scale_launch(bool to_cor, const void* A, int64_t lda, int64_t n, void* sd_ws, int from_diag, void* out,
                 int64_t ldo, unsigned long long* stat, cudaStream_t s) {
    const T* a = reinterpret_cast<const T*>(A);
    T* sd = reinterpret_cast<T*>(sd_ws);
    if (from_diag) {
        diag_sqrt_kernel<T><<<(unsigned)gmg_ceil_div(n, 256), 256, 0, s>>>(a, lda, n, sd);
        GMG_LAUNCH_CHECK();
    }
    if (stat) GMG_CUDA_CHECK(cudaMemsetAsync(stat, 0, 2 * sizeof(unsigned long long), s));
    const int vec2 = (lda % 2 == 0) && (ldo % 2 == 0) && (reinterpret_cast<uintptr_t>(A) % (2 * sizeof(T)) == 0) &&
                     (reinterpret_cast<uintptr_t>(out) % (2 * sizeof(T)) == 0);
    dim3 grid((unsigned)gmg_ceil_div(gmg_ceil_div(n, 2), 256), (unsigned)(n < 2048 ? n : 2048));
    if (to_cor)
        scale_kernel<T, true><<<grid, 256, 0, s>>>(a, lda, n, sd, reinterpret_cast<T*>(out), ldo, stat, vec2);
    else
        scale_kernel<T, false><<<grid, 256, 0, s>>>(a, lda, n, sd, reinterpret_cast<T*>(out), ldo, nullptr, vec2);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

}  // namespace

extern "C" {

int gmg_cov2cor(const void* cov, int64_t ld, int64_t n, int is_f32, void* sd_ws, void* cor, int64_t ldo,
                unsigned long long* stat, gmg_stream_t stream) {
    if (!cov) return -1;
    if (ld < n) return -2;
    if (n <= 0) return -3;
    if (!sd_ws) return -5;
    if (!cor) return -6;
    if (ldo < n) return -7;
    if (!stat) return -8;
    cudaStream_t s = (cudaStream_t)stream;
    return is_f32 ? scale_launch<float>(true, cov, ld, n, sd_ws, 1, cor, ldo, stat, s)
                  : scale_launch<double>(true, cov, ld, n, sd_ws, 1, cor, ldo, stat, s);
}

int gmg_cor2cov(const void* cor, int64_t ld, int64_t n, int is_f32, const void* sd, void* cov, int64_t ldo,
                gmg_stream_t stream) {
    if (!cor) return -1;
    if (ld < n) return -2;
    if (n <= 0) return -3;
    if (!sd) return -5;
    if (!cov) return -6;
    if (ldo < n) return -7;
    cudaStream_t s = (cudaStream_t)stream;
    return is_f32 ? scale_launch<float>(false, cor, ld, n, const_cast<void*>(sd), 0, cov, ldo, nullptr, s)
                  : scale_launch<double>(false, cor, ld, n, const_cast<void*>(sd), 0, cov, ldo, nullptr, s);
}

int gmg_uncompress(const void* cov, int64_t ld, const int* inv, int64_t M, int is_f32, double fill, double diag_fill,
                   void* out, int64_t ldo, gmg_stream_t stream) {
    if (!cov) return -1;
    if (ld <= 0) return -2;
    if (!inv) return -3;
    if (M <= 0) return -4;
    if (!out) return -8;
    if (ldo < M) return -9;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)gmg_ceil_div(M, 256), (unsigned)(M < 2048 ? M : 2048));
    if (is_f32)
        uncompress_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(cov), ld, inv, M, (float)fill,
                                                      (float)diag_fill, reinterpret_cast<float*>(out), ldo);
    else
        uncompress_kernel<double><<<grid, 256, 0, s>>>(reinterpret_cast<const double*>(cov), ld, inv, M, fill,
                                                       diag_fill, reinterpret_cast<double*>(out), ldo);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

}  // extern "C"
