// Microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) vs FADD2 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed) {
    float2 acc[8];
    float2 w = make_float2(seed, seed * 0.5f), p = make_float2(seed * 0.25f, seed * 0.125f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) {  // scalar FFMA x2
                    acc[i].x = fmaf(w.x, p.x, acc[i].x);
                    acc[i].y = fmaf(w.y, p.y, acc[i].y);
                } else if (MODE == 1) {
                    acc[i] = __ffma2_rn(w, p, acc[i]);
                } else {
                    acc[i] = __fadd2_rn(acc[i], p);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, float *out) {
    const int grid = 148 * 8, iters = 2000;
    k<MODE><<<grid, 256>>>(out, 10, 1.0f);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, 256>>>(out, iters, 1.0f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double fma = (double)grid * 256 * iters * 16 * 8 * 2;
    printf("%-8s %.3f ms  %.2f T(fma or add)/s  = %.2f TFLOP/s-equivalent\n", name, ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("FFMA", out);
    run<1>("FFMA2", out);
    run<2>("FADD2", out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
