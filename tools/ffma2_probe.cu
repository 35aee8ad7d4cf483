// FP32 throughput of FFMA vs the packed FFMA2 (fma.rn.f32x2, sm_100) on a B200: 148 x 4 CTAs of 256 threads,
// 16 independent accumulator chains per thread, with and without a shared-memory load per 8 FMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ffma2_probe tools/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: FFMA, 1: FFMA2, 2: FFMA + LDS, 3: FFMA2 + LDS
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = a;
    __syncthreads();
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    float2 x = make_float2(a, a), y = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
        if (MODE >= 2) { const float4 v = *reinterpret_cast<const float4 *>(&sm[(threadIdx.x * 4 + it * 4) & 1020]); x.x = v.x; x.y = v.y; }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE & 1) acc[i] = __ffma2_rn(x, acc[i], y);
                else { acc[i].x = fmaf(x.x, acc[i].x, y.x); acc[i].y = fmaf(x.y, acc[i].y, y.y); }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE> void run(const char *name) {
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double flops = 2.0 * 148 * 8 * 256 * (double)iters * 4 * 16;
    printf("%-14s %.3f ms  %.1f TFLOP/s\n", name, best, flops / best / 1e9);
    cudaFree(d);
}
int main() { run<0>("FFMA"); run<1>("FFMA2"); run<2>("FFMA + LDS"); run<3>("FFMA2 + LDS"); return 0; }
