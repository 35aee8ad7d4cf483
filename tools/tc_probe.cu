// Probe of the tcgen05 helpers in tnmf_b200/csrc/tc_common.cuh on a real B200: one CTA, D[128 x N] = A[128 x K] * B[N x K]^T
// with 3xTF32 (hi*hi + lo*hi + hi*lo), canonical K-major no-swizzle operands, a B row-block offset, a TMEM column
// offset and a second accumulating pass.  Prints the max relative error against a double reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I tnmf_b200/csrc -I include -o tools/tc_probe tools/tc_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"

using namespace tnmf::tc;

// A: 128 x KP, B: NR x KP (row-major in global); uses B rows [jb*16, jb*16 + N); D columns at col0; terms = 1 or 3
__global__ void __launch_bounds__(128, 1) probe_kernel(const float *A, const float *B, float *out, int KP, int NR, int jb,
                                                      int N, int col0, int terms, int passes) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *a_hi = smem, *a_lo = a_hi + 128 * KP, *b_hi = a_lo + 128 * KP, *b_lo = b_hi + NR * KP;
    for (int k = 0; k < KP; ++k) {
        float hi, lo;
        split_tf32(A[tid * KP + k], hi, lo);
        a_hi[canon_offset_floats(tid, k, 128)] = hi;
        a_lo[canon_offset_floats(tid, k, 128)] = lo;
    }
    for (int idx = tid; idx < NR * KP; idx += 128) {
        const int n = idx / KP, k = idx % KP;
        float hi, lo;
        split_tf32(B[idx], hi, lo);
        b_hi[canon_offset_floats(n, k, NR)] = hi;
        b_lo[canon_offset_floats(n, k, NR)] = lo;
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = tmem_base_s;
    if (tid == 0) {
        const unsigned lbo_a = 128 * 16, lbo_b = (unsigned)NR * 16;
        const unsigned idesc = idesc_tf32(128, N);
        for (int pass = 0; pass < passes; ++pass)
            for (int ks = 0; ks < KP / 8; ++ks)
                for (int t = 0; t < terms; ++t) {
                    const float *pa = (t == 1) ? a_lo : a_hi;
                    const float *pb = (t == 2) ? b_lo : b_hi;
                    const unsigned long long da = smem_desc(smem_u32(pa) + ks * 2 * lbo_a, lbo_a, 128);
                    const unsigned long long db = smem_desc(smem_u32(pb) + jb * 256 + ks * 2 * lbo_b, lbo_b, 128);
                    mma_tf32(tmem_base + col0, da, db, idesc, (pass | ks | t) ? 1u : 0u);
                }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((unsigned)(warp * 32) << 16) + col0 + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static double run(int KP, int NR, int jb, int N, int col0, int terms, int passes) {
    std::vector<float> A(128 * KP), B(NR * KP), out(128 * N);
    for (auto &x : A) x = (float)rand() / RAND_MAX;
    for (auto &x : B) x = (float)rand() / RAND_MAX;
    float *dA, *dB, *dO;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, out.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0, out.size() * 4);
    const size_t smem = (size_t)(2 * 128 * KP + 2 * NR * KP) * 4;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dO, KP, NR, jb, N, col0, terms, passes);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < 128; ++i)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < KP; ++k) ref += (double)A[i * KP + k] * (double)B[(jb * 16 + n) * KP + k];
            ref *= passes;
            const double err = fabs(out[i * N + n] - ref) / fabs(ref);
            if (err > worst) worst = err;
        }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
    return worst;
}

// Timing: `count` MMAs of shape 128 x N x 8 issued back to back by one thread, accumulating into `ndst` alternating
// TMEM column ranges; cycles from first issue to the commit's arrival.
__global__ void __launch_bounds__(128, 1) time_kernel(long long *out, int N, int count, int ndst, int KP, int NR) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int idx = tid; idx < (128 + NR) * KP; idx += 128) smem[idx] = 1.0f;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = tmem_base_s;
    if (tid == 0) {
        const unsigned lbo_a = 128 * 16, lbo_b = (unsigned)NR * 16;
        const unsigned idesc = idesc_tf32(128, N);
        const unsigned a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * KP);
        const int ks_n = KP / 8;
        long long t0 = clock64();
        (void)ks_n;
        const unsigned long long da0 = smem_desc(a0, lbo_a, 128), db0 = smem_desc(b0, lbo_b, 128);
        const unsigned long long da1 = smem_desc(a0 + 2 * lbo_a, lbo_a, 128), db1 = smem_desc(b0 + 2 * lbo_b, lbo_b, 128);
        const unsigned d0 = tmem_base, d1 = tmem_base + (ndst > 1 ? 256u : 0u);
        mma_tf32(d0, da0, db0, idesc, 0u);
        mma_tf32(d1, da0, db0, idesc, 0u);
        for (int i = 0; i < count; i += 4) {
            mma_tf32(d0, da0, db0, idesc, 1u);
            mma_tf32(d1, da1, db1, idesc, 1u);
            mma_tf32(d0, da1, db1, idesc, 1u);
            mma_tf32(d1, da0, db0, idesc, 1u);
        }
        long long t1 = clock64();
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static void time_it(int N, int count, int ndst) {
    const int KP = 40, NR = 256;
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const size_t smem = (size_t)(128 + NR) * KP * 4;
    cudaFuncSetAttribute(time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
        time_kernel<<<1, 128, smem>>>(d, N, count, ndst, KP, NR);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d count=%4d ndst=%d : issue %6lld clk (%.1f/mma)  complete %7lld clk (%.1f/mma, ideal %.1f)\n", N, count, ndst,
           h[0], (double)h[0] / count, h[1], (double)h[1] / count, N / 2.0);
    cudaFree(d);
}

int main() {
    for (int N : {16, 32, 64, 96, 128, 176, 240, 256}) time_it(N, 400, 1);
    time_it(176, 400, 2);
    time_it(64, 400, 2);
    time_it(64, 400, 4);
    time_it(16, 400, 8);

    printf("1xTF32  KP=8   NR=16  N=16        : max rel err %.3e\n", run(8, 16, 0, 16, 0, 1, 1));
    printf("3xTF32  KP=8   NR=16  N=16        : max rel err %.3e\n", run(8, 16, 0, 16, 0, 3, 1));
    printf("3xTF32  KP=40  NR=176 N=176 col 48: max rel err %.3e\n", run(40, 176, 0, 176, 48, 3, 1));
    printf("3xTF32  KP=40  NR=176 jb=3 N=64 col 400, 2 passes: max rel err %.3e\n", run(40, 176, 3, 64, 400, 3, 2));
    printf("3xTF32  KP=16  NR=240 jb=1 N=224 col 256: max rel err %.3e\n", run(16, 240, 1, 224, 256, 3, 1));
    return 0;
}
