// Probe of tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form) on a real B200.
//   1. numerics: D[128 x N] = A[128 x K] * B[N x K]^T with 3xTF32, A written by tcgen05.st (lane = row, column = k), at an
//      arbitrary (also odd) TMEM column, B canonical K-major no-swizzle in shared memory;
//   2. timing: back-to-back MMAs of shape 128 x N x 8, A from TMEM against A from shared memory, alone and while eight
//      other warps hammer shared memory with conflict-free LDS (what the expander warps of the kernels do).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I tnmf_b200/csrc -I include -o tools/tc_probe2 tools/tc_probe2.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"

using namespace tnmf::tc;

__device__ __forceinline__ void tmem_st1(unsigned addr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};\n" ::"r"(addr), "r"(__float_as_uint(v)) : "memory");
}

// A: 128 x KP (row-major global), written to TMEM columns [acol, acol + KP) (hi) and [acol + KP, acol + 2 KP) (lo);
// the MMA reads it from column acol + shift, i.e. it multiplies A[:, shift : shift + KP'] with KP' = KP - shift rounded
// down to 8.  D goes to columns [0, N).
__global__ void __launch_bounds__(128, 1) ts_kernel(const float *A, const float *B, float *out, int KP, int N, int acol,
                                                    int shift) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *b_hi = smem, *b_lo = b_hi + N * KP;
    for (int idx = tid; idx < N * KP; idx += 128) {
        const int n = idx / KP, k = idx % KP;
        float hi, lo;
        split_tf32(B[idx], hi, lo);
        b_hi[canon_offset_floats(n, k, N)] = hi;
        b_lo[canon_offset_floats(n, k, N)] = lo;
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = tmem_base_s;
    const unsigned lane_base = (unsigned)(warp * 32) << 16;
    for (int k = 0; k < KP; ++k) {
        float hi, lo;
        split_tf32(A[tid * KP + k], hi, lo);
        tmem_st1(tmem_base + lane_base + (unsigned)(acol + k), hi);
        tmem_st1(tmem_base + lane_base + (unsigned)(acol + KP + k), lo);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int KE = (KP - shift) / 8 * 8;
    if (tid == 0) {
        const unsigned lbo_b = (unsigned)N * 16;
        const unsigned idesc = idesc_tf32(128, N);
        for (int ks = 0; ks < KE / 8; ++ks)
            for (int t = 0; t < 3; ++t) {
                const unsigned ta = tmem_base + (unsigned)(acol + shift + (t == 1 ? KP : 0) + 8 * ks);
                const float *pb = (t == 2) ? b_lo : b_hi;
                const unsigned long long db = smem_desc(smem_u32(pb) + ks * 2 * lbo_b, lbo_b, 128);
                mma_tf32_ts(tmem_base, ta, db, idesc, (ks | t) ? 1u : 0u);
            }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + lane_base + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static double run_ts(int KP, int N, int acol, int shift) {
    std::vector<float> A(128 * KP), B(N * KP), out(128 * N);
    for (auto &x : A) x = (float)rand() / RAND_MAX;
    for (auto &x : B) x = (float)rand() / RAND_MAX;
    float *dA, *dB, *dO;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, out.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0, out.size() * 4);
    const size_t smem = (size_t)(2 * N * KP) * 4;
    cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ts_kernel<<<1, 128, smem>>>(dA, dB, dO, KP, N, acol, shift);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    const int KE = (KP - shift) / 8 * 8;
    double worst = 0;
    for (int i = 0; i < 128; ++i)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < KE; ++k) ref += (double)A[i * KP + shift + k] * (double)B[n * KP + k];
            const double err = fabs(out[i * N + n] - ref) / fabs(ref);
            if (err > worst) worst = err;
        }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
    return worst;
}

// Timing.  Warp 8 issues `count` MMAs of 128 x N x 8 (one elected lane, uniform operands) into ndst alternating TMEM ranges; A
// from TMEM (ts = 1) or shared memory (ts = 0).  Warps 0..7 meanwhile generate the traffic named by `hammer` (bit 0: LDS.128,
// bit 1: STS.128, bit 2: tcgen05.st x16, bit 3: tcgen05.ld x16 - the latter two on TMEM columns the MMAs do not touch) until
// the issuer is done.  Reports cycles per MMA and the hammer rates.
__global__ void __launch_bounds__(32 * 9, 1) time_kernel(long long *out, int N, int count, int ndst, int ts, int hammer,
                                                        float *sink) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    __shared__ volatile int done;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int KP = 16, NR = 256;
    for (int idx = tid; idx < (128 + NR) * KP + 8192; idx += blockDim.x) smem[idx] = 1.0f;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); done = 0; }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);
    if (warp < 4) {                                             // A in TMEM columns [448, 464): ones
        for (int c = 0; c < 16; ++c) tmem_st1(tmem_base + ((unsigned)(warp * 32) << 16) + 448u + c, 1.0f);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 8) {
        const unsigned lbo_a = 128 * 16, lbo_b = (unsigned)NR * 16;
        const unsigned idesc = idesc_tf32(128, N);
        const unsigned a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * KP);
        const unsigned long long da0 = smem_desc(a0, lbo_a, 128), db0 = smem_desc(b0, lbo_b, 128);
        const unsigned long long da1 = smem_desc(a0 + 2 * lbo_a, lbo_a, 128), db1 = smem_desc(b0 + 2 * lbo_b, lbo_b, 128);
        const unsigned ta0 = tmem_base + 448u, ta1 = tmem_base + 456u;
        const unsigned d0 = tmem_base, d1 = tmem_base + (ndst > 1 ? 192u : 0u);
        long long t0 = clock64();
        if (elect_one()) {
            if (ts) {
                for (int i = 0; i < count; i += 4) {
                    mma_tf32_ts(d0, ta0, db0, idesc, 1u);
                    mma_tf32_ts(d1, ta1, db1, idesc, 1u);
                    mma_tf32_ts(d0, ta1, db1, idesc, 1u);
                    mma_tf32_ts(d1, ta0, db0, idesc, 1u);
                }
            } else {
                for (int i = 0; i < count; i += 4) {
                    mma_tf32(d0, da0, db0, idesc, 1u);
                    mma_tf32(d1, da1, db1, idesc, 1u);
                    mma_tf32(d0, da1, db1, idesc, 1u);
                    mma_tf32(d1, da0, db0, idesc, 1u);
                }
            }
        }
        __syncwarp();
        long long t1 = clock64();
        mma_commit_elect(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if ((tid & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; done = 1; }
    } else {
        volatile float4 *base = reinterpret_cast<volatile float4 *>(smem + (128 + NR) * KP) + (tid & 31);
        const unsigned tcol = tmem_base + ((unsigned)((warp & 3) * 32) << 16) + 384u + (unsigned)((warp >> 2) * 16);
        float acc = 0.f;
        long long n = 0;
        float v16[16];
        for (int j = 0; j < 16; ++j) v16[j] = (float)j;
        while (!done) {
            if (hammer & 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc += base[32 * ((j + warp) & 63)].x; }
            }
            if (hammer & 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { base[32 * ((j + warp) & 63)].y = acc; }
            }
            if (hammer & 4) { tmem_st16(tcol, v16); tmem_st_wait(); }
            if (hammer & 8) { tmem_ld16(tcol, v16); tmem_ld_wait(); acc += v16[3]; }
            n += 1;
        }
        if (acc == 123.f) sink[tid] = acc;
        if ((tid & 31) == 0) out[2 + warp] = n;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static void time_it(int N, int count, int ndst, int ts, int hammer) {
    long long *d, h[12] = {0};
    float *sink;
    cudaMalloc(&d, sizeof(h));
    cudaMalloc(&sink, 4096);
    cudaMemset(d, 0, sizeof(h));
    const size_t smem = (size_t)((128 + 256) * 16 + 8192) * 4;
    cudaFuncSetAttribute(time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
        time_kernel<<<1, 32 * 9, smem>>>(d, N, count, ndst, ts, hammer, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double loops = 0;
    for (int w = 0; w < 8; ++w) loops += (double)h[2 + w];
    printf("%s N=%3d ndst=%d hammer=%2d : issue %.1f clk/mma  complete %.1f clk/mma (ideal %.1f)   hammer loops/kclk (8 warps) %.1f\n",
           ts ? "TS" : "SS", N, ndst, hammer, (double)h[0] / count, (double)h[1] / count, N / 2.0, 1e3 * loops / (double)h[1]);
    cudaFree(d);
    cudaFree(sink);
}

int main(int argc, char **argv) {
    // every case in its own process (a faulting case poisons the context):  num KP N acol shift | time N ndst ts hammer
    if (argc >= 6 && argv[1][0] == 'n') {
        const int KP = atoi(argv[2]), N = atoi(argv[3]), acol = atoi(argv[4]), shift = atoi(argv[5]);
        printf("3xTF32 A-in-TMEM  KP=%d N=%d acol %d shift %d: ", KP, N, acol, shift);
        fflush(stdout);
        printf("max rel err %.3e\n", run_ts(KP, N, acol, shift));
        return 0;
    }
    if (argc >= 6 && argv[1][0] == 't') {
        time_it(atoi(argv[2]), 400, atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
        return 0;
    }
    return 1;
}
