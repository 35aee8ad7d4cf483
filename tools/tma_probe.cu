// Stand-alone probe: one TMA box load per launch, for a list of box shapes / coordinates (debug aid).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3, unsigned bytes, float *out, int n_out) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
                     ::"r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < n_out; i += blockDim.x) out[i] = smem[i];
}
int main(int argc, char **argv) {
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int DX = 72, DY = 40, C = 3, N = 2;
    std::vector<float> h((size_t)DX * DY * C * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 1 << 20);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct T { unsigned bx, by, bc; int c0, c1, c2, c3; } all[] = {
        {44, 66, 1, 2, 0, 0, 0}, {44, 66, 1, 1, 3, 0, 0}, {44, 66, 1, 0, -10, 0, 0}, {44, 66, 1, -4, 0, 0, 0}, {44, 66, 1, -8, -10, 0, 0},
        {44, 66, 1, -10, 0, 0, 0}, {44, 66, 1, -1, 0, 0, 0}, {44, 66, 1, -12, -3, 1, 1}, {44, 66, 1, 50, 30, 0, 0}, {44, 66, 1, 80, 50, 0, 0}};
    int which = argc > 1 ? atoi(argv[1]) : 0;
    T tests[1] = {all[which]};
    for (auto &t : tests) {
        CUtensorMap map;
        cuuint64_t dims[4] = {DX, DY, C, N}, str[3] = {DX * 4ull, DX * DY * 4ull, DX * DY * C * 4ull};
        cuuint32_t box[4] = {t.bx, t.by, t.bc, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        unsigned bytes = t.bx * t.by * t.bc * 4;
        probe<<<1, 64, 200 * 1024>>>(map, t.c0, t.c1, t.c2, t.c3, bytes, out, 64);
        cudaError_t e = cudaDeviceSynchronize();
        float o[4] = {0, 0, 0, 0};
        if (e == cudaSuccess) cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
        printf("box [%u,%u,%u,1] at (%d,%d,%d,%d) bytes %u: encode %d run %s first %g %g\n", t.bx, t.by, t.bc, t.c0, t.c1, t.c2,
               t.c3, bytes, (int)r, cudaGetErrorString(e), o[0], o[1]);
        if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 1; }
    }
    return 0;
}
