// Shared pieces of the register-tiled FP32-FMA kernels (rank <= 2, float).
//
// Data movement: global -> shared with cp.async (LDGSTS, 4-byte granules because the halo origin of a tile is
// not 16-byte aligned in general), zero-filled or wrapped at the borders according to the reconstruction mode,
// so the compute loops never test a boundary.  Shared -> registers with 16-byte LDS on an XOR-swizzled row
// layout that is conflict-free for "8 consecutive lanes read 16 bytes at a 32-byte pitch".
#pragma once
#include "common.cuh"

namespace tnmf {
namespace tiled {

constexpr int kCols = 8;          // consecutive output columns per thread
constexpr int kMaxSmem = 200 * 1024;

// Two-dimensional view of a problem (rank-1 problems have DY = AY = TY = 1).
struct Geo2 {
    int N, C, M;
    int DY, DX, AY, AX, TY, TX;
    int offy, offx, wrap;
    long long hsn, hsm;
    int AXC;      // atom-width chunk handled per register window (4, 8, 12 or 16)
    int NK;       // number of chunks
    int AXP;      // padded atom width = NK * AXC
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// atom-width chunk: trade padded (wasted) taps against the shared-memory loads a narrow register window costs
inline void choose_chunk(int AX, int &AXC, int &NK) {
    int best = 4;
    double best_cost = 1e30;
    for (int c = 4; c <= 16; c += 4) {
        const int pad = round_up(AX, c);
        const double cost = pad * (1.0 + 0.1 * (8.0 + c) / c);
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    AXC = best;
    NK = round_up(AX, best) / best;
}

inline Geo2 make_geo2(const Geo &g) {
    Geo2 q;
    q.N = g.N; q.C = g.C; q.M = g.M;
    q.DY = g.D[1]; q.DX = g.D[2]; q.AY = g.A[1]; q.AX = g.A[2]; q.TY = g.T[1]; q.TX = g.T[2];
    q.offy = g.off[1]; q.offx = g.off[2]; q.wrap = g.wrap;
    q.hsn = g.hsn; q.hsm = g.hsm;
    choose_chunk(q.AX, q.AXC, q.NK);
    q.AXP = q.AXC * q.NK;
    return q;
}

// ---- device helpers -----------------------------------------------------------------------------------

// element index within a row -> swizzled element index (16-byte units, unit ^= bit 3 of the unit index)
__device__ __forceinline__ int swz(int e) { return e ^ (((e >> 5) & 1) << 2); }

__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gmem_src, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 4 : 0;                       // src-size 0: the 4 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// fold a logical row/column index into [0, extent); false = implicit zero
__device__ __forceinline__ bool fold(int &i, int extent, int wrap) {
    if (wrap) {
        if (i < 0) i += extent * ((-i + extent - 1) / extent);
        else if (i >= extent) i %= extent;
        return true;
    }
    return (unsigned)i < (unsigned)extent;
}

// Stage a [rows x cols] window of a row-major [extent_y x extent_x] plane whose top-left logical coordinate is
// (gy0, gx0) into shared memory (row pitch `pitch` floats, swizzled when SWZ).  All threads of the block call.
template <bool SWZ>
__device__ __forceinline__ void stage_plane(float *dst, int pitch, const float *plane, int extent_y, int extent_x,
                                            int gy0, int gx0, int rows, int cols, int wrap, int warp, int n_warps,
                                            int lane) {
    for (int r = warp; r < rows; r += n_warps) {
        int y = gy0 + r;
        const bool row_ok = fold(y, extent_y, wrap);
        const float *src_row = plane + (long long)(row_ok ? y : 0) * extent_x;
        float *dst_row = dst + r * pitch;
        for (int c = lane; c < cols; c += 32) {
            int x = gx0 + c;
            const bool ok = fold(x, extent_x, wrap) && row_ok;
            cp_async4(dst_row + (SWZ ? swz(c) : c), src_row + (ok ? x : 0), ok);
        }
    }
}

__device__ __forceinline__ float4 lds128(const float *p) { return *reinterpret_cast<const float4 *>(p); }

__device__ __forceinline__ double block_sum(double v, double *red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v;
}

// ---- per-operation planners and launchers (one translation unit each) ----------------------------------
bool recon_plan_ok(const Geo2 &g);
int recon_launch(const Geo2 &g, const float *W, const float *H, float *R, const float *V, double *energy_partials,
                 int *n_partials, cudaStream_t st);
long long recon_grid(const Geo2 &g);

bool hupd_plan_ok(const Geo2 &g);
int hupd_launch(const Geo2 &g, const float *V, const float *R, const float *W, float *neg, float *pos, float *H,
                float reg, const float *G, float lambda, const float *Gsum, float lambda_cross, cudaStream_t st);

bool gradw_plan_ok(const Geo2 &g);
size_t gradw_workspace_bytes(const Geo2 &g);
int gradw_launch(const Geo2 &g, const float *V, const float *R, const float *H, float *neg, float *pos,
                 void *workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace tiled
}  // namespace tnmf
