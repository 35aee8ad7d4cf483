// Shared pieces of the register-tiled FP32-FMA kernels (rank <= 2, float).
//
// All three correlations of the iteration are evaluated in "window" form: a thread owns 8 consecutive
// positions along the fastest axis, keeps a register window of 8+AXC source values and applies AXC taps to
// it (AXC = atom-width chunk: 4, 8, 12 or 16), so one 16-byte shared-memory load feeds 8..32 FFMAs.
//
// Data movement: global -> shared with cp.async (LDGSTS), zero-filled or wrapped at the borders according to
// the reconstruction mode, so the compute loops never test a boundary.  Shared -> registers with 16-byte LDS on
// an XOR-swizzled layout (swz below) that is conflict-free for every lane arrangement the planners use:
// a quarter-warp (the unit a 16-byte LDS is served in) covers R = 1, 2, 4 or 8 consecutive tile rows with 8/R
// lanes per row at a 32-byte pitch.
#pragma once
#include "common.cuh"

namespace tnmf {
namespace tiled {

constexpr int kCols = 8;                // consecutive output positions per thread
constexpr int kMaxSmem = 220 * 1024;    // of the 227 KB a CTA may use on sm_100

// Two-dimensional view of a problem (rank-1 problems have DY = AY = TY = 1).
struct Geo2 {
    int N, C, M;
    int DY, DX, AY, AX, TY, TX;
    int offy, offx, wrap;
    long long hsn, hsm, hsy;
};

// Atom-width chunking: the atom row is cut into NK chunks of AXC taps (zero-padded to AXP = NK*AXC).
// drop = 1 when a single chunk carries exactly one dead tap, which the kernels then skip at compile time
// (odd atom widths 3, 7, 11, 15 cost no padding).
struct Chunking {
    int AXC, NK, AXP, drop;
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

inline Chunking choose_chunk(int AX) {
    int best = 4;
    double best_cost = 1e30;
    for (int c = 4; c <= 16; c += 4) {
        const int pad = round_up(AX, c);
        const int nk = pad / c;
        const int eff = (nk == 1 && pad - AX == 1) ? AX : pad;
        const double cost = eff * (1.0 + 0.1 * (8.0 + c) / c);
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    Chunking ch;
    ch.AXC = best;
    ch.AXP = round_up(AX, best);
    ch.NK = ch.AXP / best;
    ch.drop = (ch.NK == 1 && ch.AXP - AX == 1) ? 1 : 0;
    return ch;
}

inline Geo2 make_geo2(const Geo &g) {
    Geo2 q;
    q.N = g.N; q.C = g.C; q.M = g.M;
    q.DY = g.D[1]; q.DX = g.D[2]; q.AY = g.A[1]; q.AX = g.A[2]; q.TY = g.T[1]; q.TX = g.T[2];
    q.offy = g.off[1]; q.offx = g.off[2]; q.wrap = g.wrap;
    q.hsn = g.hsn; q.hsm = g.hsm; q.hsy = g.hsy;
    return q;
}

// Lane arrangement of a warp over a 2-D output: LX lanes along x (8 columns each), 32/LX lanes along y.
// Chooses the power of two that wastes the fewest lanes on the given extent.
inline int choose_lx(int EY, int EX, int min_lx) {
    int best = 32;
    double best_cost = 1e30;
    for (int lx = 32; lx >= min_lx; lx >>= 1) {
        const int ly = 32 / lx;
        const double cost = (double)round_up(EX, kCols * lx) * (double)round_up(EY, ly);
        if (cost < best_cost * 0.999) { best_cost = cost; best = lx; }
    }
    return best;
}

// ---- device helpers -----------------------------------------------------------------------------------

// Swizzle of a tile whose row pitch is a multiple of 32 floats: the 16-byte unit index within a 128-byte line
// is XORed with the line parity (bit 0) and a bit-permuted row index, so that the units {c + 2*lx} of the lanes
// of a quarter-warp never share a bank group, whatever the row alignment.
__device__ __forceinline__ int swz_row(int row) { return ((row & 1) | ((row & 2) << 1) | ((row & 4) >> 1)) << 2; }
__device__ __forceinline__ int swz(int e, int rowbits) { return e ^ ((((e >> 5) & 1) << 2) ^ rowbits); }

__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gmem_src, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 4 : 0;                       // src-size 0: the 4 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gmem_src, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// fold a logical row/column index into [0, extent); false = implicit zero
__device__ __forceinline__ bool fold(int &i, int extent, int wrap) {
    if (wrap) {
        i %= extent;
        if (i < 0) i += extent;
        return true;
    }
    return (unsigned)i < (unsigned)extent;
}

__device__ __forceinline__ void cp_async8(float *smem_dst, const float *gmem_src, int bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16n(float *smem_dst, const float *gmem_src, int bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}

// Stage a [rows x cols] window of a row-major [extent_y x extent_x] plane whose top-left logical coordinate is
// (gy0, gx0) into a swizzled shared-memory tile (row pitch `pitch` floats, a multiple of 32).  Rows are dealt to
// warps, 16-byte units to lanes.  Per row the widest cp.async the alignment of the row start allows is used
// (16, 8 or 4 bytes); elements outside the plane are zero-filled through the src-size operand ('wrap': folded).
// All threads of the block call.
__device__ __forceinline__ void stage_plane(float *dst, int pitch, const float *__restrict__ plane, int extent_y,
                                            int extent_x, long long src_pitch, int gy0, int gx0, int rows, int cols,
                                            int wrap, int warp, int n_warps, int lane) {
    const int units = (cols + 3) >> 2;
    for (int r = warp; r < rows; r += n_warps) {
        int y = gy0 + r;
        const bool row_ok = fold(y, extent_y, wrap);
        const float *src_row = plane + (long long)(row_ok ? y : 0) * src_pitch;
        float *dst_row = dst + r * pitch;
        const int rb = swz_row(r);
        if (wrap) {
            int x = (gx0 + lane) % extent_x;
            if (x < 0) x += extent_x;
            const int step = 32 % extent_x;
            for (int c = lane; c < cols; c += 32) {
                cp_async4(dst_row + swz(c, rb), src_row + x, true);
                x += step;
                if (x >= extent_x) x -= extent_x;
            }
            continue;
        }
        if (!row_ok) {
            for (int u = lane; u < units; u += 32) cp_async16n(dst_row + swz(4 * u, rb), plane, 0);
            continue;
        }
        const unsigned align = (unsigned)(reinterpret_cast<unsigned long long>(src_row + gx0) & 15ull);   // warp-uniform
        for (int u = lane; u < units; u += 32) {
            const int x = gx0 + 4 * u;
            float *d = dst_row + swz(4 * u, rb);
            const int n_in = extent_x - x;                   // elements of this unit left of the right border
            if (x >= 0 && align == 0) {
                const int bytes = n_in >= 4 ? 16 : (n_in > 0 ? 4 * n_in : 0);
                cp_async16n(d, bytes ? src_row + x : plane, bytes);
            } else if (x >= 0 && (align & 7) == 0) {
                const int b0 = n_in >= 2 ? 8 : (n_in > 0 ? 4 * n_in : 0);
                const int b1 = n_in >= 4 ? 8 : (n_in > 2 ? 4 * (n_in - 2) : 0);
                cp_async8(d, b0 ? src_row + x : plane, b0);
                cp_async8(d + 2, b1 ? src_row + x + 2 : plane, b1);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool ok = (unsigned)(x + e) < (unsigned)extent_x;
                    cp_async4(d + e, src_row + (ok ? x + e : 0), ok);
                }
            }
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

// ---- launch plans (tiled_kernels.cu) ----------------------------------------------------------------------

// Tile geometry shared by the reconstruction and the H-update kernels: a CTA of WX x WY warps, a warp of
// LX x LY lanes, a thread of RB rows x 8 columns (x CB channels / MB atoms).
struct TilePlan {
    Chunking ch;
    int LX, LY, WX, WY, RB;
    int NB;              // CB (reconstruction: channels per thread) or MB (H update: atoms per thread)
    int nblk;            // number of channel / atom blocks
    int tile_y, tile_x, tiles_y, tiles_x;
    int HR, WT, pitch;   // staged source rows / columns / row pitch (floats, multiple of 32)
    int plane_floats, taps_floats, stage_floats, n_stages;
    int threads;
    size_t smem;
    long long grid;
};

bool make_recon_plan(const Geo2 &g, TilePlan &p);
bool make_hupd_plan(const Geo2 &g, TilePlan &p);

// W gradient: the flattened (group, work item) space, group = (atom, channel block, tap-unit group), item =
// (sample, row chunk, column chunk), is dealt to `grid` CTAs in contiguous ranges of `chunk` positions.
struct GradWPlan {
    Chunking ch;
    int CB, ncb;         // channels per thread, channel blocks
    int BYB;             // tap units (atom row x atom-column chunk) per warp
    int units;           // AY * NK tap units in total
    int warps, ugroups;  // warps per CTA, tap-unit groups
    int LX, LY;          // lane arrangement over a work item
    int RY, XC;          // rows / columns per work item
    int ny, nx;          // work items per sample along y / x
    int pitch_h, pitch_x;
    int x_floats, h_floats, stage_floats;
    int threads;
    size_t smem;
    int grid;
    int groups;          // M * ncb * ugroups
    long long items;     // N * ny * nx
    long long chunk;     // positions per CTA
    int smax;            // partial slices per output element (CTAs that can share a group)
};

bool make_gradw_plan(const Geo2 &g, GradWPlan &p);

// ---- launchers: tiled_recon.cu, tiled_hupd.cu and tiled_gradw.cu are compiled once per atom-width chunk
// (-DTNMF_AXC=4|8|12|16) and each defines the specialisation of its *_launch_axc template for that chunk.
template <int AXC>
int recon_launch_axc(const Geo2 &g, const TilePlan &p, const float *W, const float *H, float *R, const float *V,
                     double *epart, cudaStream_t st);
template <int AXC>
int hupd_launch_axc(const Geo2 &g, const TilePlan &p, const float *V, const float *R, const float *W, float *neg,
                    float *pos, float *H, float reg, const float *G, float lambda, const float *Gsum,
                    float lambda_cross, cudaStream_t st);
template <int AXC>
int gradw_launch_axc(const Geo2 &g, const GradWPlan &p, const float *V, const float *R, const float *H,
                     float *partials, cudaStream_t st);

}  // namespace tiled
}  // namespace tnmf
