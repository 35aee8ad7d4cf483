// Shared pieces of the persistent, warp-specialised TMA kernels (tma_hupd.cu, tma_recon.cu, tma_gradw.cu).
//
// Structure common to the three kernels (rank <= 2, float, 'valid' / 'full' modes):
//   * one persistent CTA per SM walks a static list of work units (u = blockIdx.x, += gridDim.x);
//   * warp `n_consumers` is the producer: one elected lane waits for a free ring stage (mbarrier `empty`), arms the
//     stage's `full` mbarrier with the byte count and issues cp.async.bulk.tensor (TMA) box loads of the source
//     tiles - halo included, the zero boundary of the correlation supplied by TMA's out-of-bounds fill - plus one
//     cp.async.bulk copy of the pre-arranged atom slice;
//   * the consumer warps wait on `full`, run LDS.128 + FFMA only (no address arithmetic beyond one pointer bump per
//     atom row, no boundary tests), and release the stage with one mbarrier arrive per warp.
// The ring keeps flowing across work units, so the loads of the next unit overlap the FFMAs and the epilogue of the
// current one; nothing in the steady state executes a __syncthreads().
//
// TMA constraint measured on B200 (tools/tma_probe.cu): with no swizzle / no interleave the innermost box coordinate
// must be a multiple of 16 bytes (4 floats), negative or not, or the load traps with 'illegal instruction'; the other
// coordinates are free.  The planners therefore place tile origins so that every box starts on a multiple of 4.
//
// Shared-memory tiles are dense boxes [rows][pitch] with pitch a multiple of 4 floats and pitch/4 odd: a
// quarter-warp of 4 lanes x 2 rows (8 lanes, the unit a 16-byte LDS is served in) then reads 8 distinct bank groups.
#pragma once
#include <cuda.h>
#include "tiled_common.cuh"

namespace tnmf {
namespace tma {

using tiled::Chunking;
using tiled::Geo2;
using tiled::kCols;

constexpr int kLX = 4, kLY = 8;          // lanes of a warp: 4 along x (8 columns each) x 8 along y
constexpr int kConsumersMax = 12;     // consumer warps of a CTA (+ 1 producer warp = 416 threads, <= 152 registers)
constexpr int kMaxSmem = 224 * 1024;

// ---- device: mbarrier / TMA primitives (PTX ISA 8.x, sm_90+) ----------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TNMF_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TNMF_DONE;\n"
        "bra TNMF_WAIT;\n"
        "TNMF_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// Producer-side wait: the ring gives the producer several stage periods of slack, so it polls at a low rate instead
// of competing with the consumer warps of its scheduler for issue slots.
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long *bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(256);
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
            "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];\n" ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// plain bulk copy global -> shared (16-byte aligned addresses, size a multiple of 16)
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                     "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

// Ring position shared by the producer and the consumers: stage index and the parity of the current pass.
struct Ring {
    int stage = 0;
    unsigned phase = 0;
    __device__ __forceinline__ void advance(int n_stages) {
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
    }
};

__device__ __forceinline__ float4 lds128(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// ---- packed FP32 pairs (sm_100: fma.rn.f32x2 = FFMA2, two FMAs per issue slot) ------------------------------------------
// A pair lives in one 64-bit register (an aligned register pair).  pk2_fma multiplies a pair by a scalar - ptxas encodes
// the {w, w} operand as a broadcast selector (R.F32), no duplication is executed.  Pairs must come out of 8/16-byte
// loads as they are: a pair assembled from two unrelated registers is re-created by ptxas with two MOVs in front of
// every FFMA2 that uses it (measured on the reconstruction and on the W gradient: it cancels the packing).
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk2_make(float lo, float hi) {
    pk2 r;
    asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ pk2 pk2_fma(pk2 x, float w, pk2 c) {
    pk2 d;
    asm("{\n.reg .b64 t;\nmov.b64 t, {%2, %2};\nfma.rn.f32x2 %0, %1, t, %3;\n}\n" : "=l"(d) : "l"(x), "f"(w), "l"(c));
    return d;
}
__device__ __forceinline__ float pk2_lo(pk2 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;\n" : "=f"(lo), "=f"(hi) : "l"(v));
    (void)hi;
    return lo;
}
__device__ __forceinline__ float pk2_hi(pk2 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;\n" : "=f"(lo), "=f"(hi) : "l"(v));
    (void)lo;
    return hi;
}

// ---- host: tensor maps -------------------------------------------------------------------------------------------
// Encodes a float32 tiled tensor map: dims / box are listed fastest axis first, strides (bytes) belong to axes 1..rank-1
// and must be multiples of 16, the base address must be 16-byte aligned.  Out-of-bounds elements read as zero.
int encode_map(CUtensorMap *map, const void *base, int rank, const unsigned long long *dims,
               const unsigned long long *strides_bytes, const unsigned *box);
bool aligned16(const void *p);
int sm_count();

inline int odd_pitch(int floats) {       // smallest pitch >= floats with pitch % 4 == 0 and (pitch / 4) odd
    int p = tiled::round_up(floats, 4);
    if (((p >> 2) & 1) == 0) p += 4;
    return p;
}

// ---- launch plans (tma_kernels.cu) -----------------------------------------------------------------------------------
struct HupdPlan {
    Chunking ch;
    int MB, nblk;                 // atoms per thread, atom blocks
    int WX, WY, consumers;        // consumer warps of a CTA: WX x WY, each 32 columns x 8 rows
    int tile_y, tile_x, tiles_y, tiles_x;
    int HR, pitch;                // staged rows, row pitch (floats)
    int wide;                     // 1: the 9-warp / 224-register build of the kernel (hupd_needs_wide)
    int x_shift;                  // tile origin along x: -3..0, makes the TMA box start a multiple of 4 elements
    int plane_floats, taps_floats, stage_floats, n_stages;
    long long units;              // (sample, tile, atom block)
    int grid, threads;
    size_t smem;
};
struct ReconPlan {
    Chunking ch;
    int CB, nblk, RB;             // channels per thread, channel blocks, rows per thread
    int WX, WY, consumers;
    int tile_y, tile_x, tiles_y, tiles_x;
    int HR, pitch;
    int plane_floats, taps_floats, stage_floats, n_stages;
    long long units;              // (sample, tile, channel block)
    int grid, threads;
    size_t smem;
};
struct GradWPlan {
    Chunking ch;
    int CB, ncb;                  // channels per thread, channel blocks
    int BYB;                      // tap units (atom row x column chunk) per warp
    int units;                    // AY * NK
    int UW, RW, consumers;        // warps along the tap units x warps along the rows of a work item
    int ugroups;                  // tap-unit groups (passes over the data when one CTA cannot hold all units)
    int RY, XC, ny, nx;           // rows / columns of a work item, items per sample
    int by_span;                  // extra H rows a unit group needs
    int pitch_x, pitch_h, hrows;
    int x_floats, h_floats, stage_floats, n_stages;
    int groups;                   // M * ncb * ugroups
    long long items;              // N * ny * nx
    long long chunk;              // positions of the flattened (group, item) space per CTA
    int smax;                     // partial slices per output element
    int grid, threads;
    size_t smem;
};

// accumulators 16*MB + two windows 2*(8+AXC) + ~30 for taps and addressing: beyond ~136 the 128-register build spills
constexpr int hupd_needs_wide(int axc, int mb) { return 16 * mb + 2 * (kCols + axc) + 30 > 136 ? 1 : 0; }
bool make_hupd_plan(const Geo2 &g, HupdPlan &p);
bool make_recon_plan(const Geo2 &g, ReconPlan &p);
bool make_gradw_plan(const Geo2 &g, GradWPlan &p);

// ---- launchers: one specialisation per atom-width chunk (-DTNMF_AXC=4|8|12|16) -----------------------------------
struct HupdArgs {
    const float *Wt;              // atom slices pre-arranged by prepare_taps_hupd: [c][mb][ay][q][i][4]
    float *neg, *pos, *H;
    float reg, lambda, lambda_cross;
    const float *G, *Gsum;
};
template <int AXC>
int hupd_launch_axc(const Geo2 &g, const HupdPlan &p, const CUtensorMap &mapV, const CUtensorMap &mapR,
                    const HupdArgs &a, cudaStream_t st);

struct ReconArgs {
    const float *Wt;              // flipped atom slices pre-arranged by prepare_taps_recon: [m][cb][by][q][c][4]
    float *R;
    const float *V;
    double *epart;                // grid * consumers partial energies (or null)
};
template <int AXC>
int recon_launch_axc(const Geo2 &g, const ReconPlan &p, const CUtensorMap &mapH, const ReconArgs &a, cudaStream_t st);

template <int AXC>
int gradw_launch_axc(const Geo2 &g, const GradWPlan &p, const CUtensorMap &mapV, const CUtensorMap &mapR,
                     const CUtensorMap &mapH, float *partials, cudaStream_t st);

}  // namespace tma
}  // namespace tnmf
