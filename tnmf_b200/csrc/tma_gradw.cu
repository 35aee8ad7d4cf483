// Persistent TMA kernel: W gradient, a split-K reduction over samples and positions.
//
//   neg[m,c,ay,ax] = sum_n sum_{y,x} Hext[n,m,y+offy-ay,x+offx-ax] * V[n,c,y,x]          (tnmf/backends/NumPy.py:77-85)
//   pos[m,c,ay,ax] = the same with R                                                     (tnmf/backends/NumPy.py:80,87-90)
//
// The accumulators are indexed by the tap: consumer warp (uw, rw) owns BYB tap units (atom row x chunk of AXC atom
// columns) and the rows [rw*RY/RW, (rw+1)*RY/RW) of every work item; a lane owns strips of 8 consecutive sample
// columns.  Per strip the lane loads 8 values of V and of R for CB channels and, per tap unit, a register window of
// 8+AXC activations, then issues 2*CB*AXC*8 FFMAs into accumulators that stay live across all work items of a group.
// Work item = (sample, RY rows, XC columns) = one ring stage: V and R strips (two 4-D TMA boxes over the channel
// block) and the H tile with halo (one 4-D TMA box, zero outside [0,T)).  The flattened (group, item) space,
// group = (atom, channel block, tap-unit group), is dealt to the persistent CTAs in equal contiguous ranges; at a
// group boundary every consumer warp reduces its accumulators over the lanes with shuffles and writes one partial
// slice; finish_gradient_w sums the slices in a fixed order in double (deterministic: no atomics).
// Bound: FP32 FMA pipe.
//
// Compiled once per atom-width chunk: -DTNMF_AXC=4|8|12|16.
#include "tma_common.cuh"

#ifndef TNMF_AXC
#error "compile with -DTNMF_AXC=4|8|12|16"
#endif

namespace tnmf {
namespace tma {

template <int AXC, int DROP, int CB, int BYB>
__global__ void __launch_bounds__(32 * 12, 1)
gradw_tma_kernel(const Geo2 g, const GradWPlan p, const __grid_constant__ CUtensorMap mapV,
                 const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapH,
                 float *__restrict__ partials) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long full_bar[8], empty_bar[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_stages = p.n_stages;
    const int NK = p.ch.NK;

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], (unsigned)p.consumers);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const long long total = (long long)p.groups * p.items;
    const long long pos0 = (long long)blockIdx.x * p.chunk;
    const long long pos1 = pos0 + p.chunk < total ? pos0 + p.chunk : total;
    if (pos0 >= pos1) return;

    // first atom row of unit group ug
    auto by_lo_of = [&](int ug) { return (ug * p.UW * BYB) / NK; };

    // position in the flattened (group, item) space, advanced without divisions
    struct Cursor {
        int grp, m, cb, ug, n, yc, xc;
    };
    auto decode = [&](long long pos) {
        Cursor c;
        c.grp = (int)(pos / p.items);
        const long long it = pos - (long long)c.grp * p.items;
        c.ug = c.grp % p.ugroups;
        c.cb = (c.grp / p.ugroups) % p.ncb;
        c.m = c.grp / (p.ugroups * p.ncb);
        c.xc = (int)(it % p.nx);
        c.yc = (int)((it / p.nx) % p.ny);
        c.n = (int)(it / ((long long)p.nx * p.ny));
        return c;
    };
    auto advance = [&](Cursor &c) {
        if (++c.xc < p.nx) return;
        c.xc = 0;
        if (++c.yc < p.ny) return;
        c.yc = 0;
        if (++c.n < g.N) return;
        c.n = 0;
        ++c.grp;
        if (++c.ug < p.ugroups) return;
        c.ug = 0;
        if (++c.cb < p.ncb) return;
        c.cb = 0;
        ++c.m;
    };

    Ring ring;
    const unsigned stage_bytes = (unsigned)((p.x_floats + p.hrows * p.pitch_h) * sizeof(float));

    if (warp == p.consumers) {
        // ---------------- producer ----------------
        if (lane == 0) {
            prefetch_map(&mapV);
            prefetch_map(&mapR);
            prefetch_map(&mapH);
            Cursor cur = decode(pos0);
            const int g0y = g.offy - (g.AY - 1), g0x = g.offx - (g.AX - 1);
            for (long long pos = pos0; pos < pos1; ++pos) {
                mbar_wait_relaxed(&empty_bar[ring.stage], ring.phase ^ 1u);
                float *sx = smem + (size_t)ring.stage * p.stage_floats;
                float *sh = sx + p.x_floats;
                const int y_base = cur.yc * p.RY, x_base = cur.xc * p.XC;
                mbar_arrive_expect_tx(&full_bar[ring.stage], stage_bytes);
                tma_load_4d(sx, &mapV, &full_bar[ring.stage], x_base, y_base, cur.cb * CB, cur.n);
                tma_load_4d(sx + p.x_floats / 2, &mapR, &full_bar[ring.stage], x_base, y_base, cur.cb * CB, cur.n);
                tma_load_4d(sh, &mapH, &full_bar[ring.stage], x_base + g0x, y_base + g0y + by_lo_of(cur.ug), cur.m,
                            cur.n);
                ring.advance(n_stages);
                advance(cur);
            }
        }
        return;
    }
    if (warp > p.consumers) return;

    // ---------------- consumers ----------------
    const int rw = warp / p.UW, uw = warp - rw * p.UW;
    const int ly = lane / kLX, lx = lane - ly * kLX;
    const int ryw = p.RY / p.RW;                                  // rows of a work item this warp visits
    const long long count = (long long)g.M * g.C * g.AY * g.AX;

    float acc[BYB][2][CB][AXC];
#pragma unroll
    for (int i = 0; i < BYB; ++i)
#pragma unroll
        for (int X = 0; X < 2; ++X)
#pragma unroll
            for (int c = 0; c < CB; ++c)
#pragma unroll
                for (int bx = 0; bx < AXC; ++bx) acc[i][X][c][bx] = 0.f;

    auto flush = [&](int grp) {
        // sum over the lanes, then every lane stores the values whose index is congruent to it
        const int ug = grp % p.ugroups;
        const int cb = (grp / p.ugroups) % p.ncb;
        const int m = grp / (p.ugroups * p.ncb);
        const long long first_cta = ((long long)grp * p.items) / p.chunk;
        float *slot = partials + (((long long)blockIdx.x - first_cta) * p.RW + rw) * 2 * count;
        const int ubase = (ug * p.UW + uw) * BYB;
#pragma unroll
        for (int i = 0; i < BYB; ++i) {
            const int u = ubase + i;
            const int by = u / NK, k = u % NK;
#pragma unroll
            for (int X = 0; X < 2; ++X)
#pragma unroll
                for (int c = 0; c < CB; ++c)
#pragma unroll
                    for (int bx = 0; bx < AXC; ++bx) {
                        float v = acc[i][X][c][bx];
                        acc[i][X][c][bx] = 0.f;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        const int flat = ((i * 2 + X) * CB + c) * AXC + bx;
                        const int ax = g.AX - 1 - (k * AXC + bx);
                        const int ch = cb * CB + c;
                        if (lane == (flat & 31) && u < p.units && ax >= 0 && ch < g.C) {
                            const int ay = g.AY - 1 - by;
                            slot[X * count + (((long long)m * g.C + ch) * g.AY + ay) * g.AX + ax] = v;
                        }
                    }
        }
    };

    Cursor cur = decode(pos0);
    int cur_grp = cur.grp;
    const int plane_x = p.RY * p.pitch_x;                         // floats of one (X, c) strip plane
    for (long long pos = pos0; pos < pos1; ++pos) {
        if (pos != pos0) advance(cur);
        if (cur.grp != cur_grp) {
            flush(cur_grp);
            cur_grp = cur.grp;
        }
        mbar_wait(&full_bar[ring.stage], ring.phase);
        const int ug = cur.ug;
        const int by_lo = by_lo_of(ug);
        const int ubase = (ug * p.UW + uw) * BYB;
        int hoff[BYB];                                            // window origin of tap unit i inside the H tile
        bool valid[BYB];
#pragma unroll
        for (int i = 0; i < BYB; ++i) {
            const int u = ubase + i;
            valid[i] = u < p.units;
            hoff[i] = (u / NK - by_lo) * p.pitch_h + (u % NK) * AXC;
        }
        if (valid[0]) {
            const float *sx = smem + (size_t)ring.stage * p.stage_floats;
            const float *sh = sx + p.x_floats;
            const int rows_here = min(p.RY, g.DY - cur.yc * p.RY);
            const int r_end = min((rw + 1) * ryw, rows_here);
            const int strips_here = (min(p.XC, g.DX - cur.xc * p.XC) + kCols - 1) / kCols;
            for (int r = rw * ryw + ly; r < r_end; r += kLY) {
                const float *xrow = sx + r * p.pitch_x;
                const float *hrow = sh + r * p.pitch_h;
                for (int s = lx; s < strips_here; s += kLX) {
                    const float *xs = xrow + kCols * s;
                    const float *hs = hrow + kCols * s;
                    if constexpr (BYB == 1) {
                        // one tap unit: keep the activation window, stream V then R through the registers
                        float win[kCols + AXC];
#pragma unroll
                        for (int q = 0; q < (kCols + AXC) / 4; ++q) {
                            const float4 v = lds128(hs + hoff[0] + 4 * q);
                            win[4 * q] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
                        }
#pragma unroll
                        for (int X = 0; X < 2; ++X) {
                            float xv[CB][kCols];
#pragma unroll
                            for (int c = 0; c < CB; ++c) {
                                const float4 a4 = lds128(xs + (X * CB + c) * plane_x);
                                const float4 b4 = lds128(xs + (X * CB + c) * plane_x + 4);
                                xv[c][0] = a4.x; xv[c][1] = a4.y; xv[c][2] = a4.z; xv[c][3] = a4.w;
                                xv[c][4] = b4.x; xv[c][5] = b4.y; xv[c][6] = b4.z; xv[c][7] = b4.w;
                            }
#pragma unroll
                            for (int bx = 0; bx < AXC - DROP; ++bx)
#pragma unroll
                                for (int j = 0; j < kCols; ++j)
#pragma unroll
                                    for (int c = 0; c < CB; ++c)
                                        acc[0][X][c][bx] = fmaf(win[j + bx], xv[c][j], acc[0][X][c][bx]);
                        }
                    } else {
                        // several tap units: keep V and R, stream the activation windows
                        float xv[2][CB][kCols];
#pragma unroll
                        for (int X = 0; X < 2; ++X)
#pragma unroll
                            for (int c = 0; c < CB; ++c) {
                                const float4 a4 = lds128(xs + (X * CB + c) * plane_x);
                                const float4 b4 = lds128(xs + (X * CB + c) * plane_x + 4);
                                xv[X][c][0] = a4.x; xv[X][c][1] = a4.y; xv[X][c][2] = a4.z; xv[X][c][3] = a4.w;
                                xv[X][c][4] = b4.x; xv[X][c][5] = b4.y; xv[X][c][6] = b4.z; xv[X][c][7] = b4.w;
                            }
#pragma unroll
                        for (int i = 0; i < BYB; ++i) {
                            if (!valid[i]) continue;                    // warp-uniform
                            float win[kCols + AXC];
#pragma unroll
                            for (int q = 0; q < (kCols + AXC) / 4; ++q) {
                                const float4 v = lds128(hs + hoff[i] + 4 * q);
                                win[4 * q] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
                            }
#pragma unroll
                            for (int bx = 0; bx < AXC - DROP; ++bx)
#pragma unroll
                                for (int j = 0; j < kCols; ++j)
#pragma unroll
                                    for (int X = 0; X < 2; ++X)
#pragma unroll
                                        for (int c = 0; c < CB; ++c)
                                            acc[i][X][c][bx] = fmaf(win[j + bx], xv[X][c][j], acc[i][X][c][bx]);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[ring.stage]);
        ring.advance(n_stages);
    }
    flush(cur_grp);
}

template <int AXC, int DROP, int CB, int BYB>
static int launch_one(const Geo2 &g, const GradWPlan &p, const CUtensorMap &mapV, const CUtensorMap &mapR,
                      const CUtensorMap &mapH, float *partials, cudaStream_t st) {
    auto kern = gradw_tma_kernel<AXC, DROP, CB, BYB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    kern<<<(unsigned)p.grid, p.threads, p.smem, st>>>(g, p, mapV, mapR, mapH, partials);
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}

template <>
int gradw_launch_axc<TNMF_AXC>(const Geo2 &g, const GradWPlan &p, const CUtensorMap &mapV, const CUtensorMap &mapR,
                               const CUtensorMap &mapH, float *partials, cudaStream_t st) {
#define TNMF_GRADW_CASE(cb, byb)                                                                      \
    if (p.CB == cb && p.BYB == byb)                                                                   \
        return p.ch.drop ? launch_one<TNMF_AXC, 1, cb, byb>(g, p, mapV, mapR, mapH, partials, st)     \
                         : launch_one<TNMF_AXC, 0, cb, byb>(g, p, mapV, mapR, mapH, partials, st);
    TNMF_GRADW_CASE(1, 1)
    TNMF_GRADW_CASE(1, 2)
    TNMF_GRADW_CASE(1, 3)
    TNMF_GRADW_CASE(2, 1)
    TNMF_GRADW_CASE(3, 1)
#if TNMF_AXC <= 12
    TNMF_GRADW_CASE(2, 2)
#endif
#undef TNMF_GRADW_CASE
    return TNMF_EUNSUPPORTED;
}

}  // namespace tma
}  // namespace tnmf
