// Register-tiled W gradient: split-K reduction over samples and positions.
//
//   neg[m,c,ay,ax] = sum_n sum_{y,x} Hext[n,m,y+offy-ay,x+offx-ax] * V[n,c,y,x]          (tnmf/backends/NumPy.py:77-85)
//   pos[m,c,ay,ax] = the same with R                                                     (tnmf/backends/NumPy.py:80,87-90)
//
// The accumulators are indexed by the tap, not by the position: a warp owns BYB "tap units" (one atom row `by` x one
// chunk of AXC atom columns), a lane owns strips of 8 consecutive sample columns.  Per strip the lane loads the 8
// values of V and of R for CB channels and, per tap unit, a register window of 8+AXC activations, then performs
// 2*CB*AXC*8 FFMAs entirely from registers into 2*CB*AXC*BYB accumulators that stay live across all samples the CTA
// visits.  The K dimension (samples x rows x columns) is cut into work items (sample, row chunk, column chunk) that
// stream through shared memory with a two-stage cp.async pipeline.  The flattened (group, item) space, group =
// (atom, channel block, unit group), is dealt to the CTAs in equal contiguous ranges; at every group boundary a CTA
// reduces its accumulators over the lanes with warp shuffles and writes one partial slice.  The partial slices are
// summed in a fixed order in double by finish_gradient_w (deterministic: no atomics).  Bound: FP32 FMA pipe.
//
// Compiled once per atom-width chunk: -DTNMF_AXC=4|8|12|16.
#include "tiled_common.cuh"

#ifndef TNMF_AXC
#error "compile with -DTNMF_AXC=4|8|12|16"
#endif

namespace tnmf {
namespace tiled {

template <int AXC, int DROP, int CB, int BYB>
__global__ void __launch_bounds__(384, 1)
gradw_kernel(const Geo2 g, const GradWPlan p, const float *__restrict__ V, const float *__restrict__ R,
             const float *__restrict__ H, float *__restrict__ partials) {
    extern __shared__ __align__(128) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    const int NK = p.ch.NK, AXP = p.ch.AXP;
    const long long total = (long long)p.groups * p.items;
    const long long pos0 = (long long)blockIdx.x * p.chunk;
    const long long pos1 = pos0 + p.chunk < total ? pos0 + p.chunk : total;
    if (pos0 >= pos1) return;

    const int lx = lane % p.LX, ly = lane / p.LX;
    const int g0y = g.offy - (g.AY - 1), g0x = g.offx - (g.AX - 1);
    const long long dvol = (long long)g.DY * g.DX;
    const long long count = (long long)g.M * g.C * g.AY * g.AX;
    const bool vec_x = (g.DX & 3) == 0;

    float acc[BYB][2][CB][AXC];
#pragma unroll
    for (int i = 0; i < BYB; ++i)
#pragma unroll
        for (int X = 0; X < 2; ++X)
#pragma unroll
            for (int c = 0; c < CB; ++c)
#pragma unroll
                for (int bx = 0; bx < AXC; ++bx) acc[i][X][c][bx] = 0.f;

    // by range covered by unit group ug (all warps of the CTA)
    auto by_lo_of = [&](int ug) { return (ug * p.warps * BYB) / NK; };
    auto by_hi_of = [&](int ug) {
        int last = (ug + 1) * p.warps * BYB - 1;
        if (last > p.units - 1) last = p.units - 1;
        return last / NK;
    };

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

    auto issue = [&](int stage, const Cursor &cur) {
        const int ug = cur.ug, cb = cur.cb, m = cur.m, xc = cur.xc, yc = cur.yc, n = cur.n;
        float *tx = smem + stage * p.stage_floats;
        float *th = tx + p.x_floats;
        const int y_base = yc * p.RY, x_base = xc * p.XC;
        // V and R strips: plane row (X*CB + c)*RY + r, columns [x_base, x_base + XC), zero beyond the sample
        const int prow_n = 2 * CB * p.RY;
        for (int pr = warp; pr < prow_n; pr += n_warps) {
            const int r = pr % p.RY;
            const int xc_i = pr / p.RY;
            const int c = xc_i % CB, X = xc_i / CB;
            const int y = y_base + r;
            const bool row_ok = y < g.DY && (cb * CB + c) < g.C;
            const float *src = (X ? R : V) + ((long long)n * g.C + (row_ok ? cb * CB + c : 0)) * dvol +
                               (long long)(row_ok ? y : 0) * g.DX;
            float *dst = tx + pr * p.pitch_x;
            const int rbits = swz_row(pr);
            if (vec_x) {
                for (int u = lane; u < (p.XC >> 2); u += 32) {
                    const int x = x_base + 4 * u;
                    const bool ok = row_ok && x < g.DX;
                    cp_async16(dst + swz(4 * u, rbits), src + (ok ? x : 0), ok);
                }
            } else {
                for (int cidx = lane; cidx < p.XC; cidx += 32) {
                    const int x = x_base + cidx;
                    const bool ok = row_ok && x < g.DX;
                    cp_async4(dst + swz(cidx, rbits), src + (ok ? x : 0), ok);
                }
            }
        }
        const int by_lo = by_lo_of(ug);
        const int hrows = p.RY + by_hi_of(ug) - by_lo;
        stage_plane(th, p.pitch_h, H + n * g.hsn + m * g.hsm, g.TY, g.TX, g.hsy, y_base + g0y + by_lo, x_base + g0x, hrows,
                    p.XC + AXP - 1, g.wrap, warp, n_warps, lane);
        cp_async_commit();
    };

    auto flush = [&](int grp) {
        // sum over the lanes, then every lane stores the values whose index is congruent to it
        const int ug = grp % p.ugroups;
        const int cb = (grp / p.ugroups) % p.ncb;
        const int m = grp / (p.ugroups * p.ncb);
        const long long first_cta = ((long long)grp * p.items) / p.chunk;
        float *slot = partials + ((long long)blockIdx.x - first_cta) * 2 * count;
        const int ubase = (ug * p.warps + warp) * BYB;
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

    Cursor cur = decode(pos0), ahead = cur;
    issue(0, ahead);
    int cur_grp = cur.grp;
    for (long long pos = pos0; pos < pos1; ++pos) {
        const int stage = (int)((pos - pos0) & 1);
        if (pos + 1 < pos1) {
            advance(ahead);
            issue(stage ^ 1, ahead);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (pos != pos0) advance(cur);
        const int grp = cur.grp;
        if (grp != cur_grp) {
            flush(cur_grp);
            cur_grp = grp;
        }
        const int ug = cur.ug, xc = cur.xc, yc = cur.yc;
        const int by_lo = by_lo_of(ug);
        const int ubase = (ug * p.warps + warp) * BYB;
        int by_i[BYB], k_i[BYB];
        bool valid[BYB];
#pragma unroll
        for (int i = 0; i < BYB; ++i) {
            const int u = ubase + i;
            valid[i] = u < p.units;
            by_i[i] = u / NK - by_lo;
            k_i[i] = (u % NK) * AXC;
        }
        if (valid[0]) {
            const float *tx = smem + stage * p.stage_floats;
            const float *th = tx + p.x_floats;
            const int rows_here = min(p.RY, g.DY - yc * p.RY);
            const int strips_here = (min(p.XC, g.DX - xc * p.XC) + kCols - 1) / kCols;
            for (int r = ly; r < rows_here; r += p.LY) {
                for (int s = lx; s < strips_here; s += p.LX) {
                    if constexpr (BYB == 1) {
                        // one tap unit: keep the activation window, stream V then R through the registers
                        float win[kCols + AXC];
                        {
                            const int hrow = r + by_i[0];
                            const int rbits = swz_row(hrow);
                            const float *trow = th + hrow * p.pitch_h;
#pragma unroll
                            for (int q = 0; q < (kCols + AXC) / 4; ++q) {
                                const float4 v = lds128(trow + swz(kCols * s + k_i[0] + 4 * q, rbits));
                                win[4 * q] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
                            }
                        }
#pragma unroll
                        for (int X = 0; X < 2; ++X) {
                            float xv[CB][kCols];
#pragma unroll
                            for (int c = 0; c < CB; ++c) {
                                const int prow = (X * CB + c) * p.RY + r;
                                const int rbits = swz_row(prow);
                                const float *xrow = tx + prow * p.pitch_x;
                                const float4 a = lds128(xrow + swz(kCols * s, rbits));
                                const float4 b4 = lds128(xrow + swz(kCols * s + 4, rbits));
                                xv[c][0] = a.x; xv[c][1] = a.y; xv[c][2] = a.z; xv[c][3] = a.w;
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
                                const int prow = (X * CB + c) * p.RY + r;
                                const int rbits = swz_row(prow);
                                const float *xrow = tx + prow * p.pitch_x;
                                const float4 a = lds128(xrow + swz(kCols * s, rbits));
                                const float4 b4 = lds128(xrow + swz(kCols * s + 4, rbits));
                                xv[X][c][0] = a.x; xv[X][c][1] = a.y; xv[X][c][2] = a.z; xv[X][c][3] = a.w;
                                xv[X][c][4] = b4.x; xv[X][c][5] = b4.y; xv[X][c][6] = b4.z; xv[X][c][7] = b4.w;
                            }
#pragma unroll
                        for (int i = 0; i < BYB; ++i) {
                            if (!valid[i]) continue;                    // warp-uniform
                            float win[kCols + AXC];
                            const int hrow = r + by_i[i];
                            const int rbits = swz_row(hrow);
                            const float *trow = th + hrow * p.pitch_h;
#pragma unroll
                            for (int q = 0; q < (kCols + AXC) / 4; ++q) {
                                const float4 v = lds128(trow + swz(kCols * s + k_i[i] + 4 * q, rbits));
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
        __syncthreads();
    }
    flush(cur_grp);
}

template <int AXC, int DROP, int CB, int BYB>
static int launch_one(const Geo2 &g, const GradWPlan &p, const float *V, const float *R, const float *H,
                      float *partials, cudaStream_t st) {
    auto kern = gradw_kernel<AXC, DROP, CB, BYB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    kern<<<(unsigned)p.grid, p.threads, p.smem, st>>>(g, p, V, R, H, partials);
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}

template <>
int gradw_launch_axc<TNMF_AXC>(const Geo2 &g, const GradWPlan &p, const float *V, const float *R, const float *H,
                               float *partials, cudaStream_t st) {
#define TNMF_GRADW_CASE(cb, byb)                                                                   \
    if (p.CB == cb && p.BYB == byb)                                                                \
        return p.ch.drop ? launch_one<TNMF_AXC, 1, cb, byb>(g, p, V, R, H, partials, st)           \
                         : launch_one<TNMF_AXC, 0, cb, byb>(g, p, V, R, H, partials, st);
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

}  // namespace tiled
}  // namespace tnmf
