// Persistent TMA kernel: H gradient with the fused multiplicative update.
//
//   neg[n,m,t] = sum_c sum_{ay,ax} W[m,c,ay,ax] * Vext[n,c,ty-offy+ay,tx-offx+ax]      (tnmf/backends/NumPy.py:101-109)
//   pos[n,m,t] = the same with R                                                       (tnmf/backends/NumPy.py:111-119)
//   epilogue (H given):  pos += lambda*(G-H); pos += lambda_c*(Gsum-G); pos += reg; H = (H*neg)/pos
//                                                           (tnmf/TransformInvariantNMF.py:217-235,246-271)
//
// Work unit = (sample, tile of activation positions, block of MB atoms); ring stage = one channel of the unit: the V
// and the R tile with halo (two TMA boxes, zero-filled outside the sample) and the atom slices W[m0..m0+MB, c]
// (one bulk copy out of the pre-arranged buffer).  A consumer thread owns 8 consecutive positions x MB atoms x
// {neg, pos}; per atom row it loads the V and the R register windows (8+AXC values each, LDS.128 at immediate
// offsets) and the taps (warp-uniform LDS.128) and issues 2*MB*AXC*8 FFMAs.  neg and pos never touch HBM in the
// fused form.  Bound: FP32 FMA pipe.
//
// Compiled once per atom-width chunk: -DTNMF_AXC=4|8|12|16.
#include "tma_common.cuh"

#ifndef TNMF_AXC
#error "compile with -DTNMF_AXC=4|8|12|16"
#endif

namespace tnmf {
namespace tma {

// WIDE = 1: at most 8 consumer warps + the producer (9 warps, up to 224 registers) for the register-hungry
// combinations; WIDE = 0: up to 12 consumer warps + the producer (13 warps: 128 registers, the register file of an
// SM sub-partition divided by the 4 warps one of the schedulers then hosts).
template <int AXC, int DROP, int MB, int WIDE>
__global__ void __launch_bounds__(WIDE ? 32 * 9 : 32 * (kConsumersMax + 1), 1)
hupd_tma_kernel(const Geo2 g, const HupdPlan p, const __grid_constant__ CUtensorMap mapV,
                const __grid_constant__ CUtensorMap mapR, const HupdArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long full_bar[8], empty_bar[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_stages = p.n_stages;
    constexpr int QC = AXC / 4;
    const int NK = p.ch.NK;

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], (unsigned)p.consumers);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int units_per_sample = p.tiles_y * p.tiles_x * p.nblk;
    const unsigned stage_bytes = (unsigned)((2 * p.pitch * p.HR + g.AY * p.ch.AXP * MB) * sizeof(float));
    Ring ring;

    if (warp == p.consumers) {
        // ---------------- producer ----------------
        if (lane == 0) {
            prefetch_map(&mapV);
            prefetch_map(&mapR);
            for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
                const int n = (int)(u / units_per_sample);
                int r = (int)(u - (long long)n * units_per_sample);
                const int mb = r % p.nblk; r /= p.nblk;
                const int tx_i = r % p.tiles_x, ty_i = r / p.tiles_x;
                // x0 = tx_i*tile_x + x_shift is congruent to offx mod 4, so the box starts on a multiple of 4
                const int gx0 = tx_i * p.tile_x + p.x_shift - g.offx, gy0 = ty_i * p.tile_y - g.offy;
                for (int c = 0; c < g.C; ++c) {
                    mbar_wait_relaxed(&empty_bar[ring.stage], ring.phase ^ 1u);
                    float *sv = smem + (size_t)ring.stage * p.stage_floats;
                    float *sr = sv + p.plane_floats;
                    float *sw = sr + p.plane_floats;
                    mbar_arrive_expect_tx(&full_bar[ring.stage], stage_bytes);
                    tma_load_4d(sv, &mapV, &full_bar[ring.stage], gx0, gy0, c, n);
                    tma_load_4d(sr, &mapR, &full_bar[ring.stage], gx0, gy0, c, n);
                    bulk_load(sw, a.Wt + (size_t)(c * p.nblk + mb) * (g.AY * p.ch.AXP * MB),
                              (unsigned)(g.AY * p.ch.AXP * MB * sizeof(float)), &full_bar[ring.stage]);
                    ring.advance(n_stages);
                }
            }
        }
        return;
    }
    if (warp > p.consumers) return;

    // ---------------- consumers ----------------
    const int wy = warp / p.WX, wx = warp - wy * p.WX;
    const int ly = lane / kLX, lx = lane - ly * kLX;
    const int ry0 = wy * kLY + ly, rx0 = (wx * kLX + lx) * kCols;
    const long long tvol = (long long)g.TY * g.TX;

    for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int n = (int)(u / units_per_sample);
        int r = (int)(u - (long long)n * units_per_sample);
        const int mb = r % p.nblk; r /= p.nblk;
        const int tx_i = r % p.tiles_x, ty_i = r / p.tiles_x;
        const int x0 = tx_i * p.tile_x + p.x_shift, y0 = ty_i * p.tile_y, m0 = mb * MB;
        const bool warp_active = (y0 + wy * kLY < g.TY) && (x0 + wx * kLX * kCols < g.TX);

        float neg[MB][kCols], pos[MB][kCols];
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
            for (int j = 0; j < kCols; ++j) { neg[i][j] = 0.f; pos[i][j] = 0.f; }

        for (int c = 0; c < g.C; ++c) {
            mbar_wait(&full_bar[ring.stage], ring.phase);
            if (warp_active) {
                const float *tv = smem + (size_t)ring.stage * p.stage_floats + ry0 * p.pitch + rx0;
                const float *tr = tv + p.plane_floats;
                const float4 *wq = reinterpret_cast<const float4 *>(smem + (size_t)ring.stage * p.stage_floats +
                                                                    2 * p.plane_floats);
                for (int ay = 0; ay < g.AY; ++ay) {
                    for (int k = 0; k < NK; ++k) {
                        // both windows live at once so that every tap quad is loaded once and feeds neg and pos
                        // (measured on cfg2: 7% faster than one window at a time, even where it spills a few words)
                        float wv[kCols + AXC], wr[kCols + AXC];
#pragma unroll
                        for (int q = 0; q < (kCols + AXC) / 4; ++q) {
                            const float4 v = lds128(tv + k * AXC + 4 * q);
                            wv[4 * q] = v.x; wv[4 * q + 1] = v.y; wv[4 * q + 2] = v.z; wv[4 * q + 3] = v.w;
                            const float4 rr = lds128(tr + k * AXC + 4 * q);
                            wr[4 * q] = rr.x; wr[4 * q + 1] = rr.y; wr[4 * q + 2] = rr.z; wr[4 * q + 3] = rr.w;
                        }
#pragma unroll
                        for (int q = 0; q < QC; ++q) {
#pragma unroll
                            for (int i = 0; i < MB; ++i) {
                                const float4 w = wq[(k * QC + q) * MB + i];
#pragma unroll
                                for (int j = 0; j < kCols; ++j) {
                                    float s = neg[i][j], t = pos[i][j];
                                    s = fmaf(w.x, wv[4 * q + j], s);
                                    t = fmaf(w.x, wr[4 * q + j], t);
                                    s = fmaf(w.y, wv[4 * q + 1 + j], s);
                                    t = fmaf(w.y, wr[4 * q + 1 + j], t);
                                    s = fmaf(w.z, wv[4 * q + 2 + j], s);
                                    t = fmaf(w.z, wr[4 * q + 2 + j], t);
                                    if (!(DROP && q == QC - 1)) {            // dead last tap of an odd atom width
                                        s = fmaf(w.w, wv[4 * q + 3 + j], s);
                                        t = fmaf(w.w, wr[4 * q + 3 + j], t);
                                    }
                                    neg[i][j] = s; pos[i][j] = t;
                                }
                            }
                        }
                    }
                    tv += p.pitch;
                    tr += p.pitch;
                    wq += NK * QC * MB;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[ring.stage]);
            ring.advance(n_stages);
        }

        // ---- epilogue of the unit ----
        // tx may be negative (shifted tile origin) and is congruent to offx mod 4: rows of H are read / written
        // with the widest access the address allows (16, 8 or 4 bytes)
        const int ty = y0 + ry0, tx = x0 + rx0;
        if (!warp_active || ty >= g.TY || tx >= g.TX || tx + kCols <= 0) continue;
        const long long tin_h = (long long)ty * g.hsy + tx;       // H may carry a padded row pitch
        const long long tin = (long long)ty * g.TX + tx;          // neg / pos / G / Gsum are dense
        const bool whole = tx >= 0 && tx + kCols <= g.TX;
        if (a.H) {
#pragma unroll
            for (int i = 0; i < MB; ++i) {
                const int m = m0 + i;
                if (m >= g.M) continue;
                float *hp = a.H + n * g.hsn + m * g.hsm + tin_h;
                const unsigned mis = (unsigned)(reinterpret_cast<unsigned long long>(hp) & 15ull);
                float hv[kCols];
                if (whole && mis == 0) {
                    const float4 h0 = *reinterpret_cast<const float4 *>(hp);
                    const float4 h1 = *reinterpret_cast<const float4 *>(hp + 4);
                    hv[0] = h0.x; hv[1] = h0.y; hv[2] = h0.z; hv[3] = h0.w;
                    hv[4] = h1.x; hv[5] = h1.y; hv[6] = h1.z; hv[7] = h1.w;
                } else if (whole && mis == 8) {
#pragma unroll
                    for (int j = 0; j < kCols; j += 2) {
                        const float2 h2 = *reinterpret_cast<const float2 *>(hp + j);
                        hv[j] = h2.x; hv[j + 1] = h2.y;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < kCols; ++j) hv[j] = ((unsigned)(tx + j) < (unsigned)g.TX) ? hp[j] : 0.f;
                }
                const long long cidx = ((long long)n * g.M + m) * tvol + tin;
                float out[kCols];
#pragma unroll
                for (int j = 0; j < kCols; ++j) {
                    const bool ok = (unsigned)(tx + j) < (unsigned)g.TX;
                    const float h = hv[j];
                    float ps = pos[i][j];
                    if (a.G) {
                        const float gi = ok ? a.G[cidx + j] : 0.f;
                        if (a.lambda != 0.f) { float tmp = gi - h; tmp *= a.lambda; ps += tmp; }
                        if (a.Gsum) {
                            const float gs = ok ? a.Gsum[(long long)n * tvol + tin + j] : 0.f;
                            float tmp = -gi + gs; tmp *= a.lambda_cross; ps += tmp;
                        }
                    }
                    ps += a.reg;
                    float hn = h * neg[i][j];
                    hn /= ps;
                    out[j] = hn;
                }
                if (whole && mis == 0) {
                    *reinterpret_cast<float4 *>(hp) = make_float4(out[0], out[1], out[2], out[3]);
                    *reinterpret_cast<float4 *>(hp + 4) = make_float4(out[4], out[5], out[6], out[7]);
                } else if (whole && mis == 8) {
#pragma unroll
                    for (int j = 0; j < kCols; j += 2) *reinterpret_cast<float2 *>(hp + j) = make_float2(out[j], out[j + 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < kCols; ++j)
                        if ((unsigned)(tx + j) < (unsigned)g.TX) hp[j] = out[j];
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < MB; ++i) {
                const int m = m0 + i;
                if (m >= g.M) continue;
                const long long cidx = ((long long)n * g.M + m) * tvol + tin;
#pragma unroll
                for (int j = 0; j < kCols; ++j) {
                    if ((unsigned)(tx + j) >= (unsigned)g.TX) continue;
                    a.neg[cidx + j] = neg[i][j];
                    a.pos[cidx + j] = pos[i][j];
                }
            }
        }
    }
}

template <int AXC, int DROP, int MB, int WIDE>
static int launch_one(const Geo2 &g, const HupdPlan &p, const CUtensorMap &mapV, const CUtensorMap &mapR,
                      const HupdArgs &a, cudaStream_t st) {
    auto kern = hupd_tma_kernel<AXC, DROP, MB, WIDE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    kern<<<(unsigned)p.grid, p.threads, p.smem, st>>>(g, p, mapV, mapR, a);
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}

template <>
int hupd_launch_axc<TNMF_AXC>(const Geo2 &g, const HupdPlan &p, const CUtensorMap &mapV, const CUtensorMap &mapR,
                              const HupdArgs &a, cudaStream_t st) {
#define TNMF_HUPD_CASE(mb)                                                                         \
    if (p.MB == mb && p.wide == hupd_needs_wide(TNMF_AXC, mb))                                     \
        return p.ch.drop ? launch_one<TNMF_AXC, 1, mb, hupd_needs_wide(TNMF_AXC, mb)>(g, p, mapV, mapR, a, st)  \
                         : launch_one<TNMF_AXC, 0, mb, hupd_needs_wide(TNMF_AXC, mb)>(g, p, mapV, mapR, a, st);
    TNMF_HUPD_CASE(1)
    TNMF_HUPD_CASE(2)
    TNMF_HUPD_CASE(3)
    TNMF_HUPD_CASE(4)
#undef TNMF_HUPD_CASE
    return TNMF_EUNSUPPORTED;
}

}  // namespace tma
}  // namespace tnmf
