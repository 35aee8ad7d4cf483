// Persistent TMA kernel: reconstruction  R[n,c,y,x] = sum_m sum_{ay,ax} W[m,c,ay,ax] * Hext[n,m,y+offy-ay,x+offx-ax]
// (tnmf/backends/_Backend.py:120-122, NumPy.py:122-132), optionally fused with the energy reduction
// 0.5*sum (V-R)^2 (tnmf/backends/_Backend.py:127-130).
//
// Work unit = (sample, output tile, block of CB channels); ring stage = one atom m of the unit: the H[n,m] tile with
// halo (one TMA box; activations outside [0,T) read as zero, which is the 'full' mode's padding) and the flipped
// atom slice W[m, c0..c0+CB] (one bulk copy out of the pre-arranged buffer).  A consumer thread owns RB consecutive
// rows x 8 consecutive columns x CB channels; per staged H row it loads one register window of 8+AXC activations
// and applies up to RB atom rows x CB channels x AXC taps to it.  Bound: FP32 FMA pipe.
//
// Compiled once per atom-width chunk: -DTNMF_AXC=4|8|12|16.
#include "tma_common.cuh"

#ifndef TNMF_AXC
#error "compile with -DTNMF_AXC=4|8|12|16"
#endif

namespace tnmf {
namespace tma {

template <int AXC, int DROP, int CB, int RB>
__global__ void __launch_bounds__(32 * (kConsumersMax + 1), 1)
recon_tma_kernel(const Geo2 g, const ReconPlan p, const __grid_constant__ CUtensorMap mapH, const ReconArgs a) {
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
    // floats of one (atom, channel block) slice: plain taps, or (RB even) pairs over two atom rows - see below
    const int taps_slice = RB > 1 ? 2 * (g.AY + 1) * p.ch.AXP * CB : g.AY * p.ch.AXP * CB;
    const unsigned stage_bytes = (unsigned)((p.pitch * p.HR + taps_slice) * sizeof(float));
    Ring ring;

    if (warp == p.consumers) {
        // ---------------- producer ----------------
        if (lane == 0) {
            prefetch_map(&mapH);
            for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
                const int n = (int)(u / units_per_sample);
                int r = (int)(u - (long long)n * units_per_sample);
                const int cb = r % p.nblk; r /= p.nblk;
                const int tx_i = r % p.tiles_x, ty_i = r / p.tiles_x;
                const int gx0 = tx_i * p.tile_x + g.offx - (g.AX - 1), gy0 = ty_i * p.tile_y + g.offy - (g.AY - 1);
                for (int m = 0; m < g.M; ++m) {
                    mbar_wait_relaxed(&empty_bar[ring.stage], ring.phase ^ 1u);
                    float *sh = smem + (size_t)ring.stage * p.stage_floats;
                    float *sw = sh + p.plane_floats;
                    mbar_arrive_expect_tx(&full_bar[ring.stage], stage_bytes);
                    tma_load_4d(sh, &mapH, &full_bar[ring.stage], gx0, gy0, m, n);
                    bulk_load(sw, a.Wt + (size_t)(m * p.nblk + cb) * taps_slice, (unsigned)(taps_slice * sizeof(float)),
                              &full_bar[ring.stage]);
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
    const int ry0 = (wy * kLY + ly) * RB, rx0 = (wx * kLX + lx) * kCols;
    double e_local = 0.0;

    for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int n = (int)(u / units_per_sample);
        int r = (int)(u - (long long)n * units_per_sample);
        const int cb = r % p.nblk; r /= p.nblk;
        const int tx_i = r % p.tiles_x, ty_i = r / p.tiles_x;
        const int x0 = tx_i * p.tile_x, y0 = ty_i * p.tile_y, c0 = cb * CB;
        const bool warp_active = (y0 + wy * kLY * RB < g.DY) && (x0 + wx * kLX * kCols < g.DX);

        // RB even: packed FP32 (FFMA2, fma.rn.f32x2, sm_100).  The two halves of a 64-bit register pair are the same
        // column of two vertically adjacent output rows (2p, 2p+1); a staged H row hrr reaches them through the atom
        // rows t = hrr - 2p and t - 1, whose taps prepare_taps_recon stores side by side (zeros past either end), so a
        // 16-byte load yields two ready-made tap pairs and the window value is a scalar-broadcast operand (R.F32).
        // One issue slot carries two FMAs and no pack / unpack instruction is executed; the price is one padded atom
        // row in AY + 1.
        constexpr int RP = RB > 1 ? RB / 2 : 1;
        pk2 acc2[RP][CB][kCols];
        float acc[RB == 1 ? 1 : 1][RB == 1 ? CB : 1][RB == 1 ? kCols : 1];
        if constexpr (RB > 1) {
#pragma unroll
            for (int pr = 0; pr < RP; ++pr)
#pragma unroll
                for (int c = 0; c < CB; ++c)
#pragma unroll
                    for (int j = 0; j < kCols; ++j) acc2[pr][c][j] = 0ull;
        } else {
#pragma unroll
            for (int c = 0; c < CB; ++c)
#pragma unroll
                for (int j = 0; j < kCols; ++j) acc[0][c][j] = 0.f;
        }

        for (int m = 0; m < g.M; ++m) {
            mbar_wait(&full_bar[ring.stage], ring.phase);
            if (warp_active) {
                const float *trow = smem + (size_t)ring.stage * p.stage_floats + ry0 * p.pitch + rx0;
                const float4 *wf = reinterpret_cast<const float4 *>(smem + (size_t)ring.stage * p.stage_floats +
                                                                    p.plane_floats);
                const int rows = g.AY + RB - 1;
                for (int hrr = 0; hrr < rows; ++hrr) {
                    for (int k = 0; k < NK; ++k) {
                        float win[kCols + AXC];
#pragma unroll
                        for (int q = 0; q < (kCols + AXC) / 4; ++q) {
                            const float4 v = lds128(trow + k * AXC + 4 * q);
                            win[4 * q] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
                        }
                        if constexpr (RB > 1) {
#pragma unroll
                            for (int pr = 0; pr < RP; ++pr) {
                                const int t = hrr - 2 * pr;                  // pair table row: (atom row t, atom row t - 1)
                                if (t < 0 || t > g.AY) continue;             // warp-uniform
                                const float4 *wq = wf + (t * NK + k) * QC * CB * 2;
#pragma unroll
                                for (int q = 0; q < QC; ++q) {
#pragma unroll
                                    for (int c = 0; c < CB; ++c) {
                                        const float4 wa = wq[(q * CB + c) * 2], wb = wq[(q * CB + c) * 2 + 1];
                                        const pk2 t0 = pk2_make(wa.x, wa.y), t1 = pk2_make(wa.z, wa.w);
                                        const pk2 t2 = pk2_make(wb.x, wb.y), t3 = pk2_make(wb.z, wb.w);
#pragma unroll
                                        for (int j = 0; j < kCols; ++j) {
                                            pk2 sacc = acc2[pr][c][j];
                                            sacc = pk2_fma(t0, win[4 * q + j], sacc);
                                            sacc = pk2_fma(t1, win[4 * q + 1 + j], sacc);
                                            sacc = pk2_fma(t2, win[4 * q + 2 + j], sacc);
                                            if (!(DROP && q == QC - 1)) sacc = pk2_fma(t3, win[4 * q + 3 + j], sacc);
                                            acc2[pr][c][j] = sacc;
                                        }
                                    }
                                }
                            }
                        } else {
                            const int by = hrr;
                            const float4 *wq = wf + (by * NK + k) * QC * CB;
#pragma unroll
                            for (int q = 0; q < QC; ++q) {
#pragma unroll
                                for (int c = 0; c < CB; ++c) {
                                    const float4 w = wq[q * CB + c];
#pragma unroll
                                    for (int j = 0; j < kCols; ++j) {
                                        float sacc = acc[0][c][j];
                                        sacc = fmaf(w.x, win[4 * q + j], sacc);
                                        sacc = fmaf(w.y, win[4 * q + 1 + j], sacc);
                                        sacc = fmaf(w.z, win[4 * q + 2 + j], sacc);
                                        if (!(DROP && q == QC - 1)) sacc = fmaf(w.w, win[4 * q + 3 + j], sacc);
                                        acc[0][c][j] = sacc;
                                    }
                                }
                            }
                        }
                    }
                    trow += p.pitch;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[ring.stage]);
            ring.advance(n_stages);
        }
        // value of output row rr, channel c, column j
        auto out_val = [&](int rr, int c, int j) -> float {
            if constexpr (RB > 1) return (rr & 1) ? pk2_hi(acc2[rr >> 1][c][j]) : pk2_lo(acc2[rr >> 1][c][j]);
            else return acc[0][c][j];
        };

        // ---- epilogue of the unit: store R and / or accumulate the energy ----
        const int x = x0 + rx0;
        if (!warp_active || x >= g.DX) continue;
        const bool vec = (g.DX & 3) == 0 && x + kCols <= g.DX;
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
            const int y = y0 + ry0 + rr;
            if (y >= g.DY) continue;
#pragma unroll
            for (int c = 0; c < CB; ++c) {
                if (c0 + c >= g.C) continue;
                const long long base = (((long long)n * g.C + c0 + c) * g.DY + y) * g.DX + x;
                if (vec) {
                    if (a.R) {
                        *reinterpret_cast<float4 *>(a.R + base) =
                            make_float4(out_val(rr, c, 0), out_val(rr, c, 1), out_val(rr, c, 2), out_val(rr, c, 3));
                        *reinterpret_cast<float4 *>(a.R + base + 4) =
                            make_float4(out_val(rr, c, 4), out_val(rr, c, 5), out_val(rr, c, 6), out_val(rr, c, 7));
                    }
                    if (a.V) {
                        const float4 v0 = *reinterpret_cast<const float4 *>(a.V + base);
                        const float4 v1 = *reinterpret_cast<const float4 *>(a.V + base + 4);
                        const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                        for (int j = 0; j < kCols; ++j) {
                            const double d = (double)vv[j] - (double)out_val(rr, c, j);
                            e_local += d * d;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < kCols; ++j) {
                        if (x + j < g.DX) {
                            if (a.R) a.R[base + j] = out_val(rr, c, j);
                            if (a.V) {
                                const double d = (double)a.V[base + j] - (double)out_val(rr, c, j);
                                e_local += d * d;
                            }
                        }
                    }
                }
            }
        }
    }
    if (a.epart) {                                             // one partial per consumer warp
        for (int o = 16; o > 0; o >>= 1) e_local += __shfl_xor_sync(0xffffffffu, e_local, o);
        if (lane == 0) a.epart[(long long)blockIdx.x * p.consumers + warp] = e_local;
    }
}

template <int AXC, int DROP, int CB, int RB>
static int launch_one(const Geo2 &g, const ReconPlan &p, const CUtensorMap &mapH, const ReconArgs &a,
                      cudaStream_t st) {
    auto kern = recon_tma_kernel<AXC, DROP, CB, RB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    kern<<<(unsigned)p.grid, p.threads, p.smem, st>>>(g, p, mapH, a);
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}

template <>
int recon_launch_axc<TNMF_AXC>(const Geo2 &g, const ReconPlan &p, const CUtensorMap &mapH, const ReconArgs &a,
                               cudaStream_t st) {
#define TNMF_RECON_CASE(cb, rb)                                                              \
    if (p.CB == cb && p.RB == rb)                                                            \
        return p.ch.drop ? launch_one<TNMF_AXC, 1, cb, rb>(g, p, mapH, a, st)                \
                         : launch_one<TNMF_AXC, 0, cb, rb>(g, p, mapH, a, st);
    TNMF_RECON_CASE(1, 1)
    TNMF_RECON_CASE(1, 2)
    TNMF_RECON_CASE(1, 4)
    TNMF_RECON_CASE(2, 1)
    TNMF_RECON_CASE(2, 2)
    TNMF_RECON_CASE(3, 1)
    TNMF_RECON_CASE(3, 2)
    TNMF_RECON_CASE(4, 1)
    TNMF_RECON_CASE(4, 2)
#undef TNMF_RECON_CASE
    return TNMF_EUNSUPPORTED;
}

}  // namespace tma
}  // namespace tnmf
