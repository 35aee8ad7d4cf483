// Register-tiled reconstruction  R[n,c,y,x] = sum_m sum_{ay,ax} W[m,c,ay,ax] * Hext[n,m,y+offy-ay,x+offx-ax]
// (tnmf/backends/_Backend.py:120-122, NumPy.py:122-132), optionally fused with the energy reduction
// 0.5*sum (V-R)^2 (tnmf/backends/_Backend.py:127-130).
//
// One CTA owns an output tile of one sample for a block of CB channels and walks the atoms m with a two-stage
// cp.async pipeline: stage = H[n,m] tile with halo + the flipped, zero-padded atom slice.  A thread owns RB
// consecutive rows x 8 consecutive columns x CB channels of accumulators; per H-tile row it loads a register
// window of 8+AXC values (LDS.128, swizzled, conflict-free) and applies AXC taps x RB rows x CB channels of FFMAs,
// with the taps arriving as warp-uniform (broadcast) LDS.128.  Bound: FP32 FMA pipe.
//
// Compiled once per atom-width chunk: -DTNMF_AXC=4|8|12|16.
#include "tiled_common.cuh"

#ifndef TNMF_AXC
#error "compile with -DTNMF_AXC=4|8|12|16"
#endif

namespace tnmf {
namespace tiled {

template <int AXC, int DROP, int CB, int RB>
__global__ void __launch_bounds__(256, 2)
recon_kernel(const Geo2 g, const TilePlan p, const float *__restrict__ W, const float *__restrict__ H,
             float *__restrict__ R, const float *__restrict__ V, double *__restrict__ epart) {
    extern __shared__ __align__(128) float smem[];
    __shared__ double red[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    const int NK = p.ch.NK, AXP = p.ch.AXP;

    long long b = blockIdx.x;
    const int tx_i = (int)(b % p.tiles_x); b /= p.tiles_x;
    const int ty_i = (int)(b % p.tiles_y); b /= p.tiles_y;
    const int cb = (int)(b % p.nblk);
    const int n = (int)(b / p.nblk);
    const int x0 = tx_i * p.tile_x, y0 = ty_i * p.tile_y, c0 = cb * CB;
    const int wy = warp / p.WX, wx = warp % p.WX, ly = lane / p.LX, lx = lane % p.LX;
    const int ry0 = (wy * p.LY + ly) * RB, rx0 = (wx * p.LX + lx) * kCols;
    const int gy0 = y0 + g.offy - (g.AY - 1), gx0 = x0 + g.offx - (g.AX - 1);   // 'valid': the tile origin itself
    const bool warp_active = (y0 + wy * p.LY * RB < g.DY) && (x0 + wx * p.LX * kCols < g.DX);
    constexpr int QC = AXC / 4;                      // tap quads per chunk

    float acc[RB][CB][kCols];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int c = 0; c < CB; ++c)
#pragma unroll
            for (int j = 0; j < kCols; ++j) acc[r][c][j] = 0.f;

    auto issue = [&](int stage, int m) {
        float *tile = smem + stage * p.stage_floats;
        float *wf = tile + p.plane_floats;
        stage_plane(tile, p.pitch, H + n * g.hsn + m * g.hsm, g.TY, g.TX, g.hsy, gy0, gx0, p.HR, p.WT, g.wrap, warp,
                    n_warps, lane);
        // flipped atom slice, zero-padded at the end: wf[by][bx/4][c][bx%4] = W[m][c0+c][AY-1-by][AX-1-bx]
        const int qpr = AXP >> 2;
        const int groups = g.AY * qpr * CB;
        for (int i = tid; i < groups; i += blockDim.x) {
            const int c = i % CB;
            const int t = i / CB;
            const int q = t % qpr, by = t / qpr;
            const bool c_ok = (c0 + c) < g.C;
            const float *wrow = W + (((long long)m * g.C + (c_ok ? c0 + c : 0)) * g.AY + (g.AY - 1 - by)) * g.AX;
            float *d = wf + ((by * qpr + q) * CB + c) * 4;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ax = g.AX - 1 - (4 * q + e);
                const bool ok = c_ok && ax >= 0;
                cp_async4(d + e, wrow + (ok ? ax : 0), ok);
            }
        }
        cp_async_commit();
    };

    issue(0, 0);
    for (int m = 0; m < g.M; ++m) {
        if (m + 1 < g.M) {
            issue((m + 1) & 1, m + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (warp_active) {
            const float *tile = smem + (m & 1) * p.stage_floats;
            const float4 *wf = reinterpret_cast<const float4 *>(tile + p.plane_floats);
            const int rows = g.AY + RB - 1;
            for (int hrr = 0; hrr < rows; ++hrr) {
                const int row = ry0 + hrr;
                const int rbits = swz_row(row);
                const float *trow = tile + row * p.pitch;
                for (int k = 0; k < NK; ++k) {
                    float win[kCols + AXC];
#pragma unroll
                    for (int q = 0; q < (kCols + AXC) / 4; ++q) {
                        const float4 v = lds128(trow + swz(rx0 + k * AXC + 4 * q, rbits));
                        win[4 * q] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
                    }
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        const int by = hrr - r;
                        if (by < 0 || by >= g.AY) continue;          // warp-uniform
                        const float4 *wq = wf + (by * NK + k) * QC * CB;
#pragma unroll
                        for (int q = 0; q < QC; ++q) {
#pragma unroll
                            for (int c = 0; c < CB; ++c) {
                                const float4 w = wq[q * CB + c];
#pragma unroll
                                for (int j = 0; j < kCols; ++j) {
                                    float a = acc[r][c][j];
                                    a = fmaf(w.x, win[4 * q + j], a);
                                    a = fmaf(w.y, win[4 * q + 1 + j], a);
                                    a = fmaf(w.z, win[4 * q + 2 + j], a);
                                    if (!(DROP && q == QC - 1)) a = fmaf(w.w, win[4 * q + 3 + j], a);   // dead last tap
                                    acc[r][c][j] = a;
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    // epilogue: store R and / or accumulate the energy
    double e_local = 0.0;
    const int x = x0 + rx0;
    const bool vec = (g.DX & 3) == 0 && x + kCols <= g.DX;
    if (warp_active) {
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const int y = y0 + ry0 + r;
            if (y >= g.DY || x >= g.DX) continue;
#pragma unroll
            for (int c = 0; c < CB; ++c) {
                if (c0 + c >= g.C) continue;
                const long long base = (((long long)n * g.C + c0 + c) * g.DY + y) * g.DX + x;
                if (vec) {
                    if (R) {
                        *reinterpret_cast<float4 *>(R + base) =
                            make_float4(acc[r][c][0], acc[r][c][1], acc[r][c][2], acc[r][c][3]);
                        *reinterpret_cast<float4 *>(R + base + 4) =
                            make_float4(acc[r][c][4], acc[r][c][5], acc[r][c][6], acc[r][c][7]);
                    }
                    if (V) {
                        const float4 v0 = *reinterpret_cast<const float4 *>(V + base);
                        const float4 v1 = *reinterpret_cast<const float4 *>(V + base + 4);
                        const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                        for (int j = 0; j < kCols; ++j) {
                            const double d = (double)vv[j] - (double)acc[r][c][j];
                            e_local += d * d;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < kCols; ++j) {
                        if (x + j < g.DX) {
                            if (R) R[base + j] = acc[r][c][j];
                            if (V) {
                                const double d = (double)V[base + j] - (double)acc[r][c][j];
                                e_local += d * d;
                            }
                        }
                    }
                }
            }
        }
    }
    if (epart) {
        const double s = block_sum(e_local, red);
        if (tid == 0) epart[blockIdx.x] = s;
    }
}

template <int AXC, int DROP, int CB, int RB>
static int launch_one(const Geo2 &g, const TilePlan &p, const float *W, const float *H, float *R, const float *V,
                      double *epart, cudaStream_t st) {
    auto kern = recon_kernel<AXC, DROP, CB, RB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    kern<<<(unsigned)p.grid, p.threads, p.smem, st>>>(g, p, W, H, R, V, epart);
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}

template <>
int recon_launch_axc<TNMF_AXC>(const Geo2 &g, const TilePlan &p, const float *W, const float *H, float *R,
                               const float *V, double *epart, cudaStream_t st) {
#define TNMF_RECON_CASE(cb, rb)                                                                         \
    if (p.NB == cb && p.RB == rb)                                                                       \
        return p.ch.drop ? launch_one<TNMF_AXC, 1, cb, rb>(g, p, W, H, R, V, epart, st)                 \
                         : launch_one<TNMF_AXC, 0, cb, rb>(g, p, W, H, R, V, epart, st);
    TNMF_RECON_CASE(1, 1)
    TNMF_RECON_CASE(1, 2)
    TNMF_RECON_CASE(1, 4)
    TNMF_RECON_CASE(2, 1)
    TNMF_RECON_CASE(2, 2)
    TNMF_RECON_CASE(3, 1)
    TNMF_RECON_CASE(4, 1)
#undef TNMF_RECON_CASE
    return TNMF_EUNSUPPORTED;
}

}  // namespace tiled
}  // namespace tnmf
