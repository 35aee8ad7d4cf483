// Register-tiled H gradient with the fused multiplicative update.
//
//   neg[n,m,t] = sum_c sum_{ay,ax} W[m,c,ay,ax] * Vext[n,c,ty-offy+ay,tx-offx+ax]      (tnmf/backends/NumPy.py:101-109)
//   pos[n,m,t] = the same with R                                                       (tnmf/backends/NumPy.py:111-119)
//   epilogue (H given):  pos += lambda*(G-H); pos += lambda_c*(Gsum-G); pos += reg; H = (H*neg)/pos
//                                                           (tnmf/TransformInvariantNMF.py:217-235,246-271)
//
// One CTA owns a tile of activation positions of one sample for a block of MB atoms and walks the channels c with a
// two-stage cp.async pipeline: stage = V[n,c] and R[n,c] tiles with halo + the atom slices W[m0..m0+MB, c].  A thread
// owns 8 consecutive positions x MB atoms x {neg, pos}; per tile row it loads the V and the R register windows
// (8+AXC values each) once and applies AXC taps x MB atoms to both, so V and R are read from shared memory once per
// atom block and neither neg nor pos ever touches HBM in the fused form.  Bound: FP32 FMA pipe.
//
// Compiled once per atom-width chunk: -DTNMF_AXC=4|8|12|16.
#include "tiled_common.cuh"

#ifndef TNMF_AXC
#error "compile with -DTNMF_AXC=4|8|12|16"
#endif

namespace tnmf {
namespace tiled {

template <int AXC, int DROP, int MB>
__global__ void __launch_bounds__(256, 2)
hupd_kernel(const Geo2 g, const TilePlan p, const float *__restrict__ V, const float *__restrict__ R,
            const float *__restrict__ W, float *__restrict__ neg_out, float *__restrict__ pos_out,
            float *__restrict__ H, float reg, const float *__restrict__ G, float lambda,
            const float *__restrict__ Gsum, float lambda_cross) {
    extern __shared__ __align__(128) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    const int NK = p.ch.NK, AXP = p.ch.AXP;

    long long b = blockIdx.x;
    const int tx_i = (int)(b % p.tiles_x); b /= p.tiles_x;
    const int ty_i = (int)(b % p.tiles_y); b /= p.tiles_y;
    const int mb = (int)(b % p.nblk);
    const int n = (int)(b / p.nblk);
    const int x0 = tx_i * p.tile_x, y0 = ty_i * p.tile_y, m0 = mb * MB;
    const int wy = warp / p.WX, wx = warp % p.WX, ly = lane / p.LX, lx = lane % p.LX;
    const int ry0 = wy * p.LY + ly, rx0 = (wx * p.LX + lx) * kCols;
    const int gy0 = y0 - g.offy, gx0 = x0 - g.offx;
    const bool warp_active = (y0 + wy * p.LY < g.TY) && (x0 + wx * p.LX * kCols < g.TX);
    constexpr int QC = AXC / 4;
    const long long dvol = (long long)g.DY * g.DX;

    float neg[MB][kCols], pos[MB][kCols];
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < kCols; ++j) { neg[i][j] = 0.f; pos[i][j] = 0.f; }

    auto issue = [&](int stage, int c) {
        float *tv = smem + stage * p.stage_floats;
        float *tr = tv + p.plane_floats;
        float *wt = tr + p.plane_floats;
        const long long plane = ((long long)n * g.C + c) * dvol;
        stage_plane(tv, p.pitch, V + plane, g.DY, g.DX, g.DX, gy0, gx0, p.HR, p.WT, g.wrap, warp, n_warps, lane);
        stage_plane(tr, p.pitch, R + plane, g.DY, g.DX, g.DX, gy0, gx0, p.HR, p.WT, g.wrap, warp, n_warps, lane);
        // zero-padded atom slices: wt[ay][ax/4][i][ax%4] = W[m0+i][c][ay][ax]
        const int qpr = AXP >> 2;
        const int groups = g.AY * qpr * MB;
        for (int i = tid; i < groups; i += blockDim.x) {
            const int im = i % MB;
            const int t = i / MB;
            const int q = t % qpr, ay = t / qpr;
            const bool m_ok = (m0 + im) < g.M;
            const float *wrow = W + (((long long)(m_ok ? m0 + im : 0) * g.C + c) * g.AY + ay) * g.AX;
            float *d = wt + ((ay * qpr + q) * MB + im) * 4;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ax = 4 * q + e;
                const bool ok = m_ok && ax < g.AX;
                cp_async4(d + e, wrow + (ok ? ax : 0), ok);
            }
        }
        cp_async_commit();
    };

    issue(0, 0);
    for (int c = 0; c < g.C; ++c) {
        if (c + 1 < g.C) {
            issue((c + 1) & 1, c + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (warp_active) {
            const float *tv = smem + (c & 1) * p.stage_floats;
            const float *tr = tv + p.plane_floats;
            const float4 *wt = reinterpret_cast<const float4 *>(tr + p.plane_floats);
            for (int ay = 0; ay < g.AY; ++ay) {
                const int row = ry0 + ay;
                const int rbits = swz_row(row);
                const int roff = row * p.pitch;
                for (int k = 0; k < NK; ++k) {
                    const float4 *wq = wt + (ay * NK + k) * QC * MB;
                    // the V window feeds neg, then the R window feeds pos: one window live at a time
#pragma unroll
                    for (int X = 0; X < 2; ++X) {
                        const float *src = X ? tr : tv;
                        float win[kCols + AXC];
#pragma unroll
                        for (int q = 0; q < (kCols + AXC) / 4; ++q) {
                            const float4 a = lds128(src + roff + swz(rx0 + k * AXC + 4 * q, rbits));
                            win[4 * q] = a.x; win[4 * q + 1] = a.y; win[4 * q + 2] = a.z; win[4 * q + 3] = a.w;
                        }
#pragma unroll
                        for (int q = 0; q < QC; ++q) {
#pragma unroll
                            for (int i = 0; i < MB; ++i) {
                                const float4 w = wq[q * MB + i];
#pragma unroll
                                for (int j = 0; j < kCols; ++j) {
                                    float a = X ? pos[i][j] : neg[i][j];
                                    a = fmaf(w.x, win[4 * q + j], a);
                                    a = fmaf(w.y, win[4 * q + 1 + j], a);
                                    a = fmaf(w.z, win[4 * q + 2 + j], a);
                                    if (!(DROP && q == QC - 1)) a = fmaf(w.w, win[4 * q + 3 + j], a);   // dead tap
                                    if (X) pos[i][j] = a; else neg[i][j] = a;
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    if (!warp_active) return;
    const int ty = y0 + ry0, tx = x0 + rx0;
    if (ty >= g.TY || tx >= g.TX) return;
    const long long tvol = (long long)g.TY * g.TX;
    const long long tin = (long long)ty * g.TX + tx;          // neg / pos / G / Gsum are dense
    const long long tin_h = (long long)ty * g.hsy + tx;       // H may carry a padded row pitch
    float hv[MB][kCols];
    if (H) {
#pragma unroll
        for (int i = 0; i < MB; ++i) {                               // all H loads in flight before the first use
            const float *hp = H + n * g.hsn + (m0 + i < g.M ? m0 + i : m0) * g.hsm + tin_h;
#pragma unroll
            for (int j = 0; j < kCols; ++j) hv[i][j] = (tx + j < g.TX) ? hp[j] : 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < MB; ++i) {
        const int m = m0 + i;
        if (m >= g.M) continue;
        const long long cidx = ((long long)n * g.M + m) * tvol + tin;        // contiguous [n, m, T] tensors
        if (H) {
            float *hp = H + n * g.hsn + m * g.hsm + tin_h;
            float gi[kCols], gs[kCols];
#pragma unroll
            for (int j = 0; j < kCols; ++j) {
                const bool ok = tx + j < g.TX;
                gi[j] = (ok && G) ? G[cidx + j] : 0.f;
                gs[j] = (ok && Gsum) ? Gsum[(long long)n * tvol + tin + j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < kCols; ++j) {
                if (tx + j >= g.TX) continue;
                const float h = hv[i][j];
                float ps = pos[i][j];
                if (G) {
                    if (lambda != 0.f) { float tmp = gi[j] - h; tmp *= lambda; ps += tmp; }
                    if (Gsum) { float tmp = -gi[j] + gs[j]; tmp *= lambda_cross; ps += tmp; }
                }
                ps += reg;
                float hn = h * neg[i][j];
                hn /= ps;
                hp[j] = hn;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kCols; ++j) {
                if (tx + j >= g.TX) continue;
                neg_out[cidx + j] = neg[i][j];
                pos_out[cidx + j] = pos[i][j];
            }
        }
    }
}

template <int AXC, int DROP, int MB>
static int launch_one(const Geo2 &g, const TilePlan &p, const float *V, const float *R, const float *W, float *neg,
                      float *pos, float *H, float reg, const float *G, float lambda, const float *Gsum,
                      float lambda_cross, cudaStream_t st) {
    auto kern = hupd_kernel<AXC, DROP, MB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return status_from_cuda(e);
    kern<<<(unsigned)p.grid, p.threads, p.smem, st>>>(g, p, V, R, W, neg, pos, H, reg, G, lambda, Gsum, lambda_cross);
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}

template <>
int hupd_launch_axc<TNMF_AXC>(const Geo2 &g, const TilePlan &p, const float *V, const float *R, const float *W,
                              float *neg, float *pos, float *H, float reg, const float *G, float lambda,
                              const float *Gsum, float lambda_cross, cudaStream_t st) {
#define TNMF_HUPD_CASE(mb)                                                                                       \
    if (p.NB == mb)                                                                                              \
        return p.ch.drop                                                                                         \
                   ? launch_one<TNMF_AXC, 1, mb>(g, p, V, R, W, neg, pos, H, reg, G, lambda, Gsum, lambda_cross, st) \
                   : launch_one<TNMF_AXC, 0, mb>(g, p, V, R, W, neg, pos, H, reg, G, lambda, Gsum, lambda_cross, st);
    TNMF_HUPD_CASE(1)
    TNMF_HUPD_CASE(2)
    TNMF_HUPD_CASE(3)
    TNMF_HUPD_CASE(4)
#undef TNMF_HUPD_CASE
    return TNMF_EUNSUPPORTED;
}

}  // namespace tiled
}  // namespace tnmf
