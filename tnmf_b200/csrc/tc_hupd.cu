// Tensor-core (tcgen05, 3xTF32) H gradient with the fused multiplicative update.
//
//   neg[n,m,ty,tx] = sum_c sum_{ay,ax} W[m,c,ay,ax] * Vext[n,c,ty-offy+ay,tx-offx+ax]   (tnmf/backends/NumPy.py:101-109)
//   pos[n,m,ty,tx] = the same with R                                                    (tnmf/backends/NumPy.py:111-119)
//   epilogue (H given):  pos += lambda*(G-H); pos += lambda_c*(Gsum-G); pos += reg; H = (H*neg)/pos
//                                                           (tnmf/TransformInvariantNMF.py:217-235,246-271)
//
// Formulation.  A CTA owns a tile of 128 activation COLUMNS - column q = (sample n, position tx), the flattened
// [N x TX] space cut into runs of 128, so partial rows never waste MMA lanes - and walks down the rows.  For a source
// row r of X (V or R) the "expander" warps build, once,
//       A_r[column i, k = (c, ax)] = Xext[n_i, c, r, tx_i - offx + ax]          128 x KP,  KP = roundup(C*AX, 8)
// in shared memory (canonical K-major no-swizzle layout, hi/lo TF32 split).  Source row r contributes to the output
// rows ty = r + offy - ay, ay = 0..AY-1, through
//       D_ty[column, m] += A_r[column, :] . Wt_ay[m, :]
// Output rows live in a ring of 16 TMEM slots (16 atoms x {neg, pos} = 32 columns each, 512 columns in all) and
// consecutive output rows occupy consecutive slots, so ONE tcgen05.mma with N = 16 * (number of live rows) serves all
// atom rows of the source row (the atom-row blocks of Wt are stored in that order).  3xTF32: hi*hi + lo*hi + hi*lo,
// FP32 accumulation in TMEM.  When the last source row of an output row has been issued, a tcgen05.commit hands the
// slot to the epilogue warps, which read it back (tcgen05.ld), apply the update and release the slot.
//
// Roles (288 threads): warps 0-3 expanders (thread = column), warps 4-7 epilogue (thread = column = TMEM lane),
// warp 8 = one elected thread issuing the MMAs.  mbarriers: a_full/a_empty per operand stage, row_done/slot_free per
// TMEM slot.  Atoms are processed in blocks of 16 (one launch per block).
// Bound: tensor pipe at the TF32 rate / 3.
#include "tc_common.cuh"

namespace tnmf {
namespace tc {

using tiled::ceil_div;
using tiled::Geo2;
using tiled::round_up;

constexpr int kTile = 128;          // activation columns per CTA tile = MMA M
constexpr int kNB = 16;             // atoms per launch = MMA N granule
constexpr int kSlots = 16;          // TMEM ring: 16 output rows x (16 neg + 16 pos columns)
constexpr int kMaxStages = 6;
constexpr int kThreads = 32 * 9;
constexpr int kMaxSmem = 226 * 1024;

struct TcHupdPlan {
    int KP, ksteps;                 // padded contraction length C*AX -> multiple of 8
    int tiles, rblocks, rows_per_block;
    long long units;
    int n_stages, stage_floats, w_floats;
    int grid;
    size_t smem;
};

struct TcHupdArgs {
    const float *V, *R, *W;
    float *neg, *pos, *H;
    float reg, lambda, lambda_cross;
    const float *G, *Gsum;
    int m0;                         // first atom of this launch
};

bool make_tc_hupd_plan(const Geo2 &g, TcHupdPlan &p) {
    p = TcHupdPlan();
    if (g.AY > kSlots - 1 || g.AY < 1) return false;
    p.KP = round_up(g.C * g.AX, 8);
    p.ksteps = p.KP / 8;
    p.stage_floats = 2 * kTile * p.KP;                     // hi + lo
    p.w_floats = 2 * g.AY * kNB * p.KP;                    // hi + lo, atom-row blocks in ring order
    const size_t fixed = (size_t)p.w_floats * 4 + 1024;
    if (fixed + 2 * (size_t)p.stage_floats * 4 > (size_t)kMaxSmem) return false;
    p.n_stages = (int)(((size_t)kMaxSmem - fixed) / ((size_t)p.stage_floats * 4));
    if (p.n_stages > kMaxStages) p.n_stages = kMaxStages;
    p.smem = fixed + (size_t)p.n_stages * p.stage_floats * 4;
    const long long cols = (long long)g.N * g.TX;
    if (cols <= 0 || cols >= (1ll << 31) - kTile) return false;
    p.tiles = (int)((cols + kTile - 1) / kTile);
    // row blocks: minimise waves * (rows + per-unit overhead)
    const int sms = tma::sm_count();
    double best = -1;
    for (int rb = 1; rb <= g.TY && rb <= 64; ++rb) {
        const int rows = ceil_div(g.TY, rb);
        if (ceil_div(g.TY, rows) != rb) continue;
        const long long units = (long long)p.tiles * rb;
        const double waves = (double)((units + sms - 1) / sms);
        const double cost = waves * (rows + 0.3 * (g.AY - 1) + 1.0);
        if (best < 0 || cost < best * 0.999) { best = cost; p.rblocks = rb; p.rows_per_block = rows; }
    }
    p.units = (long long)p.tiles * p.rblocks;
    p.grid = (int)(p.units < sms ? p.units : sms);
    return true;
}

struct Unit {
    int tile, ty0, ty1, r_lo, r_hi;
};
__device__ __forceinline__ Unit make_unit(long long u, const Geo2 &g, const TcHupdPlan &p) {
    Unit w;
    const int rb = (int)(u / p.tiles);
    w.tile = (int)(u - (long long)rb * p.tiles);
    w.ty0 = rb * p.rows_per_block;
    w.ty1 = min(g.TY, w.ty0 + p.rows_per_block);
    w.r_lo = max(0, w.ty0 - g.offy);
    w.r_hi = min(g.DY - 1, w.ty1 - 1 - g.offy + g.AY - 1);
    return w;
}

__global__ void __launch_bounds__(kThreads, 1)
hupd_tc_kernel(const Geo2 g, const TcHupdPlan p, const TcHupdArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long a_full[kMaxStages], a_empty[kMaxStages], row_done[kSlots],
        slot_free[kSlots];
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KP = p.KP, AY = g.AY, AX = g.AX, C = g.C;
    const int NR = AY * kNB;                               // rows of the atom operand
    float *w_hi = smem, *w_lo = smem + NR * KP;
    float *stages = smem + p.w_floats;

    if (tid == 0) {
        for (int s = 0; s < p.n_stages; ++s) { mbar_init(&a_full[s], kTile); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kSlots; ++s) { mbar_init(&row_done[s], 1); mbar_init(&slot_free[s], kTile); }
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(&tmem_base_s, 512);
    // atom operand: row n = j*16 + ml  <->  atom m0+ml, atom row ay = AY-1-j;  k = c*AX + ax
    for (int idx = tid; idx < NR * KP; idx += kThreads) {
        const int n = idx / KP, k = idx - n * KP;
        const int j = n / kNB, ml = n - j * kNB;
        const int m = a.m0 + ml, ay = AY - 1 - j;
        float v = 0.f;
        if (m < g.M && k < C * AX) {
            const int c = k / AX, ax = k - c * AX;
            v = a.W[(((long long)m * C + c) * AY + ay) * AX + ax];
        }
        float hi, lo;
        split_tf32(v, hi, lo);
        const size_t o = canon_offset_floats(n, k, NR);
        w_hi[o] = hi;
        w_lo[o] = lo;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = tmem_base_s;
    const long long total_cols = (long long)g.N * g.TX;

    if (warp < 4) {
        // ------------------------------------ expanders ------------------------------------
        const int i = tid;
        int st = 0;
        unsigned ph = 0;
        const long long plane = (long long)g.DY * g.DX;
        const int KG = KP >> 2;
        for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
            const Unit w = make_unit(u, g, p);
            const long long q = (long long)w.tile * kTile + i;
            const bool active = q < total_cols;
            const int n = active ? (int)(q / g.TX) : 0;
            const int xs = (active ? (int)(q - (long long)n * g.TX) : 0) - g.offx;
            for (int r = w.r_lo; r <= w.r_hi; ++r) {
                for (int x = 0; x < 2; ++x) {
                    const float *src = (x ? a.R : a.V) + (long long)n * C * plane + (long long)r * g.DX;
                    mbar_wait(&a_empty[st], ph ^ 1u);
                    float *d_hi = stages + (size_t)st * p.stage_floats + (size_t)(i >> 3) * 32 + (size_t)(i & 7) * 4;
                    float *d_lo = d_hi + kTile * KP;
                    int c = 0, ax = 0;
                    for (int kg0 = 0; kg0 < KG; kg0 += 4) {
                        float v[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const int xx = xs + ax;
                            const bool ok = active && c < C && (unsigned)xx < (unsigned)g.DX && (kg0 * 4 + e) < KP;
                            v[e] = ok ? __ldg(src + (long long)c * plane + xx) : 0.f;
                            if (++ax == AX) { ax = 0; ++c; }
                        }
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {
                            if (kg0 + gq < KG) {
                                float4 hi, lo;
                                split_tf32(v[4 * gq + 0], hi.x, lo.x);
                                split_tf32(v[4 * gq + 1], hi.y, lo.y);
                                split_tf32(v[4 * gq + 2], hi.z, lo.z);
                                split_tf32(v[4 * gq + 3], hi.w, lo.w);
                                *reinterpret_cast<float4 *>(d_hi + (size_t)(kg0 + gq) * (kTile * 4)) = hi;
                                *reinterpret_cast<float4 *>(d_lo + (size_t)(kg0 + gq) * (kTile * 4)) = lo;
                            }
                        }
                    }
                    fence_proxy_async();
                    mbar_arrive(&a_full[st]);
                    if (++st == p.n_stages) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp < 8) {
        // ------------------------------------ epilogue ------------------------------------
        const int i = tid - kTile;
        const unsigned lane_base = (unsigned)((warp & 3) * 32) << 16;
        const long long tvol = (long long)g.TY * g.TX;
        long long g_base = 0;
        for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
            const Unit w = make_unit(u, g, p);
            const long long q = (long long)w.tile * kTile + i;
            const bool active = q < total_cols;
            const int n = active ? (int)(q / g.TX) : 0;
            const int tx = active ? (int)(q - (long long)n * g.TX) : 0;
            for (int ty = w.ty0; ty < w.ty1; ++ty) {
                const long long gi = g_base + (ty - w.ty0);
                const int s = (int)(gi & (kSlots - 1));
                const unsigned par = (unsigned)((gi >> 4) & 1);
                // the activations of this row do not depend on the accumulators: fetch them while the MMAs run
                float hv[kNB];
                float *hp = a.H ? a.H + (long long)n * g.hsn + (long long)ty * g.hsy + tx : nullptr;
                if (a.H) {
#pragma unroll
                    for (int ml = 0; ml < kNB; ++ml)
                        hv[ml] = (active && a.m0 + ml < g.M) ? hp[(long long)(a.m0 + ml) * g.hsm] : 0.f;
                }
                mbar_wait(&row_done[s], par);
                tc_fence_after();
                float neg[kNB], pos[kNB];
                tmem_ld16(tmem_base + lane_base + (unsigned)(s * kNB), neg);
                tmem_ld16(tmem_base + lane_base + 256u + (unsigned)(s * kNB), pos);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&slot_free[s]);
                if (!active) continue;
                const long long tin = (long long)ty * g.TX + tx;
#pragma unroll
                for (int ml = 0; ml < kNB; ++ml) {
                    const int m = a.m0 + ml;
                    if (m >= g.M) continue;
                    const long long cidx = ((long long)n * g.M + m) * tvol + tin;
                    if (a.H) {
                        const float h = hv[ml];
                        float ps = pos[ml];
                        if (a.G) {
                            const float gv = a.G[cidx];
                            if (a.lambda != 0.f) { float tmp = gv - h; tmp *= a.lambda; ps += tmp; }
                            if (a.Gsum) {
                                const float gs = a.Gsum[(long long)n * tvol + tin];
                                float tmp = -gv + gs; tmp *= a.lambda_cross; ps += tmp;
                            }
                        }
                        ps += a.reg;
                        float hn = h * neg[ml];
                        hn /= ps;
                        hp[(long long)m * g.hsm] = hn;
                    } else {
                        a.neg[cidx] = neg[ml];
                        a.pos[cidx] = pos[ml];
                    }
                }
            }
            g_base += w.ty1 - w.ty0;
        }
    } else {
      if (lane == 0) {
        // ------------------------------------ MMA issuer ------------------------------------
        const unsigned lbo_a = kTile * 16, lbo_b = (unsigned)NR * 16;
        const unsigned w_hi_addr = smem_u32(w_hi), w_lo_addr = smem_u32(w_lo);
        const unsigned stage_addr0 = smem_u32(stages);
        int st = 0;
        unsigned ph = 0;
        long long g_base = 0;
        for (long long u = blockIdx.x; u < p.units; u += gridDim.x) {
            const Unit w = make_unit(u, g, p);
            int next_new = w.ty0, next_done = w.ty0;
            for (int r = w.r_lo; r <= w.r_hi; ++r) {
                const int ay_lo = max(0, r + g.offy - (w.ty1 - 1)), ay_hi = min(AY - 1, r + g.offy - w.ty0);
                const int t_a = r + g.offy - ay_hi, t_b = r + g.offy - ay_lo;
                const int j0 = r + g.offy - AY + 1;                 // output row of atom-row block 0
                // output rows that receive their first contribution from this source row: [first_new, t_b]
                const int first_new = next_new;
                for (; next_new <= t_b; ++next_new) {
                    const long long gi = g_base + (next_new - w.ty0);
                    if (gi >= kSlots) mbar_wait(&slot_free[gi & (kSlots - 1)], (unsigned)(((gi >> 4) - 1) & 1));
                }
                tc_fence_after();
                for (int x = 0; x < 2; ++x) {
                    mbar_wait(&a_full[st], ph);
                    tc_fence_after();
                    const unsigned col_base = tmem_base + (x ? 256u : 0u);
                    const unsigned a_hi_addr = stage_addr0 + (unsigned)st * (unsigned)p.stage_floats * 4u;
                    const unsigned a_lo_addr = a_hi_addr + kTile * KP * 4u;
                    for (int ks = 0; ks < p.ksteps; ++ks) {
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            const unsigned long long da =
                                smem_desc((t == 1 ? a_lo_addr : a_hi_addr) + ks * 2 * lbo_a, lbo_a, 128);
                            const unsigned b_addr = (t == 2 ? w_lo_addr : w_hi_addr) + ks * 2 * lbo_b;
                            // rows [lo, hi] of the window, accumulate flag
                            auto issue = [&](int lo, int hi, unsigned acc) {
                                if (lo > hi) return;
                                const int cnt = hi - lo + 1;
                                const int s = (int)((g_base + (lo - w.ty0)) & (kSlots - 1));
                                const int first = min(cnt, kSlots - s);
                                mma_tf32(col_base + (unsigned)(s * kNB), da,
                                         smem_desc(b_addr + (unsigned)(lo - j0) * 256u, lbo_b, 128),
                                         idesc_tf32(kTile, kNB * first), acc);
                                if (cnt > first)
                                    mma_tf32(col_base, da,
                                             smem_desc(b_addr + (unsigned)(lo - j0 + first) * 256u, lbo_b, 128),
                                             idesc_tf32(kTile, kNB * (cnt - first)), acc);
                            };
                            if (ks == 0 && t == 0 && first_new <= t_b) {
                                issue(t_a, first_new - 1, 1u);
                                issue(max(first_new, t_a), t_b, 0u);
                            } else {
                                issue(t_a, t_b, 1u);
                            }
                        }
                    }
                    mma_commit(&a_empty[st]);
                    if (++st == p.n_stages) { st = 0; ph ^= 1u; }
                }
                // output rows whose last source row this was
                for (; next_done < w.ty1 && min(g.DY - 1, next_done - g.offy + AY - 1) <= r; ++next_done)
                    mma_commit(&row_done[(g_base + (next_done - w.ty0)) & (kSlots - 1)]);
            }
            g_base += w.ty1 - w.ty0;
        }
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc

// ---- dispatch ----------------------------------------------------------------------------------------------------------
bool tc_hupd_supported(const Geo &g, int dtype) {
    if (dtype != TNMF_F32 || g.wrap) return false;
    if (g.D[0] != 1 || g.A[0] != 1 || g.T[0] != 1) return false;      // rank <= 2
    if (g.D[1] == 1 && g.A[1] == 1) return false;                     // rank 1: the FP32 kernels serve it
    if (g.N < 1) return false;
    tc::TcHupdPlan p;
    return tc::make_tc_hupd_plan(tiled::make_geo2(g), p);
}

int tc_gradient_h(const Geo &g, const float *V, const float *R, const float *W, float *neg, float *pos, float *H,
                  double reg, const float *G, double lambda, const float *Gsum, double lambda_cross, cudaStream_t st) {
    const tiled::Geo2 q = tiled::make_geo2(g);
    tc::TcHupdPlan p;
    if (!tc::make_tc_hupd_plan(q, p)) return TNMF_EUNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(tc::hupd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kMaxSmem);
    if (e != cudaSuccess) return status_from_cuda(e);
    tc::TcHupdArgs a;
    a.V = V; a.R = R; a.W = W; a.neg = neg; a.pos = pos; a.H = H;
    a.reg = (float)reg; a.lambda = (float)lambda; a.lambda_cross = (float)lambda_cross;
    a.G = G; a.Gsum = Gsum;
    for (int m0 = 0; m0 < g.M; m0 += tc::kNB) {
        a.m0 = m0;
        tc::hupd_tc_kernel<<<(unsigned)p.grid, tc::kThreads, p.smem, st>>>(q, p, a);
        TNMF_CHECK_LAUNCH();
    }
    return TNMF_OK;
}

int tc_hupd_launches(const Geo &g) { return tiled::ceil_div(g.M, tc::kNB); }

}  // namespace tnmf
