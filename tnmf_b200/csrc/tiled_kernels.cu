// Launch planning and dispatch of the register-tiled FP32-FMA kernels (tiled_recon.cu, tiled_hupd.cu,
// tiled_gradw.cu).  The planners pick, per problem geometry, the atom-width chunking, the lane / warp arrangement
// that wastes the fewest lanes on the output extent, the shared-memory tile and the grid; they are pure host code
// and are exercised on the CPU by tests/test_cabi.py through tnmf_uses_tiled_path / tnmf_workspace_bytes.
#include <cstdlib>
#include "tiled_common.cuh"

namespace tnmf {
namespace tiled {

static constexpr int kSMs = 148;

// CTA shape search shared by the two position-tiled kernels: 8 warps arranged WX x WY, chosen to minimise the
// staged area (halo included) over all tiles of one plane.
static bool finish_tile_plan(TilePlan &p, int EY, int EX, int AY, int stage_planes, int taps_floats, int n_items,
                             int max_warps) {
    const int AXP = p.ch.AXP;
    long long best_cost = -1;
    TilePlan best = p;
    const int wtile_y = p.LY * p.RB, wtile_x = p.LX * kCols;
    const int need_wy = ceil_div(EY, wtile_y), need_wx = ceil_div(EX, wtile_x);
    for (int wx = 1; wx <= max_warps; wx <<= 1) {
        TilePlan q = p;
        q.WX = wx < need_wx ? wx : need_wx;
        q.WY = max_warps / wx < need_wy ? max_warps / wx : need_wy;
        q.threads = 32 * q.WX * q.WY;
        q.tile_y = q.WY * wtile_y;
        q.tile_x = q.WX * wtile_x;
        q.tiles_y = ceil_div(EY, q.tile_y);
        q.tiles_x = ceil_div(EX, q.tile_x);
        q.HR = q.tile_y + AY - 1;
        q.WT = q.tile_x + AXP - 1;
        q.pitch = round_up(q.tile_x + AXP, 32);
        q.plane_floats = q.HR * q.pitch;
        q.taps_floats = round_up(taps_floats, 32);
        q.stage_floats = stage_planes * q.plane_floats + q.taps_floats;
        q.n_stages = n_items > 1 ? 2 : 1;
        q.smem = (size_t)q.n_stages * q.stage_floats * sizeof(float);
        if (q.smem > (size_t)kMaxSmem) continue;
        // staged floats per plane plus a penalty for idle warps in ragged tiles
        const long long staged = (long long)q.tiles_y * q.tiles_x * q.plane_floats;
        const long long warp_slots = (long long)q.tiles_y * q.tiles_x * q.WX * q.WY;
        const long long cost = staged + warp_slots * wtile_y * wtile_x / 4;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = q; }
    }
    if (best_cost < 0) return false;
    p = best;
    return true;
}

bool make_recon_plan(const Geo2 &g, TilePlan &p) {
    p = TilePlan();
    p.ch = choose_chunk(g.AX);
    const bool one_d = g.DY == 1 && g.AY == 1;
    const int ncb = ceil_div(g.C, 4);
    p.NB = ceil_div(g.C, ncb);                    // balanced channel blocks of at most 4
    p.nblk = ncb;
    int rb = one_d ? 1 : (p.NB == 1 ? 4 : (p.NB == 2 ? 2 : 1));
    for (; rb >= 1; rb >>= 1) {
        p.RB = rb;
        if (one_d) {
            p.LX = 32;
        } else if (rb > 1) {
            p.LX = 8;                             // RB consecutive rows per thread need a quarter-warp in one row
            if (round_up(g.DY, 4 * rb) > g.DY + g.DY / 6 && rb > 1) continue;   // too much row padding
        } else {
            p.LX = choose_lx(g.DY, g.DX, 1);
        }
        p.LY = 32 / p.LX;
        if (finish_tile_plan(p, g.DY, g.DX, g.AY, 1, g.AY * p.ch.AXP * p.NB, g.M, 8)) break;
    }
    if (rb < 1) return false;
    p.grid = (long long)p.tiles_x * p.tiles_y * p.nblk * g.N;
    return p.grid > 0 && p.grid < 0x7fffffffLL;
}

bool make_hupd_plan(const Geo2 &g, TilePlan &p) {
    p = TilePlan();
    p.ch = choose_chunk(g.AX);
    const bool one_d = g.TY == 1 && g.AY == 1;
    const int mb_max = 4;
    const int nmb = ceil_div(g.M, mb_max);
    p.NB = ceil_div(g.M, nmb);
    p.nblk = nmb;
    p.RB = 1;
    p.LX = one_d ? 32 : choose_lx(g.TY, g.TX, 1);
    p.LY = 32 / p.LX;
    if (!finish_tile_plan(p, g.TY, g.TX, g.AY, 2, g.AY * p.ch.AXP * p.NB, g.C, 8)) return false;
    p.grid = (long long)p.tiles_x * p.tiles_y * p.nblk * g.N;
    return p.grid > 0 && p.grid < 0x7fffffffLL;
}

bool make_gradw_plan(const Geo2 &g, GradWPlan &p) {
    p = GradWPlan();
    p.ch = choose_chunk(g.AX);
    const int AXC = p.ch.AXC;
    p.ncb = ceil_div(g.C, 3);
    p.CB = ceil_div(g.C, p.ncb);
    p.units = g.AY * p.ch.NK;
    // tap units per warp: as many as the accumulator budget (96) allows while keeping at least 8 warps busy
    p.BYB = 1;
    if (p.CB <= 2) {
        for (int b = 3; b >= 2; --b) {
            if (2 * p.CB * AXC * b > 96) continue;
            if (p.CB == 2 && b != 2) continue;                   // instantiated combinations
            if (ceil_div(p.units, b) >= 8) { p.BYB = b; break; }
        }
    }
    const int max_warps = 12;
    const int wunits = ceil_div(p.units, p.BYB);                 // warps' worth of tap units
    p.ugroups = ceil_div(wunits, max_warps);
    p.warps = ceil_div(wunits, p.ugroups);
    p.threads = 32 * p.warps;
    p.LX = (g.DY == 1) ? 32 : choose_lx(g.DY, g.DX, 1);
    p.LY = 32 / p.LX;
    p.XC = round_up(g.DX < 2048 ? g.DX : 2048, kCols * p.LX);
    p.nx = ceil_div(g.DX, p.XC);
    p.pitch_x = round_up(p.XC, 32);
    p.pitch_h = round_up(p.XC + p.ch.AXP, 32);
    // the widest range of atom rows one unit group touches
    int by_span = 0;
    for (int ug = 0; ug < p.ugroups; ++ug) {
        const int lo = (ug * p.warps * p.BYB) / p.ch.NK;
        int last = (ug + 1) * p.warps * p.BYB - 1;
        if (last > p.units - 1) last = p.units - 1;
        const int span = last / p.ch.NK - lo;
        if (span > by_span) by_span = span;
    }
    // rows per work item: as many as two stages allow, a multiple of LY, at most 16
    int ry = round_up(g.DY < 16 ? g.DY : 16, p.LY);
    for (;; ry -= p.LY) {
        if (ry < p.LY) return false;
        p.RY = ry;
        p.x_floats = 2 * p.CB * p.RY * p.pitch_x;
        p.h_floats = (p.RY + by_span) * p.pitch_h;
        p.stage_floats = p.x_floats + p.h_floats;
        p.smem = (size_t)2 * p.stage_floats * sizeof(float);
        if (p.smem <= (size_t)kMaxSmem) break;
    }
    p.ny = ceil_div(g.DY, p.RY);
    p.items = (long long)g.N * p.ny * p.nx;
    p.groups = g.M * p.ncb * p.ugroups;
    const long long total = (long long)p.groups * p.items;
    if (total <= 0) return false;
    long long ctas = kSMs;
    if (ctas > total) ctas = total;
    p.chunk = (total + ctas - 1) / ctas;
    p.grid = (int)((total + p.chunk - 1) / p.chunk);
    p.smax = (int)((p.items + p.chunk - 1) / p.chunk) + 1;
    return true;
}

}  // namespace tiled

using namespace tiled;

bool tiled_supported(const Geo &g, int dtype) {
    if (dtype != TNMF_F32) return false;
    if (g.D[0] != 1 || g.A[0] != 1 || g.T[0] != 1) return false;    // rank <= 2
    if (g.N < 1) return true;
    const Geo2 q = make_geo2(g);
    TilePlan tp;
    GradWPlan gp;
    return make_recon_plan(q, tp) && make_hupd_plan(q, tp) && make_gradw_plan(q, gp);
}

size_t tiled_workspace_bytes(const Geo &g) {
    const Geo2 q = make_geo2(g);
    size_t bytes = 0;
    TilePlan tp;
    if (g.N >= 1 && make_recon_plan(q, tp)) bytes = sizeof(double) * (size_t)tp.grid;    // energy partials
    GradWPlan gp;
    if (g.N >= 1 && make_gradw_plan(q, gp)) {
        const size_t w = (size_t)gp.smax * 2 * (size_t)g.M * g.C * g.A[1] * g.A[2] * sizeof(float);
        if (w > bytes) bytes = w;
    }
    return bytes;
}

int tiled_reconstruct(const Geo &g, const float *W, const float *H, float *R, const float *V,
                      double *energy_partials, int *n_partials, cudaStream_t st) {
    const Geo2 q = make_geo2(g);
    TilePlan p;
    if (!make_recon_plan(q, p)) return TNMF_EUNSUPPORTED;
    if (n_partials) *n_partials = (int)p.grid;
    switch (p.ch.AXC) {
        case 4: return recon_launch_axc<4>(q, p, W, H, R, V, energy_partials, st);
        case 8: return recon_launch_axc<8>(q, p, W, H, R, V, energy_partials, st);
        case 12: return recon_launch_axc<12>(q, p, W, H, R, V, energy_partials, st);
        case 16: return recon_launch_axc<16>(q, p, W, H, R, V, energy_partials, st);
        default: return TNMF_EUNSUPPORTED;
    }
}

int tiled_gradient_h(const Geo &g, const float *V, const float *R, const float *W, float *neg, float *pos,
                     float *H, double reg, const float *G, double lambda, const float *Gsum, double lambda_cross,
                     cudaStream_t st) {
    const Geo2 q = make_geo2(g);
    TilePlan p;
    if (!make_hupd_plan(q, p)) return TNMF_EUNSUPPORTED;
    const float fr = (float)reg, fl = (float)lambda, fc = (float)lambda_cross;
    switch (p.ch.AXC) {
        case 4: return hupd_launch_axc<4>(q, p, V, R, W, neg, pos, H, fr, G, fl, Gsum, fc, st);
        case 8: return hupd_launch_axc<8>(q, p, V, R, W, neg, pos, H, fr, G, fl, Gsum, fc, st);
        case 12: return hupd_launch_axc<12>(q, p, V, R, W, neg, pos, H, fr, G, fl, Gsum, fc, st);
        case 16: return hupd_launch_axc<16>(q, p, V, R, W, neg, pos, H, fr, G, fl, Gsum, fc, st);
        default: return TNMF_EUNSUPPORTED;
    }
}

int tiled_gradient_w(const Geo &g, const float *V, const float *R, const float *H, float *neg, float *pos,
                     void *workspace, size_t workspace_bytes, cudaStream_t st) {
    const Geo2 q = make_geo2(g);
    GradWPlan p;
    if (!make_gradw_plan(q, p)) return TNMF_EUNSUPPORTED;
    const long long count = (long long)g.M * g.C * g.A[1] * g.A[2];
    const size_t need = (size_t)p.smax * 2 * (size_t)count * sizeof(float);
    if (!workspace || workspace_bytes < need) return TNMF_EWORKSPACE;
    float *partials = (float *)workspace;
    cudaError_t e = cudaMemsetAsync(partials, 0, need, st);      // slices a group's CTAs do not reach stay zero
    if (e != cudaSuccess) return status_from_cuda(e);
    int s;
    switch (p.ch.AXC) {
        case 4: s = gradw_launch_axc<4>(q, p, V, R, H, partials, st); break;
        case 8: s = gradw_launch_axc<8>(q, p, V, R, H, partials, st); break;
        case 12: s = gradw_launch_axc<12>(q, p, V, R, H, partials, st); break;
        case 16: s = gradw_launch_axc<16>(q, p, V, R, H, partials, st); break;
        default: s = TNMF_EUNSUPPORTED;
    }
    if (s) return s;
    return finish_gradient_w<float>(partials, p.smax, count, neg, pos, st);
}

}  // namespace tnmf
