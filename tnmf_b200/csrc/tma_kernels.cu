// Launch planning, tensor-map construction and dispatch of the persistent TMA kernels (tma_recon.cu, tma_hupd.cu,
// tma_gradw.cu), plus the two tiny kernels that pre-arrange the atoms into the order the consumer warps read them.
//
// A problem takes this path when it is float, has at most two shift axes, is not 'circular' (TMA cannot wrap) and
// every tensor a TMA box is cut from has a 16-byte aligned base and row / plane strides that are multiples of 16
// bytes: V and R need DX % 4 == 0, H needs h_pitch % 4 == 0 (B200_Backend allocates H with a padded row pitch for
// that).  Everything else falls back to the cp.async kernels of tiled_kernels.cu.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "tma_common.cuh"

namespace tnmf {
namespace tma {

using tiled::ceil_div;
using tiled::choose_chunk;
using tiled::round_up;

// ---- driver entry point (no link-time dependency on libcuda) --------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

int encode_map(CUtensorMap *map, const void *base, int rank, const unsigned long long *dims,
               const unsigned long long *strides_bytes, const unsigned *box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return TNMF_EUNSUPPORTED;
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr,
                          bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? TNMF_OK : TNMF_EINVAL;
}

bool aligned16(const void *p) { return (reinterpret_cast<unsigned long long>(p) & 15ull) == 0; }

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            n = v;
        else
            return 148;                 // B200; also the answer when planning without a device (CPU tests)
    }
    return n;
}

// ---- plans -----------------------------------------------------------------------------------------------------------
// CTA shape shared by the two position-tiled kernels: WX x WY consumer warps (at most 12), each covering 32 columns x
// (8*RB) rows.  Cost model: a tile keeps its SM busy for max(active warps, 8) warp-tile periods (fewer than 8 active
// warps cannot hide the LDS latency of the FFMA stream), plus a small charge for the staged bytes (halo included).
template <typename Plan>
static bool finish_tile_plan(Plan &p, int EY, int EX_in, int AY, int RB, int planes, int taps_floats, int x_shift = 0,
                             int max_consumers = kConsumersMax) {
    const int EX = EX_in - x_shift;                               // columns to cover when the origin is shifted left
    const int AXP = p.ch.AXP;
    const int wtile_y = kLY * RB, wtile_x = kLX * kCols;
    const int need_wy = ceil_div(EY, wtile_y), need_wx = ceil_div(EX, wtile_x);
    const int force_wx = 0, force_wy = 0;      // tuning hook: a forced warp grid
    double best_cost = -1;
    Plan best = p;
    for (int wx = 1; wx <= max_consumers && wx <= need_wx; ++wx) {
        for (int wy = 1; wx * wy <= max_consumers && wy <= need_wy; ++wy) {
            if (force_wx && force_wy && (wx != (force_wx < need_wx ? force_wx : need_wx) ||
                                         wy != (force_wy < need_wy ? force_wy : need_wy)))
                continue;
            // warps are dealt round-robin to the 4 schedulers of an SM: a consumer count that is not a multiple of 4
            // leaves one scheduler with an extra warp and everyone waits for it (measured: 3x3 warps 30% slower)
            if (!(force_wx && force_wy) && (wx * wy) % 4 != 0 && (wx < need_wx || wy < need_wy)) continue;
            Plan q = p;
            q.WX = wx;
            q.WY = wy;
            q.consumers = wx * wy;
            q.tile_y = wy * wtile_y;
            q.tile_x = wx * wtile_x;
            q.tiles_y = ceil_div(EY, q.tile_y);
            q.tiles_x = ceil_div(EX, q.tile_x);
            q.HR = q.tile_y + AY - 1;
            q.pitch = odd_pitch(q.tile_x + AXP);
            if (q.pitch > 256 || q.HR > 256) continue;            // TMA box limits
            q.plane_floats = round_up(q.HR * q.pitch, 32);
            q.taps_floats = round_up(taps_floats, 32);
            q.stage_floats = planes * q.plane_floats + q.taps_floats;
            const size_t stage_bytes = (size_t)q.stage_floats * sizeof(float);
            q.n_stages = (int)(kMaxSmem / stage_bytes);
            if (q.n_stages > 6) q.n_stages = 6;
            if (q.n_stages < 2) continue;
            q.smem = (size_t)q.n_stages * stage_bytes;
            // active warps of the last tile column / row
            const int last_wx = need_wx - (q.tiles_x - 1) * wx, last_wy = need_wy - (q.tiles_y - 1) * wy;
            auto busy = [](int active) { return active < 8 ? 8 : active; };
            double periods = (double)(q.tiles_x - 1) * (q.tiles_y - 1) * busy(wx * wy) +
                             (double)(q.tiles_y - 1) * busy(last_wx * wy) + (double)(q.tiles_x - 1) * busy(wx * last_wy) +
                             busy(last_wx * last_wy);
            const double staged = (double)q.tiles_y * q.tiles_x * q.HR * q.pitch / ((double)EY * EX);   // ~1.3 .. 3
            const double cost = periods * (1.0 + 0.02 * staged) * (q.n_stages < 3 ? 1.15 : 1.0);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = q; }
        }
    }
    if (best_cost < 0) return false;
    p = best;
    p.threads = 32 * (p.consumers + 1);
    return true;
}

bool make_hupd_plan(const Geo2 &g, HupdPlan &p) {
    p = HupdPlan();
    p.ch = choose_chunk(g.AX);
    const int nmb = ceil_div(g.M, 4);
    p.MB = ceil_div(g.M, nmb);                      // balanced atom blocks of at most 4
    p.nblk = nmb;
    // box origin = x0 - offx must be a multiple of 4 elements: start the tile grid at x_shift = offx mod 4 (- 4)
    p.x_shift = (g.offx & 3) ? (g.offx & 3) - 4 : 0;
    p.wide = hupd_needs_wide(p.ch.AXC, p.MB);
    if (!finish_tile_plan(p, g.TY, g.TX, g.AY, 1, 2, g.AY * p.ch.AXP * p.MB, p.x_shift, p.wide ? 8 : kConsumersMax))
        return false;
    p.units = (long long)g.N * p.tiles_y * p.tiles_x * p.nblk;
    if (p.units <= 0 || p.units >= 0x7fffffffLL) return false;
    p.grid = (int)(p.units < sm_count() ? p.units : sm_count());
    return true;
}

// Taps of one (atom, channel block) slice of the reconstruction: plain atom rows for RB = 1, pairs over two adjacent
// atom rows (AY + 1 table rows, zeros past either end) for the packed-FP32 kernels (RB even).
static int recon_taps_floats(int AY, int AXP, int CB, int RB) { return RB > 1 ? 2 * (AY + 1) * AXP * CB : AY * AXP * CB; }

bool make_recon_plan(const Geo2 &g, ReconPlan &p) {
    p = ReconPlan();
    p.ch = choose_chunk(g.AX);
    const int ncb = ceil_div(g.C, 4);
    p.CB = ceil_div(g.C, ncb);                      // balanced channel blocks of at most 4
    p.nblk = ncb;
    // rows per thread: as many as keep the accumulators at <= 48 registers and the row padding small
    int rb = p.CB == 1 ? 4 : 2;
    // one-row atoms (1-D batches on the rows view): the packed kernel pairs the taps of two atom rows, and with A_y = 1
    // one half of every pair is the zero row - the scalar kernel does the same useful work in half the FMA-pipe time
    // (cfg4: 5.38 -> 2.82 ms per launch)
    if (g.AY == 1) rb = 1;
    for (; rb >= 1; rb >>= 1) {
        if (rb > 1 && round_up(g.DY, kLY * rb) > g.DY + g.DY / 6) continue;
        p.RB = rb;
        if (finish_tile_plan(p, g.DY, g.DX, g.AY, rb, 1, recon_taps_floats(g.AY, p.ch.AXP, p.CB, rb))) break;
    }
    if (rb < 1) return false;
    p.units = (long long)g.N * p.tiles_y * p.tiles_x * p.nblk;
    if (p.units <= 0 || p.units >= 0x7fffffffLL) return false;
    p.grid = (int)(p.units < sm_count() ? p.units : sm_count());
    return true;
}

bool make_gradw_plan(const Geo2 &g, GradWPlan &p) {
    p = GradWPlan();
    p.ch = choose_chunk(g.AX);
    const int AXC = p.ch.AXC;
    p.ncb = ceil_div(g.C, 3);
    p.CB = ceil_div(g.C, p.ncb);
    p.units = g.AY * p.ch.NK;
    // tap units per warp: as many as the accumulator budget (96) allows while keeping at least 4 warps busy
    p.BYB = 1;
    if (p.CB <= 2) {
        for (int b = 3; b >= 2; --b) {
            if (2 * p.CB * AXC * b > 96) continue;
            if (p.CB == 2 && b != 2) continue;                   // instantiated combinations
            if (p.CB == 2 && AXC > 12) continue;
            if (ceil_div(p.units, b) >= 4) { p.BYB = b; break; }
        }
    }
    const int max_warps = 11;
    const int wunits = ceil_div(p.units, p.BYB);                 // warps' worth of tap units
    p.ugroups = ceil_div(wunits, max_warps);
    p.UW = ceil_div(wunits, p.ugroups);
    p.RW = (2 * p.UW <= max_warps && g.DY >= 2 * kLY) ? 2 : 1;
    p.consumers = p.UW * p.RW;
    p.threads = 32 * (p.consumers + 1);
    // the widest range of atom rows one unit group touches
    p.by_span = 0;
    for (int ug = 0; ug < p.ugroups; ++ug) {
        const int lo = (ug * p.UW * p.BYB) / p.ch.NK;
        int last = (ug + 1) * p.UW * p.BYB - 1;
        if (last > p.units - 1) last = p.units - 1;
        const int span = last / p.ch.NK - lo;
        if (span > p.by_span) p.by_span = span;
    }
    // columns per work item: strips of 8 dealt to 4 lanes, box width <= 256
    int xc = round_up(g.DX < 128 ? g.DX : 128, kLX * kCols);
    while (xc > kLX * kCols && odd_pitch(xc + p.ch.AXP) > 256) xc -= kLX * kCols;
    if (odd_pitch(xc + p.ch.AXP) > 256) return false;
    p.XC = xc;
    p.nx = ceil_div(g.DX, p.XC);
    p.pitch_x = odd_pitch(p.XC);
    p.pitch_h = odd_pitch(p.XC + p.ch.AXP);
    // rows per work item: 16 or 8 per row-warp, whatever leaves at least 3 ring stages
    bool ok = false;
    for (int per = 16; per >= 8; per -= 8) {
        p.RY = per * p.RW;
        if (per > 8 && p.RY > round_up(g.DY, kLY * p.RW)) continue;
        p.hrows = p.RY + p.by_span;
        if (p.hrows > 256) continue;
        p.x_floats = 2 * p.CB * p.RY * p.pitch_x;
        p.h_floats = round_up(p.hrows * p.pitch_h, 32);
        p.stage_floats = p.x_floats + p.h_floats;
        const size_t stage_bytes = (size_t)p.stage_floats * sizeof(float);
        p.n_stages = (int)(kMaxSmem / stage_bytes);
        if (p.n_stages > 6) p.n_stages = 6;
        if (p.n_stages >= 3 || (per == 8 && p.n_stages >= 2)) { ok = true; p.smem = p.n_stages * stage_bytes; break; }
    }
    if (!ok) return false;
    p.ny = ceil_div(g.DY, p.RY);
    p.items = (long long)g.N * p.ny * p.nx;
    p.groups = g.M * p.ncb * p.ugroups;
    const long long total = (long long)p.groups * p.items;
    if (total <= 0) return false;
    long long ctas = sm_count();
    if (ctas > total) ctas = total;
    p.chunk = (total + ctas - 1) / ctas;
    p.grid = (int)((total + p.chunk - 1) / p.chunk);
    p.smax = ((int)((p.items + p.chunk - 1) / p.chunk) + 1) * p.RW;
    return true;
}

// ---- atom pre-arrangement ----------------------------------------------------------------------------------------------
// hupd: Wt[c][mb][ay][q][i][e] = W[mb*MB+i][c][ay][4q+e]   (zero beyond the atom / the atom count)
__global__ void prepare_taps_hupd_kernel(const float *__restrict__ W, float *__restrict__ Wt, int M, int C, int AY,
                                         int AX, int AXP, int MB, int nblk) {
    const int total = C * nblk * AY * AXP * MB;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int r = idx;
        const int e = r & 3; r >>= 2;
        const int i = r % MB; r /= MB;
        const int q = r % (AXP >> 2); r /= (AXP >> 2);
        const int ay = r % AY; r /= AY;
        const int mb = r % nblk;
        const int c = r / nblk;
        const int m = mb * MB + i, ax = 4 * q + e;
        Wt[idx] = (m < M && ax < AX) ? W[(((long long)m * C + c) * AY + ay) * AX + ax] : 0.f;
    }
}
// recon: Wt[m][cb][by][q][c][e] = W[m][cb*CB+c][AY-1-by][AX-1-(4q+e)]   (zero beyond the atom / the channel count)
// paired (packed-FP32 kernels): Wt[m][cb][t][q][c][e][h] = the same with by = t - h, t = 0..AY, zero for by outside [0, AY)
__global__ void prepare_taps_recon_kernel(const float *__restrict__ W, float *__restrict__ Wt, int M, int C, int AY,
                                          int AX, int AXP, int CB, int nblk, int paired) {
    const int rows = paired ? AY + 1 : AY;
    const int total = M * nblk * rows * AXP * CB * (paired ? 2 : 1);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int r = idx;
        int h = 0;
        if (paired) { h = r & 1; r >>= 1; }
        const int e = r & 3; r >>= 2;
        const int c = r % CB; r /= CB;
        const int q = r % (AXP >> 2); r /= (AXP >> 2);
        const int t = r % rows; r /= rows;
        const int cb = r % nblk;
        const int m = r / nblk;
        const int by = t - h;
        const int ch = cb * CB + c, ax = AX - 1 - (4 * q + e);
        Wt[idx] = (ch < C && ax >= 0 && by >= 0 && by < AY) ? W[(((long long)m * C + ch) * AY + (AY - 1 - by)) * AX + ax]
                                                            : 0.f;
    }
}

}  // namespace tma

// ---- dispatch ----------------------------------------------------------------------------------------------------------
using namespace tma;

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

static bool tma_geometry_ok(const Geo &g, int dtype) {
    if (dtype != TNMF_F32 || g.wrap) return false;
    if (g.D[0] != 1 || g.A[0] != 1 || g.T[0] != 1) return false;      // rank <= 2
    if (g.D[1] == 1 && g.A[1] == 1) return false;                     // rank 1: the cp.async kernels serve it
    if (g.flags & TNMF_FLAG_NO_TMA) return false;
    return g.N >= 1;
}

// V and R boxes need DX % 4 == 0; H boxes need 16-byte multiples for all of H's strides.
static bool vr_strides_ok(const Geo &g) { return (g.D[2] & 3) == 0; }
static bool h_strides_ok(const Geo &g) { return (g.hsy & 3) == 0 && (g.hsm & 3) == 0 && (g.hsn & 3) == 0; }
// H boxes of the reconstruction and of the W gradient start at x0 + off - (A_x - 1) with x0 a multiple of 4
static bool h_origin_ok(const Geo &g) { return ((g.off[2] - (g.A[2] - 1)) & 3) == 0; }

bool tma_hupd_supported(const Geo &g, int dtype) {
    HupdPlan p;
    return tma_geometry_ok(g, dtype) && vr_strides_ok(g) && make_hupd_plan(tiled::make_geo2(g), p);
}
bool tma_recon_supported(const Geo &g, int dtype) {
    ReconPlan p;
    return tma_geometry_ok(g, dtype) && h_strides_ok(g) && h_origin_ok(g) && make_recon_plan(tiled::make_geo2(g), p);
}
bool tma_gradw_supported(const Geo &g, int dtype) {
    GradWPlan p;
    return tma_geometry_ok(g, dtype) && vr_strides_ok(g) && h_strides_ok(g) && h_origin_ok(g) &&
           make_gradw_plan(tiled::make_geo2(g), p);
}

// Workspace layout: [ atom slices | energy partials or W-gradient partial slices ]
static size_t taps_region_bytes(const Geo &g) {
    const Geo2 q = tiled::make_geo2(g);
    size_t bytes = 0;
    HupdPlan hp;
    if (make_hupd_plan(q, hp)) bytes = (size_t)g.C * hp.nblk * q.AY * hp.ch.AXP * hp.MB * sizeof(float);
    ReconPlan rp;
    if (make_recon_plan(q, rp)) {
        const size_t b = (size_t)g.M * rp.nblk * recon_taps_floats(q.AY, rp.ch.AXP, rp.CB, rp.RB) * sizeof(float);
        if (b > bytes) bytes = b;
    }
    return align256(bytes);
}

size_t tma_workspace_bytes(const Geo &g, int dtype) {
    if (!tma_geometry_ok(g, dtype)) return 0;
    const Geo2 q = tiled::make_geo2(g);
    size_t rest = 0;
    ReconPlan rp;
    if (make_recon_plan(q, rp)) rest = sizeof(double) * (size_t)sm_count() * kConsumersMax;
    GradWPlan gp;
    if (make_gradw_plan(q, gp)) {
        const size_t w = (size_t)gp.smax * 2 * (size_t)g.M * g.C * g.A[1] * g.A[2] * sizeof(float);
        if (w > rest) rest = w;
    }
    return taps_region_bytes(g) + align256(rest);
}

static int map_vr4(CUtensorMap *map, const Geo2 &q, const float *X, int box_x, int box_y, int box_c) {
    const unsigned long long dims[4] = {(unsigned long long)q.DX, (unsigned long long)q.DY, (unsigned long long)q.C,
                                        (unsigned long long)q.N};
    const unsigned long long strides[3] = {(unsigned long long)q.DX * 4, (unsigned long long)q.DX * q.DY * 4,
                                           (unsigned long long)q.DX * q.DY * q.C * 4};
    const unsigned box[4] = {(unsigned)box_x, (unsigned)box_y, (unsigned)box_c, 1};
    return encode_map(map, X, 4, dims, strides, box);
}
static int map_h(CUtensorMap *map, const Geo2 &q, const float *H, int box_x, int box_y) {
    const unsigned long long dims[4] = {(unsigned long long)q.TX, (unsigned long long)q.TY, (unsigned long long)q.M,
                                        (unsigned long long)q.N};
    // a size-1 axis may carry any stride; give it a valid one
    const unsigned long long hsm = q.M > 1 ? (unsigned long long)q.hsm : (unsigned long long)q.hsy * q.TY;
    const unsigned long long hsn = q.N > 1 ? (unsigned long long)q.hsn : hsm * q.M;
    const unsigned long long strides[3] = {(unsigned long long)q.hsy * 4, hsm * 4, hsn * 4};
    const unsigned box[4] = {(unsigned)box_x, (unsigned)box_y, 1, 1};
    return encode_map(map, H, 4, dims, strides, box);
}

int tma_reconstruct(const Geo &g, const float *W, const float *H, float *R, const float *V, double *energy_partials,
                    int *n_partials, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    const Geo2 q = tiled::make_geo2(g);
    ReconPlan p;
    if (!make_recon_plan(q, p)) return TNMF_EUNSUPPORTED;
    const size_t taps_bytes = (size_t)g.M * p.nblk * recon_taps_floats(q.AY, p.ch.AXP, p.CB, p.RB) * sizeof(float);
    if (!workspace || workspace_bytes < tma_workspace_bytes(g, TNMF_F32)) return TNMF_EWORKSPACE;
    float *Wt = (float *)workspace;
    const int total = (int)(taps_bytes / sizeof(float));
    prepare_taps_recon_kernel<<<ceil_div(total, 256) < 64 ? ceil_div(total, 256) : 64, 256, 0, st>>>(
        W, Wt, g.M, g.C, q.AY, q.AX, p.ch.AXP, p.CB, p.nblk, p.RB > 1 ? 1 : 0);
    TNMF_CHECK_LAUNCH();
    CUtensorMap mapH;
    int s = map_h(&mapH, q, H, p.pitch, p.HR);
    if (s) return s;
    ReconArgs a;
    a.Wt = Wt; a.R = R; a.V = V;
    a.epart = energy_partials;
    if (n_partials) *n_partials = p.grid * p.consumers;
    switch (p.ch.AXC) {
        case 4: return recon_launch_axc<4>(q, p, mapH, a, st);
        case 8: return recon_launch_axc<8>(q, p, mapH, a, st);
        case 12: return recon_launch_axc<12>(q, p, mapH, a, st);
        case 16: return recon_launch_axc<16>(q, p, mapH, a, st);
        default: return TNMF_EUNSUPPORTED;
    }
}

double *tma_energy_partials(const Geo &g, void *workspace) {
    return (double *)((char *)workspace + taps_region_bytes(g));
}

int tma_gradient_h(const Geo &g, const float *V, const float *R, const float *W, float *neg, float *pos, float *H,
                   double reg, const float *G, double lambda, const float *Gsum, double lambda_cross, void *workspace,
                   size_t workspace_bytes, cudaStream_t st) {
    const Geo2 q = tiled::make_geo2(g);
    HupdPlan p;
    if (!make_hupd_plan(q, p)) return TNMF_EUNSUPPORTED;
    if (!workspace || workspace_bytes < tma_workspace_bytes(g, TNMF_F32)) return TNMF_EWORKSPACE;
    float *Wt = (float *)workspace;
    const int total = g.C * p.nblk * q.AY * p.ch.AXP * p.MB;
    prepare_taps_hupd_kernel<<<ceil_div(total, 256) < 64 ? ceil_div(total, 256) : 64, 256, 0, st>>>(
        W, Wt, g.M, g.C, q.AY, q.AX, p.ch.AXP, p.MB, p.nblk);
    TNMF_CHECK_LAUNCH();
    CUtensorMap mapV, mapR;
    int s = map_vr4(&mapV, q, V, p.pitch, p.HR, 1);
    if (s) return s;
    s = map_vr4(&mapR, q, R, p.pitch, p.HR, 1);
    if (s) return s;
    HupdArgs a;
    a.Wt = Wt; a.neg = neg; a.pos = pos; a.H = H;
    a.reg = (float)reg; a.lambda = (float)lambda; a.lambda_cross = (float)lambda_cross;
    a.G = G; a.Gsum = Gsum;
    switch (p.ch.AXC) {
        case 4: return hupd_launch_axc<4>(q, p, mapV, mapR, a, st);
        case 8: return hupd_launch_axc<8>(q, p, mapV, mapR, a, st);
        case 12: return hupd_launch_axc<12>(q, p, mapV, mapR, a, st);
        case 16: return hupd_launch_axc<16>(q, p, mapV, mapR, a, st);
        default: return TNMF_EUNSUPPORTED;
    }
}

int tma_gradient_w(const Geo &g, const float *V, const float *R, const float *H, float *neg, float *pos,
                   void *workspace, size_t workspace_bytes, cudaStream_t st) {
    const Geo2 q = tiled::make_geo2(g);
    GradWPlan p;
    if (!make_gradw_plan(q, p)) return TNMF_EUNSUPPORTED;
    const long long count = (long long)g.M * g.C * g.A[1] * g.A[2];
    const size_t need = (size_t)p.smax * 2 * (size_t)count * sizeof(float);
    if (!workspace || workspace_bytes < tma_workspace_bytes(g, TNMF_F32)) return TNMF_EWORKSPACE;
    float *partials = (float *)((char *)workspace + taps_region_bytes(g));
    cudaError_t e = cudaMemsetAsync(partials, 0, need, st);      // slices a group's CTAs do not reach stay zero
    if (e != cudaSuccess) return status_from_cuda(e);
    CUtensorMap mapV, mapR, mapH;
    int s = map_vr4(&mapV, q, V, p.pitch_x, p.RY, p.CB);
    if (s) return s;
    s = map_vr4(&mapR, q, R, p.pitch_x, p.RY, p.CB);
    if (s) return s;
    s = map_h(&mapH, q, H, p.pitch_h, p.hrows);
    if (s) return s;
    switch (p.ch.AXC) {
        case 4: s = gradw_launch_axc<4>(q, p, mapV, mapR, mapH, partials, st); break;
        case 8: s = gradw_launch_axc<8>(q, p, mapV, mapR, mapH, partials, st); break;
        case 12: s = gradw_launch_axc<12>(q, p, mapV, mapR, mapH, partials, st); break;
        case 16: s = gradw_launch_axc<16>(q, p, mapV, mapR, mapH, partials, st); break;
        default: s = TNMF_EUNSUPPORTED;
    }
    if (s) return s;
    return finish_gradient_w<float>(partials, p.smax, count, neg, pos, st);
}

}  // namespace tnmf
