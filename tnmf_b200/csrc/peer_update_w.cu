// W-gradient all-reduce FUSED with the W update, over NVLink peer memory (one kernel, no NCCL call on the step).
//
//   grad_sum[m,c,a] = sum_ranks grad_r[m,c,a]           (tnmf/TransformInvariantNMF.py:457-465: Cyclic_MU's sum over sample
//                                                        blocks - here the blocks live on different GPUs)
//   W = (W * neg_sum) / (pos_sum + eps);  W[m,c,:] /= sum_a W[m,c,a]        (tnmf/TransformInvariantNMF.py:217-244,
//                                                                            tnmf/backends/_Backend.py:75-77)
//
// The payload is tiny (2 * M*C*prod(A) numbers: 46 KB on cfg2) and the step waits for it, so what counts is latency, not
// bandwidth: an NCCL all-reduce costs a kernel launch of its own plus 25-40 us of protocol between the W gradient and the
// W update of every iteration.  Here every rank owns a SYMMETRIC buffer (same layout on every GPU, mapped into every
// peer's address space):   data[2 parities][world][2 * count]   flags[world][pairs]
// and ONE kernel does the whole exchange:  a block takes (atom, channel) pairs; for each it
//   1. PUSHES its rank's neg/pos segment of the pair into slot [parity][rank] of EVERY rank's buffer (plain stores over
//      NVLink; NVSwitch gives every peer full bandwidth at once),
//   2. makes them visible (__threadfence_system) and raises flag[rank][pair] = epoch on every rank (st.release.sys),
//   3. waits until all `world` flags of the pair in its OWN buffer show the epoch (ld.acquire.sys),
//   4. sums the `world` segments in RANK ORDER in double - every rank adds the same numbers in the same order, so the
//      dictionary stays bit-identical on all ranks - and applies the multiplicative update and the normalisation.
// No grid-wide barrier: pair p of rank A depends only on pair p of the other ranks.  All pushes of a block precede its
// first wait, so blocks never wait for each other in a cycle.
// Reuse: the epoch counter lives in device memory (`state[0]`, advanced by the last block of each call), so a CUDA graph
// can replay the kernel.  Slots alternate with the epoch's parity: slot parity e is rewritten in call e + 2, which a rank
// only reaches after its call e + 1 saw the flags of every peer's call e + 1 - issued after that peer finished reading e.
#include "common.cuh"

namespace tnmf {

namespace {

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct Peers {
    int world, rank;
    void *buf[TNMF_MAX_PEERS];
};

template <typename T>
__global__ void __launch_bounds__(256) allreduce_update_w_kernel(T *__restrict__ W, const T *__restrict__ grad,
                                                                const Peers pw, unsigned *__restrict__ state, T eps,
                                                                long long avol, int pairs, long long flags_offset) {
    __shared__ double red[256];
    __shared__ T total;
    const int tid = threadIdx.x, world = pw.world, rank = pw.rank;
    const long long count = (long long)pairs * avol;
    const unsigned epoch = *(volatile unsigned *)state + 1u;
    const long long slot = ((long long)(epoch & 1u) * world) * 2 * count;

    // ---- push: this rank's segment of every pair of the block -> slot [parity][rank] of every rank ----
    for (int pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
        const T *ng = grad + (long long)pair * avol, *ps = ng + count;
        for (int i = 0; i < world; ++i) {
            const int r = (rank + i) % world;                   // start with the own buffer, spread the links
            T *dst = (T *)pw.buf[r] + slot + (long long)rank * 2 * count + (long long)pair * avol;
            for (long long a = tid; a < avol; a += 256) {
                dst[a] = ng[a];
                dst[count + a] = ps[a];
            }
        }
        __threadfence_system();
        __syncthreads();
        if (tid < world)
            st_release_sys((unsigned *)((char *)pw.buf[tid] + flags_offset) + (long long)rank * pairs + pair, epoch);
    }

    // ---- wait, sum in rank order, update, normalise ----
    const T *mine = (const T *)pw.buf[rank] + slot;
    const unsigned *flags = (const unsigned *)((const char *)pw.buf[rank] + flags_offset);
    for (int pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
        if (tid < world)
            while ((int)(ld_acquire_sys(flags + (long long)tid * pairs + pair) - epoch) < 0) __nanosleep(20);
        __syncthreads();
        T *w = W + (long long)pair * avol;
        double s = 0.0;
        for (long long a = tid; a < avol; a += 256) {
            double n = 0.0, p = 0.0;
            for (int r = 0; r < world; ++r) {                   // written by the peers: read past L1
                const T *seg = mine + (long long)r * 2 * count + (long long)pair * avol;
                n += (double)__ldcg(seg + a);
                p += (double)__ldcg(seg + count + a);
            }
            T pp = (T)p;
            pp += eps;
            T v = w[a] * (T)n;
            v /= pp;
            w[a] = v;
            s += (double)v;
        }
        red[tid] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        if (tid == 0) total = (T)red[0];
        __syncthreads();
        const T t = total;
        for (long long a = tid; a < avol; a += 256) w[a] /= t;
        __syncthreads();
    }

    // ---- the last block to leave advances the epoch (every block has read it by then) ----
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&state[1], 1u) == gridDim.x - 1) {
            state[1] = 0u;
            __threadfence();
            *(volatile unsigned *)state = epoch;
        }
    }
}

}  // namespace

size_t peer_flags_offset(const Geo &g, int dtype, int world) {
    const size_t count = (size_t)g.M * g.C * vol3(g.A);
    const size_t data = (size_t)2 * world * 2 * count * (dtype == TNMF_F32 ? 4 : 8);
    return (data + 127) & ~(size_t)127;
}

size_t peer_buffer_bytes(const Geo &g, int dtype, int world) {
    return peer_flags_offset(g, dtype, world) + (size_t)world * g.M * g.C * sizeof(unsigned);
}

template <typename T>
int allreduce_update_w(const Geo &g, int dtype, T *W, const T *grad, const tnmf_peer_world *pw, unsigned *state,
                       double eps, cudaStream_t st) {
    Peers p;
    p.world = pw->world;
    p.rank = pw->rank;
    for (int i = 0; i < TNMF_MAX_PEERS; ++i) p.buf[i] = i < pw->world ? pw->buffers[i] : nullptr;
    const int pairs = g.M * g.C;
    const int sms = sm_count_cached();
    const int grid = pairs < sms ? pairs : sms;                  // co-resident: a block only ever waits for peers
    allreduce_update_w_kernel<T><<<grid, 256, 0, st>>>(W, grad, p, state, (T)eps, vol3(g.A), pairs,
                                                        (long long)peer_flags_offset(g, dtype, pw->world));
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}
template int allreduce_update_w<float>(const Geo &, int, float *, const float *, const tnmf_peer_world *, unsigned *, double,
                                       cudaStream_t);
template int allreduce_update_w<double>(const Geo &, int, double *, const double *, const tnmf_peer_world *, unsigned *,
                                        double, cudaStream_t);

}  // namespace tnmf
