// Small finishing / elementwise kernels: final reductions of the split-K partials, the W update with the
// per-(atom, channel) normalisation, the separable inhibition convolution, and the FP32 peak probe.
// All of them are HBM/L2-bound and tiny next to the correlations.
#include "common.cuh"

namespace tnmf {

int sm_count_cached() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
    }
    return sms;
}

// ---------------------------------------------------------------------------------------------------------
// energy: fixed-order sum of the per-block partials, E = 0.5 * sum          (tnmf/backends/_Backend.py:127-130)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) finish_energy_kernel(const double *__restrict__ partials, int n,
                                                           double *__restrict__ energy) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *energy = 0.5 * red[0];
}

int finish_energy(const double *partials, int n, double *energy, cudaStream_t st) {
    finish_energy_kernel<<<1, 256, 0, st>>>(partials, n, energy);
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}

// ---------------------------------------------------------------------------------------------------------
// W gradient: partials[p][2][count] -> neg[count], pos[count], summed over p in a fixed order in double
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) finish_gradient_w_kernel(const T *__restrict__ partials, int n_partials,
                                                               long long count, T *__restrict__ neg,
                                                               T *__restrict__ pos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * count) return;
    double s = 0.0;
    for (int p = 0; p < n_partials; ++p) s += (double)partials[(long long)p * 2 * count + i];
    if (i < count) neg[i] = (T)s; else pos[i - count] = (T)s;
}

template <typename T>
int finish_gradient_w(const T *partials, int n_partials, long long count, T *neg, T *pos, cudaStream_t st) {
    const long long blocks = (2 * count + 255) / 256;
    finish_gradient_w_kernel<T><<<(int)blocks, 256, 0, st>>>(partials, n_partials, count, neg, pos);
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}
template int finish_gradient_w<float>(const float *, int, long long, float *, float *, cudaStream_t);
template int finish_gradient_w<double>(const double *, int, long long, double *, double *, cudaStream_t);

// ---------------------------------------------------------------------------------------------------------
// W update: pos += eps; W *= neg; W /= pos; W[m,c,:] /= sum_a W[m,c,a]
// (tnmf/TransformInvariantNMF.py:217-244, tnmf/backends/_Backend.py:75-77).  One block per (atom, channel).
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) update_w_kernel(T *__restrict__ W, const T *__restrict__ neg,
                                                      const T *__restrict__ pos, T eps, long long avol) {
    __shared__ double red[256];
    __shared__ T total;
    T *w = W + (long long)blockIdx.x * avol;
    const T *ng = neg + (long long)blockIdx.x * avol;
    const T *ps = pos + (long long)blockIdx.x * avol;
    double s = 0.0;
    for (long long a = threadIdx.x; a < avol; a += 256) {
        T p = ps[a];
        p += eps;
        T v = w[a] * ng[a];
        v /= p;
        w[a] = v;
        s += (double)v;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) total = (T)red[0];
    __syncthreads();
    const T t = total;
    for (long long a = threadIdx.x; a < avol; a += 256) w[a] /= t;
}

template <typename T>
int update_w(const Geo &g, T *W, const T *neg, const T *pos, double eps, cudaStream_t st) {
    update_w_kernel<T><<<g.M * g.C, 256, 0, st>>>(W, neg, pos, (T)eps, vol3(g.A));
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}
template int update_w<float>(const Geo &, float *, const float *, const float *, double, cudaStream_t);
template int update_w<double>(const Geo &, double *, const double *, const double *, double, cudaStream_t);

// ---------------------------------------------------------------------------------------------------------
// normalize: arr[o,:,i] /= sum_l arr[o,l,i]                                (tnmf/backends/_Backend.py:75-77)
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) normalize_axis_kernel(T *__restrict__ arr, long long len, long long inner) {
    __shared__ double red[128];
    __shared__ T total;
    const long long o = blockIdx.x / inner, i = blockIdx.x % inner;
    T *base = arr + o * len * inner + i;
    double s = 0.0;
    for (long long l = threadIdx.x; l < len; l += 128) s += (double)base[l * inner];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 64; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) total = (T)red[0];
    __syncthreads();
    const T t = total;
    for (long long l = threadIdx.x; l < len; l += 128) base[l * inner] /= t;
}

template <typename T>
int normalize_axis(T *arr, long long outer, long long len, long long inner, cudaStream_t st) {
    const long long blocks = outer * inner;
    if (blocks <= 0 || blocks > 0x7fffffffLL) return TNMF_EINVAL;
    normalize_axis_kernel<T><<<(int)blocks, 128, 0, st>>>(arr, len, inner);
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}
template int normalize_axis<float>(float *, long long, long long, long long, cudaStream_t);
template int normalize_axis<double>(double *, long long, long long, long long, cudaStream_t);

// ---------------------------------------------------------------------------------------------------------
// one axis of the separable inhibition convolution, zero boundary, centred odd kernel
// (tnmf/backends/_NumPyBackend.py:56-64: scipy.ndimage.convolve1d(mode='constant', cval=0), which accumulates
//  in double and rounds to the array type once per axis)
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) convolve_axis_kernel(const T *__restrict__ in, T *__restrict__ out,
                                                           long long total, long long len, long long inner,
                                                           const double *__restrict__ taps, int n_taps) {
    const int r = (n_taps - 1) / 2;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx % inner;
        const long long l = (idx / inner) % len;
        const long long o = idx / (inner * len);
        const T *base = in + o * len * inner + i;
        double acc = 0.0;
        for (int j = -r; j <= r; ++j) {
            const long long src = l - j;
            if (src >= 0 && src < len) acc += taps[r + j] * (double)base[src * inner];
        }
        out[idx] = (T)acc;
    }
}

template <typename T>
int convolve_axis(const T *in, T *out, long long outer, long long len, long long inner, const double *taps,
                  int n_taps, cudaStream_t st) {
    if (n_taps < 1 || (n_taps & 1) == 0) return TNMF_EINVAL;
    const long long total = outer * len * inner;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
    convolve_axis_kernel<T><<<(int)blocks, 256, 0, st>>>(in, out, total, len, inner, taps, n_taps);
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}
template int convolve_axis<float>(const float *, float *, long long, long long, long long, const double *, int,
                                  cudaStream_t);
template int convolve_axis<double>(const double *, double *, long long, long long, long long, const double *, int,
                                   cudaStream_t);

// ---------------------------------------------------------------------------------------------------------
// Gsum[n,0,i] = sum_m G[n,m,i]                                   (tnmf/TransformInvariantNMF.py:263)
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) sum_atoms_kernel(const T *__restrict__ G, T *__restrict__ Gsum,
                                                       long long n, long long m, long long inner) {
    const long long total = n * inner;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx % inner, s = idx / inner;
        const T *base = G + s * m * inner + i;
        T acc = 0;
        for (long long k = 0; k < m; ++k) acc += base[k * inner];
        Gsum[idx] = acc;
    }
}

template <typename T>
int sum_atoms(const T *G, T *Gsum, long long n, long long m, long long inner, cudaStream_t st) {
    long long blocks = (n * inner + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
    sum_atoms_kernel<T><<<(int)blocks, 256, 0, st>>>(G, Gsum, n, m, inner);
    TNMF_CHECK_LAUNCH();
    return TNMF_OK;
}
template int sum_atoms<float>(const float *, float *, long long, long long, long long, cudaStream_t);
template int sum_atoms<double>(const double *, double *, long long, long long, long long, cudaStream_t);

// ---------------------------------------------------------------------------------------------------------
// FP32-FMA pipe probe: 16 independent FFMA chains per thread, 2 CTAs of 512 threads per SM
// ---------------------------------------------------------------------------------------------------------
static constexpr int kProbeChains = 16;
static constexpr int kProbeUnroll = 16;
static constexpr int kProbeThreads = 512;
static constexpr int kProbeBlocks = 148 * 4;

__global__ void __launch_bounds__(kProbeThreads) fp32_peak_probe_kernel(float *sink, int iterations) {
    float acc[kProbeChains];
#pragma unroll
    for (int k = 0; k < kProbeChains; ++k) acc[k] = 1.0f + 1e-3f * (float)(threadIdx.x + k);
    const float a = 0.999f + 1e-7f * (float)blockIdx.x, b = 1e-4f;
    for (int it = 0; it < iterations; ++it) {
#pragma unroll
        for (int u = 0; u < kProbeUnroll; ++u) {
#pragma unroll
            for (int k = 0; k < kProbeChains; ++k) acc[k] = fmaf(acc[k], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kProbeChains; ++k) s += acc[k];
    if (s == 123.456f) sink[0] = s;   // practically never true: keeps the loop alive without memory traffic
}

int fp32_peak_probe(void *sink, int iterations, double *flops_out, cudaStream_t st) {
    fp32_peak_probe_kernel<<<kProbeBlocks, kProbeThreads, 0, st>>>((float *)sink, iterations);
    TNMF_CHECK_LAUNCH();
    if (flops_out)
        *flops_out = 2.0 * (double)kProbeBlocks * kProbeThreads * (double)iterations * kProbeUnroll * kProbeChains;
    return TNMF_OK;
}

}  // namespace tnmf
