// The zero-phase high-pass at the head of VC.pipeline — `audio = signal.filtfilt(bh, ah, audio)` (main/inference/convert.py:402, coefficients :29) — on the
// device, in float64 like scipy: odd extension by padlen = 3 max(len a, len b) samples, forward pass with the steady-state initial state zi * ext[0],
// backward pass with zi * y[-1], the padding dropped.  scipy's lfilter is the direct-form-II-transposed recurrence
//     y = b0 x + z0;   z_k = (b_{k+1} x + z_{k+1}) - a_{k+1} y
// one sample after the other (137 - 259 ms per 10 minutes of audio on the host).  Here the signal is cut into blocks of 4096 samples, one thread each:
// (1) every block runs the recurrence from a zero state and keeps its end state, (2) one thread chains the blocks' true start states,
// s_{c+1} = M^block s_c + e_c (M^block comes from the host, obtained by running the recurrence itself), (3) every block runs again from its true state and
// writes its outputs.  Inside a block the operations and their order are scipy's (separate roundings, no fma); what differs is the rounding of the
// state hand-off, which the filter's companion form amplifies to ~2e-8 absolute on a signal of amplitude 0.5 — the size of the rounding noise scipy's own
// evaluation carries (tests/test_gpu_filtfilt.py).
#include "common.cuh"

namespace {

constexpr int MAXS = 8;        // filter order <= 8
constexpr int LCMAX = 4096;    // samples per block (rvc_filtfilt_block)
constexpr int PF = 16;         // inputs requested ahead of the recurrence

struct Coef {
    double b[MAXS + 1], a[MAXS + 1], zi[MAXS], ML[MAXS * MAXS];
    int ns;                    // number of states = order
};

// the filter's input at extended index i (0 <= i < N = n + 2 pl): scipy.signal._arraytools.odd_ext, evaluated in the signal's dtype
template <typename T>
__device__ __forceinline__ double ext_at(const T* __restrict__ x, int64_t n, int pl, int64_t i);
template <>
__device__ __forceinline__ double ext_at<float>(const float* __restrict__ x, int64_t n, int pl, int64_t i) {
    if (i < pl) return (double)__fsub_rn(__fmul_rn(2.f, x[0]), x[pl - i]);
    if (i >= pl + n) return (double)__fsub_rn(__fmul_rn(2.f, x[n - 1]), x[n - 2 - (i - pl - n)]);
    return (double)x[i - pl];
}
template <>
__device__ __forceinline__ double ext_at<double>(const double* __restrict__ x, int64_t n, int pl, int64_t i) {
    if (i < pl) return __dsub_rn(__dmul_rn(2.0, x[0]), x[pl - i]);
    if (i >= pl + n) return __dsub_rn(__dmul_rn(2.0, x[n - 1]), x[n - 2 - (i - pl - n)]);
    return x[i - pl];
}

// BACK = false: input = the extended signal; BACK = true: input = the forward pass's output read backwards
template <typename T, bool BACK>
__device__ __forceinline__ double input_at(const T* __restrict__ x, const double* __restrict__ y1, int64_t n, int pl, int64_t N, int64_t i) {
    return BACK ? y1[N - 1 - i] : ext_at<T>(x, n, pl, i);
}

// NS > 0: the order is a compile-time constant (NS = 5 is the reference's filter: no predicated-off operations in the dependent chain); NS = 0: c.ns
template <int NS>
__device__ __forceinline__ double step(const Coef& c, double (&z)[MAXS], double x) {
    const double y = __dadd_rn(__dmul_rn(c.b[0], x), z[0]);
#pragma unroll
    for (int k = 0; k < (NS ? NS : MAXS); ++k) {
        if (NS || k < c.ns) {
            const double zn = k + 1 < (NS ? NS : MAXS) ? z[k + 1] : 0.0;      // (the state behind the last one is zero)
            z[k] = __dsub_rn(__dadd_rn(__dmul_rn(c.b[k + 1], x), zn), __dmul_rn(c.a[k + 1], y));
        }
    }
    return y;
}

// (1) end state of every block but the last from a zero start
template <typename T, bool BACK, int NS>
__global__ void __launch_bounds__(32) iir_block_end_kernel(const T* __restrict__ x, const double* __restrict__ y1, int64_t n, int pl, int64_t N, Coef c,
                                                           int nblk, int LC, double* __restrict__ ends) {
    const int blk = blockIdx.x * 32 + threadIdx.x;
    if (blk >= nblk - 1) return;
    double z[MAXS];
#pragma unroll
    for (int k = 0; k < MAXS; ++k) z[k] = 0.0;
    const int64_t i0 = (int64_t)blk * LC;
    // the recurrence waits for nothing but its own state: the next PF inputs are requested while the current ones are consumed
    double cur[PF], nxt[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) cur[j] = input_at<T, BACK>(x, y1, n, pl, N, i0 + j);
    for (int i = 0; i < LC; i += PF) {
        if (i + PF < LC) {
#pragma unroll
            for (int j = 0; j < PF; ++j) nxt[j] = input_at<T, BACK>(x, y1, n, pl, N, i0 + i + PF + j);
        }
#pragma unroll
        for (int j = 0; j < PF; ++j) step<NS>(c, z, cur[j]);
#pragma unroll
        for (int j = 0; j < PF; ++j) cur[j] = nxt[j];
    }
#pragma unroll
    for (int k = 0; k < MAXS; ++k) ends[(size_t)blk * MAXS + k] = z[k];
}

// (2) true start states: s_0 = zi * input[0], s_{c+1} = M^LC s_c + e_c.  One thread walks the chain; the block stages the end states through
// shared memory in tiles (a global load per link would cost more than the link itself).
constexpr int CT = 512;        // blocks per tile
template <typename T, bool BACK, int NS>
__global__ void __launch_bounds__(256) iir_chain_kernel(const T* __restrict__ x, const double* __restrict__ y1, int64_t n, int pl, int64_t N, Coef c,
                                                        int nblk, const double* __restrict__ ends, double* __restrict__ starts) {
    __shared__ double tile[CT * MAXS];
    double s[MAXS];
    if (threadIdx.x == 0) {
        const double x0 = input_at<T, BACK>(x, y1, n, pl, N, 0);
#pragma unroll
        for (int k = 0; k < MAXS; ++k) s[k] = k < c.ns ? __dmul_rn(c.zi[k], x0) : 0.0;
    }
    for (int b0 = 0; b0 < nblk; b0 += CT) {
        const int cnt = min(CT, nblk - b0);
        for (int idx = threadIdx.x; idx < cnt * MAXS; idx += 256)
            tile[idx] = (b0 + idx / MAXS) < nblk - 1 ? ends[(size_t)b0 * MAXS + idx] : 0.0;      // (the last block's end state is never needed)
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int b = 0; b < cnt; ++b) {
                double t[MAXS];
#pragma unroll
                for (int k = 0; k < MAXS; ++k) {
                    if (NS && k >= NS) {              // (states beyond the order stay zero)
                        t[k] = 0.0;
                        tile[b * MAXS + k] = 0.0;
                        continue;
                    }
                    double acc = tile[b * MAXS + k];
#pragma unroll
                    for (int j = 0; j < (NS ? NS : MAXS); ++j)
                        if (NS || j < c.ns) acc += c.ML[k * MAXS + j] * s[j];
                    t[k] = (NS || k < c.ns) ? acc : 0.0;
                    tile[b * MAXS + k] = s[k];        // the start state of block b0 + b takes the slot of its end state
                }
#pragma unroll
                for (int k = 0; k < MAXS; ++k) s[k] = t[k];
            }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < cnt * MAXS; idx += 256) starts[(size_t)b0 * MAXS + idx] = tile[idx];
        __syncthreads();
    }
}

// (3) the outputs.  Forward: y1[i]; backward: the final signal, out[N - 1 - i - pl] for the samples that are not padding
template <typename T, bool BACK, int NS>
__global__ void __launch_bounds__(32) iir_block_apply_kernel(const T* __restrict__ x, const double* __restrict__ y1, int64_t n, int pl, int64_t N, Coef c,
                                                             int nblk, int LC, const double* __restrict__ starts, double* __restrict__ out) {
    const int blk = blockIdx.x * 32 + threadIdx.x;
    if (blk >= nblk) return;
    double z[MAXS];
#pragma unroll
    for (int k = 0; k < MAXS; ++k) z[k] = starts[(size_t)blk * MAXS + k];
    const int64_t i0 = (int64_t)blk * LC;
    const int cnt = (int)min((int64_t)LC, N - i0);
    double cur[PF], nxt[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) cur[j] = j < cnt ? input_at<T, BACK>(x, y1, n, pl, N, i0 + j) : 0.0;
    for (int i = 0; i < cnt; i += PF) {
#pragma unroll
        for (int j = 0; j < PF; ++j) nxt[j] = i + PF + j < cnt ? input_at<T, BACK>(x, y1, n, pl, N, i0 + i + PF + j) : 0.0;
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            if (i + j < cnt) {
                const double y = step<NS>(c, z, cur[j]);
                if (BACK) {
                    const int64_t q = N - 1 - (i0 + i + j) - pl;
                    if (q >= 0 && q < n) out[q] = y;
                } else {
                    out[i0 + i + j] = y;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < PF; ++j) cur[j] = nxt[j];
    }
}

template <typename T, int NS>
int run(const T* x, int64_t n, int pl, int LC, const Coef& c, double* y, double* ws, cudaStream_t st) {
    const int64_t N = n + 2 * pl;
    const int nblk = (int)((N + LC - 1) / LC);
    double* y1 = ws;
    double* ends = y1 + N;
    double* starts = ends + (size_t)nblk * MAXS;
    const int grid = (nblk + 31) / 32;
    iir_block_end_kernel<T, false, NS><<<grid, 32, 0, st>>>(x, y1, n, pl, N, c, nblk, LC, ends);
    iir_chain_kernel<T, false, NS><<<1, 256, 0, st>>>(x, y1, n, pl, N, c, nblk, ends, starts);
    iir_block_apply_kernel<T, false, NS><<<grid, 32, 0, st>>>(x, y1, n, pl, N, c, nblk, LC, starts, y1);
    iir_block_end_kernel<T, true, NS><<<grid, 32, 0, st>>>(x, y1, n, pl, N, c, nblk, LC, ends);
    iir_chain_kernel<T, true, NS><<<1, 256, 0, st>>>(x, y1, n, pl, N, c, nblk, ends, starts);
    iir_block_apply_kernel<T, true, NS><<<grid, 32, 0, st>>>(x, y1, n, pl, N, c, nblk, LC, starts, y);
    return rvc::check_launch("rvc_filtfilt");
}

}  // namespace

static int pick_block(int64_t) {
    // 4 passes of `block` dependent steps against 2 chains of n / block links favour block ~ sqrt(1.3 n), but the hand-off s_{c+1} = M^block s_c + e_c is
    // only harmless once M^block has decayed (the companion form is far from normal): measured max |y - scipy| at amplitude 0.5 — blocks of 256 - 1024:
    // up to 1e-3 (split points moved); 2048 (|M^2048| ~ 2e2): 1.6e-7; 4096 (|M^4096| ~ 2e-3): 4e-8, the filter's own noise floor
    return LCMAX;
}

RVC_API int32_t rvc_filtfilt_block(int64_t n) { return n > 0 ? pick_block(n) : 0; }

RVC_API int64_t rvc_filtfilt_workspace(int64_t n, int32_t ncoef) {
    if (n <= 0 || ncoef < 2 || ncoef > MAXS + 1) return 0;
    const int64_t N = n + 2 * 3 * (int64_t)ncoef;
    const int64_t nblk = (N + pick_block(n) - 1) / pick_block(n);
    return (N + 2 * nblk * MAXS) * (int64_t)sizeof(double);
}

RVC_API int rvc_filtfilt(const void* x, int32_t is_f64, int64_t n, const double* b, const double* a, int32_t ncoef, const double* zi, const double* ML,
                         double* y, void* ws, void* stream) {
    if (ncoef < 2 || ncoef > MAXS + 1) return rvc::fail(RVC_ESHAPE, "rvc_filtfilt: 2 <= len(b) = len(a) <= %d", MAXS + 1);
    if (!b || !a || !zi || !ML) return rvc::fail(RVC_EINVAL, "rvc_filtfilt: null coefficient pointer");
    if (a[0] != 1.0) return rvc::fail(RVC_EINVAL, "rvc_filtfilt: normalise the coefficients (a[0] == 1) first, as scipy does");
    const int pl = 3 * ncoef;                      // scipy's default padlen
    if (n <= pl) return rvc::fail(RVC_ESHAPE, "rvc_filtfilt: the signal must be longer than padlen = %d (scipy raises too)", pl);
    if (!x || !y || !ws) return rvc::fail(RVC_EINVAL, "rvc_filtfilt: null pointer");
    Coef c = {};
    c.ns = ncoef - 1;
    for (int k = 0; k < ncoef; ++k) { c.b[k] = b[k]; c.a[k] = a[k]; }
    for (int k = 0; k < c.ns; ++k) c.zi[k] = zi[k];
    for (int k = 0; k < c.ns; ++k)
        for (int j = 0; j < c.ns; ++j) c.ML[k * MAXS + j] = ML[k * c.ns + j];
    const int lc = pick_block(n);
    cudaStream_t st = (cudaStream_t)stream;
    if (c.ns == 5)                                 // the reference's Butterworth (convert.py:29)
        return is_f64 ? run<double, 5>((const double*)x, n, pl, lc, c, y, (double*)ws, st) : run<float, 5>((const float*)x, n, pl, lc, c, y, (double*)ws, st);
    return is_f64 ? run<double, 0>((const double*)x, n, pl, lc, c, y, (double*)ws, st) : run<float, 0>((const float*)x, n, pl, lc, c, y, (double*)ws, st);
}
