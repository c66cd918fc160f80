// The split-point search of VC.pipeline (main/inference/convert.py:404-412), the first statement block of the chunk loop around Synthesizer.infer:
//   audio_pad = np.pad(audio, (window/2, window/2), mode="reflect");  audio_sum = zeros_like(audio)
//   for i in range(window): audio_sum += audio_pad[i : i - window]
//   for t in range(t_center, len(audio), t_center):
//       seg = abs(audio_sum[t - t_query : t + t_query]);  opt_ts.append(t - t_query + where(seg == seg.min())[0][0])
// On the host this is `window` (160) passes over the whole signal: 1.5 s for a 10-minute file, ten times the synthesis of that file on this GPU.
// Here only the 2 t_query positions around every candidate are evaluated, each as the SAME sequential sum (i = 0 .. window-1, starting from 0, one
// rounding per addition, in the signal's own dtype — float64 behind the reference's filtfilt, float32 for a raw float32 signal), so the sums and
// therefore the chosen indices are bit-identical to numpy's; ties resolve to the first index like np.where(...)[0][0].
#include "common.cuh"

namespace {

constexpr int PB = 2048;       // positions per block
constexpr int MAXW = 512;      // window <= MAXW

__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

struct Best {
    double v;      // |audio_sum| (a float32 sum widens exactly)
    int64_t i;     // absolute sample index
};
__device__ __forceinline__ bool better(const Best& a, const Best& b) { return a.v < b.v || (a.v == b.v && a.i < b.i); }

__device__ __forceinline__ Best block_min(Best b, Best* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best x;
        x.v = __shfl_xor_sync(0xffffffffu, b.v, o);
        x.i = __shfl_xor_sync(0xffffffffu, b.i, o);
        if (better(x, b)) b = x;
    }
    if (lane == 0) sh[warp] = b;
    __syncthreads();
    if (warp == 0) {
        b = lane < (int)(blockDim.x >> 5) ? sh[lane] : Best{INFINITY, INT64_MAX};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Best x;
            x.v = __shfl_xor_sync(0xffffffffu, b.v, o);
            x.i = __shfl_xor_sync(0xffffffffu, b.i, o);
            if (better(x, b)) b = x;
        }
    }
    return b;   // valid in thread 0
}

// grid (blocks per candidate, candidates); candidate c: t = (c + 1) t_center, positions [t - t_query, min(t + t_query, n))
template <typename T>
__global__ void __launch_bounds__(256) split_partial_kernel(const T* __restrict__ audio, int64_t n, int window, int64_t t_center, int64_t t_query,
                                                            Best* __restrict__ ws) {
    __shared__ T tile[PB + MAXW];
    __shared__ Best red[8];
    const int64_t t = (int64_t)(blockIdx.y + 1) * t_center;
    const int64_t lo = t - t_query, hi = min(t + t_query, n);
    const int64_t p0 = lo + (int64_t)blockIdx.x * PB;
    const int count = (int)max((int64_t)0, min((int64_t)PB, hi - p0));
    const int half = window / 2;
    // audio_pad[p0 .. p0 + count + window - 1), audio_pad[j] = audio[reflect(j - window/2)]
    for (int j = threadIdx.x; j < count + window - 1; j += 256) {
        int64_t s = p0 + j - half;
        if (s < 0) s = -s;
        if (s >= n) s = 2 * (n - 1) - s;
        tile[j] = audio[s];
    }
    __syncthreads();
    Best best{INFINITY, INT64_MAX};
    for (int q = threadIdx.x; q < count; q += 256) {      // ascending positions per thread: a strict `<` keeps the first minimum
        T acc = (T)0;
        for (int i = 0; i < window; ++i) acc = add_rn(acc, tile[q + i]);
        const double a = fabs((double)acc);
        if (a < best.v) { best.v = a; best.i = p0 + q; }
    }
    best = block_min(best, red);
    if (threadIdx.x == 0) ws[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = best;
}

__global__ void __launch_bounds__(256) split_final_kernel(const Best* __restrict__ ws, int nblk, int64_t* __restrict__ opt_ts) {
    __shared__ Best red[8];
    Best best{INFINITY, INT64_MAX};
    for (int b = threadIdx.x; b < nblk; b += 256) {
        const Best x = ws[(size_t)blockIdx.x * nblk + b];
        if (better(x, best)) best = x;
    }
    best = block_min(best, red);
    if (threadIdx.x == 0) opt_ts[blockIdx.x] = best.i;
}

template <typename T>
int run(const T* audio, int64_t n, int window, int64_t t_center, int64_t t_query, int n_t, int nblk, int64_t* opt_ts, void* ws, cudaStream_t st) {
    split_partial_kernel<T><<<dim3(nblk, n_t), 256, 0, st>>>(audio, n, window, t_center, t_query, (Best*)ws);
    if (int rc = rvc::check_launch("rvc_split_points")) return rc;
    split_final_kernel<<<n_t, 256, 0, st>>>((const Best*)ws, nblk, opt_ts);
    return rvc::check_launch("rvc_split_points");
}

}  // namespace

RVC_API int64_t rvc_split_points_workspace(int64_t n, int64_t t_center, int64_t t_query) {
    if (n <= 0 || t_center <= 0 || t_query <= 0) return 0;
    const int64_t n_t = (n - 1) / t_center;                      // len(range(t_center, n, t_center))
    return n_t * ((2 * t_query + PB - 1) / PB) * (int64_t)sizeof(Best);
}

RVC_API int rvc_split_points(const void* audio, int32_t is_f64, int64_t n, int32_t window, int64_t t_center, int64_t t_query, int64_t* opt_ts, void* ws,
                             void* stream) {
    if (window < 2 || window > MAXW || (window & 1)) return rvc::fail(RVC_ESHAPE, "rvc_split_points: window must be even, 2..%d", MAXW);
    if (n <= window / 2) return rvc::fail(RVC_ESHAPE, "rvc_split_points: the reflect padding needs more than window/2 samples");
    if (t_query <= 0 || t_center < t_query) return rvc::fail(RVC_ESHAPE, "rvc_split_points: need 0 < t_query <= t_center");
    const int64_t n_t = (n - 1) / t_center;
    if (n_t == 0) return 0;
    if (n_t > 65535) return rvc::fail(RVC_ESHAPE, "rvc_split_points: too many split candidates");
    if (!audio || !opt_ts || !ws) return rvc::fail(RVC_EINVAL, "rvc_split_points: null pointer");
    const int nblk = (int)((2 * t_query + PB - 1) / PB);
    if (is_f64) return run((const double*)audio, n, window, t_center, t_query, (int)n_t, nblk, opt_ts, ws, (cudaStream_t)stream);
    return run((const float*)audio, n, window, t_center, t_query, (int)n_t, nblk, opt_ts, ws, (cudaStream_t)stream);
}
