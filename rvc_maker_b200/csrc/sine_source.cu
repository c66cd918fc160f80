// SineGen + SourceModuleHnNSF (harmonic_num = 0): frame-level wrapped phase prefix sum, then one bandwidth-bound
// per-sample pass (4 B noise in, 4 B out per sample).  Reference: synthesizers.py:79-100, :111-112.
//
// The reference's CPU `cumsum` accumulates in double and rounds every prefix to fp32; the scan below does the same
// (fp64 warp-shuffle block scan, one CTA per batch row), so the phase fed to sin() matches the oracle to rounding.
// Arithmetic that feeds fmod/sin uses explicit round-to-nearest intrinsics so that nvcc does not contract mul+add.
#include "common.cuh"

namespace {

constexpr int SCAN_NT = 256;

__global__ void __launch_bounds__(SCAN_NT) sine_phase_scan_kernel(const float* __restrict__ f0, float* __restrict__ phase, int T,
                                                                  int upp, float sr) {
    rvc::pdl_prologue();
    __shared__ double warp_tot[SCAN_NT / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* fb = f0 + (int64_t)b * T;
    float* pb = phase + (int64_t)b * T;
    const int per = (T + SCAN_NT - 1) / SCAN_NT;
    const int beg = tid * per, end = min(beg + per, T);
    // term_t = fmod(f0_t/sr*upp + 0.5, 1) - 0.5   (rad[:, :-1, -1:], synthesizers.py:84)
    double local = 0.0;
    for (int t = beg; t < end; ++t) {
        const float r = __fmul_rn(__fdiv_rn(fb[t], sr), (float)upp);
        local += (double)__fsub_rn(fmodf(__fadd_rn(r, 0.5f), 1.0f), 0.5f);
    }
    // exclusive block scan of `local`
    double incl = local;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    double base = 0.0;
    for (int w = 0; w < wid; ++w) base += warp_tot[w];
    double run = base + incl - local;  // sum of terms of all frames before `beg`
    for (int t = beg; t < end; ++t) {
        // phase_t = fmod((float) sum_{t' < t} term_t', 1); frame 0 gets 0 (F.pad, :84)
        pb[t] = fmodf((float)run, 1.0f);
        const float r = __fmul_rn(__fdiv_rn(fb[t], sr), (float)upp);
        run += (double)__fsub_rn(fmodf(__fadd_rn(r, 0.5f), 1.0f), 0.5f);
    }
}

__global__ void __launch_bounds__(256) sine_synth_kernel(const float* __restrict__ f0, const float* __restrict__ phase,
                                                         const float* __restrict__ noise, const float* __restrict__ lin_w,
                                                         const float* __restrict__ lin_b, float* __restrict__ har, int T, int upp,
                                                         float sr, int64_t n) {
    rvc::pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t frame = i / upp;              // b*T + t
    const int nn = (int)(i - frame * upp) + 1;  // 1..upp
    const float f = f0[frame];
    const float rad = __fadd_rn(__fmul_rn(__fdiv_rn(f, sr), (float)nn), phase[frame]);
    const float s = __fmul_rn(sinf(__fmul_rn(rad, 6.2831855f)), 0.1f);
    const float uv = f > 0.f ? 1.f : 0.f;
    const float amp = __fadd_rn(__fmul_rn(uv, 0.003f), __fdiv_rn(__fmul_rn(1.f - uv, 0.1f), 3.f));
    const float v = __fadd_rn(__fmul_rn(s, uv), __fmul_rn(amp, noise[i]));
    har[i] = tanhf(__fadd_rn(__fmul_rn(v, lin_w[0]), lin_b[0]));
}

}  // namespace

RVC_API int rvc_sine_source_f32(const float* f0, const float* noise, const float* lin_w, const float* lin_b, float* scratch,
                                   float* har, int32_t B, int32_t T, int32_t upp, float sr, void* stream) {
    if (!f0 || !noise || !lin_w || !lin_b || !scratch || !har) return rvc::fail(RVC_EINVAL, "rvc_sine_source_f32: null pointer");
    if (B <= 0 || T < 0 || upp <= 0 || !(sr > 0.f)) return rvc::fail(RVC_ESHAPE, "rvc_sine_source_f32: bad shape");
    if (T == 0) return 0;
    rvc::launch_pdl(sine_phase_scan_kernel, dim3(B), dim3(SCAN_NT), (size_t)(0), (cudaStream_t)stream, f0, scratch, T, upp, sr);
    int rc = rvc::check_launch("rvc_sine_source_f32(scan)");
    if (rc) return rc;
    const int64_t n = (int64_t)B * T * upp;
    rvc::launch_pdl(sine_synth_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, f0, scratch, noise, lin_w, lin_b, har, T, upp, sr, n);
    return rvc::check_launch("rvc_sine_source_f32(synth)");
}
