// The f0 tail between the RMVPE network and Synthesizer.infer (post-network part), on the device so that the salience map
// never leaves HBM: RMVPE.decode / to_local_average_cents (main/library/predictors/RMVPE.py:217-221, :237-252) and the tail of VC.get_f0
// (main/inference/convert.py:308-323: autotune, pitch shift, coarse mel quantiser) with the two casts of :437-439.
//
// The reference does this arithmetic in numpy float64 (float32 for the salience sums).  Every product, sum and quotient below is a single
// IEEE operation in the reference's order (no fma contraction; the 9-term sums in numpy's pairwise order), so `cents` is bit-identical;
// exp2 / log are the CUDA double routines (<= 1 ulp, like the host's libm), which the float32 cast of pitchf and the rint of the coarse
// pitch absorb.
#include "common.cuh"

namespace {

struct TailArgs {
    const double* notes;   // autotune note table (convert.py:200) or nullptr
    int n_notes;
    double strength;       // f0_autotune_strength
    double shift;          // pow(2, pitch / 12), computed by the caller in double like :311
    double mel_min, mel_max;   // 1127 * ln(1 + f0_min / 700), ... (:198-199)
    double gate_lo, gate_hi;   // "rmvpe-legacy": f0 outside [gate_lo, gate_hi] -> 0 (RMVPE.py:231); off when gate_hi <= 0
};

// numpy's pairwise reduction for n = 9: eight accumulators combined as a tree, then the ninth term.
__device__ __forceinline__ double sum9(const double* t) {
    const double a = __dadd_rn(__dadd_rn(t[0], t[1]), __dadd_rn(t[2], t[3]));
    const double b = __dadd_rn(__dadd_rn(t[4], t[5]), __dadd_rn(t[6], t[7]));
    return __dadd_rn(__dadd_rn(a, b), t[8]);
}
__device__ __forceinline__ float sum9(const float* t) {
    const float a = __fadd_rn(__fadd_rn(t[0], t[1]), __fadd_rn(t[2], t[3]));
    const float b = __fadd_rn(__fadd_rn(t[4], t[5]), __fadd_rn(t[6], t[7]));
    return __fadd_rn(__fadd_rn(a, b), t[8]);
}

// convert.py:308-323 for one frame; returns the shifted f0, *coarse = the 1..255 index.
__device__ __forceinline__ double f0_tail(double f0, const TailArgs& a, int* coarse) {
    if (a.notes) {   // Autotune.autotune_f0 (:176-177): nearest note, the first one on ties; applies to unvoiced frames too
        double best = a.notes[0], bd = fabs(__dsub_rn(a.notes[0], f0));
        for (int i = 1; i < a.n_notes; ++i) {
            const double d = fabs(__dsub_rn(a.notes[i], f0));
            if (d < bd) { bd = d; best = a.notes[i]; }
        }
        f0 = __dadd_rn(f0, __dmul_rn(__dsub_rn(best, f0), a.strength));
    }
    f0 = __dmul_rn(f0, a.shift);                                                            // :311
    double mel = __dmul_rn(1127.0, log(__dadd_rn(1.0, __ddiv_rn(f0, 700.0))));              // :318
    if (mel > 0.0) mel = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(mel, a.mel_min), 254.0), __dsub_rn(a.mel_max, a.mel_min)), 1.0);   // :319
    if (mel <= 1.0) mel = 1.0;                                                              // :320
    if (mel > 255.0) mel = 255.0;                                                           // :321
    *coarse = (int)rint(mel);                                                               // :323 (half to even)
    return f0;
}

// One warp per frame.  salience: [T][360] fp32, row stride in elements.
template <bool TAIL>
__global__ void __launch_bounds__(256) rmvpe_decode_kernel(const float* __restrict__ sal, int64_t row_stride, int T, float thred, TailArgs a,
                                                           double* __restrict__ f0_out, int64_t* __restrict__ pitch, float* __restrict__ pitchf) {
    rvc::pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= T) return;
    const float* row = sal + (int64_t)t * row_stride;
    // np.argmax (:238): the first maximum
    float m = row[lane];
    int mi = lane;
#pragma unroll
    for (int i = 1; i < 12; ++i) {
        const int c = lane + 32 * i;
        if (c < 360) {
            const float v = row[c];
            if (v > m) { m = v; mi = c; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
    }
    // the 9 bins around it; outside [0, 360) both the salience and the cents mapping are the zero padding (:239, RMVPE.py:208)
    const int c = mi - 4 + lane;
    const bool in = lane < 9 && c >= 0 && c < 360;
    const float s = in ? row[c] : 0.f;
    const double cents = in ? __dadd_rn((double)(20 * c), 1997.3794084376191) : 0.0;
    const double prod = __dmul_rn((double)s, cents);
    double tp[9];
    float ts[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        tp[j] = __shfl_sync(0xffffffffu, prod, j);
        ts[j] = __shfl_sync(0xffffffffu, s, j);
    }
    if (lane != 0) return;
    double dev = __ddiv_rn(sum9(tp), (double)sum9(ts));                                    // :250
    if (fmaxf(m, 0.f) <= thred) dev = 0.0;                                                 // :251 (the maximum of the padded row)
    double f0 = __dmul_rn(10.0, exp2(__ddiv_rn(dev, 1200.0)));                             // :218
    if (f0 == 10.0) f0 = 0.0;                                                              // :219
    if (a.gate_hi > 0.0 && (f0 < a.gate_lo || f0 > a.gate_hi)) f0 = 0.0;                   // :231
    if (TAIL) {
        int coarse;
        f0 = f0_tail(f0, a, &coarse);
        pitch[t] = coarse;
        pitchf[t] = __double2float_rn(f0);
    }
    if (f0_out) f0_out[t] = f0;
}

__global__ void __launch_bounds__(256) f0_coarse_kernel(const double* __restrict__ f0_in, int64_t n, TailArgs a, double* __restrict__ f0_out,
                                                        int64_t* __restrict__ pitch, float* __restrict__ pitchf) {
    rvc::pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int coarse;
    const double f0 = f0_tail(f0_in[i], a, &coarse);
    pitch[i] = coarse;
    pitchf[i] = __double2float_rn(f0);
    if (f0_out) f0_out[i] = f0;
}

int check_tail(const char* who, const double* notes, int n_notes, double shift, double mel_min, double mel_max) {
    if (notes && n_notes <= 0) return rvc::fail(RVC_EINVAL, "%s: autotune needs n_notes > 0", who);
    if (!(shift > 0.0)) return rvc::fail(RVC_EINVAL, "%s: shift must be pow(2, semitones / 12) > 0", who);
    if (!(mel_max > mel_min)) return rvc::fail(RVC_EINVAL, "%s: need mel_max > mel_min", who);
    return 0;
}

}  // namespace

RVC_API int rvc_rmvpe_decode_f32(const float* salience, int64_t row_stride, int32_t T, float thred, double gate_lo, double gate_hi, double* f0,
                                 void* stream) {
    if (!salience || !f0) return rvc::fail(RVC_EINVAL, "rvc_rmvpe_decode_f32: null pointer");
    if (T < 0 || row_stride < 360) return rvc::fail(RVC_ESHAPE, "rvc_rmvpe_decode_f32: need T >= 0 and rows of 360 bins (row_stride >= 360)");
    if (T == 0) return 0;
    TailArgs a = {};
    a.gate_lo = gate_lo;
    a.gate_hi = gate_hi;
    rmvpe_decode_kernel<false><<<rvc::cdiv(T, 8), 256, 0, (cudaStream_t)stream>>>(salience, row_stride, T, thred, a, f0, nullptr, nullptr);
    return rvc::check_launch("rvc_rmvpe_decode_f32");
}

RVC_API int rvc_f0_coarse_f64(const double* f0, int64_t n, const double* notes, int32_t n_notes, double strength, double shift, double mel_min,
                              double mel_max, int64_t* pitch, float* pitchf, double* f0_out, void* stream) {
    if (!f0 || !pitch || !pitchf) return rvc::fail(RVC_EINVAL, "rvc_f0_coarse_f64: null pointer");
    if (n < 0 || n > ((int64_t)1 << 38)) return rvc::fail(RVC_ESHAPE, "rvc_f0_coarse_f64: bad length");
    if (int rc = check_tail("rvc_f0_coarse_f64", notes, n_notes, shift, mel_min, mel_max)) return rc;
    if (n == 0) return 0;
    TailArgs a = {notes, n_notes, strength, shift, mel_min, mel_max, 0.0, 0.0};
    f0_coarse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(f0, n, a, f0_out, pitch, pitchf);
    return rvc::check_launch("rvc_f0_coarse_f64");
}

RVC_API int rvc_rmvpe_pitch_f32(const float* salience, int64_t row_stride, int32_t T, float thred, double gate_lo, double gate_hi,
                                const double* notes, int32_t n_notes, double strength, double shift, double mel_min, double mel_max,
                                int64_t* pitch, float* pitchf, double* f0_out, void* stream) {
    if (!salience || !pitch || !pitchf) return rvc::fail(RVC_EINVAL, "rvc_rmvpe_pitch_f32: null pointer");
    if (T < 0 || row_stride < 360) return rvc::fail(RVC_ESHAPE, "rvc_rmvpe_pitch_f32: need T >= 0 and rows of 360 bins (row_stride >= 360)");
    if (int rc = check_tail("rvc_rmvpe_pitch_f32", notes, n_notes, shift, mel_min, mel_max)) return rc;
    if (T == 0) return 0;
    TailArgs a = {notes, n_notes, strength, shift, mel_min, mel_max, gate_lo, gate_hi};
    rmvpe_decode_kernel<true><<<rvc::cdiv(T, 8), 256, 0, (cudaStream_t)stream>>>(salience, row_stride, T, thred, a, f0_out, pitch, pitchf);
    return rvc::check_launch("rvc_rmvpe_pitch_f32");
}
