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

// One warp per frame, FPW frames per warp one after the other; lane j < FPW keeps frame j's sums and the warp runs the scalar float64 tail
// (exp2, log, ~400 instructions) once for all of them — with FPW = 1 the tail's issue slots, not HBM, bound the kernel on long tracks.
// salience: [T][360] fp32, row stride in elements.
template <bool TAIL, int FPW>
__global__ void __launch_bounds__(256) rmvpe_decode_kernel(const float* __restrict__ sal, int64_t row_stride, int T, float thred, TailArgs a,
                                                           double* __restrict__ f0_out, int64_t* __restrict__ pitch, float* __restrict__ pitchf) {
    rvc::pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int t0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * FPW;
    if (t0 >= T) return;
    double my_num = 0.0;
    float my_den = 0.f, my_max = 0.f;
#pragma unroll 2
    for (int f = 0; f < FPW; ++f) {
        if (t0 + f >= T) break;                                                            // warp-uniform
        const float* row = sal + (int64_t)(t0 + f) * row_stride;
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
        // numpy's pairwise reduction for n = 9 across lanes 0..8 (IEEE addition commutes, so the xor tree gives every lane of 0..7 the
        // same ((t0+t1)+(t2+t3))+((t4+t5)+(t6+t7))), then the ninth term
        double num = __dadd_rn(prod, __shfl_xor_sync(0xffffffffu, prod, 1));
        float den = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
        num = __dadd_rn(num, __shfl_xor_sync(0xffffffffu, num, 2));
        den = __fadd_rn(den, __shfl_xor_sync(0xffffffffu, den, 2));
        num = __dadd_rn(num, __shfl_xor_sync(0xffffffffu, num, 4));
        den = __fadd_rn(den, __shfl_xor_sync(0xffffffffu, den, 4));
        num = __dadd_rn(num, __shfl_sync(0xffffffffu, prod, 8));
        den = __fadd_rn(den, __shfl_sync(0xffffffffu, s, 8));
        if (lane == f) { my_num = num; my_den = den; my_max = m; }                          // FPW <= 8: lanes 0..7 all hold the sums
    }
    const int t = t0 + lane;
    if (lane >= FPW || t >= T) return;
    double dev = __ddiv_rn(my_num, (double)my_den);                                        // :250
    if (fmaxf(my_max, 0.f) <= thred) dev = 0.0;                                            // :251 (the maximum of the padded row)
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

template <bool TAIL>
void launch_decode(const float* sal, int64_t row_stride, int T, float thred, const TailArgs& a, double* f0, int64_t* pitch, float* pitchf,
                   cudaStream_t st) {
    // short tracks: one frame per warp (latency); long tracks: eight, so that the scalar tail is issued once per eight frames (HBM-bound); in
    // between (a ten-minute file is 60 001 frames) four, which still gives a few waves of CTAs to overlap loads, reductions and tails
    if (T < 8192)
        rmvpe_decode_kernel<TAIL, 1><<<rvc::cdiv(T, 8), 256, 0, st>>>(sal, row_stride, T, thred, a, f0, pitch, pitchf);
    else if (T < 200000)
        rmvpe_decode_kernel<TAIL, 4><<<rvc::cdiv(T, 32), 256, 0, st>>>(sal, row_stride, T, thred, a, f0, pitch, pitchf);
    else
        rmvpe_decode_kernel<TAIL, 8><<<rvc::cdiv(T, 64), 256, 0, st>>>(sal, row_stride, T, thred, a, f0, pitch, pitchf);
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
    if (T < 0 || row_stride < 360) return rvc::fail(RVC_ESHAPE, "rvc_rmvpe_decode_f32: need T >= 0 and rows of 360 bins (row_stride >= 360)");
    if (T == 0) return 0;
    if (!salience || !f0) return rvc::fail(RVC_EINVAL, "rvc_rmvpe_decode_f32: null pointer");
    TailArgs a = {};
    a.gate_lo = gate_lo;
    a.gate_hi = gate_hi;
    launch_decode<false>(salience, row_stride, T, thred, a, f0, nullptr, nullptr, (cudaStream_t)stream);
    return rvc::check_launch("rvc_rmvpe_decode_f32");
}

RVC_API int rvc_f0_coarse_f64(const double* f0, int64_t n, const double* notes, int32_t n_notes, double strength, double shift, double mel_min,
                              double mel_max, int64_t* pitch, float* pitchf, double* f0_out, void* stream) {
    if (n < 0 || n > ((int64_t)1 << 38)) return rvc::fail(RVC_ESHAPE, "rvc_f0_coarse_f64: bad length");
    if (int rc = check_tail("rvc_f0_coarse_f64", notes, n_notes, shift, mel_min, mel_max)) return rc;
    if (n == 0) return 0;
    if (!f0 || !pitch || !pitchf) return rvc::fail(RVC_EINVAL, "rvc_f0_coarse_f64: null pointer");
    TailArgs a = {notes, n_notes, strength, shift, mel_min, mel_max, 0.0, 0.0};
    f0_coarse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(f0, n, a, f0_out, pitch, pitchf);
    return rvc::check_launch("rvc_f0_coarse_f64");
}

RVC_API int rvc_rmvpe_pitch_f32(const float* salience, int64_t row_stride, int32_t T, float thred, double gate_lo, double gate_hi,
                                const double* notes, int32_t n_notes, double strength, double shift, double mel_min, double mel_max,
                                int64_t* pitch, float* pitchf, double* f0_out, void* stream) {
    if (T < 0 || row_stride < 360) return rvc::fail(RVC_ESHAPE, "rvc_rmvpe_pitch_f32: need T >= 0 and rows of 360 bins (row_stride >= 360)");
    if (int rc = check_tail("rvc_rmvpe_pitch_f32", notes, n_notes, shift, mel_min, mel_max)) return rc;
    if (T == 0) return 0;
    if (!salience || !pitch || !pitchf) return rvc::fail(RVC_EINVAL, "rvc_rmvpe_pitch_f32: null pointer");
    TailArgs a = {notes, n_notes, strength, shift, mel_min, mel_max, gate_lo, gate_hi};
    launch_decode<true>(salience, row_stride, T, thred, a, f0_out, pitch, pitchf, (cudaStream_t)stream);
    return rvc::check_launch("rvc_rmvpe_pitch_f32");
}
