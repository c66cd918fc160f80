// The RMVPE mel front end (pre-network part) as ONE kernel: MelSpectrogram.forward for keyshift = 0, speed = 1, center = True
// (main/library/predictors/RMVPE.py:161-181: torch.stft with a periodic Hann window of 1024, hop 160, reflect padding; magnitude; mel_basis matmul;
// log(clamp(., 1e-5))) plus the reflect padding of the frame axis to a multiple of 32 that mel2hidden applies before the network (:212-213).
// The reference runs this as ~8 library kernels (pad, cuFFT, abs, GEMM, clamp, log, pad) with the [513, T] spectrum in HBM in between; here a block
// takes eight frames: window + radix-2 FFT in shared memory (a real frame as one complex transform of half the length, per-stage twiddle tables), magnitudes parked in
// shared memory, then the filterbank product over each band's non-zero bins only (rvc_mel_basis_ranges; the dense product is the fallback).  Floating point: fp32 like the reference, different summation order than cuFFT / cuBLAS (tolerance in
// tests/test_gpu_mel_front.py).
#include "common.cuh"

namespace {

constexpr int NFFT = 1024, NBIN = NFFT / 2 + 1, NMEL = 128, HOP = 160, FPB = 8, MAGLD = 516;

// [lo, hi) of the non-zero bins of every filterbank row (triangular filters: 2 - 45 of the 513 bins); an all-zero row gives lo = hi = 0
__global__ void __launch_bounds__(32) mel_ranges_kernel(const float* __restrict__ basis, int* __restrict__ ranges) {
    const int m = blockIdx.x, lane = threadIdx.x;
    int lo = NBIN, hi = 0;
    for (int k = lane; k < NBIN; k += 32)
        if (basis[(size_t)m * NBIN + k] != 0.f) { lo = min(lo, k); hi = max(hi, k + 1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { ranges[2 * m] = hi ? lo : 0; ranges[2 * m + 1] = hi; }
}

__global__ void __launch_bounds__(256) mel_front_kernel(const float* __restrict__ audio, int64_t n, const float* __restrict__ window,
                                                        const float* __restrict__ basis, const int* __restrict__ ranges, float* __restrict__ out,
                                                        int64_t ld, int T, int Tp, float clamp) {
    rvc::pdl_prologue();
    __shared__ float2 buf[NFFT / 2];
    __shared__ float2 tw[NFFT];                        // per stage, contiguous: tw[half + pos] = exp(-i pi pos / half), half = 1 .. 512 (conflict-free reads)
    __shared__ float mag[FPB][MAGLD];
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * FPB;
    const int nf = min(FPB, Tp - c0);
    for (int k = tid; k < NFFT; k += 256) {
        if (k == 0) { tw[0] = make_float2(1.f, 0.f); continue; }
        const int half = 1 << (31 - __clz(k));
        float s, c;
        sincospif(-(float)(k - half) / (float)half, &s, &c);
        tw[k] = make_float2(c, s);
    }
    // a real frame of 1024 as ONE complex transform of 512: z[m] = x[2m] + i x[2m+1];  E[k] = (Z[k] + conj Z[512-k]) / 2 (even samples),
    // O[k] = -i (Z[k] - conj Z[512-k]) / 2 (odd samples),  X[k] = E[k] + exp(-2 pi i k / 1024) O[k].  (A silent frame stays exactly zero: no other
    // frame shares its transform.)
    for (int f = 0; f < FPB; ++f) {
        if (f >= nf) {                                 // (block-uniform) unused frame slots: zeros, never written out
            for (int k = tid; k < NBIN; k += 256) mag[f][k] = 0.f;
            continue;
        }
        const int c = c0 + f;
        const int t = c < T ? c : 2 * (T - 1) - c;     // F.pad(mel, (0, Tp - T), mode="reflect"): column T + i = column T - 2 - i
        const int64_t base = (int64_t)t * HOP - NFFT / 2;
        for (int m = tid; m < NFFT / 2; m += 256) {
            int64_t i0 = base + 2 * m, i1 = i0 + 1;    // torch.stft(center=True, pad_mode="reflect")
            if (i0 < 0) i0 = -i0;
            if (i0 >= n) i0 = 2 * (n - 1) - i0;
            if (i1 < 0) i1 = -i1;
            if (i1 >= n) i1 = 2 * (n - 1) - i1;
            const float2 w = reinterpret_cast<const float2*>(window)[m];
            buf[__brev((unsigned)m) >> 23] = make_float2(audio[i0] * w.x, audio[i1] * w.y);
        }
        __syncthreads();
#pragma unroll 1
        for (int half = 1; half < NFFT / 2; half <<= 1) {   // decimation in time: 9 stages, one butterfly per thread
            const int pos = tid & (half - 1);
            const int i0 = ((tid - pos) << 1) + pos, i1 = i0 + half;
            const float2 w = tw[half + pos], a = buf[i0], b = buf[i1];
            const float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
            buf[i0] = make_float2(a.x + bw.x, a.y + bw.y);
            buf[i1] = make_float2(a.x - bw.x, a.y - bw.y);
            __syncthreads();
        }
        for (int k = tid; k < NBIN; k += 256) {        // magnitude (:168)
            const float2 z = buf[k & 511], zm = buf[(512 - k) & 511];
            const float er = 0.5f * (z.x + zm.x), ei = 0.5f * (z.y - zm.y);
            const float orr = 0.5f * (z.y + zm.y), oi = -0.5f * (z.x - zm.x);
            const float2 w = k < 512 ? tw[512 + k] : make_float2(-1.f, 0.f);
            const float xr = er + (orr * w.x - oi * w.y), xi = ei + (orr * w.y + oi * w.x);
            mag[f][k] = sqrtf(xr * xr + xi * xi);
        }
        __syncthreads();
    }
    __syncthreads();
    // mel_basis @ magnitude (:178), log(clamp) (:181)
    if (ranges) {
        // thread = (band, frame): only the band's non-zero bins, summed in order
        for (int o = tid; o < NMEL * FPB; o += 256) {
            const int f = o & (FPB - 1), m = o >> 3;
            const int lo = ranges[2 * m], hi = ranges[2 * m + 1];
            const float* row = basis + (size_t)m * NBIN;
            float acc = 0.f;
            for (int k = lo; k < hi; ++k) acc = fmaf(row[k], mag[f][k], acc);
            if (f < nf) out[(size_t)m * ld + c0 + f] = logf(fmaxf(acc, clamp));
        }
        return;
    }
    // dense: warp w owns bands 16 w .. 16 w + 15, each filterbank row read once for the block's frames
    const int lane = tid & 31, warp = tid >> 5;
    for (int m = warp * (NMEL / 8); m < (warp + 1) * (NMEL / 8); ++m) {
        const float* row = basis + (size_t)m * NBIN;
        float acc[FPB];
#pragma unroll
        for (int f = 0; f < FPB; ++f) acc[f] = 0.f;
        for (int k = lane; k < NBIN; k += 32) {
            const float b = row[k];
#pragma unroll
            for (int f = 0; f < FPB; ++f) acc[f] = fmaf(b, mag[f][k], acc[f]);
        }
        float mine = 0.f;
#pragma unroll
        for (int f = 0; f < FPB; ++f) {
            float v = acc[f];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == f) mine = v;
        }
        if (lane < nf) out[(size_t)m * ld + c0 + lane] = logf(fmaxf(mine, clamp));
    }
}

}  // namespace

RVC_API int rvc_mel_basis_ranges(const float* mel_basis, int32_t* ranges, void* stream) {
    if (!mel_basis || !ranges) return rvc::fail(RVC_EINVAL, "rvc_mel_basis_ranges: null pointer");
    mel_ranges_kernel<<<NMEL, 32, 0, (cudaStream_t)stream>>>(mel_basis, ranges);
    return rvc::check_launch("rvc_mel_basis_ranges");
}

RVC_API int rvc_mel_spectrogram_f32(const float* audio, int64_t n, const float* window, const float* mel_basis, const int32_t* ranges, float* out,
                                    int64_t ld, int32_t T, int32_t Tp, float clamp, void* stream) {
    if (!audio || !window || !mel_basis || !out) return rvc::fail(RVC_EINVAL, "rvc_mel_spectrogram_f32: null pointer");
    if ((uintptr_t)window & 7) return rvc::fail(RVC_EALIGN, "rvc_mel_spectrogram_f32: window must be 8-byte aligned");
    if (n <= NFFT / 2) return rvc::fail(RVC_ESHAPE, "rvc_mel_spectrogram_f32: reflect padding needs more than %d samples (torch.stft raises too)", NFFT / 2);
    if (T != 1 + n / HOP) return rvc::fail(RVC_ESHAPE, "rvc_mel_spectrogram_f32: T must be 1 + n / %d", HOP);
    if (Tp < T || Tp - T >= T || ld < Tp) return rvc::fail(RVC_ESHAPE, "rvc_mel_spectrogram_f32: need T <= Tp <= ld and Tp - T < T (reflect padding of the frame axis)");
    mel_front_kernel<<<rvc::cdiv(Tp, FPB), 256, 0, (cudaStream_t)stream>>>(audio, n, window, mel_basis, ranges, out, ld, T, Tp, clamp);
    return rvc::check_launch("rvc_mel_spectrogram_f32");
}
