// The RMVPE mel front end (pre-network part) as ONE kernel: MelSpectrogram.forward for keyshift = 0, speed = 1, center = True
// (main/library/predictors/RMVPE.py:161-181: torch.stft with a periodic Hann window of 1024, hop 160, reflect padding; magnitude; mel_basis matmul;
// log(clamp(., 1e-5))) plus the reflect padding of the frame axis to a multiple of 32 that mel2hidden applies before the network (:212-213).
// The reference runs this as ~8 library kernels (pad, cuFFT, abs, GEMM, clamp, log, pad) with the [513, T] spectrum in HBM in between; here a block
// takes eight frames: window + radix-2 FFT in shared memory, magnitudes parked in shared memory, then every warp owns 16 mel bands and reads their
// filterbank rows once for the eight frames.  Floating point: fp32 like the reference, different summation order than cuFFT / cuBLAS (tolerance in
// tests/test_gpu_mel_front.py).
#include "common.cuh"

namespace {

constexpr int NFFT = 1024, NBIN = NFFT / 2 + 1, NMEL = 128, HOP = 160, FPB = 8, MAGLD = 516;

__global__ void __launch_bounds__(256) mel_front_kernel(const float* __restrict__ audio, int64_t n, const float* __restrict__ window,
                                                        const float* __restrict__ basis, float* __restrict__ out, int64_t ld, int T, int Tp,
                                                        float clamp) {
    rvc::pdl_prologue();
    __shared__ float2 buf[NFFT];
    __shared__ float2 tw[NFFT / 2];
    __shared__ float mag[FPB][MAGLD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = blockIdx.x * FPB;
    const int nf = min(FPB, Tp - c0);
    for (int k = tid; k < NFFT / 2; k += 256) {        // exp(-2 pi i k / 1024)
        float s, c;
        sincospif(-(float)k * (1.0f / 512.0f), &s, &c);
        tw[k] = make_float2(c, s);
    }
    for (int f = 0; f < FPB; ++f) {
        if (f >= nf) {                                 // (block-uniform) unused frame slots: zeros, never written out
            for (int k = tid; k < NBIN; k += 256) mag[f][k] = 0.f;
            continue;
        }
        const int c = c0 + f;
        const int t = c < T ? c : 2 * (T - 1) - c;     // F.pad(mel, (0, Tp - T), mode="reflect"): column T + i = column T - 2 - i
        const int64_t base = (int64_t)t * HOP - NFFT / 2;
        for (int i = tid; i < NFFT; i += 256) {
            int64_t idx = base + i;                    // torch.stft(center=True, pad_mode="reflect")
            if (idx < 0) idx = -idx;
            if (idx >= n) idx = 2 * (n - 1) - idx;
            buf[__brev((unsigned)i) >> 22] = make_float2(audio[idx] * window[i], 0.f);
        }
        __syncthreads();
#pragma unroll 1
        for (int half = 1; half < NFFT; half <<= 1) {  // decimation in time, 10 stages of 512 butterflies
            const int tstep = (NFFT / 2) / half;
            for (int j = tid; j < NFFT / 2; j += 256) {
                const int pos = j & (half - 1);
                const int i0 = ((j - pos) << 1) + pos, i1 = i0 + half;
                const float2 w = tw[pos * tstep], a = buf[i0], b = buf[i1];
                const float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
                buf[i0] = make_float2(a.x + bw.x, a.y + bw.y);
                buf[i1] = make_float2(a.x - bw.x, a.y - bw.y);
            }
            __syncthreads();
        }
        for (int k = tid; k < NBIN; k += 256) mag[f][k] = sqrtf(buf[k].x * buf[k].x + buf[k].y * buf[k].y);    // :168
        __syncthreads();
    }
    __syncthreads();
    // mel_basis @ magnitude (:178), log(clamp) (:181): warp w owns bands 16 w .. 16 w + 15, each filterbank row read once for the block's frames
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

RVC_API int rvc_mel_spectrogram_f32(const float* audio, int64_t n, const float* window, const float* mel_basis, float* out, int64_t ld, int32_t T,
                                    int32_t Tp, float clamp, void* stream) {
    if (!audio || !window || !mel_basis || !out) return rvc::fail(RVC_EINVAL, "rvc_mel_spectrogram_f32: null pointer");
    if (n <= NFFT / 2) return rvc::fail(RVC_ESHAPE, "rvc_mel_spectrogram_f32: reflect padding needs more than %d samples (torch.stft raises too)", NFFT / 2);
    if (T != 1 + n / HOP) return rvc::fail(RVC_ESHAPE, "rvc_mel_spectrogram_f32: T must be 1 + n / %d", HOP);
    if (Tp < T || Tp - T >= T || ld < Tp) return rvc::fail(RVC_ESHAPE, "rvc_mel_spectrogram_f32: need T <= Tp <= ld and Tp - T < T (reflect padding of the frame axis)");
    mel_front_kernel<<<rvc::cdiv(Tp, FPB), 256, 0, (cudaStream_t)stream>>>(audio, n, window, mel_basis, out, ld, T, Tp, clamp);
    return rvc::check_launch("rvc_mel_spectrogram_f32");
}
