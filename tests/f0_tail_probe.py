"""Timing of the f0 tail kernels (csrc/f0_post.cu) and the mel front end (csrc/mel_front.cu) with CUDA events, and of the reference's path beside them.

    python tests/f0_tail_probe.py [--json out.json]

Algorithmic bytes of rvc_rmvpe_pitch_f32: the salience map read once (T x 360 x 4 B) + 12 B of pitch / pitchf written per frame."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import f0_oracle  # noqa: E402  (timing baseline only)
from rvc_maker_b200 import pitch as P  # noqa: E402


def time_gpu(fn, iters=50, flush=None):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for T in (1001, 60001, 1_000_000):
        sal_h = f0_oracle.synthetic_salience(min(T, 60001), 3)
        sal = torch.from_numpy(sal_h).cuda()
        if T > sal.shape[0]:
            sal = sal.repeat(-(-T // sal.shape[0]), 1)[:T].contiguous()
        nbytes = T * 360 * 4 + T * 12
        med, best = time_gpu(lambda: P.rmvpe_pitch(sal, pitch=2), flush=flush)
        med_t, _ = time_gpu(lambda: P.rmvpe_pitch(sal, pitch=2, f0_autotune=True), flush=flush)
        f0 = P.decode(sal)
        med_c, _ = time_gpu(lambda: P.get_f0_tail(f0, pitch=2), flush=flush)
        row = dict(T=T, audio_s=T / 100.0, rmvpe_pitch_us=med, rmvpe_pitch_best_us=best, rmvpe_pitch_autotune_us=med_t, f0_coarse_us=med_c,
                   gbs=nbytes / med / 1e3, hbm_frac=nbytes / med / 1e3 / peaks["hbm_gbs"])
        if T <= 60001:
            t0 = time.perf_counter()
            with np.errstate(all="ignore"):
                f0_oracle.get_f0_tail(f0_oracle.decode(sal_h), pitch=2)
            row["numpy_host_us"] = (time.perf_counter() - t0) * 1e6
            t0 = time.perf_counter()
            h = sal.cpu().numpy()                                     # the D2H copy the reference pays first (RMVPE.py:226)
            row["d2h_copy_us"] = (time.perf_counter() - t0) * 1e6
        rows.append(row)
        print(json.dumps(row))
    # the mel front end: one launch against the reference's torch calls (oracle/mel_oracle.py on CUDA tensors = cuFFT + cuBLAS + elementwise kernels)
    from oracle import mel_oracle
    basis = torch.from_numpy(mel_oracle.mel_filterbank()).cuda()
    for secs in (10, 600):
        audio = mel_oracle.synthetic_audio(secs * 16000, 5).cuda()
        med, best = time_gpu(lambda: P.mel_spectrogram(audio, basis), flush=flush)
        med_ref, _ = time_gpu(lambda: mel_oracle.pad_frames(mel_oracle.mel_spectrogram(audio, basis)), iters=20, flush=flush)
        T = 1 + audio.numel() // 160
        row = dict(kernel="mel_spectrogram", audio_s=secs, frames=T, mel_us=med, mel_best_us=best, torch_eager_us=med_ref,
                   gflops=(5.0 * 1024 * 10 + 2.0 * 128 * 513) * T / med / 1e3)
        rows.append(row)
        print(json.dumps(row))
    # the split-point search of VC.pipeline: one launch against the reference's 160 host passes (oracle/chunk_oracle.py), after the reference's filtfilt
    from scipy import signal
    from oracle import chunk_oracle
    from rvc_maker_b200 import pipeline as pl
    bh, ah = signal.butter(N=5, Wn=48, btype="high", fs=16000)
    for secs in (100, 600):
        x = (np.random.default_rng(secs).standard_normal(secs * 16000) * 0.1).astype(np.float32)
        t0 = time.perf_counter()
        y = np.ascontiguousarray(signal.filtfilt(bh, ah, x))
        t_filt = time.perf_counter() - t0
        t0 = time.perf_counter()
        want = chunk_oracle.reference_chunk_loop(y, lambda a, f0, f1: np.zeros(3 * 48000, np.float32), 48000)[1]
        t_host = time.perf_counter() - t0
        yd = torch.from_numpy(y).cuda()
        assert pl.split_points(yd) == want
        med, best = time_gpu(lambda: pl.split_points(yd), iters=20, flush=flush)      # (includes the .tolist() read-back of the indices)
        yp = torch.from_numpy(y).pin_memory()
        t0 = time.perf_counter()
        yd2 = yp.cuda(non_blocking=True); torch.cuda.synchronize()
        t_h2d = time.perf_counter() - t0
        xd = torch.from_numpy(x).cuda()
        pl.filtfilt(xd, bh, ah)
        med_f, _ = time_gpu(lambda: pl.filtfilt(xd, bh, ah), iters=10, flush=flush)
        err = float((pl.filtfilt(xd, bh, ah).cpu() - torch.from_numpy(y)).abs().max())
        rows.append(dict(kernel="filtfilt", audio_s=secs, device_us=med_f, host_scipy_us=t_filt * 1e6, max_abs_diff_vs_scipy=err))
        print(json.dumps(rows[-1]))
        row = dict(kernel="split_points", audio_s=secs, candidates=len(want), device_us=med, device_best_us=best, h2d_pinned_us=t_h2d * 1e6,
                   host_search_us=t_host * 1e6, host_filtfilt_us=t_filt * 1e6)
        rows.append(row)
        print(json.dumps(row))
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
