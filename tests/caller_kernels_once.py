"""One launch of every caller-side kernel at the ten-minute-stream size (for `ncu --set full`; not a test):
rvc_rmvpe_pitch_f32 (60 001 frames and 10^6 frames), rvc_mel_spectrogram_f32 (600 s), rvc_split_points (600 s, float64), rvc_filtfilt (600 s)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy import signal  # noqa: E402

from oracle import f0_oracle, mel_oracle  # noqa: E402
from rvc_maker_b200 import pipeline as pl  # noqa: E402
from rvc_maker_b200 import pitch as P  # noqa: E402

sal = torch.from_numpy(f0_oracle.synthetic_salience(60001, 3)).cuda()
big = sal.repeat(17, 1)[:1_000_000].contiguous()
audio = mel_oracle.synthetic_audio(600 * 16000, 5).cuda()
basis = torch.from_numpy(mel_oracle.mel_filterbank()).cuda()
bh, ah = signal.butter(N=5, Wn=48, btype="high", fs=16000)
torch.cuda.synchronize()
P.rmvpe_pitch(sal, pitch=2)
P.rmvpe_pitch(big, pitch=2)
P.mel_spectrogram(audio, basis)
y = pl.filtfilt(audio[0], bh, ah)
print(pl.split_points(y)[:3])
torch.cuda.synchronize()
