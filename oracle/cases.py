"""Deterministic synthetic weights and inputs for the parity suite (TEST INFRASTRUCTURE).

Everything is drawn from ``numpy.random.Generator(PCG64(seed))`` so that the build
container (where the unmodified reference runs and writes ``tests/golden``) and the
GPU box (where only the oracle restatement and the CUDA path exist) regenerate
bit-identical tensors without shipping 110 MB of weights.

State-dict key names and shapes follow the reference module tree:
``/root/reference/main/library/algorithm/synthesizers.py:396-432`` (Synthesizer),
``:350-364`` (TextEncoder), ``:317-338`` (Encoder), ``:183-219`` (MultiHeadAttention),
``:286-300`` (FFN), ``:114-142`` (GeneratorNSF), ``:21-37`` (Generator),
``residuals.py:23-30`` (ResBlock), ``:108-125`` (ResidualCouplingLayer),
``modules.py:10-33`` (WaveNet).  ``oracle/make_golden.py`` proves the key set by loading
these dicts into the reference with ``strict=True`` (minus the train-only ``enc_q``).
"""
from __future__ import annotations

import math

import numpy as np
import torch

_RB = dict(resblock="1", resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])
_ENC = dict(inter_channels=192, hidden_channels=192, filter_channels=768, n_heads=2, n_layers=6, kernel_size=3, p_dropout=0,
            upsample_initial_channel=512, spk_embed_dim=109, gin_channels=256)

# main/configs/v2/40000.json:23-41, v2/48000.json:23-41, v1/32000.json:27-45
CONFIGS = {
    "v2_40k": dict(spec_channels=1025, segment_size=32, **_ENC, **_RB, upsample_rates=[10, 10, 2, 2],
                   upsample_kernel_sizes=[16, 16, 4, 4], sr=40000, text_enc_hidden_dim=768),
    "v2_48k": dict(spec_channels=1025, segment_size=36, **_ENC, **_RB, upsample_rates=[12, 10, 2, 2],
                   upsample_kernel_sizes=[24, 20, 4, 4], sr=48000, text_enc_hidden_dim=768),
    "v1_32k": dict(spec_channels=513, segment_size=40, **_ENC, **_RB, upsample_rates=[10, 4, 2, 2, 2],
                   upsample_kernel_sizes=[16, 16, 4, 4, 4], sr=32000, text_enc_hidden_dim=256),
    # reduced-width model whose full weights + outputs fit in a committed fixture
    "tiny": dict(spec_channels=33, segment_size=8, inter_channels=16, hidden_channels=32, filter_channels=64, n_heads=2,
                 n_layers=2, kernel_size=3, p_dropout=0, resblock="1", resblock_kernel_sizes=[3, 5],
                 resblock_dilation_sizes=[[1, 2], [1, 3]], upsample_rates=[4, 2], upsample_initial_channel=32,
                 upsample_kernel_sizes=[8, 4], spk_embed_dim=4, gin_channels=16, sr=8000, text_enc_hidden_dim=24),
}

CTOR_ORDER = ["spec_channels", "segment_size", "inter_channels", "hidden_channels", "filter_channels", "n_heads", "n_layers",
              "kernel_size", "p_dropout", "resblock", "resblock_kernel_sizes", "resblock_dilation_sizes", "upsample_rates",
              "upsample_initial_channel", "upsample_kernel_sizes", "spk_embed_dim", "gin_channels", "sr"]


def ctor_args(cfg):
    """Positional list in the order of ``cpt["config"]`` (train.py:730)."""
    return [cfg[k] for k in CTOR_ORDER]


def upp_of(cfg):
    return int(math.prod(cfg["upsample_rates"]))


# ----------------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------------
class _Gen:
    def __init__(self, seed):
        self.rng = np.random.Generator(np.random.PCG64(seed))

    def normal(self, shape, std):
        return (self.rng.standard_normal(size=shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    def uniform(self, shape, lo, hi):
        return (self.rng.random(size=shape, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


def synth_state_dict(cfg, use_f0=True, seed=1234, heated=True):
    """Random state dict in the reference's *in-memory* key format (``parametrizations.weight.original{0,1}``).

    ``heated=False`` reproduces the statistics of the reference's own initialisation (ResBlock / ups weights
    N(0, 0.01), ``commons.py:5-6``; ``post`` convs zero, ``residuals.py:121-122``) — a DC-dominated output.
    ``heated=True`` draws fan-in-scaled weights so that every stage contributes to the waveform (SURVEY section 7,
    hard part 1).
    """
    g = _Gen(seed)
    sd = {}
    H, I, Fc = cfg["hidden_channels"], cfg["inter_channels"], cfg["filter_channels"]
    D, G, U = cfg["text_enc_hidden_dim"], cfg["gin_channels"], cfg["upsample_initial_channel"]
    nh, ks = cfg["n_heads"], cfg["kernel_size"]
    dk = H // nh

    def plain(name, cout, cin, k, gain=1.0, bias=True, transposed=False):
        fan_in = cin * k
        std = gain / math.sqrt(fan_in) if heated else 1.0 / math.sqrt(3 * fan_in)
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        sd[name + ".weight"] = g.normal(shape, std)
        if bias:
            sd[name + ".bias"] = g.uniform((cout,), -0.1, 0.1) if heated else g.uniform((cout,), -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    def wnorm(name, cout, cin, k, gain=1.0, cold_std=None, transposed=False):
        # weight_norm(dim=0): original0 = g [dim0,1,1], original1 = v [dim0,dim1,k]
        fan_in = cin * k
        if transposed:
            fan_in = cin * max(1, k // 2)  # k/u taps hit each output; only a scale heuristic
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        if heated:
            v = g.normal(shape, 1.0 / math.sqrt(fan_in))
            n = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
            gg = n * g.uniform((shape[0], 1, 1), 0.7, 1.3) * np.float32(gain)
        else:
            v = g.normal(shape, cold_std if cold_std is not None else 1.0 / math.sqrt(3 * fan_in))
            n = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
            gg = n.copy()  # weight_norm initialises g = ||v||
        sd[name + ".parametrizations.weight.original0"] = gg.astype(np.float32)
        sd[name + ".parametrizations.weight.original1"] = v
        sd[name + ".bias"] = g.uniform((cout,), -0.1, 0.1) if heated else g.uniform((cout,), -1 / math.sqrt(cin * k), 1 / math.sqrt(cin * k))

    # ---- enc_p (synthesizers.py:350-364)
    sd["enc_p.emb_phone.weight"] = g.normal((H, D), (0.5 if heated else 0.577) / math.sqrt(D))
    sd["enc_p.emb_phone.bias"] = g.uniform((H,), -0.05, 0.05)
    if use_f0:
        sd["enc_p.emb_pitch.weight"] = g.normal((256, H), 0.05 if heated else 1.0)
    for i in range(cfg["n_layers"]):
        a = f"enc_p.encoder.attn_layers.{i}"
        sd[a + ".emb_rel_k"] = g.normal((1, 21, dk), dk ** -0.5)
        sd[a + ".emb_rel_v"] = g.normal((1, 21, dk), dk ** -0.5)
        for nm in ("conv_q", "conv_k", "conv_v", "conv_o"):
            plain(f"{a}.{nm}", H, H, 1, gain=1.5 if nm in ("conv_q", "conv_k") else 1.0)
        for nl in ("norm_layers_1", "norm_layers_2"):
            sd[f"enc_p.encoder.{nl}.{i}.gamma"] = g.uniform((H,), 0.8, 1.2) if heated else np.ones((H,), np.float32)
            sd[f"enc_p.encoder.{nl}.{i}.beta"] = g.uniform((H,), -0.1, 0.1) if heated else np.zeros((H,), np.float32)
        plain(f"enc_p.encoder.ffn_layers.{i}.conv_1", Fc, H, ks, gain=1.4)
        plain(f"enc_p.encoder.ffn_layers.{i}.conv_2", H, Fc, ks)
    plain("enc_p.proj", 2 * I, H, 1, gain=0.5)

    # ---- flow (residuals.py:71-125, modules.py:10-33)
    half = I // 2
    nl = 3
    for f in range(4):
        p = f"flow.flows.{2 * f}"
        plain(p + ".pre", H, half, 1)
        for l in range(nl):
            wnorm(f"{p}.enc.in_layers.{l}", 2 * H, H, 5)
            wnorm(f"{p}.enc.res_skip_layers.{l}", 2 * H if l < nl - 1 else H, H, 1, gain=0.7)
        wnorm(p + ".enc.cond_layer", 2 * H * nl, G, 1, gain=0.5)
        if heated:
            plain(p + ".post", half, H, 1, gain=0.5)
        else:
            sd[p + ".post.weight"] = np.zeros((half, H, 1), np.float32)
            sd[p + ".post.bias"] = np.zeros((half,), np.float32)

    # ---- dec (synthesizers.py:114-142 / :21-37)
    rates, uks = cfg["upsample_rates"], cfg["upsample_kernel_sizes"]
    rks, rds = cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]
    n_up = len(rates)
    chans = [U // (2 ** (i + 1)) for i in range(n_up)]
    plain("dec.conv_pre", U, I, 7)
    plain("dec.cond", U, G, 1, gain=0.5)

    def resblock(prefix, ch, k, dil):
        for m, _d in enumerate(dil):
            wnorm(f"{prefix}.convs1.{m}", ch, ch, k, gain=0.9, cold_std=0.01)
        for m, _d in enumerate(dil):
            wnorm(f"{prefix}.convs2.{m}", ch, ch, k, gain=0.5, cold_std=0.01)

    if use_f0:
        sd["dec.m_source.l_linear.weight"] = np.array([[0.9]], np.float32) if heated else g.uniform((1, 1), -1, 1)
        sd["dec.m_source.l_linear.bias"] = np.array([0.05], np.float32) if heated else g.uniform((1,), -1, 1)
        for i in range(n_up):
            wnorm(f"dec.ups.{i}", chans[i], U // (2 ** i), uks[i], transposed=True, cold_std=0.01, gain=1.2)
            stride = int(math.prod(rates[i + 1:])) if i + 1 < n_up else 1
            kn = 1 if stride == 1 else stride * 2 - stride % 2
            plain(f"dec.noise_convs.{i}", chans[i], 1, kn, gain=2.0)
        j = 0
        for i in range(n_up):
            for k, dil in zip(rks, rds):
                resblock(f"dec.resblocks.{j}", chans[i], k, dil)
                j += 1
    else:
        n = 0
        for i in range(n_up):
            wnorm(f"dec.ups_and_resblocks.{n}", chans[i], U // (2 ** i), uks[i], transposed=True, cold_std=0.01, gain=1.2)
            n += 1
            for k, dil in zip(rks, rds):
                resblock(f"dec.ups_and_resblocks.{n}", chans[i], k, dil)
                n += 1
    sd["dec.conv_post.weight"] = g.normal((1, chans[-1], 7), (0.6 if heated else 0.577) / math.sqrt(chans[-1] * 7))
    sd["emb_g.weight"] = g.normal((cfg["spk_embed_dim"], G), 1.0)
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def to_disk_form(sd, half=True):
    """Key/dtype form written by the reference's exporter (train.py:729, :742): ``weight_g``/``weight_v`` names, fp16."""
    out = {}
    for k, v in sd.items():
        k = k.replace(".parametrizations.weight.original1", ".weight_v").replace(".parametrizations.weight.original0", ".weight_g")
        out[k] = v.half() if half else v.clone()
    return out


# ----------------------------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------------------------
F0_MIN, F0_MAX = 50.0, 1100.0  # convert.py:196-199


def coarse_pitch(f0):
    """Coarse-pitch quantiser, float64, ``np.rint`` (half-to-even) — restates convert.py:318-323."""
    f0 = np.asarray(f0, dtype=np.float64)
    mel_min = 1127.0 * np.log(1.0 + F0_MIN / 700.0)
    mel_max = 1127.0 * np.log(1.0 + F0_MAX / 700.0)
    mel = 1127.0 * np.log(1.0 + f0 / 700.0)
    pos = mel > 0
    mel[pos] = (mel[pos] - mel_min) * 254.0 / (mel_max - mel_min) + 1.0
    mel[mel <= 1] = 1
    mel[mel > 255] = 255
    return np.rint(mel).astype(np.int32)


def synth_f0(rng, B, T):
    """Piecewise-smooth contour in 80–600 Hz with ~30 % unvoiced frames in runs >= 5."""
    out = np.zeros((B, T), np.float64)
    for b in range(B):
        steps = rng.standard_normal(T + 9) * 0.03
        lf = np.log(200.0) + np.cumsum(steps)
        lf = np.convolve(lf, np.ones(10) / 10.0, mode="valid")[:T]
        f = np.clip(np.exp(lf), 80.0, 600.0)
        t = 0
        while t < T:
            run = int(rng.integers(5, 40))
            if rng.random() < 0.3:
                f[t:t + run] = 0.0
            t += run
        out[b] = f
    return out


def synth_inputs(cfg, B, T, seed=0, ragged=False, use_f0=True):
    rng = np.random.Generator(np.random.PCG64(seed))
    D, I = cfg["text_enc_hidden_dim"], cfg["inter_channels"]
    upp = upp_of(cfg)
    phone = rng.standard_normal(size=(B, T, D), dtype=np.float32)
    f0 = synth_f0(rng, B, T)
    lengths = np.full((B,), T, np.int64)
    if ragged and B > 1:
        lengths[1:] = rng.integers(T // 2, T + 1, size=B - 1)
    eps_z = rng.standard_normal(size=(B, I, T), dtype=np.float32)
    eps_src = rng.standard_normal(size=(B, T * upp, 1), dtype=np.float32)
    d = dict(
        phone=torch.from_numpy(phone),
        phone_lengths=torch.from_numpy(lengths),
        sid=torch.from_numpy((np.arange(B) % cfg["spk_embed_dim"]).astype(np.int64)),
        eps_z=torch.from_numpy(eps_z),
    )
    if use_f0:
        d["pitch"] = torch.from_numpy(coarse_pitch(f0).astype(np.int64))
        d["nsff0"] = torch.from_numpy(f0.astype(np.float32))
        d["eps_src"] = torch.from_numpy(eps_src)
    else:
        d["pitch"] = None
        d["nsff0"] = None
        d["eps_src"] = None
    return d


def metrics(cand, ref):
    """max-abs, total-power SNR and SNR of the mean-removed (AC) signal, in dB (SURVEY section 7 hard part 1)."""
    c = cand.detach().double().cpu().flatten()
    r = ref.detach().double().cpu().flatten()
    err = c - r
    mx = float(err.abs().max()) if err.numel() else 0.0
    pe = float((err ** 2).sum())
    snr = 10.0 * math.log10(float((r ** 2).sum()) / pe) if pe > 0 else float("inf")
    rac = r - r.mean()
    snr_ac = 10.0 * math.log10(max(float((rac ** 2).sum()), 1e-300) / pe) if pe > 0 else float("inf")
    return dict(max_abs=mx, snr_db=snr, snr_ac_db=snr_ac)
