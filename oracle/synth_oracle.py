"""Functional CPU restatement of the reference ``Synthesizer.infer`` (TEST INFRASTRUCTURE — see oracle/__init__.py).

Plain PyTorch ops on whatever dtype the state dict is cast to (fp32 by default, fp64 for the precision floor).
Each function cites the reference lines it follows (paths relative to /root/reference/main/library/algorithm/).
Weight-norm is folded up front (w = g * v / ||v||, norm over all dims but 0 — the semantics of
``torch.nn.utils.parametrizations.weight_norm`` the reference applies at synthesizers.py:30,133, residuals.py:17,
modules.py:22,28,32); everything else keeps the reference's operation order so that fp32 results agree to rounding.

Pinned against the unmodified reference by tests/golden (oracle/make_golden.py, tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # residuals.py:14


def fold_state_dict(sd, dtype=torch.float32):
    """Accept in-memory (``parametrizations.weight.original{0,1}``) and on-disk (``weight_g``/``weight_v``, fp16,
    train.py:729,742) forms; return plain ``.weight`` tensors in ``dtype``."""
    out = {}
    pend = {}
    for k, v in sd.items():
        if k.startswith("enc_q."):
            continue
        v = v.detach().to(dtype)
        for suf_g, suf_v in ((".parametrizations.weight.original0", ".parametrizations.weight.original1"), (".weight_g", ".weight_v")):
            if k.endswith(suf_g):
                pend.setdefault(k[: -len(suf_g)], {})["g"] = v
                break
            if k.endswith(suf_v):
                pend.setdefault(k[: -len(suf_v)], {})["v"] = v
                break
        else:
            out[k] = v
    for base, gv in pend.items():
        v, g = gv["v"], gv["g"]
        norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
        out[base + ".weight"] = v * (g / norm)
    return out


def sequence_mask(length, max_length):
    # commons.py:43-46
    return torch.arange(max_length, dtype=length.dtype).unsqueeze(0) < length.unsqueeze(1)


def layer_norm_c(x, gamma, beta, eps=1e-5):
    # synthesizers.py:179-181 — LayerNorm over the channel axis of [B,C,T]
    return F.layer_norm(x.transpose(1, -1), (x.size(1),), gamma, beta, eps).transpose(1, -1)


def rel_attention(w, p, x, attn_mask, n_heads, window=10):
    """MultiHeadAttention with shared-head relative position tables — synthesizers.py:221-284.

    The reference's pad/reshape skew (:259-279) is restated in closed form (verified bit-exact in SURVEY 8a a4):
    scores[i,j] += q_i . E_k[j-i+w] and out_i += sum_j p[i,j] E_v[j-i+w] for |j-i| <= w.
    """
    b, d, t = x.shape
    dk = d // n_heads
    q = F.conv1d(x, w[p + ".conv_q.weight"], w[p + ".conv_q.bias"])
    k = F.conv1d(x, w[p + ".conv_k.weight"], w[p + ".conv_k.bias"])
    v = F.conv1d(x, w[p + ".conv_v.weight"], w[p + ".conv_v.bias"])
    q = q.view(b, n_heads, dk, t).transpose(2, 3)
    k = k.view(b, n_heads, dk, t).transpose(2, 3)
    v = v.view(b, n_heads, dk, t).transpose(2, 3)
    qs = q / math.sqrt(dk)
    scores = torch.matmul(qs, k.transpose(-2, -1))  # :226
    ek = w[p + ".emb_rel_k"][0]  # [2w+1, dk]
    ev = w[p + ".emb_rel_v"][0]
    idx = torch.arange(t)
    rel = idx.unsqueeze(0) - idx.unsqueeze(1) + window  # [i,j] -> j-i+w
    band = (rel >= 0) & (rel <= 2 * window)
    relc = rel.clamp(0, 2 * window)
    qe = torch.matmul(qs, ek.t())  # [b,h,t,2w+1]   (:230, _matmul_with_relative_keys)
    scores = scores + torch.gather(qe, 3, relc.expand(b, n_heads, t, t)) * band
    scores = scores.masked_fill(attn_mask == 0, -1e4)  # :242
    p_attn = F.softmax(scores, dim=-1)
    out = torch.matmul(p_attn, v)
    # :248 relative values: sum_j p[i,j] E_v[j-i+w]
    pb = p_attn * band
    pr = torch.zeros(b, n_heads, t, 2 * window + 1, dtype=x.dtype)
    pr.scatter_add_(3, relc.expand(b, n_heads, t, t), pb)
    out = out + torch.matmul(pr, ev)
    out = out.transpose(2, 3).contiguous().view(b, d, t)
    return F.conv1d(out, w[p + ".conv_o.weight"], w[p + ".conv_o.bias"])


def ffn(w, p, x, x_mask, ks):
    # synthesizers.py:302-315 (relu activation, "same" padding)
    pad = ((ks - 1) // 2, ks // 2)
    h = F.conv1d(F.pad(x * x_mask, pad), w[p + ".conv_1.weight"], w[p + ".conv_1.bias"])
    h = torch.relu(h)
    return F.conv1d(F.pad(h * x_mask, pad), w[p + ".conv_2.weight"], w[p + ".conv_2.bias"]) * x_mask


def text_encoder(w, cfg, phone, pitch, lengths):
    # synthesizers.py:366-371 + Encoder.forward :340-348
    H = cfg["hidden_channels"]
    x = F.linear(phone, w["enc_p.emb_phone.weight"], w["enc_p.emb_phone.bias"])
    if pitch is not None:
        x = x + F.embedding(pitch, w["enc_p.emb_pitch.weight"])
    x = F.leaky_relu(x * math.sqrt(H), 0.1).transpose(1, -1)
    x_mask = sequence_mask(lengths, x.size(2)).unsqueeze(1).to(x.dtype)
    x = x * x_mask
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    x = x * x_mask
    for i in range(cfg["n_layers"]):
        y = rel_attention(w, f"enc_p.encoder.attn_layers.{i}", x, attn_mask, cfg["n_heads"])
        x = layer_norm_c(x + y, w[f"enc_p.encoder.norm_layers_1.{i}.gamma"], w[f"enc_p.encoder.norm_layers_1.{i}.beta"])
        y = ffn(w, f"enc_p.encoder.ffn_layers.{i}", x, x_mask, cfg["kernel_size"])
        x = layer_norm_c(x + y, w[f"enc_p.encoder.norm_layers_2.{i}.gamma"], w[f"enc_p.encoder.norm_layers_2.{i}.beta"])
    x = x * x_mask
    stats = F.conv1d(x, w["enc_p.proj.weight"], w["enc_p.proj.bias"]) * x_mask
    m, logs = torch.split(stats, cfg["inter_channels"], dim=1)
    return m, logs, x_mask


def wavenet(w, p, x, x_mask, g, H, n_layers=3, ks=5):
    # modules.py:35-51, gate = commons.py:35-41
    out = torch.zeros_like(x)
    gc = F.conv1d(g, w[p + ".cond_layer.weight"], w[p + ".cond_layer.bias"])
    for i in range(n_layers):
        x_in = F.conv1d(x, w[f"{p}.in_layers.{i}.weight"], w[f"{p}.in_layers.{i}.bias"], padding=(ks - 1) // 2)
        a = x_in + gc[:, i * 2 * H:(i + 1) * 2 * H, :]
        acts = torch.tanh(a[:, :H]) * torch.sigmoid(a[:, H:])
        rs = F.conv1d(acts, w[f"{p}.res_skip_layers.{i}.weight"], w[f"{p}.res_skip_layers.{i}.bias"])
        if i < n_layers - 1:
            x = (x + rs[:, :H]) * x_mask
            out = out + rs[:, H:]
        else:
            out = out + rs
    return out * x_mask


def flow_reverse(w, cfg, x, x_mask, g):
    # residuals.py:91-93 (reverse order: Flip, C3, Flip, C2, Flip, C1, Flip, C0), coupling :127-137, Flip :53-58
    H, half = cfg["hidden_channels"], cfg["inter_channels"] // 2
    for f in (3, 2, 1, 0):
        x = torch.flip(x, [1])
        p = f"flow.flows.{2 * f}"
        x0, x1 = x[:, :half], x[:, half:]
        h = F.conv1d(x0, w[p + ".pre.weight"], w[p + ".pre.bias"]) * x_mask
        h = wavenet(w, p + ".enc", h, x_mask, g, H)
        m = F.conv1d(h, w[p + ".post.weight"], w[p + ".post.bias"]) * x_mask
        x1 = (x1 - m) * x_mask  # mean_only: logs = 0, exp(-0) = 1
        x = torch.cat([x0, x1], 1)
    return x


def sine_source(w, f0, upp, sr, eps_src, sine_amp=0.1, noise_std=0.003):
    """SineGen.forward + SourceModuleHnNSF.forward, harmonic_num=0 — synthesizers.py:79-100, :111-112.

    Returns har_source [B,1,T*upp].  ``rand_ini`` (:87) is zeroed for the fundamental (:88) so it never contributes.
    CPU ``cumsum`` accumulates in double and rounds each prefix to the tensor dtype; this restatement keeps that.
    """
    f0 = f0.unsqueeze(-1)  # [B,T,1]
    rad = f0 / sr * torch.arange(1, upp + 1, dtype=f0.dtype)  # [B,T,upp]
    rad2 = torch.fmod(rad[:, :-1, -1:].float() + 0.5, 1.0) - 0.5
    rad_acc = rad2.cumsum(dim=1).fmod(1.0).to(f0)
    rad = rad + F.pad(rad_acc, (0, 0, 1, 0))
    rad = rad.reshape(f0.shape[0], -1, 1)
    sine = torch.sin(2 * np.pi * rad) * sine_amp
    uv = (f0 > 0).to(f0.dtype)
    uv = F.interpolate(uv.transpose(2, 1), scale_factor=float(upp), mode="nearest").transpose(2, 1)
    sine = sine * uv + (uv * noise_std + (1 - uv) * sine_amp / 3) * eps_src
    har = torch.tanh(F.linear(sine, w["dec.m_source.l_linear.weight"], w["dec.m_source.l_linear.bias"]))
    return har.transpose(1, 2)


def res_block(w, p, x, dilations, ks):
    # residuals.py:32-36 with x_mask=None
    for m, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, w[f"{p}.convs1.{m}.weight"], w[f"{p}.convs1.{m}.bias"], dilation=d, padding=(ks * d - d) // 2)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, w[f"{p}.convs2.{m}.weight"], w[f"{p}.convs2.{m}.bias"], padding=(ks - 1) // 2)
        x = xt + x
    return x


def generator(w, cfg, z, f0, g, eps_src, use_f0=True, taps=None):
    """GeneratorNSF.forward (synthesizers.py:144-161) / the non-f0 Generator body (:39-56)."""
    rates, uks = cfg["upsample_rates"], cfg["upsample_kernel_sizes"]
    rks, rds = cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]
    nk = len(rks)
    upp = int(math.prod(rates))
    har = sine_source(w, f0, upp, cfg["sr"], eps_src) if use_f0 else None
    x = F.conv1d(z, w["dec.conv_pre.weight"], w["dec.conv_pre.bias"], padding=3)
    x = x + F.conv1d(g, w["dec.cond.weight"], w["dec.cond.bias"])
    if taps is not None:
        taps["conv_pre"] = x
        taps["har_source"] = har
    for i, (u, k) in enumerate(zip(rates, uks)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        if use_f0:
            x = F.conv_transpose1d(x, w[f"dec.ups.{i}.weight"], w[f"dec.ups.{i}.bias"], stride=u, padding=(k - u) // 2)
            stride = int(math.prod(rates[i + 1:])) if i + 1 < len(rates) else 1
            kn = 1 if stride == 1 else stride * 2 - stride % 2
            x = x + F.conv1d(har, w[f"dec.noise_convs.{i}.weight"], w[f"dec.noise_convs.{i}.bias"], stride=stride,
                             padding=0 if stride == 1 else (kn - stride) // 2)
            names = [f"dec.resblocks.{i * nk + j}" for j in range(nk)]
        else:
            base = i * (nk + 1)
            x = F.conv_transpose1d(x, w[f"dec.ups_and_resblocks.{base}.weight"], w[f"dec.ups_and_resblocks.{base}.bias"], stride=u, padding=(k - u) // 2)
            names = [f"dec.ups_and_resblocks.{base + 1 + j}" for j in range(nk)]
        if taps is not None:
            taps[f"ups{i}"] = x
        xs = None
        for nm, kk, dd in zip(names, rks, rds):
            y = res_block(w, nm, x, dd, kk)
            xs = y if xs is None else xs + y
        x = xs / nk
        if taps is not None:
            taps[f"stage{i}"] = x
    return torch.tanh(F.conv1d(F.leaky_relu(x), w["dec.conv_post.weight"], None, padding=3))  # slope 0.01 (:161)


def infer(sd, cfg, phone, phone_lengths, pitch=None, nsff0=None, sid=None, rate=None, eps_z=None, eps_src=None,
          use_f0=True, dtype=torch.float32, folded=None, taps=None):
    """``Synthesizer.infer`` — synthesizers.py:446-465.  Noise tensors are explicit; when ``None`` they are drawn in
    the reference's order and shapes (:449 randn_like(m_p); :87 rand(1,1,1); :98 randn_like(sine_waves))."""
    w = folded if folded is not None else fold_state_dict(sd, dtype)
    with torch.no_grad():
        phone = phone.to(dtype)
        g = F.embedding(sid, w["emb_g.weight"]).unsqueeze(-1)
        m_p, logs_p, x_mask = text_encoder(w, cfg, phone, pitch if use_f0 else None, phone_lengths)
        if eps_z is None:
            eps_z = torch.randn_like(m_p)
        z_p = (m_p + torch.exp(logs_p) * eps_z.to(dtype) * 0.66666) * x_mask
        if use_f0:
            nsff0 = nsff0.to(dtype)
        if rate is not None:
            head = int(z_p.shape[2] * (1.0 - float(rate)))
            z_p, x_mask = z_p[:, :, head:], x_mask[:, :, head:]
            if use_f0:
                nsff0 = nsff0[:, head:]
        z = flow_reverse(w, cfg, z_p, x_mask, g)
        if use_f0 and eps_src is None:
            torch.rand(1, 1, 1)
            eps_src = torch.randn(nsff0.shape[0], nsff0.shape[1] * int(math.prod(cfg["upsample_rates"])), 1, dtype=dtype)
        if use_f0:
            eps_src = eps_src[:, : nsff0.shape[1] * int(math.prod(cfg["upsample_rates"]))]
        o = generator(w, cfg, z * x_mask, nsff0, g, eps_src.to(dtype) if use_f0 else None, use_f0=use_f0, taps=taps)
    return o, x_mask, (z, z_p, m_p, logs_p)
