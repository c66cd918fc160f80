"""CPU study (not a test): which fp16 operand roundings of the TextEncoder produce the error of m_p / logs_p.  The oracle's encoder is
re-run with operands rounded to fp16 at selected GEMM sites (fp32 accumulate), against the unrounded fp32 run."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import cases, synth_oracle as so  # noqa: E402

SITES = ("emb", "qkv", "qk", "pv", "o", "ffn1", "ffn2", "proj")


def enc(w, cfg, phone, pitch, lengths, rs, split=()):
    """rs: sites whose operands are rounded to fp16; split: sites computed with the 3-term hi/lo expansion instead"""
    def r(t, site):
        return t.half().float() if site in rs else t

    def mm(site, fn, a, b):
        # fn(a, b) bilinear; fp16 rounding of both operands, or the 3-term split a_hi*b_hi + a_lo*b_hi + a_hi*b_lo
        if site in split:
            ah, bh = a.half().float(), b.half().float()
            al, bl = (a - ah).half().float(), (b - bh).half().float()
            return fn(ah, bh) + fn(al, bh) + fn(ah, bl)
        return fn(r(a, site), r(b, site))

    H, nh = cfg["hidden_channels"], cfg["n_heads"]
    x = mm("emb", lambda a, b: F.linear(a, b), phone, w["enc_p.emb_phone.weight"]) + w["enc_p.emb_phone.bias"]
    x = x + F.embedding(pitch, w["enc_p.emb_pitch.weight"])
    x = F.leaky_relu(x * math.sqrt(H), 0.1).transpose(1, -1)
    x_mask = so.sequence_mask(lengths, x.size(2)).unsqueeze(1).to(x.dtype)
    x = x * x_mask
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    b, d, t = x.shape
    dk = d // nh
    window = 10
    idx = torch.arange(t)
    rel = idx.unsqueeze(0) - idx.unsqueeze(1) + window
    band = (rel >= 0) & (rel <= 2 * window)
    relc = rel.clamp(0, 2 * window)
    for i in range(cfg["n_layers"]):
        p = f"enc_p.encoder.attn_layers.{i}"
        wqkv = torch.cat([w[p + ".conv_q.weight"], w[p + ".conv_k.weight"], w[p + ".conv_v.weight"]], 0)
        bqkv = torch.cat([w[p + ".conv_q.bias"], w[p + ".conv_k.bias"], w[p + ".conv_v.bias"]], 0)
        qkv = mm("qkv", lambda a, bb: F.conv1d(a, bb), x, wqkv) + bqkv.view(1, -1, 1)
        q, k, v = [u.view(b, nh, dk, t).transpose(2, 3) for u in qkv.split(d, dim=1)]
        qs = q / math.sqrt(dk)
        scores = mm("qk", lambda a, bb: torch.matmul(a, bb.transpose(-2, -1)), q, k) / math.sqrt(dk)
        ek, ev = w[p + ".emb_rel_k"][0], w[p + ".emb_rel_v"][0]
        qe = torch.matmul(qs, ek.t())
        scores = scores + torch.gather(qe, 3, relc.expand(b, nh, t, t)) * band
        scores = scores.masked_fill(attn_mask == 0, -1e4)
        pa = F.softmax(scores, dim=-1)
        out = mm("pv", lambda a, bb: torch.matmul(a, bb), pa, v)
        pr = torch.zeros(b, nh, t, 2 * window + 1, dtype=x.dtype)
        pr.scatter_add_(3, relc.expand(b, nh, t, t), r(pa, "pv") * band)
        out = out + torch.matmul(pr, ev)
        out = out.transpose(2, 3).contiguous().view(b, d, t)
        y = mm("o", lambda a, bb: F.conv1d(a, bb), out, w[p + ".conv_o.weight"]) + w[p + ".conv_o.bias"].view(1, -1, 1)
        x = so.layer_norm_c(x + y, w[f"enc_p.encoder.norm_layers_1.{i}.gamma"], w[f"enc_p.encoder.norm_layers_1.{i}.beta"])
        fp = f"enc_p.encoder.ffn_layers.{i}"
        h = mm("ffn1", lambda a, bb: F.conv1d(F.pad(a, (1, 1)), bb), x * x_mask, w[fp + ".conv_1.weight"]) + w[fp + ".conv_1.bias"].view(1, -1, 1)
        h = torch.relu(h)
        y = (mm("ffn2", lambda a, bb: F.conv1d(F.pad(a, (1, 1)), bb), h * x_mask, w[fp + ".conv_2.weight"]) + w[fp + ".conv_2.bias"].view(1, -1, 1)) * x_mask
        x = so.layer_norm_c(x + y, w[f"enc_p.encoder.norm_layers_2.{i}.gamma"], w[f"enc_p.encoder.norm_layers_2.{i}.beta"])
    x = x * x_mask
    stats = (mm("proj", lambda a, bb: F.conv1d(a, bb), x, w["enc_p.proj.weight"]) + w["enc_p.proj.bias"].view(1, -1, 1)) * x_mask
    return stats


def main():
    cname, B, T = "v2_48k", 2, int(sys.argv[1]) if len(sys.argv) > 1 else 600
    cfg = cases.CONFIGS[cname]
    sd = cases.synth_state_dict(cfg, seed=778, heated=True)
    w = so.fold_state_dict(sd)
    inp = cases.synth_inputs(cfg, B, T, seed=5, ragged=True)
    ref = enc(w, cfg, inp["phone"], inp["pitch"], inp["phone_lengths"], ())
    m_ref, _, _ = so.text_encoder(w, cfg, inp["phone"], inp["pitch"], inp["phone_lengths"])
    print("closed form vs oracle:", float((ref[:, :cfg["inter_channels"]] - m_ref).abs().max()))

    def report(tag, rs, split=()):
        s = enc(w, cfg, inp["phone"], inp["pitch"], inp["phone_lengths"], rs, split)
        e = (s - ref).abs()
        print(f"{tag:34s} stats max_abs={float(e.max()):.3e}  rms={float(e.pow(2).mean().sqrt()):.3e}  (max|ref|={float(ref.abs().max()):.2f})", flush=True)

    report("all sites fp16", SITES)
    for s in SITES:
        report(f"only {s}", (s,))
    for s in SITES:
        report(f"all but {s}", tuple(x for x in SITES if x != s))
    report("all fp16, qk split", tuple(x for x in SITES if x != "qk"), ("qk",))
    report("all fp16, qkv+qk split", tuple(x for x in SITES if x not in ("qk", "qkv")), ("qk", "qkv"))
    report("all split", (), SITES)
    report("all split but pv", ("pv",), tuple(x for x in SITES if x != "pv"))


if __name__ == "__main__":
    main()
