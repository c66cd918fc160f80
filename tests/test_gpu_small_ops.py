"""Op-level checks of the encoder's non-GEMM kernels through the C ABI: LayerNorm (+ residual, mask, packed copy) and the
attention softmax stage, against plain torch fp32/fp64 restatements of synthesizers.py:179-181 and :226-248."""
import math

import pytest
import torch
import torch.nn.functional as F

from rvc_maker_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def unpack16(buf, B, C_, L, rows):
    t = buf.view(torch.float16).view(B, C_ // 8, rows, 8)[:, :, 64:64 + L].float()
    return t.permute(0, 1, 3, 2).reshape(B, C_, L)


@pytest.mark.parametrize("C_,T", [(192, 1000), (64, 37), (128, 33), (256, 500), (320, 100), (192, 1)])
@pytest.mark.parametrize("with_res", [True, False])
def test_layernorm_matches_torch(C_, T, with_res):
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    B = 2
    g = torch.Generator().manual_seed(C_ + T)
    x = (torch.randn(B, C_, T, generator=g) * 3 + 1).to(DEV)
    r = torch.randn(B, C_, T, generator=g).to(DEV) if with_res else None
    gamma, beta = torch.randn(C_, generator=g).to(DEV), torch.randn(C_, generator=g).to(DEV)
    mask = (torch.rand(B, T, generator=g) > 0.3).float().to(DEV)
    rows = lib.rvc_packed_rows(T)
    for out_mask in (0, 1):
        y = torch.empty_like(x)
        ya = torch.zeros(B * C_ * rows * 2, dtype=torch.uint8, device=DEV)
        _lib.check(lib.rvc_layernorm_f32(x.data_ptr(), r.data_ptr() if with_res else None, None, None, gamma.data_ptr(), beta.data_ptr(), mask.data_ptr(),
                                         y.data_ptr(), B, C_, T, 1e-5, out_mask, ya.data_ptr(), 2, st), "layernorm")
        torch.cuda.synchronize()
        xin = (x + r) if with_res else x
        ref = F.layer_norm(xin.double().transpose(1, 2), (C_,), gamma.double(), beta.double(), 1e-5).transpose(1, 2)
        want = ref * mask.double().unsqueeze(1) if out_mask else ref
        assert float((y.double() - want).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
        packed = unpack16(ya, B, C_, T, rows)
        assert float((packed.double() - ref * mask.double().unsqueeze(1)).abs().max()) <= 2e-3 * max(1.0, float(ref.abs().max()))
    if with_res:   # split-K partial sums r + r2 + r3
        r2, r3 = torch.randn(B, C_, T, generator=g).to(DEV), torch.randn(B, C_, T, generator=g).to(DEV)
        y3 = torch.empty_like(x)
        _lib.check(lib.rvc_layernorm_f32(x.data_ptr(), r.data_ptr(), r2.data_ptr(), r3.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mask.data_ptr(),
                                         y3.data_ptr(), B, C_, T, 1e-5, 0, None, 0, st), "layernorm r2 r3")
        torch.cuda.synchronize()
        ref3 = F.layer_norm((x + r + r2 + r3).double().transpose(1, 2), (C_,), gamma.double(), beta.double(), 1e-5).transpose(1, 2)
        assert float((y3.double() - ref3).abs().max()) <= 2e-5 * max(1.0, float(ref3.abs().max()))
        assert lib.rvc_layernorm_f32(x.data_ptr(), r.data_ptr(), r2.data_ptr(), None, gamma.data_ptr(), beta.data_ptr(), mask.data_ptr(),
                                     y3.data_ptr(), B, C_, T, 1e-5, 0, None, 0, st) != 0
    # in place (how the encoder calls it)
    x2 = x.clone()
    _lib.check(lib.rvc_layernorm_f32(x2.data_ptr(), r.data_ptr() if with_res else None, None, None, gamma.data_ptr(), beta.data_ptr(), mask.data_ptr(),
                                     x2.data_ptr(), B, C_, T, 1e-5, 1, None, 0, st), "layernorm in place")
    torch.cuda.synchronize()
    assert torch.equal(x2, y)


@pytest.mark.parametrize("B,T", [(1, 1000), (2, 100), (3, 257), (1, 4000)])
def test_attn_softmax_stage(B, T):
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    H, nh, W = 192, 2, 10
    dk, NR = H // nh, 2 * W + 1
    Tk = (T + 15) // 16 * 16
    g = torch.Generator().manual_seed(T)
    qkv = torch.randn(B, 3 * H, T, generator=g).to(DEV)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    lens[0] = T
    mask = (torch.arange(T)[None] < lens[:, None]).float().to(DEV)
    emb_k, emb_v = (torch.randn(NR, dk, generator=g) * 0.3).to(DEV), (torch.randn(NR, dk, generator=g) * 0.3).to(DEV)
    q = qkv[:, :H].view(B, nh, dk, T).double()
    k = qkv[:, H:2 * H].view(B, nh, dk, T).double()
    v = qkv[:, 2 * H:].view(B, nh, dk, T).double()
    # S[b,h,j,i] = q_i . k_j (what the QK^T launch produces); rows >= T are never read
    S = torch.zeros(B, nh, Tk, T, device=DEV)
    S[:, :, :T] = torch.einsum("bhdi,bhdj->bhji", q, k).float()
    S_blocked = S.view(B, nh, Tk // 8, 8, T).permute(0, 1, 2, 4, 3).contiguous()      # [b][h][key/8][query][8], as the GEMM writes it
    rows = lib.rvc_packed_rows(T)
    P = torch.zeros(B * nh * (Tk // 8) * rows * 16, dtype=torch.uint8, device=DEV)
    R = torch.zeros(B, H, T, device=DEV)
    vw = torch.zeros(B * nh * (Tk // 8) * dk * 16, dtype=torch.uint8, device=DEV)
    _lib.check(lib.rvc_attn_softmax(S_blocked.data_ptr(), qkv.data_ptr(), mask.data_ptr(), emb_k.data_ptr(), emb_v.data_ptr(), P.data_ptr(), R.data_ptr(),
                                    vw.data_ptr(), B, H, nh, T, Tk, W, st), "attn_softmax")
    torch.cuda.synchronize()
    # reference (synthesizers.py:226-248 in closed form)
    scores = S[:, :, :T].double().transpose(2, 3) / math.sqrt(dk)                       # [b,h,i,j]
    i_idx, j_idx = torch.arange(T, device=DEV)[:, None], torch.arange(T, device=DEV)[None]
    rel = j_idx - i_idx + W
    band = (rel >= 0) & (rel < NR)
    qe = torch.einsum("bhdi,rd->bhir", q / math.sqrt(dk), emb_k.double())                # [b,h,i,r]
    scores = scores + torch.where(band, qe.gather(3, rel.clamp(0, NR - 1).expand(B, nh, T, T)), torch.zeros((), device=DEV, dtype=torch.float64))
    mm = mask[:, None, :, None] * mask[:, None, None, :]
    scores = scores.masked_fill(mm == 0, -1e4)
    p = torch.softmax(scores, -1)                                                        # [b,h,i,j]
    Pg = P.view(torch.float16).view(B, nh, Tk // 8, rows, 8)[:, :, :, 64:64 + T].float()  # [b,h,jc,i,8]
    Pg = Pg.permute(0, 1, 3, 2, 4).reshape(B, nh, T, Tk)[..., :T]
    assert float((Pg.double() - p).abs().max()) <= 1e-3
    pr = torch.where(band, p, torch.zeros((), device=DEV, dtype=torch.float64))
    Rref = torch.zeros(B, nh, T, dk, device=DEV, dtype=torch.float64)
    for r_ in range(NR):
        jj = torch.arange(T, device=DEV) + r_ - W
        ok = (jj >= 0) & (jj < T)
        pv = torch.where(ok[None, None], p[:, :, torch.arange(T, device=DEV), jj.clamp(0, T - 1)], torch.zeros((), device=DEV, dtype=torch.float64))
        Rref += pv.unsqueeze(-1) * emb_v.double()[r_]
    Rg = R.view(B, nh, dk, T).transpose(2, 3).double()
    assert float((Rg - Rref).abs().max()) <= 1e-5
    vg = vw.view(torch.float16).view(B, nh, Tk // 8, dk, 8).float().permute(0, 1, 3, 2, 4).reshape(B, nh, dk, Tk)
    assert torch.equal(vg[..., :T], v.float().half().float())
    assert float(vg[..., T:].abs().max()) == 0.0 if Tk > T else True
