"""Op-level parity of the tcgen05 Conv1d (rvc_conv1d_tc) through the C ABI.

Reference = fp64 torch conv1d on operands rounded exactly as the kernel rounds them (bf16 / fp16, round-to-nearest), so the
only difference left is fp32 accumulation order: tolerance 2e-4 * max|ref| (up to 2816 products per output).
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16, FP16 = 0, 2
DEV = "cuda:0"


def rnd(x, dtype):
    return x.to(torch.bfloat16 if dtype == BF16 else torch.float16).to(torch.float32)


def unpack(buf, B, C_, L, dtype, rows):
    t = buf.view(torch.bfloat16 if dtype == BF16 else torch.float16).view(B, C_ // 8, rows, 8)
    return t[:, :, 64:64 + L, :].permute(0, 1, 3, 2).reshape(B, C_, L).float()


def pad_rows_zero(buf, B, C_, L, dtype, rows):
    t = buf.view(torch.bfloat16 if dtype == BF16 else torch.float16).view(B, C_ // 8, rows, 8)
    return float(t[:, :, :64].abs().max()) == 0.0 and float(t[:, :, 64 + L:].abs().max()) == 0.0


def gate_perm(H, nt):
    """weight-column order for a gated conv: per N tile, nt/2 tanh rows then the matching nt/2 sigmoid rows."""
    idx = []
    for t in range(2 * H // nt):
        idx += list(range(t * nt // 2, (t + 1) * nt // 2)) + list(range(H + t * nt // 2, H + (t + 1) * nt // 2))
    return torch.tensor(idx)


def run_case(dtype, Cin, Cout, K, dil, B, L, seed=0, njobs=1, gated=False, relu=False, use_mask=False, use_res=True, alpha=1.0,
             slope=0.1, nt=0, mt=0, use_bias_b=False):
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    esz = 2
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.nt, d.mt = njobs, B, Cin, Cout, L, dtype, nt, mt
    Ceff = Cout // 2 if gated else Cout
    keep, checks = [], []
    mask = (torch.rand(B, L, generator=g) > 0.2).float().to(DEV) if use_mask else None
    for j in range(njobs):
        Kj, dj = (K, dil) if j == 0 else ((3, 1) if j == 1 else (7, 3))
        x = torch.randn(B, Cin, L, generator=g).to(DEV)
        w = (torch.randn(Cout, Cin, Kj, generator=g) / (Cin * Kj) ** 0.5).to(DEV)
        bias = torch.randn(Cout, generator=g).to(DEV)
        bias_b = torch.randn(B, Cout + 5, generator=g).to(DEV) if use_bias_b else None
        res_full = torch.randn(B, Ceff + 3, L, generator=g).to(DEV) if use_res else None   # strided residual (channel slice)
        a = torch.zeros(B * Cin * rows * esz, dtype=torch.uint8, device=DEV)
        wp = torch.zeros(Kj * Cin * Cout * esz, dtype=torch.uint8, device=DEV)
        y_act = torch.zeros(B * Ceff * rows * esz, dtype=torch.uint8, device=DEV)
        y_full = torch.zeros(B, Ceff + 2, L, device=DEV)                                    # strided output (channel slice)
        _lib.check(lib.rvc_pack_act(x.data_ptr(), None, None, 1.0, Cin * L, L, None, a.data_ptr(), B, Cin, L, dtype, slope, st), "pack_act")
        if gated:
            perm = gate_perm(Cout // 2, nt).to(DEV)
            wk, bk = w[perm].contiguous(), bias[perm].contiguous()
            bbk = torch.cat([bias_b[:, :Cout][:, perm], bias_b[:, Cout:]], 1).contiguous() if use_bias_b else None
        else:
            wk, bk, bbk = w, bias, bias_b
        _lib.check(lib.rvc_pack_weight(wk.data_ptr(), wp.data_ptr(), Cout, Cin, Kj, dtype, st), "pack_weight")
        jb = d.jobs[j]
        jb.a, jb.w, jb.bias = a.data_ptr(), wp.data_ptr(), bk.data_ptr()
        if use_bias_b:
            jb.bias_b, jb.bias_b_bs = bbk.data_ptr(), Cout + 5
        if use_mask:
            jb.mask = mask.data_ptr()
        if use_res:
            jb.res, jb.res_bs, jb.res_cs = res_full.data_ptr() + 4 * L, (Ceff + 3) * L, L   # channels 1..Ceff
        jb.y_f32, jb.y_bs, jb.y_cs = y_full.data_ptr() + 4 * L, (Ceff + 2) * L, L           # channels 1..Ceff
        jb.y_act = y_act.data_ptr()
        jb.K, jb.dil, jb.pad, jb.gated, jb.relu, jb.alpha, jb.act_slope = Kj, dj, (Kj - 1) // 2 * dj, int(gated), int(relu), alpha, slope
        keep += [x, w, bias, bias_b, res_full, a, wp, y_act, y_full, wk, bk, bbk]
        checks.append((x, w, bias, bias_b, res_full, a, y_act, y_full, Kj, dj))
    _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "conv1d_tc")
    torch.cuda.synchronize()
    for x, w, bias, bias_b, res_full, a, y_act, y_full, Kj, dj in checks:
        xa = rnd(F.leaky_relu(x, slope), dtype)
        assert torch.equal(unpack(a, B, Cin, L, dtype, rows), xa), "pack_act mismatch"
        ref = F.conv1d(xa.double(), rnd(w, dtype).double(), bias.double(), dilation=dj, padding=(Kj - 1) // 2 * dj)
        if use_bias_b:
            ref = ref + bias_b[:, :Cout].double().unsqueeze(-1)
        if gated:
            H = Cout // 2
            ref = torch.tanh(ref[:, :H]) * torch.sigmoid(ref[:, H:])
        if relu:
            ref = torch.relu(ref)
        if use_res:
            ref = res_full[:, 1:1 + Ceff].double() + alpha * ref
        if use_mask:
            ref = ref * mask.double().unsqueeze(1)
        y = y_full[:, 1:1 + Ceff]
        scale = max(1.0, float(ref.abs().max()))
        err = float((y.double() - ref).abs().max())
        assert err <= 2e-4 * scale, (dtype, Cin, Cout, Kj, dj, B, L, err, scale)
        assert float(y_full[:, 0].abs().max()) == 0.0 and float(y_full[:, -1].abs().max()) == 0.0, "wrote outside the channel slice"
        assert torch.equal(unpack(y_act, B, Ceff, L, dtype, rows), rnd(F.leaky_relu(y, slope), dtype)), "y_act mismatch"
        assert pad_rows_zero(y_act, B, Ceff, L, dtype, rows)


@pytest.mark.parametrize("dtype", [BF16, FP16])
@pytest.mark.parametrize("Cc", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("K,dil", [(3, 1), (7, 3), (11, 5)])
def test_resblock_shapes(dtype, Cc, K, dil):
    run_case(dtype, Cc, Cc, K, dil, B=2, L=1000)


@pytest.mark.parametrize("L", [1, 127, 128, 129, 255, 256, 257, 777, 5000])
def test_ragged_lengths(L):
    run_case(BF16, 128, 128, 7, 1, B=1, L=L, use_res=False)
    run_case(FP16, 64, 64, 11, 5, B=3, L=L)


@pytest.mark.parametrize("nt,mt", [(32, 1), (32, 2), (32, 4), (64, 1), (64, 2), (64, 4), (128, 1), (128, 2)])
def test_forced_tiles(nt, mt):
    run_case(FP16, 128, 128, 5, 2, B=2, L=1500, nt=nt, mt=mt)


def test_three_jobs_many_tiles():
    run_case(BF16, 128, 128, 11, 5, B=2, L=40000, njobs=3)
    run_case(FP16, 32, 32, 11, 1, B=1, L=100000, njobs=3)          # auto-selects 4 sub-tiles per CTA tile
    run_case(FP16, 64, 64, 11, 5, B=1, L=300001, njobs=3)
    run_case(FP16, 16, 16, 11, 5, B=2, L=160000, njobs=3)           # v1-32k last stage


@pytest.mark.parametrize("Cin,Cout,K", [(192, 576, 1), (192, 768, 3), (768, 192, 3), (192, 384, 1), (96, 192, 1), (192, 96, 1), (192, 512, 7)])
def test_encoder_flow_shapes(Cin, Cout, K):
    """TextEncoder / flow / conv_pre shapes: Cin != Cout, streamed input-channel stages (Cin = 768), mask, relu, alpha."""
    run_case(FP16, Cin, Cout, K, 1, B=2, L=1000, relu=(Cout == 768), use_mask=True, alpha=-1.0, slope=1.0, use_bias_b=(Cout == 512))


@pytest.mark.parametrize("nt", [64, 128])
def test_gated_wavenet_conv(nt):
    """WaveNet in_layer: 192 -> 2*192, k = 5, tanh * sigmoid gate with per-batch conditioning bias (modules.py:43-45)."""
    run_case(FP16, 192, 384, 5, 1, B=2, L=1000, gated=True, use_res=False, slope=1.0, nt=nt, use_bias_b=True)


@pytest.mark.parametrize("Cin,Cout,k,u,Lin,dtype", [(512, 256, 24, 12, 1000, BF16), (256, 128, 20, 10, 3000, BF16), (128, 64, 4, 2, 5001, FP16),
                                                    (64, 32, 4, 2, 777, FP16), (32, 16, 4, 2, 100000, FP16), (512, 256, 16, 10, 300, FP16), (256, 128, 16, 4, 500, FP16)])
def test_conv_transpose_phase_mode(Cin, Cout, k, u, Lin, dtype):
    """ConvTranspose1d(k, stride u, pad (k-u)/2) as u phase GEMMs, with pre-lrelu input, residual (noise conv) and packed output."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    B = 2
    g = torch.Generator(device="cpu").manual_seed(5)
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(B, Cin, Lin, generator=g).to(DEV)
    w = (torch.randn(Cin, Cout, k, generator=g) / (Cin * k / u) ** 0.5).to(DEV)   # ConvTranspose1d weight layout
    bias = torch.randn(Cout, generator=g).to(DEV)
    Lout = Lin * u
    res = torch.randn(B, Cout, Lout, generator=g).to(DEV)
    M = -(-k // u)
    # phase-major conv weight [Cout][Cin][u*M]: block r, tap t = w[:, :, r + (M-1-t)*u]
    wk = torch.zeros(Cout, Cin, u * M, device=DEV)
    for r in range(u):
        for t in range(M):
            kk = r + (M - 1 - t) * u
            if kk < k:
                wk[:, :, r * M + t] = w[:, :, kk].t()
    rows_in, rows_out = lib.rvc_packed_rows(Lin), lib.rvc_packed_rows(Lout)
    a = torch.zeros(B * Cin * rows_in * 2, dtype=torch.uint8, device=DEV)
    wp = torch.zeros(u * M * Cin * Cout * 2, dtype=torch.uint8, device=DEV)
    y_act = torch.zeros(B * Cout * rows_out * 2, dtype=torch.uint8, device=DEV)
    y = torch.zeros(B, Cout, Lout, device=DEV)
    _lib.check(lib.rvc_pack_act(x.data_ptr(), None, None, 1.0, Cin * Lin, Lin, None, a.data_ptr(), B, Cin, Lin, dtype, 0.1, st), "pack_act")
    _lib.check(lib.rvc_pack_weight(wk.data_ptr(), wp.data_ptr(), Cout, Cin, u * M, dtype, st), "pack_weight")
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.up, d.pshift = 1, B, Cin, Cout, Lin, dtype, u, (k - u) // 2
    jb = d.jobs[0]
    jb.a, jb.w, jb.bias = a.data_ptr(), wp.data_ptr(), bias.data_ptr()
    jb.res, jb.res_bs, jb.res_cs = res.data_ptr(), Cout * Lout, Lout
    jb.y_f32, jb.y_bs, jb.y_cs = y.data_ptr(), Cout * Lout, Lout
    jb.y_act = y_act.data_ptr()
    jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = M, 1, M - 1, 1.0, 0.1
    _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "conv1d_tc(up)")
    torch.cuda.synchronize()
    xa = rnd(F.leaky_relu(x, 0.1), dtype)
    ref = F.conv_transpose1d(xa.double(), rnd(w, dtype).double(), bias.double(), stride=u, padding=(k - u) // 2) + res.double()
    scale = max(1.0, float(ref.abs().max()))
    err = float((y.double() - ref).abs().max())
    assert err <= 2e-4 * scale, (err, scale)
    assert torch.equal(unpack(y_act, B, Cout, Lout, dtype, rows_out), rnd(F.leaky_relu(y, 0.1), dtype))
    assert pad_rows_zero(y_act, B, Cout, Lout, dtype, rows_out)


def test_sum_jobs_stage_average():
    """sum_jobs: three convs (k = 11/7/3) accumulate into one tile; v = alpha * (sum conv + sum bias + sum res)."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    B, Cc, L, dtype = 2, 64, 3001, FP16
    g = torch.Generator(device="cpu").manual_seed(9)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.sum_jobs = 3, B, Cc, Cc, L, dtype, 1
    y = torch.zeros(B, Cc, L, device=DEV)
    y_act = torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
    keep, ref = [], 0.0
    for j, (k, dil) in enumerate(((11, 5), (7, 3), (3, 1))):
        x = torch.randn(B, Cc, L, generator=g).to(DEV)
        w = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).to(DEV)
        bias = torch.randn(Cc, generator=g).to(DEV)
        res = torch.randn(B, Cc, L, generator=g).to(DEV)
        a = torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
        wp = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
        _lib.check(lib.rvc_pack_act(x.data_ptr(), None, None, 1.0, Cc * L, L, None, a.data_ptr(), B, Cc, L, dtype, 0.1, st), "pack_act")
        _lib.check(lib.rvc_pack_weight(w.data_ptr(), wp.data_ptr(), Cc, Cc, k, dtype, st), "pack_weight")
        jb = d.jobs[j]
        jb.a, jb.w, jb.bias = a.data_ptr(), wp.data_ptr(), bias.data_ptr()
        jb.res, jb.res_bs, jb.res_cs = res.data_ptr(), Cc * L, L
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = k, dil, (k - 1) // 2 * dil, 1.0 / 3.0, 0.1
        if j == 0:
            jb.y_f32, jb.y_bs, jb.y_cs, jb.y_act = y.data_ptr(), Cc * L, L, y_act.data_ptr()
        keep += [x, w, bias, res, a, wp]
        ref = ref + F.conv1d(rnd(F.leaky_relu(x, 0.1), dtype).double(), rnd(w, dtype).double(), bias.double(), dilation=dil,
                             padding=(k - 1) // 2 * dil) + res.double()
    _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "conv1d_tc(sum)")
    torch.cuda.synchronize()
    ref = ref / 3.0
    assert float((y.double() - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))
    assert torch.equal(unpack(y_act, B, Cc, L, dtype, rows), rnd(F.leaky_relu(y, 0.1), dtype))


def test_shape_errors_are_loud():
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype = 1, 1, 24, 64, 100, FP16
    d.jobs[0].a = d.jobs[0].w = 1
    d.jobs[0].K, d.jobs[0].dil, d.jobs[0].pad = 3, 1, 1
    assert lib.rvc_conv1d_tc(C.byref(d), None) == -2
    assert b"multiple of 16" in lib.rvc_last_error()


def to_blocked(x):
    """[B, C, L] -> channel-blocked [B, C/8, L, 8] (rvc_conv1d_tc_desc.f32_blocked)."""
    B, C_, L = x.shape
    return x.view(B, C_ // 8, 8, L).permute(0, 1, 3, 2).contiguous()


def from_blocked(xb):
    B, nb, L, _ = xb.shape
    return xb.permute(0, 1, 3, 2).reshape(B, nb * 8, L)


@pytest.mark.parametrize("Cc,K,L,blocked", [(128, 7, 3001, 3), (32, 11, 20000, 3), (16, 3, 5000, 3), (64, 3, 777, 1), (256, 7, 1200, 2)])
def test_blocked_fp32_streams(Cc, K, L, blocked):
    """Channel-blocked residual / output (the generator's residual streams): same numbers as the plain layout, bit for bit."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    B, dtype = 2, FP16
    g = torch.Generator(device="cpu").manual_seed(21)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    x = torch.randn(B, Cc, L, generator=g).to(DEV)
    w = (torch.randn(Cc, Cc, K, generator=g) / (Cc * K) ** 0.5).to(DEV)
    bias = torch.randn(Cc, generator=g).to(DEV)
    res = torch.randn(B, Cc, L, generator=g).to(DEV)
    a = torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
    wp = torch.zeros(K * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
    _lib.check(lib.rvc_pack_act(x.data_ptr(), None, None, 1.0, Cc * L, L, None, a.data_ptr(), B, Cc, L, dtype, 0.1, st), "pack_act")
    _lib.check(lib.rvc_pack_weight(w.data_ptr(), wp.data_ptr(), Cc, Cc, K, dtype, st), "pack_weight")
    outs = []
    for blk in (0, blocked):
        d = _lib.Conv1dTcDesc()
        d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.f32_blocked = 1, B, Cc, Cc, L, dtype, blk
        r_in = to_blocked(res) if blk & 1 else res
        y = torch.zeros(B, Cc // 8, L, 8, device=DEV) if blk & 2 else torch.zeros(B, Cc, L, device=DEV)
        y_act = torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
        jb = d.jobs[0]
        jb.a, jb.w, jb.bias = a.data_ptr(), wp.data_ptr(), bias.data_ptr()
        jb.res, jb.res_bs, jb.res_cs = r_in.data_ptr(), Cc * L, (8 * L if blk & 1 else L)
        jb.y_f32, jb.y_bs, jb.y_cs = y.data_ptr(), Cc * L, (8 * L if blk & 2 else L)
        jb.y_act = y_act.data_ptr()
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = K, 1, (K - 1) // 2, 1.0, 0.1
        _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "conv1d_tc")
        torch.cuda.synchronize()
        outs.append((from_blocked(y) if blk & 2 else y, y_act))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    ref = F.conv1d(rnd(F.leaky_relu(x, 0.1), dtype).double(), rnd(w, dtype).double(), bias.double(), padding=(K - 1) // 2) + res.double()
    assert float((outs[1][0].double() - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))


def test_blocked_phase_mode_and_stage_average():
    """Blocked residual + output through the ConvTranspose1d phase mode, then blocked residuals into a plain sum_jobs average."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    B, Cin, Cout, k, u, Lin, dtype = 2, 128, 64, 4, 2, 2501, FP16
    g = torch.Generator(device="cpu").manual_seed(22)
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(B, Cin, Lin, generator=g).to(DEV)
    w = (torch.randn(Cin, Cout, k, generator=g) / (Cin * k / u) ** 0.5).to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    Lout = Lin * u
    res = torch.randn(B, Cout, Lout, generator=g).to(DEV)
    M = -(-k // u)
    wk = torch.zeros(Cout, Cin, u * M, device=DEV)
    for r in range(u):
        for t in range(M):
            kk = r + (M - 1 - t) * u
            if kk < k:
                wk[:, :, r * M + t] = w[:, :, kk].t()
    rows_in, rows_out = lib.rvc_packed_rows(Lin), lib.rvc_packed_rows(Lout)
    a = torch.zeros(B * Cin * rows_in * 2, dtype=torch.uint8, device=DEV)
    wp = torch.zeros(u * M * Cin * Cout * 2, dtype=torch.uint8, device=DEV)
    y_act = torch.zeros(B * Cout * rows_out * 2, dtype=torch.uint8, device=DEV)
    yb = to_blocked(res)                      # in place: residual in, result out (as the engine does with the noise conv)
    _lib.check(lib.rvc_pack_act(x.data_ptr(), None, None, 1.0, Cin * Lin, Lin, None, a.data_ptr(), B, Cin, Lin, dtype, 0.1, st), "pack_act")
    _lib.check(lib.rvc_pack_weight(wk.data_ptr(), wp.data_ptr(), Cout, Cin, u * M, dtype, st), "pack_weight")
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.up, d.pshift, d.f32_blocked = 1, B, Cin, Cout, Lin, dtype, u, (k - u) // 2, 3
    jb = d.jobs[0]
    jb.a, jb.w, jb.bias = a.data_ptr(), wp.data_ptr(), bias.data_ptr()
    jb.res, jb.res_bs, jb.res_cs = yb.data_ptr(), Cout * Lout, 8 * Lout
    jb.y_f32, jb.y_bs, jb.y_cs = yb.data_ptr(), Cout * Lout, 8 * Lout
    jb.y_act = y_act.data_ptr()
    jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = M, 1, M - 1, 1.0, 0.1
    _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "conv1d_tc(up, blocked)")
    torch.cuda.synchronize()
    y = from_blocked(yb)
    ref = F.conv_transpose1d(rnd(F.leaky_relu(x, 0.1), dtype).double(), rnd(w, dtype).double(), bias.double(), stride=u, padding=(k - u) // 2) + res.double()
    assert float((y.double() - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))
    assert torch.equal(unpack(y_act, B, Cout, Lout, dtype, rows_out), rnd(F.leaky_relu(y, 0.1), dtype))

    # stage average: blocked residuals in, plain [B, C, L] average out
    Cc, L = Cout, Lout
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.sum_jobs, d.f32_blocked = 3, B, Cc, Cc, L, dtype, 1, 1
    avg = torch.zeros(B, Cc, L, device=DEV)
    keep, ref = [], 0.0
    for j, (kk, dil) in enumerate(((11, 1), (7, 1), (3, 1))):
        xj = torch.randn(B, Cc, L, generator=g).to(DEV)
        wj = (torch.randn(Cc, Cc, kk, generator=g) / (Cc * kk) ** 0.5).to(DEV)
        bj = torch.randn(Cc, generator=g).to(DEV)
        rj = torch.randn(B, Cc, L, generator=g).to(DEV)
        rjb = to_blocked(rj)
        aj = torch.zeros(B * Cc * rows_out * 2, dtype=torch.uint8, device=DEV)
        wpj = torch.zeros(kk * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
        _lib.check(lib.rvc_pack_act(xj.data_ptr(), None, None, 1.0, Cc * L, L, None, aj.data_ptr(), B, Cc, L, dtype, 0.1, st), "pack_act")
        _lib.check(lib.rvc_pack_weight(wj.data_ptr(), wpj.data_ptr(), Cc, Cc, kk, dtype, st), "pack_weight")
        jb = d.jobs[j]
        jb.a, jb.w, jb.bias = aj.data_ptr(), wpj.data_ptr(), bj.data_ptr()
        jb.res, jb.res_bs, jb.res_cs = rjb.data_ptr(), Cc * L, 8 * L
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = kk, dil, (kk - 1) // 2, 1.0 / 3.0, 0.1
        if j == 0:
            jb.y_f32, jb.y_bs, jb.y_cs = avg.data_ptr(), Cc * L, L
        keep += [xj, wj, bj, rj, rjb, aj, wpj]
        ref = ref + F.conv1d(rnd(F.leaky_relu(xj, 0.1), dtype).double(), rnd(wj, dtype).double(), bj.double(), padding=(kk - 1) // 2) + rj.double()
    _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "conv1d_tc(sum, blocked res)")
    torch.cuda.synchronize()
    ref = ref / 3.0
    assert float((avg.double() - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))


def test_noise_conv_blocked_output():
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    B, Cc, Lout, sn = 2, 64, 3000, 4
    kn, pn = sn * 2, sn // 2
    g = torch.Generator(device="cpu").manual_seed(23)
    st = torch.cuda.current_stream().cuda_stream
    har = torch.randn(B, Lout * sn, generator=g).to(DEV)
    w = torch.randn(Cc, kn, generator=g).to(DEV)
    bias = torch.randn(Cc, generator=g).to(DEV)
    y0 = torch.zeros(B, Cc, Lout, device=DEV)
    y1 = torch.zeros(B, Cc // 8, Lout, 8, device=DEV)
    _lib.check(lib.rvc_noise_conv_f32(har.data_ptr(), Lout * sn, w.data_ptr(), bias.data_ptr(), y0.data_ptr(), Cc * Lout, B, Cc, Lout, kn, sn, pn, 0, st), "noise")
    _lib.check(lib.rvc_noise_conv_f32(har.data_ptr(), Lout * sn, w.data_ptr(), bias.data_ptr(), y1.data_ptr(), Cc * Lout, B, Cc, Lout, kn, sn, pn, 1, st), "noise")
    torch.cuda.synchronize()
    assert torch.equal(y0, from_blocked(y1))
    ref = F.conv1d(har.unsqueeze(1), w.unsqueeze(1), bias, stride=sn, padding=pn)
    assert float((y0 - ref).abs().max()) <= 1e-4
    assert lib.rvc_noise_conv_f32(har.data_ptr(), Lout * sn, w.data_ptr(), bias.data_ptr(), y1.data_ptr(), Cc * Lout, B, 12, Lout, kn, sn, pn, 1, st) != 0


@pytest.mark.parametrize("Cc,L,B", [(64, 3001, 2), (32, 20000, 1), (64, 246, 1), (32, 1, 1), (64, 50000, 1), (32, 777, 3)])
def test_fused_resblock_pair_equals_two_launches(Cc, L, B):
    """rvc_resblock_pair_tc (conv1 -> lrelu -> conv2 + residual in one kernel, intermediate in shared memory) against the two
    rvc_conv1d_tc launches it replaces: same operands, same accumulation order -> bit-identical fp32 and packed outputs."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    dtype = FP16
    g = torch.Generator(device="cpu").manual_seed(31 + Cc + L)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    pd = _lib.ResblockPairDesc()
    pd.njobs, pd.B, pd.C, pd.L, pd.dtype, pd.slope = 3, B, Cc, L, dtype, 0.1
    d1, d2 = _lib.Conv1dTcDesc(), _lib.Conv1dTcDesc()
    for d in (d1, d2):
        d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype = 3, B, Cc, Cc, L, dtype
    d2.f32_blocked = 3
    keep, outs = [], []
    for j, (k, dil) in enumerate(((11, 5), (7, 3), (3, 1))):
        x = torch.randn(B, Cc, L, generator=g).to(DEV)
        w1 = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).to(DEV)
        w2 = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).to(DEV)
        b1, b2 = torch.randn(Cc, generator=g).to(DEV), torch.randn(Cc, generator=g).to(DEV)
        xb = to_blocked(x)
        a = torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
        w1p = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
        w2p = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
        _lib.check(lib.rvc_pack_act(x.data_ptr(), None, None, 1.0, Cc * L, L, None, a.data_ptr(), B, Cc, L, dtype, 0.1, st), "pack_act")
        _lib.check(lib.rvc_pack_weight(w1.data_ptr(), w1p.data_ptr(), Cc, Cc, k, dtype, st), "pack_weight")
        _lib.check(lib.rvc_pack_weight(w2.data_ptr(), w2p.data_ptr(), Cc, Cc, k, dtype, st), "pack_weight")
        ta = torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
        y_ref, ya_ref = torch.zeros_like(xb), torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
        y_fused, ya_fused = torch.zeros_like(xb), torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
        jb = d1.jobs[j]
        jb.a, jb.w, jb.bias, jb.y_act = a.data_ptr(), w1p.data_ptr(), b1.data_ptr(), ta.data_ptr()
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = k, dil, (k - 1) // 2 * dil, 1.0, 0.1
        jb = d2.jobs[j]
        jb.a, jb.w, jb.bias, jb.y_act = ta.data_ptr(), w2p.data_ptr(), b2.data_ptr(), ya_ref.data_ptr()
        jb.res, jb.res_bs, jb.res_cs = xb.data_ptr(), Cc * L, 8 * L
        jb.y_f32, jb.y_bs, jb.y_cs = y_ref.data_ptr(), Cc * L, 8 * L
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = k, 1, (k - 1) // 2, 1.0, 0.1
        pj = pd.jobs[j]
        pj.a, pj.w1, pj.b1, pj.w2, pj.b2 = a.data_ptr(), w1p.data_ptr(), b1.data_ptr(), w2p.data_ptr(), b2.data_ptr()
        pj.res, pj.y, pj.y_act = xb.data_ptr(), y_fused.data_ptr(), ya_fused.data_ptr()
        pj.K1, pj.dil1, pj.K2 = k, dil, k
        keep += [x, w1, w2, b1, b2, xb, a, w1p, w2p, ta]
        outs.append((y_ref, ya_ref, y_fused, ya_fused, x, w1, w2, b1, b2, dil, k))
    _lib.check(lib.rvc_conv1d_tc(C.byref(d1), st), "conv1")
    _lib.check(lib.rvc_conv1d_tc(C.byref(d2), st), "conv2")
    _lib.check(lib.rvc_resblock_pair_tc(C.byref(pd), st), "pair")
    torch.cuda.synchronize()
    for y_ref, ya_ref, y_fused, ya_fused, x, w1, w2, b1, b2, dil, k in outs:
        assert torch.equal(y_fused, y_ref)
        assert torch.equal(ya_fused, ya_ref)
        t = rnd(F.leaky_relu(F.conv1d(rnd(F.leaky_relu(x, 0.1), dtype).double(), rnd(w1, dtype).double(), b1.double(), dilation=dil,
                                      padding=(k - 1) // 2 * dil).float(), 0.1), dtype)
        ref = F.conv1d(t.double(), rnd(w2, dtype).double(), b2.double(), padding=(k - 1) // 2) + x.double()
        assert float((from_blocked(y_fused).double() - ref).abs().max()) <= 3e-4 * max(1.0, float(ref.abs().max()))


def test_fused_resblock_pair_rejects_wide_stages():
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    pd = _lib.ResblockPairDesc()
    pd.njobs, pd.B, pd.C, pd.L, pd.dtype, pd.slope = 1, 1, 128, 1000, FP16, 0.1
    assert lib.rvc_resblock_pair_tc(C.byref(pd), None) == -2
    assert b"C = 32 and 64" in lib.rvc_last_error()


@pytest.mark.parametrize("Cc,L,B,y_ncl", [(64, 3001, 2, 1), (32, 20000, 1, 0), (64, 246, 1, 0), (32, 5, 2, 1), (64, 30011, 1, 1)])
def test_fused_resblock_pair_stage_average(Cc, L, B, y_ncl):
    """sum_jobs mode of rvc_resblock_pair_tc (last ResBlock step + stage average in one kernel) against conv1 x3 followed by the
    sum_jobs conv2 launch of rvc_conv1d_tc: bit-identical average (plain or channel-blocked) and packed copy."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    dtype = FP16
    g = torch.Generator(device="cpu").manual_seed(57 + Cc + L)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    pd = _lib.ResblockPairDesc()
    pd.njobs, pd.B, pd.C, pd.L, pd.dtype, pd.slope, pd.sum_jobs, pd.y_ncl, pd.alpha = 3, B, Cc, L, dtype, 0.1, 1, y_ncl, 1.0 / 3.0
    d1, d2 = _lib.Conv1dTcDesc(), _lib.Conv1dTcDesc()
    for d in (d1, d2):
        d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype = 3, B, Cc, Cc, L, dtype
    d2.sum_jobs, d2.f32_blocked = 1, (1 if y_ncl else 3)
    shape = (B, Cc, L) if y_ncl else (B, Cc // 8, L, 8)
    y_ref, y_fused = torch.zeros(shape, device=DEV), torch.zeros(shape, device=DEV)
    ya_ref = torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
    ya_fused = torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
    keep = []
    for j, (k, dil) in enumerate(((11, 5), (7, 5), (3, 5))):
        x = torch.randn(B, Cc, L, generator=g).to(DEV)
        xin = torch.randn(B, Cc, L, generator=g).to(DEV)
        w1 = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).to(DEV)
        w2 = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).to(DEV)
        b1, b2 = torch.randn(Cc, generator=g).to(DEV), torch.randn(Cc, generator=g).to(DEV)
        xb = to_blocked(x)
        a = torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
        w1p = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
        w2p = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
        _lib.check(lib.rvc_pack_act(xin.data_ptr(), None, None, 1.0, Cc * L, L, None, a.data_ptr(), B, Cc, L, dtype, 0.1, st), "pack_act")
        _lib.check(lib.rvc_pack_weight(w1.data_ptr(), w1p.data_ptr(), Cc, Cc, k, dtype, st), "pack_weight")
        _lib.check(lib.rvc_pack_weight(w2.data_ptr(), w2p.data_ptr(), Cc, Cc, k, dtype, st), "pack_weight")
        ta = torch.zeros(B * Cc * rows * 2, dtype=torch.uint8, device=DEV)
        jb = d1.jobs[j]
        jb.a, jb.w, jb.bias, jb.y_act = a.data_ptr(), w1p.data_ptr(), b1.data_ptr(), ta.data_ptr()
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = k, dil, (k - 1) // 2 * dil, 1.0, 0.1
        jb = d2.jobs[j]
        jb.a, jb.w, jb.bias = ta.data_ptr(), w2p.data_ptr(), b2.data_ptr()
        jb.res, jb.res_bs, jb.res_cs = xb.data_ptr(), Cc * L, 8 * L
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = k, 1, (k - 1) // 2, 1.0 / 3.0, 0.1
        pj = pd.jobs[j]
        pj.a, pj.w1, pj.b1, pj.w2, pj.b2, pj.res = a.data_ptr(), w1p.data_ptr(), b1.data_ptr(), w2p.data_ptr(), b2.data_ptr(), xb.data_ptr()
        pj.K1, pj.dil1, pj.K2 = k, dil, k
        if j == 0:
            jb.y_f32, jb.y_bs, jb.y_cs, jb.y_act = y_ref.data_ptr(), Cc * L, (L if y_ncl else 8 * L), ya_ref.data_ptr()
            pj.y, pj.y_act = y_fused.data_ptr(), ya_fused.data_ptr()
        keep += [x, xin, w1, w2, b1, b2, xb, a, w1p, w2p, ta]
    _lib.check(lib.rvc_conv1d_tc(C.byref(d1), st), "conv1")
    _lib.check(lib.rvc_conv1d_tc(C.byref(d2), st), "conv2 sum")
    _lib.check(lib.rvc_resblock_pair_tc(C.byref(pd), st), "pair sum")
    torch.cuda.synchronize()
    assert float(y_ref.abs().max()) > 0.1
    assert torch.equal(y_fused, y_ref)
    assert torch.equal(ya_fused, ya_ref)


@pytest.mark.parametrize("Cin,Cout,k,u,Lin,sn", [(128, 64, 4, 2, 5001, 2), (64, 32, 4, 2, 7777, 1), (32, 16, 4, 2, 3000, 1)])
def test_fused_noise_conv_in_ups_epilogue(Cin, Cout, k, u, Lin, sn):
    """rvc_tc_job.nz_*: the short NSF noise conv computed in the ConvTranspose epilogue equals rvc_noise_conv_f32 + res, bit for bit."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    B, dtype = 2, FP16
    g = torch.Generator(device="cpu").manual_seed(77 + Cin)
    st = torch.cuda.current_stream().cuda_stream
    Lout = Lin * u
    kn = 1 if sn == 1 else 2 * sn
    pn = 0 if sn == 1 else (kn - sn) // 2
    x = torch.randn(B, Cin, Lin, generator=g).to(DEV)
    w = (torch.randn(Cin, Cout, k, generator=g) / (Cin * k / u) ** 0.5).to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    har = torch.randn(B, Lout * sn, generator=g).to(DEV)
    nzw, nzb = torch.randn(Cout, 1, kn, generator=g).to(DEV), torch.randn(Cout, generator=g).to(DEV)
    M = -(-k // u)
    wk = torch.zeros(Cout, Cin, u * M, device=DEV)
    for r in range(u):
        for t in range(M):
            kk = r + (M - 1 - t) * u
            if kk < k:
                wk[:, :, r * M + t] = w[:, :, kk].t()
    rows_in, rows_out = lib.rvc_packed_rows(Lin), lib.rvc_packed_rows(Lout)
    a = torch.zeros(B * Cin * rows_in * 2, dtype=torch.uint8, device=DEV)
    wp = torch.zeros(u * M * Cin * Cout * 2, dtype=torch.uint8, device=DEV)
    _lib.check(lib.rvc_pack_act(x.data_ptr(), None, None, 1.0, Cin * Lin, Lin, None, a.data_ptr(), B, Cin, Lin, dtype, 0.1, st), "pack_act")
    _lib.check(lib.rvc_pack_weight(wk.data_ptr(), wp.data_ptr(), Cout, Cin, u * M, dtype, st), "pack_weight")
    outs = []
    for fused in (False, True):
        yb = torch.zeros(B, Cout // 8, Lout, 8, device=DEV)
        y_act = torch.zeros(B * Cout * rows_out * 2, dtype=torch.uint8, device=DEV)
        d = _lib.Conv1dTcDesc()
        d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.up, d.pshift, d.f32_blocked = 1, B, Cin, Cout, Lin, dtype, u, (k - u) // 2, 3
        jb = d.jobs[0]
        jb.a, jb.w, jb.bias = a.data_ptr(), wp.data_ptr(), bias.data_ptr()
        jb.y_f32, jb.y_bs, jb.y_cs, jb.y_act = yb.data_ptr(), Cout * Lout, 8 * Lout, y_act.data_ptr()
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = M, 1, M - 1, 1.0, 0.1
        if fused:
            jb.nz_har, jb.nz_w, jb.nz_b, jb.nz_har_bs, jb.nz_kn, jb.nz_sn, jb.nz_pn = har.data_ptr(), nzw.data_ptr(), nzb.data_ptr(), Lout * sn, kn, sn, pn
        else:
            _lib.check(lib.rvc_noise_conv_f32(har.data_ptr(), Lout * sn, nzw.data_ptr(), nzb.data_ptr(), yb.data_ptr(), Cout * Lout, B, Cout, Lout, kn, sn, pn,
                                              1, st), "noise")
            jb.res, jb.res_bs, jb.res_cs = yb.data_ptr(), Cout * Lout, 8 * Lout
        _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "conv1d_tc(up)")
        torch.cuda.synchronize()
        outs.append((yb, y_act))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    ref = F.conv_transpose1d(rnd(F.leaky_relu(x, 0.1), dtype).double(), rnd(w, dtype).double(), bias.double(), stride=u, padding=(k - u) // 2) \
        + F.conv1d(har.unsqueeze(1).double(), nzw.double(), nzb.double(), stride=sn, padding=pn)
    assert float((from_blocked(outs[1][0]).double() - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("Cin,Cout,k,u,Lin,sn,B", [(128, 64, 4, 2, 5001, 2, 2), (64, 32, 4, 2, 7777, 1, 1), (128, 64, 4, 2, 1, 2, 1), (256, 32, 8, 4, 300, 0, 3)])
def test_conv_transpose_phase_columns_equal_phase_tiles(Cin, Cout, k, u, Lin, sn, B):
    """rvc_conv1d_tc_desc.up_cols: the u phases of a ConvTranspose as column blocks of ONE plain conv — every output (fp32, packed hi, lo plane,
    zero pad rows) bit-identical to the phase-tile mode (`up`), with and without the fused noise conv."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    dtype = FP16
    g = torch.Generator(device="cpu").manual_seed(5 + Cin + Lin)
    st = torch.cuda.current_stream().cuda_stream
    Lout = Lin * u
    fused = sn > 0
    kn = 0 if not fused else (1 if sn == 1 else 2 * sn)
    pn = 0 if sn <= 1 else (kn - sn) // 2
    x = torch.randn(B, Cin, Lin, generator=g).to(DEV)
    w = (torch.randn(Cin, Cout, k, generator=g) / (Cin * k / u) ** 0.5).to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    har = torch.randn(B, Lout * max(sn, 1), generator=g).to(DEV)
    nzw, nzb = torch.randn(Cout, 1, max(kn, 1), generator=g).to(DEV), torch.randn(Cout, generator=g).to(DEV)
    M = -(-k // u)
    wk = torch.zeros(Cout, Cin, u * M, device=DEV)
    for r in range(u):
        for t in range(M):
            kk = r + (M - 1 - t) * u
            if kk < k:
                wk[:, :, r * M + t] = w[:, :, kk].t()
    wm = wk.view(Cout, Cin, u, M).permute(2, 0, 1, 3).reshape(u * Cout, Cin, M).contiguous()      # phase columns: [u * Cout][Cin][M]
    rows_in, rows_out = lib.rvc_packed_rows(Lin), lib.rvc_packed_rows(Lout)
    a = torch.zeros(B * Cin * rows_in * 2, dtype=torch.uint8, device=DEV)
    wp, wmp = (torch.zeros(u * M * Cin * Cout * 2, dtype=torch.uint8, device=DEV) for _ in range(2))
    _lib.check(lib.rvc_pack_act(x.data_ptr(), None, None, 1.0, Cin * Lin, Lin, None, a.data_ptr(), B, Cin, Lin, dtype, 0.1, st), "pack_act")
    _lib.check(lib.rvc_pack_weight(wk.data_ptr(), wp.data_ptr(), Cout, Cin, u * M, dtype, st), "pack_weight")
    _lib.check(lib.rvc_pack_weight(wm.data_ptr(), wmp.data_ptr(), u * Cout, Cin, M, dtype, st), "pack_weight")
    bias_m = bias.repeat(u)
    outs = []
    for cols in (False, True):
        yb = torch.zeros(B, Cout // 8, Lout, 8, device=DEV)
        yh, yl = (torch.zeros(B * Cout * rows_out * 2, dtype=torch.uint8, device=DEV) for _ in range(2))
        d = _lib.Conv1dTcDesc()
        d.njobs, d.B, d.Cin, d.L, d.dtype, d.pshift, d.f32_blocked = 1, B, Cin, Lin, dtype, (k - u) // 2, 3
        if cols:
            d.Cout, d.up_cols = u * Cout, u
        else:
            d.Cout, d.up = Cout, u
        jb = d.jobs[0]
        jb.a, jb.w, jb.bias = a.data_ptr(), (wmp if cols else wp).data_ptr(), (bias_m if cols else bias).data_ptr()
        jb.y_f32, jb.y_bs, jb.y_cs, jb.y_act, jb.y_lo = yb.data_ptr(), Cout * Lout, 8 * Lout, yh.data_ptr(), yl.data_ptr()
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = M, 1, M - 1, 1.0, 0.1
        if fused:
            jb.nz_har, jb.nz_w, jb.nz_b, jb.nz_har_bs, jb.nz_kn, jb.nz_sn, jb.nz_pn = har.data_ptr(), nzw.data_ptr(), nzb.data_ptr(), Lout * sn, kn, sn, pn
        _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "conv1d_tc(up_cols)" if cols else "conv1d_tc(up)")
        torch.cuda.synchronize()
        outs.append((yb, yh, yl))
    if not torch.equal(outs[0][0], outs[1][0]):        # diagnostics: where the two modes differ
        d_ = (from_blocked(outs[0][0]) - from_blocked(outs[1][0])).abs()
        bad = torch.nonzero(d_ > 0)
        raise AssertionError(f"fp32 outputs differ at {bad.shape[0]} of {d_.numel()} elements, max {float(d_.max()):.3e}; first {bad[:6].tolist()}, "
                             f"last {bad[-3:].tolist()}; channels {sorted(set(bad[:, 1].tolist()))[:12]}; positions mod {u}: {sorted(set((bad[:, 2] % u).tolist()))}")
    for x_, y_ in zip(outs[0], outs[1]):
        assert torch.equal(x_, y_)
    assert pad_rows_zero(outs[1][1], B, Cout, Lout, FP16, rows_out)
    ref = F.conv_transpose1d(rnd(F.leaky_relu(x, 0.1), dtype).double(), rnd(w, dtype).double(), bias.double(), stride=u, padding=(k - u) // 2)
    if fused:
        ref = ref + F.conv1d(har.unsqueeze(1).double(), nzw.double(), nzb.double(), stride=sn, padding=pn)
    assert float((from_blocked(outs[1][0]).double() - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))
    # validation: residual inputs are refused, Cr must be a power of two >= 32
    d.jobs[0].nz_har = None
    d.jobs[0].res = outs[1][0].data_ptr()
    assert lib.rvc_conv1d_tc(C.byref(d), st) != 0 and b"up_cols" in lib.rvc_last_error()


# ------------------------------------------------------------------------------------------------ split-fp16 residual stream
def pack_host(t, rows):
    """fp32 [B, C, L] -> packed fp16 plane [B][C/8][rows][8] (position l at row 64 + l, other rows zero) as uint8 on the device."""
    B, C_, L = t.shape
    out = torch.zeros(B, C_ // 8, rows, 8, dtype=torch.float16)
    out[:, :, 64:64 + L, :] = t.view(B, C_ // 8, 8, L).permute(0, 1, 3, 2).to(torch.float16)
    return out.view(torch.uint8).reshape(-1).to(DEV)


def split_planes(x, rows, slope=0.1):
    """the stream's two planes for a residual value x: hi = fp16(lrelu(x)), lo = fp16(lrelu(x) - hi); and the x they decode to"""
    y = F.leaky_relu(x, slope)
    hi = y.to(torch.float16).float()
    lo = (y - hi).to(torch.float16).float()
    s = hi + lo
    return pack_host(hi, rows), pack_host(lo, rows), torch.minimum(s, s * (1.0 / torch.tensor(slope, dtype=torch.float32)))


def decode_split(hi_buf, lo_buf, B, C_, L, rows, slope=0.1):
    s = unpack(hi_buf, B, C_, L, FP16, rows) + unpack(lo_buf, B, C_, L, FP16, rows)
    return torch.minimum(s, s * (1.0 / torch.tensor(slope, dtype=torch.float32, device=s.device)))


@pytest.mark.parametrize("Cc,L,B", [(64, 3001, 2), (32, 20000, 1), (64, 246, 1), (32, 1, 1), (32, 777, 3)])
def test_fused_resblock_pair_split_stream(Cc, L, B):
    """Split-fp16 residual stream (rvc_pair_job.a_lo / y_lo): same step as the fp32-stream launch on the value the two planes decode to.
    conv results are bit-identical (same operand plane), so the two x' differ only by the output split: <= 2^-21 |lrelu(x')| + 1e-7."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(77 + Cc + L)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    n = B * Cc * rows * 2
    pd_s, pd_f = _lib.ResblockPairDesc(), _lib.ResblockPairDesc()
    for pd in (pd_s, pd_f):
        pd.njobs, pd.B, pd.C, pd.L, pd.dtype, pd.slope = 3, B, Cc, L, FP16, 0.1
    keep, outs = [], []
    for j, (k, dil) in enumerate(((11, 5), (7, 3), (3, 1))):
        x = 2.0 * torch.randn(B, Cc, L, generator=g)
        hi, lo, xdec = split_planes(x, rows)
        w1 = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).to(DEV)
        w2 = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).to(DEV)
        b1, b2 = torch.randn(Cc, generator=g).to(DEV), torch.randn(Cc, generator=g).to(DEV)
        w1p, w2p = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=DEV), torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
        _lib.check(lib.rvc_pack_weight(w1.data_ptr(), w1p.data_ptr(), Cc, Cc, k, FP16, st), "pack_weight")
        _lib.check(lib.rvc_pack_weight(w2.data_ptr(), w2p.data_ptr(), Cc, Cc, k, FP16, st), "pack_weight")
        xb = to_blocked(xdec.to(DEV))
        y_f, ya_f = torch.zeros_like(xb), torch.zeros(n, dtype=torch.uint8, device=DEV)
        yh, yl = torch.zeros(n, dtype=torch.uint8, device=DEV), lo.clone()     # the lo plane is updated in place
        for pd, split in ((pd_s, True), (pd_f, False)):
            pj = pd.jobs[j]
            pj.a, pj.w1, pj.b1, pj.w2, pj.b2 = hi.data_ptr(), w1p.data_ptr(), b1.data_ptr(), w2p.data_ptr(), b2.data_ptr()
            pj.K1, pj.dil1, pj.K2 = k, dil, k
            if split:
                pj.a_lo, pj.y_act, pj.y_lo = yl.data_ptr(), yh.data_ptr(), yl.data_ptr()
            else:
                pj.res, pj.y, pj.y_act = xb.data_ptr(), y_f.data_ptr(), ya_f.data_ptr()
        keep += [hi, lo, w1, w2, b1, b2, w1p, w2p, xb]
        outs.append((y_f, ya_f, yh, yl))
    _lib.check(lib.rvc_resblock_pair_tc(C.byref(pd_f), st), "pair fp32 stream")
    _lib.check(lib.rvc_resblock_pair_tc(C.byref(pd_s), st), "pair split stream")
    torch.cuda.synchronize()
    for y_f, ya_f, yh, yl in outs:
        assert torch.equal(yh, ya_f)                                   # hi plane == the packed operand of the fp32-stream launch
        assert pad_rows_zero(yh, B, Cc, L, FP16, rows)
        xs, xf = decode_split(yh, yl, B, Cc, L, rows), from_blocked(y_f)
        bound = 2.0 ** -21 * F.leaky_relu(xf, 0.1).abs() * 10.0 + 1e-6      # (x 10: the inverse of slope 0.1 on the negative side)
        assert bool(((xs - xf).abs() <= bound).all()), float(((xs - xf).abs() - bound).max())


@pytest.mark.parametrize("Cc,L,B,y_ncl", [(64, 3001, 2, 1), (32, 20000, 1, 0), (32, 5, 2, 1)])
def test_fused_resblock_pair_split_stream_stage_average(Cc, L, B, y_ncl):
    """sum_jobs with split-fp16 inputs: the three residuals decode from (a, a_lo); bit-identical to the fp32-stream launch fed the decoded values."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(91 + Cc + L)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    n = B * Cc * rows * 2
    pd_s, pd_f = _lib.ResblockPairDesc(), _lib.ResblockPairDesc()
    for pd in (pd_s, pd_f):
        pd.njobs, pd.B, pd.C, pd.L, pd.dtype, pd.slope, pd.sum_jobs, pd.y_ncl, pd.alpha = 3, B, Cc, L, FP16, 0.1, 1, y_ncl, 1.0 / 3.0
    shape = (B, Cc, L) if y_ncl else (B, Cc // 8, L, 8)
    y_s, y_f = torch.zeros(shape, device=DEV), torch.zeros(shape, device=DEV)
    ya_s, ya_f = torch.zeros(n, dtype=torch.uint8, device=DEV), torch.zeros(n, dtype=torch.uint8, device=DEV)
    keep = []
    for j, (k, dil) in enumerate(((11, 5), (7, 5), (3, 5))):
        hi, lo, xdec = split_planes(torch.randn(B, Cc, L, generator=g), rows)
        w1 = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).to(DEV)
        w2 = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).to(DEV)
        b1, b2 = torch.randn(Cc, generator=g).to(DEV), torch.randn(Cc, generator=g).to(DEV)
        w1p, w2p = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=DEV), torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
        _lib.check(lib.rvc_pack_weight(w1.data_ptr(), w1p.data_ptr(), Cc, Cc, k, FP16, st), "pack_weight")
        _lib.check(lib.rvc_pack_weight(w2.data_ptr(), w2p.data_ptr(), Cc, Cc, k, FP16, st), "pack_weight")
        xb = to_blocked(xdec.to(DEV))
        for pd, split in ((pd_s, True), (pd_f, False)):
            pj = pd.jobs[j]
            pj.a, pj.w1, pj.b1, pj.w2, pj.b2 = hi.data_ptr(), w1p.data_ptr(), b1.data_ptr(), w2p.data_ptr(), b2.data_ptr()
            pj.K1, pj.dil1, pj.K2 = k, dil, k
            if split:
                pj.a_lo = lo.data_ptr()
            else:
                pj.res = xb.data_ptr()
            if j == 0:
                pj.y, pj.y_act = (y_s if split else y_f).data_ptr(), (ya_s if split else ya_f).data_ptr()
        keep += [hi, lo, w1, w2, b1, b2, w1p, w2p, xb]
    _lib.check(lib.rvc_resblock_pair_tc(C.byref(pd_f), st), "pair sum fp32 stream")
    _lib.check(lib.rvc_resblock_pair_tc(C.byref(pd_s), st), "pair sum split stream")
    torch.cuda.synchronize()
    assert float(y_f.abs().max()) > 0.1
    assert torch.equal(y_s, y_f) and torch.equal(ya_s, ya_f)


def test_tc_conv_split_stream_output():
    """rvc_tc_job.y_lo: the low plane next to y_act — hi + lo reproduces leaky_relu(y_f32) to 2^-21 relative (+ the fp16 subnormal step)."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    B, Cc, L, k = 2, 64, 3001, 3
    g = torch.Generator(device="cpu").manual_seed(5)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    n = B * Cc * rows * 2
    x = torch.randn(B, Cc, L, generator=g).to(DEV)
    w = (3.0 * torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).to(DEV)
    bias = torch.randn(Cc, generator=g).to(DEV)
    a, wp = torch.zeros(n, dtype=torch.uint8, device=DEV), torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
    _lib.check(lib.rvc_pack_act(x.data_ptr(), None, None, 1.0, Cc * L, L, None, a.data_ptr(), B, Cc, L, FP16, 1.0, st), "pack_act")
    _lib.check(lib.rvc_pack_weight(w.data_ptr(), wp.data_ptr(), Cc, Cc, k, FP16, st), "pack_weight")
    y, yh, yl = torch.zeros(B, Cc, L, device=DEV), torch.zeros(n, dtype=torch.uint8, device=DEV), torch.zeros(n, dtype=torch.uint8, device=DEV)
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype = 1, B, Cc, Cc, L, FP16
    jb = d.jobs[0]
    jb.a, jb.w, jb.bias, jb.y_act, jb.y_lo = a.data_ptr(), wp.data_ptr(), bias.data_ptr(), yh.data_ptr(), yl.data_ptr()
    jb.y_f32, jb.y_bs, jb.y_cs = y.data_ptr(), Cc * L, L
    jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = k, 1, 1, 1.0, 0.1
    _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "conv")
    torch.cuda.synchronize()
    ref = F.leaky_relu(y, 0.1)
    hi, lo = unpack(yh, B, Cc, L, FP16, rows), unpack(yl, B, Cc, L, FP16, rows)
    assert torch.equal(hi, ref.to(torch.float16).float())
    assert bool(((hi + lo - ref).abs() <= 2.0 ** -21 * ref.abs() + 6e-8).all())
    assert pad_rows_zero(yl, B, Cc, L, FP16, rows)
    d.dtype = BF16
    assert lib.rvc_conv1d_tc(C.byref(d), st) != 0 and b"y_lo" in lib.rvc_last_error()


@pytest.mark.parametrize("Cc,L,B,sum_jobs", [(128, 3001, 2, 0), (256, 700, 1, 0), (128, 257, 1, 1), (256, 2100, 2, 1), (64, 1000, 1, 0), (64, 333, 2, 1)])
def test_tc_conv_split_stream_residual(Cc, L, B, sum_jobs):
    """rvc_tc_job.res_hi / res_lo: conv2 of a wide ResBlock step with its residual as the split-fp16 stream — the same launch as with the fp32
    residual the two planes decode to.  One job per tile: y_act bit-identical, decode(y_act, y_lo) within the output split of the fp32 x';
    summed group (stage average): every output bit-identical.  The lo plane is updated in place."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(123 + Cc + L)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    n = B * Cc * rows * 2
    ds, df = _lib.Conv1dTcDesc(), _lib.Conv1dTcDesc()
    for d in (ds, df):
        d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.sum_jobs = 3, B, Cc, Cc, L, FP16, sum_jobs
        d.f32_blocked = 1 if sum_jobs else 3
    keep, outs = [], []
    avg_s, avg_f = torch.zeros(B, Cc, L, device=DEV), torch.zeros(B, Cc, L, device=DEV)
    ua_s, ua_f = torch.zeros(n, dtype=torch.uint8, device=DEV), torch.zeros(n, dtype=torch.uint8, device=DEV)
    for j, k in enumerate((11, 7, 3)):
        hi, lo, xdec = split_planes(2.0 * torch.randn(B, Cc, L, generator=g), rows)
        t = torch.randn(B, Cc, L, generator=g).to(DEV)
        w = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).to(DEV)
        bias = torch.randn(Cc, generator=g).to(DEV)
        a, wp = torch.zeros(n, dtype=torch.uint8, device=DEV), torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
        _lib.check(lib.rvc_pack_act(t.data_ptr(), None, None, 1.0, Cc * L, L, None, a.data_ptr(), B, Cc, L, FP16, 0.1, st), "pack_act")
        _lib.check(lib.rvc_pack_weight(w.data_ptr(), wp.data_ptr(), Cc, Cc, k, FP16, st), "pack_weight")
        xb = to_blocked(xdec.to(DEV))
        y_f, ya_f = torch.zeros_like(xb), torch.zeros(n, dtype=torch.uint8, device=DEV)
        yh, yl = hi.clone(), lo.clone()                       # both planes are updated in place
        for d, split in ((ds, True), (df, False)):
            jb = d.jobs[j]
            jb.a, jb.w, jb.bias = a.data_ptr(), wp.data_ptr(), bias.data_ptr()
            jb.K, jb.dil, jb.pad, jb.act_slope, jb.alpha = k, 1, (k - 1) // 2, 0.1, (1.0 / 3.0 if sum_jobs else 1.0)
            if split:
                jb.res_hi, jb.res_lo = yh.data_ptr(), yl.data_ptr()
            else:
                jb.res, jb.res_bs, jb.res_cs = xb.data_ptr(), Cc * L, 8 * L
            if not sum_jobs:
                if split:
                    jb.y_act, jb.y_lo = yh.data_ptr(), yl.data_ptr()
                else:
                    jb.y_f32, jb.y_bs, jb.y_cs, jb.y_act = y_f.data_ptr(), Cc * L, 8 * L, ya_f.data_ptr()
            elif j == 0:
                jb.y_f32, jb.y_bs, jb.y_cs, jb.y_act = (avg_s if split else avg_f).data_ptr(), Cc * L, L, (ua_s if split else ua_f).data_ptr()
        keep += [hi, lo, a, wp, bias, xb, t, w]
        outs.append((y_f, ya_f, yh, yl))
    _lib.check(lib.rvc_conv1d_tc(C.byref(df), st), "conv2 fp32 residual")
    _lib.check(lib.rvc_conv1d_tc(C.byref(ds), st), "conv2 split residual")
    torch.cuda.synchronize()
    if sum_jobs:
        assert float(avg_f.abs().max()) > 0.1
        assert torch.equal(avg_s, avg_f) and torch.equal(ua_s, ua_f)
        return
    for y_f, ya_f, yh, yl in outs:
        assert torch.equal(yh, ya_f)
        assert pad_rows_zero(yh, B, Cc, L, FP16, rows)
        xs, xf = decode_split(yh, yl, B, Cc, L, rows), from_blocked(y_f)
        bound = 2.0 ** -21 * F.leaky_relu(xf, 0.1).abs() * 10.0 + 1e-6
        assert bool(((xs - xf).abs() <= bound).all()), float(((xs - xf).abs() - bound).max())


def test_tc_conv_split_stream_residual_is_validated():
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    Cc, L = 64, 300
    n = Cc * lib.rvc_packed_rows(L) * 2
    buf = [torch.zeros(n, dtype=torch.uint8, device=DEV) for _ in range(4)]
    wp = torch.zeros(3 * Cc * Cc * 2, dtype=torch.uint8, device=DEV)
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype = 1, 1, Cc, Cc, L, FP16
    jb = d.jobs[0]
    jb.a, jb.w, jb.y_act, jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = buf[0].data_ptr(), wp.data_ptr(), buf[1].data_ptr(), 3, 1, 1, 1.0, 0.1
    jb.res_hi = buf[2].data_ptr()                       # lo plane missing
    assert lib.rvc_conv1d_tc(C.byref(d), st) != 0 and b"res_hi" in lib.rvc_last_error()
    jb.res_lo = buf[3].data_ptr()
    jb.act_slope = 0.0                                  # not invertible
    assert lib.rvc_conv1d_tc(C.byref(d), st) != 0 and b"res_hi" in lib.rvc_last_error()
    jb.act_slope = 0.1
    d.dtype = BF16
    assert lib.rvc_conv1d_tc(C.byref(d), st) != 0 and b"res_hi" in lib.rvc_last_error()
    d.dtype = FP16
    _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "valid split residual launch")
    torch.cuda.synchronize()


@pytest.mark.parametrize("Cin,Cout,L,B", [(192, 576, 1000, 2), (96, 256, 300, 3), (768, 192, 257, 1)])
def test_split_precision_gemm_job_groups(Cin, Cout, L, B):
    """sum_jobs = 3 over six jobs: two independent GEMMs, each as a_hi.W_hi + a_lo.W_hi + a_hi.W_lo in one accumulator (the encoder's
    split-precision form).  Against fp64 on the UNROUNDED operands: <= 8e-6 * max|ref| — fp32 accumulation over up to 3 x 768 products —
    (a plain fp16 launch is >= 30x worse), and the
    [hi | lo] output planes reproduce the fp32 result to 2^-21."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(11 + Cin + L)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    cb = rows * 16
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.sum_jobs = 6, B, Cin, Cout, L, FP16, 3
    dp = _lib.Conv1dTcDesc()
    dp.njobs, dp.B, dp.Cin, dp.Cout, dp.L, dp.dtype = 2, B, Cin, Cout, L, FP16
    keep, outs = [], []
    for grp in range(2):
        x = torch.randn(B, Cin, L, generator=g)
        w = torch.randn(Cout, Cin, 1, generator=g) / Cin ** 0.5
        bias = torch.randn(Cout, generator=g).to(DEV)
        x_hi = x.half().float()
        a = pack_host(torch.cat([x_hi, x - x_hi], 1), rows)                 # [hi | lo] planes, 2 * Cin / 8 chunks per item
        w_hi = w.half().float()
        wh, wl = torch.zeros(Cin * Cout * 2, dtype=torch.uint8, device=DEV), torch.zeros(Cin * Cout * 2, dtype=torch.uint8, device=DEV)
        wd, wld = w.to(DEV), (w - w_hi).to(DEV)
        _lib.check(lib.rvc_pack_weight(wd.data_ptr(), wh.data_ptr(), Cout, Cin, 1, FP16, st), "pack_weight")
        _lib.check(lib.rvc_pack_weight(wld.data_ptr(), wl.data_ptr(), Cout, Cin, 1, FP16, st), "pack_weight")
        y, yp = torch.zeros(B, Cout, L, device=DEV), torch.zeros(B, Cout, L, device=DEV)
        ya = torch.zeros(B * 2 * (Cout // 8) * cb, dtype=torch.uint8, device=DEV)
        a_bs = 2 * (Cin // 8) * cb
        for t_i, (ao, wp) in enumerate(((0, wh), ((Cin // 8) * cb, wh), (0, wl))):
            jb = d.jobs[3 * grp + t_i]
            jb.a, jb.w, jb.a_bs = a.data_ptr() + ao, wp.data_ptr(), a_bs
            jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = 1, 1, 0, 1.0, 1.0
            if t_i == 0:
                jb.bias, jb.y_f32, jb.y_bs, jb.y_cs = bias.data_ptr(), y.data_ptr(), Cout * L, L
                jb.y_act, jb.y_act_bs, jb.y_lo = ya.data_ptr(), 2 * (Cout // 8) * cb, ya.data_ptr() + (Cout // 8) * cb
        jb = dp.jobs[grp]
        jb.a, jb.w, jb.a_bs, jb.bias = a.data_ptr(), wh.data_ptr(), a_bs, bias.data_ptr()
        jb.y_f32, jb.y_bs, jb.y_cs = yp.data_ptr(), Cout * L, L
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = 1, 1, 0, 1.0, 1.0
        keep += [a, wh, wl, wd, wld, bias]
        outs.append((x, w, bias, y, yp, ya))
    _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "split gemm")
    _lib.check(lib.rvc_conv1d_tc(C.byref(dp), st), "plain gemm")
    torch.cuda.synchronize()
    for x, w, bias, y, yp, ya in outs:
        ref = F.conv1d(x.double(), w.double(), bias.double().cpu())
        scale = float(ref.abs().max())
        e_split, e_plain = float((y.cpu().double() - ref).abs().max()), float((yp.cpu().double() - ref).abs().max())
        assert e_split <= 8e-6 * scale, (e_split, scale)
        assert e_plain >= 30 * e_split, (e_plain, e_split)
        planes = ya.view(torch.float16).view(B, 2 * (Cout // 8), rows, 8)[:, :, 64:64 + L, :].permute(0, 1, 3, 2).reshape(B, 2 * Cout, L).float()
        assert torch.equal(planes[:, :Cout], y.half().float())
        assert bool(((planes[:, :Cout] + planes[:, Cout:] - y).abs() <= 2.0 ** -21 * y.abs() + 6e-8).all())


def test_layernorm_and_pack_rows_split_planes():
    """RVC_DT_FP16_SPLIT outputs of rvc_layernorm_f32 / rvc_pack_rows: [hi | lo] planes, hi identical to the plain fp16 output."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    B, H, T, D = 2, 192, 333, 768
    g = torch.Generator(device="cpu").manual_seed(3)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(T)
    x, r = torch.randn(B, H, T, generator=g).to(DEV), torch.randn(B, H, T, generator=g).to(DEV)
    gamma, beta = torch.randn(H, generator=g).to(DEV), torch.randn(H, generator=g).to(DEV)
    mask = (torch.arange(T)[None, :] < torch.tensor([T, T - 40])[:, None]).float().to(DEV)
    y1, y2 = torch.zeros_like(x), torch.zeros_like(x)
    p1 = torch.zeros(B * (H // 8) * rows * 16, dtype=torch.uint8, device=DEV)
    p2 = torch.zeros(B * 2 * (H // 8) * rows * 16, dtype=torch.uint8, device=DEV)
    _lib.check(lib.rvc_layernorm_f32(x.data_ptr(), r.data_ptr(), None, None, gamma.data_ptr(), beta.data_ptr(), mask.data_ptr(), y1.data_ptr(), B, H, T, 1e-5, 0,
                                     p1.data_ptr(), FP16, st), "ln")
    _lib.check(lib.rvc_layernorm_f32(x.data_ptr(), r.data_ptr(), None, None, gamma.data_ptr(), beta.data_ptr(), mask.data_ptr(), y2.data_ptr(), B, H, T, 1e-5, 0,
                                     p2.data_ptr(), 3, st), "ln split")
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    hi1 = unpack(p1, B, H, T, FP16, rows)
    pl = p2.view(torch.float16).view(B, 2 * (H // 8), rows, 8)[:, :, 64:64 + T, :].permute(0, 1, 3, 2).reshape(B, 2 * H, T).float()
    ym = y1 * mask[:, None, :]
    assert torch.equal(pl[:, :H], hi1)
    assert bool(((pl[:, :H] + pl[:, H:] - ym).abs() <= 2.0 ** -21 * ym.abs() + 6e-8).all())
    ph = torch.randn(B, T, D, generator=g).to(DEV)
    q1 = torch.zeros(B * (D // 8) * rows * 16, dtype=torch.uint8, device=DEV)
    q2 = torch.zeros(B * 2 * (D // 8) * rows * 16, dtype=torch.uint8, device=DEV)
    _lib.check(lib.rvc_pack_rows(ph.data_ptr(), q1.data_ptr(), B, T, D, FP16, st), "pack_rows")
    _lib.check(lib.rvc_pack_rows(ph.data_ptr(), q2.data_ptr(), B, T, D, 3, st), "pack_rows split")
    torch.cuda.synchronize()
    pl = q2.view(torch.float16).view(B, 2 * (D // 8), rows, 8)[:, :, 64:64 + T, :].permute(0, 1, 3, 2).reshape(B, 2 * D, T).float()
    pt = ph.transpose(1, 2)
    assert torch.equal(pl[:, :D], unpack(q1, B, D, T, FP16, rows))
    assert bool(((pl[:, :D] + pl[:, D:] - pt).abs() <= 2.0 ** -21 * pt.abs() + 6e-8).all())


@pytest.mark.parametrize("B,T", [(1, 1000), (2, 130), (3, 64)])
def test_conv_chain_equals_separate_launches(B, T):
    """rvc_conv1d_tc_chain_*: a WaveNet-like chain (k = 5 gated conv -> res / skip 1x1 pair -> 1x1 into 96 channels from zero-padded
    weights, twice over, every phase reading what the previous one wrote, in-place residual updates) as ONE persistent kernel with grid
    barriers — bit-identical to the same descriptors launched one by one through rvc_conv1d_tc, replayed three times (the barrier
    counters must come back to zero)."""
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    H, half = 192, 96
    g = torch.Generator(device="cpu").manual_seed(1000 * B + T)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(T)
    keep = []

    def wpack(cout, cin, k, pad_to=None, perm=None):
        w = torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5
        if perm is not None:
            w = w[perm]
        if pad_to:
            w = torch.cat([w, torch.zeros(pad_to - cout, cin, k)], 0)
        wd = w.to(DEV).contiguous()
        out = torch.zeros(wd.numel() * 2, dtype=torch.uint8, device=DEV)
        _lib.check(lib.rvc_pack_weight(wd.data_ptr(), out.data_ptr(), wd.shape[0], cin, k, FP16, st), "pack_weight")
        keep.append(out)
        return out

    vec = lambda n: torch.randn(n, generator=g).to(DEV)
    mask = (torch.arange(T)[None, :] < torch.randint(T // 2, T + 1, (B,), generator=g)[:, None]).float().to(DEV)
    perm = gate_perm(H, 64)
    layers = [dict(w_in=wpack(2 * H, H, 5, perm=perm), b_in=vec(2 * H), w_res=wpack(H, H, 1), b_res=vec(H), w_skip=wpack(H, H, 1), b_skip=vec(H),
                   w_post=wpack(half, H, 1, pad_to=128), b_post=vec(half)) for _ in range(2)]
    x0 = torch.randn(B, H, T, generator=g)
    z0 = torch.randn(B, half, T, generator=g)

    def state():
        h = (x0 * mask.cpu()[:, None, :]).to(DEV).contiguous()
        hp = pack_host(x0 * mask.cpu()[:, None, :], rows)
        return dict(h=h, hp=hp, acts=torch.zeros(B * H * rows * 2, dtype=torch.uint8, device=DEV), skip=torch.zeros(B, H, T, device=DEV),
                    z=z0.to(DEV).contiguous(), za=torch.zeros(B * half * rows * 2, dtype=torch.uint8, device=DEV))

    def descs(s, chain):
        out, cvs = [], []
        for li, ly in enumerate(layers):
            d = _lib.Conv1dTcDesc()
            d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.nt, d.mt = 1, B, H, 2 * H, T, FP16, 64, 1
            jb = d.jobs[0]
            jb.a, jb.w, jb.bias, jb.y_act = s["hp"].data_ptr(), ly["w_in"].data_ptr(), ly["b_in"].data_ptr(), s["acts"].data_ptr()
            jb.K, jb.dil, jb.pad, jb.gated, jb.alpha, jb.act_slope = 5, 1, 2, 1, 1.0, 1.0
            out.append(d); cvs.append(0)
            d = _lib.Conv1dTcDesc()
            d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.nt, d.mt = 2, B, H, H, T, FP16, 64, 1
            jb = d.jobs[0]     # h <- (h + res(acts)) * mask, packed copy for the next layer
            jb.a, jb.w, jb.bias, jb.mask = s["acts"].data_ptr(), ly["w_res"].data_ptr(), ly["b_res"].data_ptr(), mask.data_ptr()
            jb.res, jb.res_bs, jb.res_cs, jb.y_f32, jb.y_bs, jb.y_cs, jb.y_act = s["h"].data_ptr(), H * T, T, s["h"].data_ptr(), H * T, T, s["hp"].data_ptr()
            jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = 1, 1, 0, 1.0, 1.0
            jb = d.jobs[1]     # skip += skip(acts)
            jb.a, jb.w, jb.bias = s["acts"].data_ptr(), ly["w_skip"].data_ptr(), ly["b_skip"].data_ptr()
            jb.res, jb.res_bs, jb.res_cs, jb.y_f32, jb.y_bs, jb.y_cs = s["skip"].data_ptr(), H * T, T, s["skip"].data_ptr(), H * T, T
            jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = 1, 1, 0, 1.0, 1.0
            out.append(d); cvs.append(0)
            d = _lib.Conv1dTcDesc()    # z <- (z - post(hp)) * mask   (96 real channels, weights padded to 128 rows)
            d.njobs, d.B, d.Cin, d.L, d.dtype = 1, B, H, T, FP16
            jb = d.jobs[0]
            jb.a, jb.w, jb.bias, jb.mask = s["hp"].data_ptr(), ly["w_post"].data_ptr(), ly["b_post"].data_ptr(), mask.data_ptr()
            jb.res, jb.res_bs, jb.res_cs, jb.y_f32, jb.y_bs, jb.y_cs, jb.y_act = s["z"].data_ptr(), half * T, T, s["z"].data_ptr(), half * T, T, s["za"].data_ptr()
            jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = 1, 1, 0, -1.0, 0.1
            if chain:
                d.Cout, d.nt, d.mt = 128, 64, 1
                cvs.append(half)
            else:
                d.Cout, d.nt, d.mt, jb.w_cs = half, 32, 1, 128
                cvs.append(0)
            out.append(d)
        return out, cvs

    ref = state()
    for d in descs(ref, False)[0]:
        _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "separate launch")
    torch.cuda.synchronize()
    assert float(ref["h"].abs().max()) > 0.5 and float(ref["skip"].abs().max()) > 0.5
    sync = torch.zeros(4, dtype=torch.int32, device=DEV)
    for rep in range(3):
        cand = state()
        dl, cvs = descs(cand, True)
        n = len(dl)
        arr, cv = (_lib.Conv1dTcDesc * n)(*dl), (C.c_int32 * n)(*cvs)
        nbytes = lib.rvc_conv1d_tc_chain_blob_bytes(n)
        blob = (C.c_uint8 * nbytes)()
        _lib.check(lib.rvc_conv1d_tc_chain_build(arr, cv, n, C.addressof(blob), nbytes), "chain build")
        dev_blob = torch.frombuffer(blob, dtype=torch.uint8).to(DEV)
        _lib.check(lib.rvc_conv1d_tc_chain_run(C.addressof(blob), dev_blob.data_ptr(), sync.data_ptr(), st), "chain run")
        torch.cuda.synchronize()
        assert sync.tolist() == [0, 0, 0, 0], (rep, sync.tolist())      # counters reset, no barrier timed out
        for k in ("h", "hp", "acts", "skip", "z", "za"):
            assert torch.equal(cand[k], ref[k]), (rep, k)
