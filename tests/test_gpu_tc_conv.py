"""Op-level parity of the tcgen05 ResBlock conv (rvc_conv1d_tc) through the C ABI.

Reference = fp64 torch conv1d on operands rounded exactly as the kernel rounds them (bf16 RN / tf32 RNA), so the only
difference left is fp32 accumulation order: tolerance 2e-4 * max|ref| (K*C up to 2816 products per output).
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rnd(x, dtype):
    if dtype == 0:
        return x.to(torch.bfloat16).to(torch.float32)
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def unpack(buf, B, C, L, dtype, rows):
    CK = 8 if dtype == 0 else 4
    t = buf.view(torch.bfloat16 if dtype == 0 else torch.float32).view(B, C // CK, rows, CK)
    return t[:, :, 64:64 + L, :].permute(0, 1, 3, 2).reshape(B, C, L).float()


def run_case(dtype, Cc, K, dil, B, L, seed=0, with_res=True, njobs=1):
    from rvc_maker_b200 import _lib

    lib = _lib.load()
    dev = "cuda:0"
    g = torch.Generator(device="cpu").manual_seed(seed)
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    CK = 8 if dtype == 0 else 4
    esz = 2 if dtype == 0 else 4
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.C, d.L, d.dtype = njobs, B, Cc, L, dtype
    keep, checks = [], []
    for j in range(njobs):
        Kj, dj = (K, dil) if j == 0 else ((3, 1) if j == 1 else (7, 3))
        x = torch.randn(B, Cc, L, generator=g).to(dev)
        w = (torch.randn(Cc, Cc, Kj, generator=g) / (Cc * Kj) ** 0.5).to(dev)
        bias = torch.randn(Cc, generator=g).to(dev)
        res = torch.randn(B, Cc, L, generator=g).to(dev) if with_res else None
        a = torch.zeros(B * (Cc // CK) * rows * CK * esz, dtype=torch.uint8, device=dev)
        wp = torch.zeros(Kj * Cc * Cc * esz, dtype=torch.uint8, device=dev)
        y_act = torch.zeros_like(a)
        y = torch.empty(B, Cc, L, device=dev)
        acc = torch.randn(B, Cc, L, generator=g).to(dev)
        acc0 = acc.clone()
        _lib.check(lib.rvc_pack_act(x.data_ptr(), Cc * L, L, a.data_ptr(), B, Cc, L, dtype, 0.1, st), "pack_act")
        _lib.check(lib.rvc_pack_weight(w.data_ptr(), wp.data_ptr(), Cc, Cc, Kj, dtype, st), "pack_weight")
        jb = d.jobs[j]
        jb.a, jb.w, jb.bias = a.data_ptr(), wp.data_ptr(), bias.data_ptr()
        jb.res = res.data_ptr() if with_res else None
        jb.y_f32, jb.y_act, jb.acc_out = y.data_ptr(), y_act.data_ptr(), acc.data_ptr()
        jb.K, jb.dil, jb.act_slope, jb.out_scale, jb.beta = Kj, dj, 0.1, 1.0 / 3.0, 1.0
        keep += [x, w, bias, res, a, wp, y_act, y, acc]
        checks.append((x, w, bias, res, a, y_act, y, acc, acc0, Kj, dj))
    _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "conv1d_tc")
    torch.cuda.synchronize()
    for x, w, bias, res, a, y_act, y, acc, acc0, Kj, dj in checks:
        xa = rnd(F.leaky_relu(x, 0.1), dtype)
        assert torch.equal(unpack(a, B, Cc, L, dtype, rows), xa), "pack_act mismatch"
        ref = F.conv1d(xa.double(), rnd(w, dtype).double(), bias.double(), dilation=dj, padding=(Kj - 1) // 2 * dj)
        if res is not None:
            ref = ref + res.double()
        scale = float(ref.abs().max())
        err = float((y.double() - ref).abs().max())
        assert err <= 2e-4 * scale, (dtype, Cc, Kj, dj, B, L, err, scale)
        assert torch.equal(unpack(y_act, B, Cc, L, dtype, rows), rnd(F.leaky_relu(y, 0.1), dtype)), "y_act mismatch"
        assert float((acc - (acc0 + y / 3.0)).abs().max()) <= 1e-5 * max(1.0, scale)
        # pad rows of the packed output stay zero
        CK = 8 if dtype == 0 else 4
        t = y_act.view(torch.bfloat16 if dtype == 0 else torch.float32).view(B, Cc // CK, rows, CK)
        assert float(t[:, :, :64].abs().max()) == 0.0 and float(t[:, :, 64 + L:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("Cc", [32, 64, 128, 256])
@pytest.mark.parametrize("K,dil", [(3, 1), (7, 3), (11, 5)])
def test_tc_conv_shapes(dtype, Cc, K, dil):
    if dtype == 1 and Cc == 256:
        # tf32 doubles the activation tile: 256 channels do not fit shared memory -> loud shape error, no fallback
        with pytest.raises(RuntimeError, match="does not fit shared memory"):
            run_case(dtype, Cc, K, dil, B=2, L=1000)
        return
    run_case(dtype, Cc, K, dil, B=2, L=1000)


@pytest.mark.parametrize("L", [1, 255, 256, 257, 777, 5000])
def test_tc_conv_ragged_lengths(L):
    run_case(0, 128, 7, 1, B=1, L=L, with_res=False)
    run_case(1, 64, 11, 5, B=3, L=L)


def test_tc_conv_three_jobs_many_tiles():
    run_case(0, 128, 11, 5, B=2, L=40000, njobs=3)
    run_case(1, 32, 11, 1, B=1, L=100000, njobs=3)
