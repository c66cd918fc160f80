"""Timing probe for the stride-2 ConvTranspose stages (rvc_conv1d_tc in phase-tile mode `up` and phase-column mode `up_cols`), with the kernel's
debug switches: which part of the launch costs what.   python tools/ups_probe.py"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from rvc_maker_b200 import _lib

lib = _lib.load()
dev = "cuda:0"


def probe(Cin, Cout, L, cols, split, noise, dbg, reps=6, u=2, k=4):
    st = torch.cuda.current_stream().cuda_stream
    Lout = L * u
    M = -(-k // u)
    rows_in, rows_out = lib.rvc_packed_rows(L), lib.rvc_packed_rows(Lout)
    a = torch.zeros(Cin * rows_in * 2, dtype=torch.uint8, device=dev)
    w = torch.zeros(u * M * Cin * Cout * 2, dtype=torch.uint8, device=dev)
    bias = torch.zeros(u * Cout, device=dev)
    har = torch.zeros(Lout * 2, device=dev)
    nzw, nzb = torch.zeros(Cout * 4, device=dev), torch.zeros(Cout, device=dev)
    y = torch.zeros(Cout * Lout, device=dev)
    yh, yl = (torch.zeros(Cout * rows_out * 2, dtype=torch.uint8, device=dev) for _ in range(2))
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.L, d.dtype, d.pshift, d.f32_blocked = 1, 1, Cin, L, 2, (k - u) // 2, 3
    if cols:
        d.Cout, d.up_cols = u * Cout, u
    else:
        d.Cout, d.up = Cout, u
    jb = d.jobs[0]
    jb.a, jb.w, jb.bias, jb.y_act = a.data_ptr(), w.data_ptr(), bias.data_ptr(), yh.data_ptr()
    if split:
        jb.y_lo = yl.data_ptr()
    else:
        jb.y_f32, jb.y_bs, jb.y_cs = y.data_ptr(), Cout * Lout, 8 * Lout
    jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope, jb.reserved = M, 1, M - 1, 1.0, 0.1, dbg
    if noise:
        jb.nz_har, jb.nz_w, jb.nz_b, jb.nz_har_bs, jb.nz_kn, jb.nz_sn, jb.nz_pn = har.data_ptr(), nzw.data_ptr(), nzb.data_ptr(), Lout * 2, 4, 2, 1
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "ups")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3


if __name__ == "__main__":
    names = {0: "full", 32: "no packed stores", 16: "no fp32 store", 48: "no stores", 1: "no epilogue I/O", 2: "no MMA", 4: "no operand loads", 7: "skeleton"}
    for Cin, Cout, L in ((128, 64, 120000), (64, 32, 240000)):
        for cols in (0, 1):
            for split in (1, 0):
                print(f"{Cin}->{Cout} L={L} {'phase columns' if cols else 'phase tiles'} {'hi+lo out' if split else 'fp32+packed out'}: " +
                      "  ".join(f"{nm} {probe(Cin, Cout, L, cols, split, 1, dbg):5.1f}" for dbg, nm in names.items()) +
                      f"  | without noise conv {probe(Cin, Cout, L, cols, split, 0, 0):5.1f}")
