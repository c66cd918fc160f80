"""Timing probe for rvc_conv1d_tc: runs one ResBlock-style launch (3 jobs, k = 11/7/3) under the kernel's debug switches."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from rvc_maker_b200 import _lib

lib = _lib.load()
dev = "cuda:0"
st = torch.cuda.current_stream().cuda_stream


def probe(Cc, L, conv2, dbg, nt=0, mt=0, reps=5, blocked=0):
    rows = lib.rvc_packed_rows(L)
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.nt, d.mt = 3, 1, Cc, Cc, L, 2, nt, mt
    d.f32_blocked = blocked if conv2 else 0
    keep = []
    for j, (k, dil) in enumerate(((11, 5), (7, 3), (3, 1))):
        a = torch.zeros(Cc * rows * 2, dtype=torch.uint8, device=dev)
        w = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=dev)
        ya = torch.zeros(Cc * rows * 2, dtype=torch.uint8, device=dev)
        bias = torch.zeros(Cc, device=dev)
        jb = d.jobs[j]
        jb.a, jb.w, jb.bias, jb.y_act = a.data_ptr(), w.data_ptr(), bias.data_ptr(), ya.data_ptr()
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = k, dil, (k - 1) // 2 * dil, 1.0, 0.1
        keep += [a, w, ya, bias]
        if conv2:
            res = torch.zeros(Cc, L, device=dev)
            y = torch.zeros(Cc, L, device=dev)
            jb.res, jb.res_bs, jb.res_cs = res.data_ptr(), Cc * L, (8 * L if blocked & 1 else L)
            jb.y_f32, jb.y_bs, jb.y_cs = y.data_ptr(), Cc * L, (8 * L if blocked & 2 else L)
            keep += [res, y]
    d.jobs[0].reserved = dbg
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.rvc_conv1d_tc(C.byref(d), st), "tc")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3


if __name__ != "__main__":
    raise SystemExit
names = {0: "full", 1: "no epilogue I/O", 2: "no MMA issue", 4: "no operand loads", 3: "no epi I/O + no MMA", 5: "no epi I/O + no loads", 6: "no MMA + no loads",
         7: "barriers only"}
for Cc, L in ((32, 480000), (64, 240000), (128, 120000)):
    for conv2 in (0, 1):
        print(f"C={Cc} L={L} {'conv2 (res + fp32 + packed out)' if conv2 else 'conv1 (packed out)'}")
        for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
            print(f"   {names[dbg]:24s} {probe(Cc, L, conv2, dbg):8.1f} us")
