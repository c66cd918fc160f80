"""Timing probe for rvc_resblock_pair_tc: one fused ResBlock step (3 jobs, k = 11/7/3) at generator-stage sizes."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from rvc_maker_b200 import _lib

lib = _lib.load()
dev = "cuda:0"
st = torch.cuda.current_stream().cuda_stream


def probe(Cc, L, reps=6, sum_jobs=0, ks=((11, 5), (7, 3), (3, 1)), dbg=0):
    rows = lib.rvc_packed_rows(L)
    d = _lib.ResblockPairDesc()
    d.njobs, d.B, d.C, d.L, d.dtype, d.slope = 3, 1, Cc, L, 2, 0.1
    d.sum_jobs, d.y_ncl, d.alpha, d.reserved = sum_jobs, 0, 1.0 / 3.0, dbg
    keep = []
    for j, (k, dil) in enumerate(ks):
        a = torch.zeros(Cc * rows * 2, dtype=torch.uint8, device=dev)
        w1 = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=dev)
        w2 = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=dev)
        ya = torch.zeros(Cc * rows * 2, dtype=torch.uint8, device=dev)
        b1, b2 = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
        x = torch.zeros(Cc, L, device=dev)
        pj = d.jobs[j]
        pj.a, pj.w1, pj.b1, pj.w2, pj.b2 = a.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
        pj.res, pj.y, pj.y_act = x.data_ptr(), (None if sum_jobs else x.data_ptr()), (ya.data_ptr() if (not sum_jobs or j == 0) else None)
        pj.K1, pj.dil1, pj.K2 = k, dil, k
        keep += [a, w1, w2, ya, b1, b2, x]
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.rvc_resblock_pair_tc(C.byref(d), st), "pair")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dbg":
    for Cc, L in ((64, 240000), (32, 480000)):
        for sm in (0, 1):
            print(f"C={Cc} {'stage-average' if sm else 'pair'}: " + "  ".join(f"{nm} {probe(Cc, L, sum_jobs=sm, dbg=dv):6.1f}" for nm, dv in (("full", 0), ("no-MMA", 1), ("no-res-loads", 4), ("no-e2-I/O", 2), ("no-MMA+no-e2-I/O", 3))))
elif __name__ == "__main__" and len(sys.argv) > 1:
    for ks in (((3, 1),) * 3, ((7, 1),) * 3, ((11, 1),) * 3, ((11, 5), (7, 3), (3, 1))):
        print(ks, "C64 pair %.1f  avg %.1f | C32 pair %.1f  avg %.1f" % (probe(64, 240000, ks=ks), probe(64, 240000, sum_jobs=1, ks=ks),
                                                                    probe(32, 480000, ks=ks), probe(32, 480000, sum_jobs=1, ks=ks)))
elif __name__ == "__main__":
    for Cc, L in ((64, 240000), (32, 480000)):
        try:
            print(f"C={Cc} L={L}: fused pair {probe(Cc, L):7.1f} us   stage-average mode {probe(Cc, L, sum_jobs=1):7.1f} us")
        except Exception as e:
            print(f"C={Cc} L={L}: {e}")
