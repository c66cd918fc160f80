"""Timeline of CTA 0 of rvc_resblock_pair_tc (debug build only: `python -m rvc_maker_b200.build --trace`, RVC_B200_LIB points at it).

Runs one fused ResBlock step (3 jobs, k = 11/7/3, split-fp16 stream as the engine uses it) at the generator-stage sizes of configs[1], reads the
(event, unit, SM clock) records of CTA 0 back and prints, per role, where the time of a steady-state unit goes: waiting on which barrier, issuing,
or working.  Event ids (resblock_pair_tc.cu): 1 producer got its free slot; 2/3 MMA passed ACC1_EMPTY / A_FULL, 4 phase 1 issued, 5/6 passed
ACC2_EMPTY / T_FULL, 7 phase 2 issued; per epilogue group eg: 10+eg unit begins, 13+eg ACC1_FULL, 16+eg T_EMPTY, 19+eg epilogue 1 done,
22+eg ACC2_FULL, 25+eg epilogue 2 done, 28+eg top of the unit loop, and for the first epilogue-2 item 31+eg residual registers arrived,
34+eg accumulator columns arrived, 37+eg item done.
"""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
os.environ.setdefault("RVC_B200_LIB", os.path.join(ROOT, "rvc_maker_b200", "librvc_b200_trace.so"))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from rvc_maker_b200 import _lib

lib = _lib.load()
dev = "cuda:0"


def run(Cc, L, sum_jobs, ks=((11, 5), (7, 3), (3, 1)), dbg=0):
    st = torch.cuda.current_stream().cuda_stream
    rows = lib.rvc_packed_rows(L)
    d = _lib.ResblockPairDesc()
    d.njobs, d.B, d.C, d.L, d.dtype, d.slope = 3, 1, Cc, L, 2, 0.1
    d.sum_jobs, d.y_ncl, d.alpha, d.reserved = sum_jobs, sum_jobs, 1.0 / 3.0, dbg
    keep = []
    avg = torch.zeros(Cc, L, device=dev)
    for j, (k, dil) in enumerate(ks):
        a, lo, ya = (torch.zeros(Cc * rows * 2, dtype=torch.uint8, device=dev) for _ in range(3))
        w1, w2 = (torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=dev) for _ in range(2))
        b1, b2 = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
        pj = d.jobs[j]
        pj.a, pj.a_lo, pj.w1, pj.b1, pj.w2, pj.b2 = a.data_ptr(), lo.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
        if not sum_jobs:
            pj.y_act, pj.y_lo = ya.data_ptr(), lo.data_ptr()
        elif j == 0:
            pj.y = avg.data_ptr()
        pj.K1, pj.dil1, pj.K2 = k, dil, k
        keep += [a, lo, ya, w1, w2, b1, b2]
    cap = 80 * 256
    buf = torch.zeros(cap, dtype=torch.int64, device=dev)
    for _ in range(2):    # warm-up without records
        _lib.check(lib.rvc_resblock_pair_tc(C.byref(d), st), "pair")
    torch.cuda.synchronize()
    lib.rvc_debug_pair_trace(buf.data_ptr(), cap)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.rvc_resblock_pair_tc(C.byref(d), st), "pair")
    e1.record()
    torch.cuda.synchronize()
    lib.rvc_debug_pair_trace(None, 0)
    rec = buf.cpu().numpy().reshape(-1, 5, 16)      # [unit][role][k]
    return rec, e0.elapsed_time(e1) * 1e3


def table(rec):
    """t[e][n] = clock of event e at unit n (relative to the first record; -1 = not reached)."""
    nu = int(np.nonzero(rec[:, 1, 0])[0].max()) + 1
    t0 = rec[rec > 0].min()
    t = {}
    evs = {1: (0, 0)}
    evs.update({2 + k: (1, k) for k in range(6)})
    evs.update({10 + 3 * k + eg: (2 + eg, k) for k in range(16) for eg in range(3)})
    for e, (role, k) in evs.items():
        col = rec[:nu, role, k]
        t[e] = np.where(col > 0, col - t0, -1)
    return t, nu


def report(name, Cc, L, sum_jobs, dbg=0):
    rec, us = run(Cc, L, sum_jobs, dbg=dbg)
    t, nu = table(rec)
    lo, hi = 3, nu - 3       # steady state
    if hi - lo < 4:
        print(f"{name}: only {nu} units in CTA 0"); return
    sl = slice(lo, hi)
    mhz = 1965.0
    cyc = lambda x: float(np.mean(x)) / mhz     # us
    period = (t[4][hi - 1] - t[4][lo]) / (hi - 1 - lo) / mhz
    print(f"== {name} dbg={dbg}: C={Cc} L={L} {'stage-average' if sum_jobs else 'pair'}: launch {us:.1f} us, {nu} units in CTA 0, {period:.2f} us per unit (steady state)")
    # MMA warp 1 (unit n: 2,3,4 ; unit n-1 phase 2: 5,6,7 recorded with unit n-1)
    d = {
        "wait ACC1_EMPTY": t[2][lo + 1:hi] - t[7][lo - 1:hi - 2],      # phase 1 of n follows phase 2 of n-2
        "wait A_FULL": t[3][sl] - t[2][sl],
        "issue phase 1": t[4][sl] - t[3][sl],
        "wait ACC2_EMPTY": t[5][lo - 1:hi - 1] - t[4][sl],              # phase 2 of n-1 follows phase 1 of n
        "wait T_FULL": t[6][sl] - t[5][sl],
        "issue phase 2": t[7][sl] - t[6][sl],
    }
    print("  MMA warp 1:   " + "  ".join(f"{k} {cyc(v):.2f}" for k, v in d.items()))
    for eg in range(3):
        b, a1, te, d1, a2, d2 = (t[x + eg] for x in (10, 13, 16, 19, 22, 25))
        parts = {
            "wait ACC1_FULL": a1[sl] - b[sl],
            "wait T_EMPTY": te[sl] - a1[sl],
            "epilogue 1": d1[sl] - te[sl],
        }
        top, rr, tr, i1 = (t[x + eg] for x in (28, 31, 34, 37))
        parts["loop top -> decoded"] = t[40 + eg][sl] - top[sl]
        parts["residual loads issued in"] = np.where(t[49 + eg][sl] > 0, t[49 + eg][sl] - t[46 + eg][sl], 0)
        if not sum_jobs:
            parts.update({"e2 item 1: residual regs": rr[lo - 1:hi - 1] - a2[lo - 1:hi - 1], "tmem": tr[lo - 1:hi - 1] - rr[lo - 1:hi - 1],
                          "math + stores": i1[lo - 1:hi - 1] - tr[lo - 1:hi - 1], "loop back": top[lo + 1:hi + 1] - d2[lo - 1:hi - 1]})
            parts.update({"wait ACC2_FULL": a2[lo - 1:hi - 1] - d1[sl], "epilogue 2": d2[lo - 1:hi - 1] - a2[lo - 1:hi - 1], "to next unit": b[lo + 1:hi + 1] - d2[lo - 1:hi - 1]})
        else:
            parts.update({"to next unit (incl. epilogue 2 every 3rd)": b[lo + 1:hi + 1] - d1[sl]})
        print(f"  epilogue g{eg}: " + "  ".join(f"{k} {cyc(v):.2f}" for k, v in parts.items()))
    # accumulator latency: phase-1 issue done (4) -> ACC1_FULL seen by g0 (13): the tensor pipe draining the queue
    print(f"  phase 1 issued -> accumulator visible: {cyc(t[13][sl] - t[4][sl]):.2f} us;  t tile ready (19 max) -> phase 2 issued: "
          f"{cyc(t[7][sl] - np.maximum(np.maximum(t[19][sl], t[20][sl]), t[21][sl])):.2f} us")
    if "--save" in sys.argv:
        np.save(os.path.join(sys.argv[sys.argv.index("--save") + 1], f"pair_trace_C{Cc}_{'avg' if sum_jobs else 'pair'}.npy"), rec)
    if "--dump" in sys.argv:
        for n in range(lo, min(lo + 4, hi)):
            print("   unit", n, {e: round(float(t[e][n]) / mhz, 2) for e in sorted(t)})


if __name__ == "__main__":
    if lib.rvc_debug_pair_trace(None, 0) < 0:
        raise SystemExit("this library has no trace points: python -m rvc_maker_b200.build --trace")
    report("stage 2", 64, 240000, 0)
    if "--quick" in sys.argv:
        report("stage 3", 32, 480000, 0)
        raise SystemExit(0)
    report("stage 2", 64, 240000, 1)
    report("stage 3", 32, 480000, 0)
    report("stage 3", 32, 480000, 1)
    for dbg in (4, 2, 1):       # no residual loads | no epilogue-2 global I/O | no MMA
        report("stage 2", 64, 240000, 0, dbg=dbg)
