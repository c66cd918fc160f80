import sys, ctypes as C, torch
sys.path.insert(0, ".")
from rvc_maker_b200 import _lib
lib = _lib.load(); dev = "cuda:0"; st = torch.cuda.current_stream().cuda_stream
def probe_sum(Cc, L, nt, mt, reps=5):
    rows = lib.rvc_packed_rows(L)
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.nt, d.mt, d.sum_jobs, d.f32_blocked = 3, 1, Cc, Cc, L, 2, nt, mt, 1, 1
    keep = []
    ya = torch.zeros(Cc * rows * 2, dtype=torch.uint8, device=dev)
    for j, k in enumerate((11, 7, 3)):
        a = torch.zeros(Cc * rows * 2, dtype=torch.uint8, device=dev); w = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=dev)
        bias = torch.zeros(Cc, device=dev); res = torch.zeros(Cc, L, device=dev)
        jb = d.jobs[j]
        jb.a, jb.w, jb.bias = a.data_ptr(), w.data_ptr(), bias.data_ptr()
        jb.res, jb.res_bs, jb.res_cs = res.data_ptr(), Cc * L, 8 * L
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = k, 1, (k - 1) // 2, 1 / 3, 0.1
        if j == 0: jb.y_act = ya.data_ptr()
        keep += [a, w, bias, res]
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = lib.rvc_conv1d_tc(C.byref(d), st); e1.record(); torch.cuda.synchronize()
        if rc: return float("nan")
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3
for Cc, L in ((256, 12000), (256, 10000), (128, 120000)):
    for nt, mt in ((0, 0), (128, 1), (128, 2), (64, 2), (64, 4), (32, 4)):
        print(f"sum_jobs C={Cc} L={L} nt={nt:3d} mt={mt}: {probe_sum(Cc, L, nt, mt):7.1f} us")
