import sys, ctypes as C, torch
sys.path.insert(0, ".")
from rvc_maker_b200 import _lib
lib = _lib.load(); dev = "cuda:0"; st = torch.cuda.current_stream().cuda_stream
def probe(contig, nt, mt, L=12000, Cc=256, reps=6):
    rows = lib.rvc_packed_rows(L)
    nb = Cc // nt if contig else 1
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.nt, d.mt = 3, nb, Cc, (nt if contig else Cc), L, 2, nt, mt
    keep = []
    for j, k in enumerate((11, 7, 3)):
        a = torch.zeros(Cc * rows * 2 + 4096, dtype=torch.uint8, device=dev); w = torch.zeros(k * Cc * Cc * 2, dtype=torch.uint8, device=dev)
        ya = torch.zeros(Cc * rows * 2, dtype=torch.uint8, device=dev); bias = torch.zeros(Cc, device=dev)
        jb = d.jobs[j]
        jb.a, jb.w, jb.bias, jb.y_act = a.data_ptr(), w.data_ptr(), bias.data_ptr(), ya.data_ptr()
        jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = k, 1, (k - 1) // 2, 1.0, 0.1
        if contig:
            jb.a_bs, jb.w_bs, jb.y_act_bs = 16, k * (Cc // 8) * nt * 16, (nt // 8) * rows * 16
        keep += [a, w, ya, bias]
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = lib.rvc_conv1d_tc(C.byref(d), st); e1.record(); torch.cuda.synchronize()
        if rc: print(lib.rvc_last_error()); return float("nan")
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3
for nt, mt in ((128, 2), (64, 4), (64, 2)):
    print(f"C=256 L=12000 conv1 nt={nt} mt={mt}: strided weight stages {probe(False, nt, mt):6.1f} us   contiguous (tile-major emulation) {probe(True, nt, mt):6.1f} us")
