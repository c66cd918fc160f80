"""Per-launch cost of small rvc_conv1d_tc launches inside a CUDA graph (back-to-back dependent launches): the numbers quoted in DESIGN.md section 4 for the short-sequence encoder / flow layers."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from rvc_maker_b200 import _lib
lib = _lib.load()
dev = "cuda:0"
def mk(Cin, Cout, K, L, nt=0, res=False, yf=True, ya=True):
    rows = lib.rvc_packed_rows(L)
    d = _lib.Conv1dTcDesc()
    d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.nt = 1, 1, Cin, Cout, L, 2, nt
    a = torch.zeros(Cin * rows * 2, dtype=torch.uint8, device=dev)
    w = torch.zeros(K * Cin * Cout * 2, dtype=torch.uint8, device=dev)
    yact = torch.zeros(Cout * rows * 2, dtype=torch.uint8, device=dev)
    y = torch.zeros(Cout, L, device=dev); r = torch.zeros(Cout, L, device=dev)
    bias = torch.zeros(Cout, device=dev)
    jb = d.jobs[0]
    jb.a, jb.w, jb.bias = a.data_ptr(), w.data_ptr(), bias.data_ptr()
    if ya: jb.y_act = yact.data_ptr()
    if yf: jb.y_f32, jb.y_bs, jb.y_cs = y.data_ptr(), Cout * L, L
    if res: jb.res, jb.res_bs, jb.res_cs = r.data_ptr(), Cout * L, L
    jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = K, 1, (K - 1) // 2, 1.0, 1.0
    return d, (a, w, yact, y, r, bias)
def time_graph(descs, n=40):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for d, _ in descs: lib.rvc_conv1d_tc(C.byref(d), s.cuda_stream)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                for d, _ in descs: lib.rvc_conv1d_tc(C.byref(d), torch.cuda.current_stream().cuda_stream)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / (n * len(descs))
for name, args in (("1 tile 192->32 k1 L=128", (192, 32, 1, 128)), ("qkv 192->576 k1 L=1000", (192, 576, 1, 1000)), ("attn_o 192->192 k1", (192, 192, 1, 1000)),
                   ("ffn1 192->768 k3", (192, 768, 3, 1000)), ("ffn2 768->192 k3", (768, 192, 3, 1000)), ("ffn2 nt=16", (768, 192, 3, 1000, 16)),
                   ("in 192->384 k5 (not gated)", (192, 384, 5, 1000, 64)), ("post 192->96 k1", (192, 96, 1, 1000)), ("embed 768->192 k1", (768, 192, 1, 1000)),
                   ("embed nt=16", (768, 192, 1, 1000, 16))):
    d = mk(*args)
    print(f"{name:32s} {time_graph([d]):6.2f} us per launch in a graph (back-to-back dependent launches)")
print("---- variants")
for name, args in (("ffn1 nt=64", (192, 768, 3, 1000, 64)), ("ffn1 nt=128", (192, 768, 3, 1000, 128)), ("ffn1 nt=32", (192, 768, 3, 1000, 32)),
                   ("qkv nt=64", (192, 576, 1, 1000, 64)), ("in nt=128", (192, 384, 5, 1000, 128))):
    d = mk(*args)
    print(f"{name:32s} {time_graph([d]):6.2f} us")
# ffn2 as 3 K-split jobs of 256 input channels in one launch
rows = lib.rvc_packed_rows(1000)
d = _lib.Conv1dTcDesc()
d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype = 3, 1, 256, 192, 1000, 2
keep = []
for j in range(3):
    a = torch.zeros(256 * rows * 2, dtype=torch.uint8, device=dev); w = torch.zeros(3 * 256 * 192 * 2, dtype=torch.uint8, device=dev)
    y = torch.zeros(192, 1000, device=dev); bias = torch.zeros(192, device=dev)
    jb = d.jobs[j]
    jb.a, jb.w, jb.bias = a.data_ptr(), w.data_ptr(), bias.data_ptr()
    jb.y_f32, jb.y_bs, jb.y_cs = y.data_ptr(), 192000, 1000
    jb.K, jb.dil, jb.pad, jb.alpha, jb.act_slope = 3, 1, 1, 1.0, 1.0
    keep += [a, w, y, bias]
print(f"{'ffn2 as 3 K-split jobs':32s} {time_graph([(d, keep)]):6.2f} us")
d.nt = 64
print(f"{'ffn2 as 3 K-split jobs nt=64':32s} {time_graph([(d, keep)]):6.2f} us")
