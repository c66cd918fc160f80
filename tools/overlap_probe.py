"""Timing probe: the encoder (+ flow) of chunk i+1 next to the generator of chunk i on two streams, the generator's persistent kernels capped to
fewer CTAs (rvc_set_grid_limit).  Timing only — the two lanes share workspaces here, so the waveform is NOT valid; the engine's overlap mode
double-buffers what crosses the lanes.

    python tools/overlap_probe.py [steps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
from rvc_maker_b200 import configs

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cfg = configs.CONFIGS["v2_48k"]
dev = torch.device("cuda", 0)
net, _ = bench.build_net(cfg, dev, "tc")
B, T = 1, 1000
inp = {k: v.to(dev) for k, v in bench.synth_inputs(cfg, B, T).items()}
for _ in range(3):
    net.infer(inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"])
torch.cuda.synchronize()
eng = net._engine
lib = eng.lib
pe, pd = eng.plan("enc", B, T), eng.plan("dec", B, T)
names = [o[0] for o in pd.ops]
print("dec ops:", names[:8], "...", len(names))


def run_ops(plan, lo, hi, st):
    for name, fn, args in plan.ops[lo:hi]:
        if name is None:
            fn()
        else:
            rc = fn(*args, st)
            assert rc == 0, (name, rc)


def capture(fn, stream):
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            fn(stream.cuda_stream)
    return g


def measure(split, limit, prio_back, label, prio_front=0, skip=()):
    sf = torch.cuda.Stream(dev, priority=prio_front)
    sb = torch.cuda.Stream(dev, priority=prio_back)
    lib.rvc_set_grid_limit(0)
    front = [i for i in range(split) if i not in skip]
    back = sorted(set(range(len(pd.ops))) - set(front))
    gf = capture(lambda st: (run_ops(pe, 0, len(pe.ops), st), [run_ops(pd, i, i + 1, st) for i in front]), sf)
    lib.rvc_set_grid_limit(limit)
    gb = capture(lambda st: [run_ops(pd, i, i + 1, st) for i in back], sb)
    lib.rvc_set_grid_limit(0)
    # serial reference: front then back on one stream
    def serial(n):
        with torch.cuda.stream(sf):
            for _ in range(n):
                gf.replay(); gb.replay()
    def piped(n):
        # front(i+1) next to back(i): back(i) waits for front(i); front(i+1) waits for back(i-1)'s hand-off (here: nothing, timing only)
        evf = [torch.cuda.Event() for _ in range(n)]
        with torch.cuda.stream(sf):
            gf.replay(); evf[0].record(sf)
        for i in range(n):
            if i + 1 < n:
                with torch.cuda.stream(sf):
                    gf.replay(); evf[i + 1].record(sf)
            with torch.cuda.stream(sb):
                sb.wait_event(evf[i])
                gb.replay()
        sf.wait_stream(sb)
    out = []
    for fn in (serial, piped):
        fn(3)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sf):
            a.record(sf)
        fn(steps)
        with torch.cuda.stream(sf):
            sf.wait_stream(sb)
            b.record(sf)
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / steps)
    print(f"{label}: front ops {[names[i] for i in front]}, limit {limit}, priorities front {prio_front} back {prio_back}: serial {out[0]:.3f} ms/step, overlapped {out[1]:.3f} ms/step")


i_flow = next(i for i, n in enumerate(names) if n and "flow" in n) + 1


def lanes_alone(split):
    """each lane's graph replayed back to back on its own: what the overlapped run could reach at best = max(front, back) ... front + back"""
    sf, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    gf = capture(lambda st: (run_ops(pe, 0, len(pe.ops), st), run_ops(pd, 0, split, st)), sf)
    gb = capture(lambda st: run_ops(pd, split, len(pd.ops), st), sb)
    out = []
    for g, s_ in ((gf, sf), (gb, sb)):
        with torch.cuda.stream(s_):
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s_)
            for _ in range(steps):
                g.replay()
            b.record(s_)
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / steps)
    print(f"alone: front lane (encoder + flow) {out[0]:.3f} ms, back lane {out[1]:.3f} ms per step")


lanes_alone(i_flow)


def measure_front_subset(frac, label, front_limit=0):
    """front lane with only a fraction of the encoder's launches (timing only): is the overlapped step bound by the front lane's progress?
    front_limit: cap on the CTAs of the front lane's persistent launches (they are latency-bound: fewer CTAs steal less SM time from the generator)"""
    sf, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    n_enc = int(len(pe.ops) * frac)
    lib.rvc_set_grid_limit(front_limit)
    gf = capture(lambda st: (run_ops(pe, 0, n_enc, st), run_ops(pd, 0, i_flow, st)), sf)
    lib.rvc_set_grid_limit(0)
    gb = capture(lambda st: run_ops(pd, i_flow, len(pd.ops), st), sb)
    def piped(n):
        evf = [torch.cuda.Event() for _ in range(n)]
        with torch.cuda.stream(sf):
            gf.replay(); evf[0].record(sf)
        for i in range(n):
            if i + 1 < n:
                with torch.cuda.stream(sf):
                    gf.replay(); evf[i + 1].record(sf)
            with torch.cuda.stream(sb):
                sb.wait_event(evf[i])
                gb.replay()
        sf.wait_stream(sb)
    piped(3)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sf):
        a.record(sf)
    piped(steps)
    with torch.cuda.stream(sf):
        sf.wait_stream(sb)
        b.record(sf)
    torch.cuda.synchronize()
    print(f"{label}: {n_enc} of {len(pe.ops)} encoder launches + flow in the front lane, front grid limit {front_limit}: overlapped {a.elapsed_time(b) / steps:.3f} ms/step")


for rep in range(2):
    for lim in (0, 96, 64, 48, 32, 24, 16, 8):
        measure_front_subset(1.0, "front grid cap", front_limit=lim)
for rep in range(1):
    measure(i_flow, 0, 0, "enc + flow")
raise SystemExit(0)

for rep in range(2):
    measure(i_flow, 0, 0, "enc + flow")
