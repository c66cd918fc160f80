"""Cost of a (B, T) shape the engine has not seen: plan build + eager run (1st call), graph capture (2nd), replay (3rd+).
The reference's energy-split chunk plan gives every chunk its own T; this is what `StreamConverter(bucket=...)` avoids.

usage (on a B200): python tools/new_shape_cost.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rvc_maker_b200 import configs  # noqa: E402

cfg = configs.CONFIGS["v2_48k"]
dev = torch.device("cuda:0")
net, _ = bench.build_net(cfg, dev)
for T in (1000, 3000, 3100, 4000):
    inp = {k: v.to(dev) for k, v in bench.synth_inputs(cfg, 1, T).items()}
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        net.infer(inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"])
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"T={T}: 1st call {ts[0]:.1f} ms (plan build + eager), 2nd {ts[1]:.1f} ms (graph capture), then {min(ts[2:]):.2f} ms; "
          f"mem {torch.cuda.memory_allocated(dev) / 2**30:.2f} GiB")
