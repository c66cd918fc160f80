"""Where the captured graph's time goes: replays sub-ranges of the launch plans of one (B, T) shape as separate CUDA graphs."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from rvc_maker_b200 import configs  # noqa: E402

cfg = dict(configs.CONFIGS["v2_48k"])
dev = torch.device("cuda:0")
net, _sd = bench.build_net(cfg, dev)
inp = {k: v.to(dev) for k, v in bench.synth_inputs(cfg, 1, 1000).items()}
for _ in range(3):
    net.infer(inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"])
torch.cuda.synchronize()
eng = net._engine
pe, pd = eng.plan("enc", 1, 1000), eng.plan("dec", 1, 1000)


def run_ops(plan, lo, hi, st):
    for name, fn, args in plan.ops[lo:hi]:
        if name is None:
            fn()
        else:
            rc = fn(*args, st)
            assert rc == 0, name


def time_range(plan, lo, hi, label, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run_ops(plan, lo, hi, s.cuda_stream)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            run_ops(plan, lo, hi, torch.cuda.current_stream().cuda_stream)
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    n = sum(1 for o in plan.ops[lo:hi] if o[0] is not None)
    print(f"{label:44s} {best * 1e3:8.1f} us  ({n} calls, {best * 1e3 / max(n, 1):.1f} us each)")
    return best


# note: ops metadata and ops lists are aligned only for named ops; use ops indices directly
ops_d = pd.ops
idx_named = [i for i, o in enumerate(ops_d) if o[0] is not None]
name_at = {i: ops_d[i][0] for i in idx_named}
first_gen = next(i for i in idx_named if name_at[i] in ("sine_source", "pack_z", "conv_pre"))
time_range(pe, 0, len(pe.ops), "encoder plan (embed .. sample_zp)")
time_range(pd, 0, first_gen, "flow (cond + 4 couplings)")
time_range(pd, first_gen, len(ops_d), "generator (source, conv_pre, 4 stages, post)")
# generator stages
marks = [i for i in idx_named if name_at[i].startswith("noise")] + [len(ops_d)]
for a, b in zip(marks[:-1], marks[1:]):
    time_range(pd, a, b, f"  stage from {name_at[a]}")
# one encoder layer
ops_e = pe.ops
q = [i for i, o in enumerate(ops_e) if o[0] and o[0].startswith("qkv")]
time_range(pe, q[0], q[1], "  one encoder layer")
