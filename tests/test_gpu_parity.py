"""GPU parity: the CUDA path (through the drop-in module -> C ABI) against the committed reference goldens and the oracle.

Tolerances (fp32 CUDA-core path): max-abs <= 2e-5 * max(1, max|ref|) and SNR >= 80 dB (reference fp32 floor
is 2.4e-7 vs fp64; the bar for the bf16 tensor-core path is max-abs <= 1e-3, SNR >= 40 dB).  Mask and shapes are exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import cases, synth_oracle
from oracle.make_golden import CASES

pytestmark = pytest.mark.gpu

FP32_TOL = 2e-5
FP32_SNR = 80.0
# bf16/tf32 tensor-core path (BASELINE.json north_star): waveform max-abs <= 1e-3 and SNR >= 40 dB
TC_TOL = 1e-3
TC_SNR = 40.0


# latent tensors (z, z_p, m_p, logs_p) on the tensor-core path: fp16 operands (10-bit mantissa) through 6 encoder layers and
# 4 couplings; bar = 1e-2 * max|ref| and SNR >= 50 dB (measured: 2e-3 .. 5.3e-3 * max|ref|, 55 .. 63 dB)
TC_LATENT_TOL = 1e-2
TC_LATENT_SNR = 50.0


def tol_for(precision, key):
    if precision.startswith("tc"):
        return (TC_TOL, TC_SNR) if key == "o" else (TC_LATENT_TOL, TC_LATENT_SNR)
    return FP32_TOL, FP32_SNR


def build_net(cfg, sd, use_f0, dtype=torch.float32, precision="fp32"):
    from rvc_maker_b200 import Synthesizer

    net = Synthesizer(*cases.ctor_args(cfg), use_f0=use_f0, text_enc_hidden_dim=cfg["text_enc_hidden_dim"], precision=precision)
    del net.enc_q  # convert.py:565
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    net = net.eval().to("cuda:0")
    return net.half() if dtype == torch.float16 else net.float()


def run_net(net, inp, rate=None):
    net._noise_override = (inp["eps_z"], inp["eps_src"])
    dev = "cuda:0"
    mv = lambda t: None if t is None else t.to(dev)
    r = None if rate is None else torch.tensor(rate)
    o, x_mask, (z, z_p, m_p, logs_p) = net.infer(mv(inp["phone"]), mv(inp["phone_lengths"]), mv(inp["pitch"]), mv(inp["nsff0"]), mv(inp["sid"]), r)
    torch.cuda.synchronize()
    return dict(o=o, x_mask=x_mask, z=z, z_p=z_p, m_p=m_p, logs_p=logs_p)


def assert_close(out, ref, precision="fp32", tag=""):
    for k, r in ref.items():
        tol, snr = tol_for(precision, k)
        if k.startswith("tap_"):
            continue
        r = torch.as_tensor(r)
        v = out[k].float().cpu()
        assert tuple(v.shape) == tuple(r.shape), (tag, k, v.shape, r.shape)
        assert torch.isfinite(v).all(), (tag, k)
        if k == "x_mask":
            assert torch.equal(v, r.float()), (tag, k)
            continue
        m = cases.metrics(v, r)
        assert m["max_abs"] <= tol * max(1.0, float(r.abs().max())), (tag, k, m)
        if float(r.abs().max()) > 0:
            assert m["snr_db"] >= snr, (tag, k, m)


@pytest.mark.parametrize("precision", ["fp32", "tc", "tc-bf16"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_reference_golden(name, precision, golden_dir):
    cname, use_f0, heated, disk, B, T, ragged, rate, wseed, iseed, _ = CASES[name]
    cfg = cases.CONFIGS[cname]
    sd = cases.synth_state_dict(cfg, use_f0=use_f0, seed=wseed, heated=heated)
    if disk:
        sd = cases.to_disk_form(sd)
    inp = cases.synth_inputs(cfg, B, T, seed=iseed, ragged=ragged, use_f0=use_f0)
    out = run_net(build_net(cfg, sd, use_f0, precision=precision), inp, rate)
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    assert_close(out, gold, precision, tag=name)


@pytest.mark.parametrize("precision", ["fp32", "tc", "tc-bf16"])
@pytest.mark.parametrize("cname,use_f0,B,T,ragged", [
    ("v2_40k", True, 1, 300, False),
    ("v2_48k", True, 2, 250, True),
    ("v1_32k", True, 3, 200, True),
    ("v2_40k", False, 2, 200, True),
])
def test_cuda_matches_oracle_midsize(cname, use_f0, B, T, ragged, precision):
    cfg = cases.CONFIGS[cname]
    sd = cases.synth_state_dict(cfg, use_f0=use_f0, seed=4321, heated=True)
    inp = cases.synth_inputs(cfg, B, T, seed=21, ragged=ragged, use_f0=use_f0)
    out = run_net(build_net(cfg, sd, use_f0, precision=precision), inp)
    ro, rmask, (rz, rzp, rm, rl) = synth_oracle.infer(sd, cfg, inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"],
                                                      eps_z=inp["eps_z"], eps_src=inp["eps_src"], use_f0=use_f0)
    ref = dict(o=ro, x_mask=rmask, z=rz, z_p=rzp, m_p=rm, logs_p=rl)
    for k in ("o", "z", "m_p", "logs_p"):
        m = cases.metrics(out[k].float().cpu(), ref[k])
        print(f"[{cname} {precision}] {k}: max_abs={m['max_abs']:.3e} (max|ref|={float(ref[k].abs().max()):.2f}) snr={m['snr_db']:.1f} dB")
    assert_close(out, ref, precision, tag=cname)


def test_half_module_protocol():
    """convert.py:569 `.half()`: inputs arrive fp16, outputs come back fp16, arithmetic stays fp32 inside."""
    cfg = cases.CONFIGS["tiny"]
    sd = cases.synth_state_dict(cfg, seed=3, heated=True)
    inp = cases.synth_inputs(cfg, 2, 32, seed=4)
    net = build_net(cfg, cases.to_disk_form(sd), True, dtype=torch.float16)
    inp16 = dict(inp, phone=inp["phone"].half(), nsff0=inp["nsff0"].half())
    out = run_net(net, inp16)
    assert out["o"].dtype == torch.float16 and out["z"].dtype == torch.float16
    sd16 = {k: v.half().float() for k, v in sd.items()}
    ro, *_ = synth_oracle.infer(sd16, cfg, inp16["phone"].float(), inp["phone_lengths"], inp["pitch"], inp16["nsff0"].float(), inp["sid"],
                                eps_z=inp["eps_z"], eps_src=inp["eps_src"])
    m = cases.metrics(out["o"].float().cpu(), ro)
    assert m["max_abs"] <= 2e-3, m  # fp16 output rounding only


def test_rng_stream_matches_reference_order():
    """Without injected noise the module draws randn[B,I,T], rand[1,1,1], randn[B,T*upp] from torch's CUDA generator."""
    cfg = cases.CONFIGS["tiny"]
    sd = cases.synth_state_dict(cfg, seed=3, heated=True)
    inp = cases.synth_inputs(cfg, 2, 32, seed=4)
    net = build_net(cfg, sd, True)
    torch.manual_seed(99)
    dev = "cuda:0"
    o1 = net.infer(inp["phone"].to(dev), inp["phone_lengths"].to(dev), inp["pitch"].to(dev), inp["nsff0"].to(dev), inp["sid"].to(dev))[0]
    torch.manual_seed(99)
    eps_z = torch.randn(2, cfg["inter_channels"], 32, device=dev)
    torch.rand(1, 1, 1, device=dev)
    eps_src = torch.randn(2, 32 * cases.upp_of(cfg), 1, device=dev)
    net._noise_override = (eps_z, eps_src)
    o2 = net.infer(inp["phone"].to(dev), inp["phone_lengths"].to(dev), inp["pitch"].to(dev), inp["nsff0"].to(dev), inp["sid"].to(dev))[0]
    assert torch.equal(o1, o2)


def test_inputs_not_mutated_and_errors():
    cfg = cases.CONFIGS["tiny"]
    sd = cases.synth_state_dict(cfg, seed=3, heated=True)
    inp = cases.synth_inputs(cfg, 1, 16, seed=4)
    net = build_net(cfg, sd, True)
    ph = inp["phone"].cuda()
    ph0 = ph.clone()
    run_net(net, dict(inp, phone=ph))
    assert torch.equal(ph, ph0)
    with pytest.raises(ValueError):
        net.infer(ph[:, :, :5], inp["phone_lengths"].cuda(), inp["pitch"].cuda(), inp["nsff0"].cuda(), inp["sid"].cuda())
    with pytest.raises(RuntimeError):
        build_net(cfg, sd, True).cpu().infer(inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"])


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_cuda_graph_replay_matches_eager(precision):
    """The second and later calls of a shape replay a captured CUDA graph; results must be bit-identical to the eager launch."""
    cfg = cases.CONFIGS["v2_40k"]
    sd = cases.synth_state_dict(cfg, seed=8, heated=True)
    inp = cases.synth_inputs(cfg, 2, 96, seed=3, ragged=True)
    net = build_net(cfg, sd, True, precision=precision)
    eager = run_net(net, inp)
    assert not net._engine.graphs
    g1 = run_net(net, inp)
    assert len(net._engine.graphs) == 1
    g2 = run_net(net, inp)
    for k in eager:
        assert torch.equal(eager[k], g1[k]) and torch.equal(eager[k], g2[k]), k
    # new inputs through the same graph
    inp2 = cases.synth_inputs(cfg, 2, 96, seed=4, ragged=True)
    a = run_net(net, inp2)
    net2 = build_net(cfg, sd, True, precision=precision)
    b = run_net(net2, inp2)
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("precision", ["tc", "fp32"])
def test_bench_workload_full_size(precision):
    """BASELINE configs[1] at its real size (v2-48k f0, B = 1, T = 1000 frames -> 480 000 samples) against the oracle."""
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=1234, heated=True)
    inp = cases.synth_inputs(cfg, 1, 1000, seed=0)
    out = run_net(build_net(cfg, sd, True, precision=precision), inp)
    ro, rmask, (rz, rzp, rm, rl) = synth_oracle.infer(sd, cfg, inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"],
                                                      eps_z=inp["eps_z"], eps_src=inp["eps_src"])
    assert tuple(out["o"].shape) == (1, 1, 480000)
    ref = dict(o=ro, x_mask=rmask, z=rz, z_p=rzp, m_p=rm, logs_p=rl)
    m = cases.metrics(out["o"].float().cpu(), ro)
    print(f"\n[config 2 full size, {precision}] o: max_abs={m['max_abs']:.3e} snr={m['snr_db']:.1f} dB ac_snr={m['snr_ac_db']:.1f} dB")
    assert_close(out, ref, precision, tag="config2")


@pytest.mark.parametrize("precision", ["tc", "fp32"])
@pytest.mark.parametrize("T", [1, 2, 9, 127, 129, 1001])
def test_edge_lengths(T, precision):
    """Lengths around the kernels' tile edges (128-position MMA tiles, 8-key chunks, 64-row pads), a single frame, and a ragged batch
    whose second item can be empty (T = 1 draws lengths [1, 0]: every key masked, softmax over a row of -1e4)."""
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=777, heated=True)
    inp = cases.synth_inputs(cfg, 2, T, seed=T, ragged=True)
    out = run_net(build_net(cfg, sd, True, precision=precision), inp)
    ro, rmask, (rz, rzp, rm, rl) = synth_oracle.infer(sd, cfg, inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"],
                                                      eps_z=inp["eps_z"], eps_src=inp["eps_src"])
    assert tuple(out["o"].shape) == (2, 1, T * 480)
    assert_close(out, dict(o=ro, x_mask=rmask, z=rz, z_p=rzp, m_p=rm, logs_p=rl), precision, tag=f"T={T}")


@pytest.mark.parametrize("B,T,rate", [(1, 1024, None), (1, 1025, None), (2, 300, 0.5)])
def test_softmax_variant_boundary_and_rate_tc(B, T, rate):
    """T = 1024 / 1025 sit on either side of the short- and long-sequence softmax kernels; `rate` (synthesizers.py:451-456) trims the
    head frames before the flow on the real architecture.  (The wide-batch case lives in tests/test_gpu_zz_wide_batch.py.)"""
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=778, heated=True)
    inp = cases.synth_inputs(cfg, B, T, seed=B + T, ragged=B > 1)
    out = run_net(build_net(cfg, sd, True, precision="tc"), inp, rate)
    ro, rmask, (rz, rzp, rm, rl) = synth_oracle.infer(sd, cfg, inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"],
                                                      rate=rate, eps_z=inp["eps_z"], eps_src=inp["eps_src"])
    assert_close(out, dict(o=ro, x_mask=rmask, z=rz, z_p=rzp, m_p=rm, logs_p=rl), "tc", tag=f"B={B} T={T} rate={rate}")


def test_plan_cache_eviction_frees_workspaces():
    """The engine keeps the 8 most recently used (B, T) shapes.  An evicted shape's workspace must go at once — not when the cyclic garbage
    collector gets round to it: energy-split streams bring a new shape per chunk — and the shape must still work when it comes back."""
    import gc
    import weakref

    cfg = cases.CONFIGS["tiny"]
    sd = cases.synth_state_dict(cfg, seed=31, heated=True)
    net = build_net(cfg, sd, True, precision="fp32")
    inp = cases.synth_inputs(cfg, 1, 16, seed=1)
    gc.collect()
    gc.disable()
    try:
        first = run_net(net, inp)["o"].clone()
        eng = net._engine
        refs = [weakref.ref(eng.plans[(kind, 1, 16)]) for kind in ("enc", "dec")]
        bufs = [weakref.ref(eng.plans[("dec", 1, 16)].out_arena)]
        for T in range(17, 17 + eng.max_plans // 2):
            run_net(net, cases.synth_inputs(cfg, 1, T, seed=T))
        assert ("dec", 1, 16) not in eng.plans and len(eng.plans) <= eng.max_plans
        assert all(r() is None for r in refs + bufs), "evicted plans are still alive without a cyclic GC pass"
        again = run_net(net, inp)["o"]
        assert torch.equal(first, again)
    finally:
        gc.enable()


def test_cold_weights_full_batch_tc():
    """Reference-like initialisation statistics (DC-dominated output), ragged batch of 4, tensor-core path."""
    cfg = cases.CONFIGS["v2_40k"]
    sd = cases.synth_state_dict(cfg, seed=99, heated=False)
    inp = cases.synth_inputs(cfg, 4, 257, seed=5, ragged=True)
    out = run_net(build_net(cfg, sd, True, precision="tc"), inp)
    ro, rmask, (rz, rzp, rm, rl) = synth_oracle.infer(sd, cfg, inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"],
                                                      eps_z=inp["eps_z"], eps_src=inp["eps_src"])
    assert_close(out, dict(o=ro, x_mask=rmask, z=rz, z_p=rzp, m_p=rm, logs_p=rl), "tc", tag="cold")


def test_infer_to_host_matches_infer():
    """Pinned read-back helper: same waveform as infer()[0], float32 on the host, complete on return, also for a half module."""
    from rvc_maker_b200 import Synthesizer

    cfg = cases.CONFIGS["tiny"]
    sd = cases.synth_state_dict(cfg, seed=7, heated=True)
    inp = cases.synth_inputs(cfg, 2, 48, seed=3)
    net = Synthesizer(*cases.ctor_args(cfg), use_f0=True, text_enc_hidden_dim=cfg["text_enc_hidden_dim"])
    del net.enc_q
    net.load_state_dict(sd, strict=True)
    net = net.eval().to("cuda:0")
    args = (inp["phone"].cuda(), inp["phone_lengths"].cuda(), inp["pitch"].cuda(), inp["nsff0"].cuda(), inp["sid"].cuda())
    for half in (False, True):
        net = net.half() if half else net.float()
        a = tuple(t.half() if (half and t.is_floating_point()) else t for t in args)
        net._noise_override = (inp["eps_z"], inp["eps_src"])
        want = net.infer(*a)[0]
        got = net.infer_to_host(*a)
        assert got.device.type == "cpu" and got.is_pinned() and got.dtype == torch.float32 and got.shape == want.shape
        assert torch.equal(got, want.float().cpu())


@pytest.mark.parametrize("B,T", [(1, 1000), (2, 173), (3, 80), (2, 81), (1, 1), (40, 64)])
def test_flow_cluster_kernel_matches_per_layer_path(B, T, monkeypatch):
    """rvc_flow_reverse_tc (one cluster kernel, operands through distributed shared memory, state in TMEM) against the per-layer path on the
    same model: same fp16 operand rounding points, only the fp32 summation order of the hidden stream differs -> z agrees to ~1e-5 * max|z|
    (tile seams at multiples of 80 frames, ragged lengths, T below one tile)."""
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=31, heated=True)
    inp = cases.synth_inputs(cfg, B, T, seed=7 * B + T, ragged=B > 1)
    monkeypatch.setenv("RVC_B200_FLOW_CLUSTER", "0")
    ref_net = build_net(cfg, sd, True, precision="tc")
    ref = run_net(ref_net, inp)
    assert not any(m and m["name"] == "flow_cluster" for m in ref_net._engine.plan("dec", B, T).meta)
    monkeypatch.setenv("RVC_B200_FLOW_CLUSTER", "1")
    net = build_net(cfg, sd, True, precision="tc")
    out = run_net(net, inp)
    assert any(m and m["name"] == "flow_cluster" for m in net._engine.plan("dec", B, T).meta)
    scale = float(ref["z"].abs().max())
    assert torch.isfinite(out["z"]).all()
    err = float((out["z"] - ref["z"]).abs().max())
    assert err <= 2e-4 * max(1.0, scale), (err, scale)
    # (the waveform differs by what a few flipped fp16 roundings of z do through the generator — the same size as either path's distance to the oracle)
    assert float((out["o"] - ref["o"]).abs().max()) <= 6e-4
    print(f"[flow cluster B={B} T={T}] z err {err:.2e} (max|z| {scale:.2f}), o diff {float((out['o'] - ref['o']).abs().max()):.2e}")
    again = run_net(net, inp)      # captured graph
    assert torch.equal(again["z"], out["z"]) and torch.equal(again["o"], out["o"])


def test_infer_stream_equals_item_by_item_calls(monkeypatch):
    """Synthesizer.infer_stream (copies of neighbouring items overlap compute, pooled staging sets) yields, in order, exactly what infer_to_host returns
    for each item — five items of two different shapes, pinned noise.  (Grouping off: with it, see test_infer_many_groups_equal_stacked_infer.)"""
    monkeypatch.setenv("RVC_B200_GROUP_FRAMES", "0")
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=5, heated=True)
    net = build_net(cfg, sd, True, precision="tc")
    shapes = [(1, 200), (2, 130), (1, 200), (1, 200), (2, 130)]
    items, want = [], []
    for n, (B, T) in enumerate(shapes):
        inp = cases.synth_inputs(cfg, B, T, seed=100 + n, ragged=B > 1)
        items.append(inp)
    for inp in items:
        net._noise_override = (inp["eps_z"], inp["eps_src"])
        want.append(net.infer_to_host(inp["phone"].pin_memory(), inp["phone_lengths"].pin_memory(), inp["pitch"].pin_memory(), inp["nsff0"].pin_memory(),
                                      inp["sid"].pin_memory()).clone())

    net._noise_override = None

    def gen():
        for inp in items:
            yield (inp["phone"].pin_memory(), inp["phone_lengths"].pin_memory(), inp["pitch"].pin_memory(), inp["nsff0"].pin_memory(), inp["sid"].pin_memory(),
                   (inp["eps_z"], inp["eps_src"]))

    got = [o.clone() for o in net.infer_stream(gen())]
    assert net._noise_override is None
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.equal(a, b)


def test_infer_many_equals_item_by_item_calls(monkeypatch):
    """Synthesizer.infer_many (two lanes: the encoder of item n+1 runs on a second stream next to the generator of item n) yields, in order, exactly the
    six tensors infer returns item by item under the same generator seed — fresh RNG draws (same draw order), shape changes, a `rate` item, repeated
    shapes (graph replays), and a plain infer call afterwards on the same engine.  Pass 1 with the batching of consecutive items off (bit-identical),
    pass 2 with it on: encoder / flow outputs still bit-identical (same RNG draws per item), the waveform within the tensor-core path's rounding."""
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=6, heated=True)
    net = build_net(cfg, sd, True, precision="tc")
    # (consecutive items of one shape run as one batch, Engine.group_size: the three (1, 300), the (2, 130) pair, nine (1, 64) = a group of 8 + 1,
    #  two (1, 1000); a `rate` item never joins its neighbours)
    shapes = [(1, 300), (2, 130), (1, 300), (1, 300), (2, 130), (1, 300), (1, 64), (1, 300), (1, 300), (1, 300), (2, 130), (2, 130)] + [(1, 64)] * 9 + \
        [(1, 1000), (1, 1000), (1, 300)]
    items = []
    for n, (B, T) in enumerate(shapes):
        inp = cases.synth_inputs(cfg, B, T, seed=200 + n, ragged=B > 1)
        rate = torch.tensor(0.75) if n in (3, 8) else None
        sid = torch.full((B,), (3 * n) % 7, dtype=torch.int64)      # a different speaker per item: stale conditioning in a lane would show
        items.append((inp["phone"].cuda(), inp["phone_lengths"].cuda(), inp["pitch"].cuda(), inp["nsff0"].cuda(), sid.cuda(), rate))

    def flat(r):
        o, x_mask, (z, z_p, m_p, logs_p) = r
        return [t.clone() for t in (o, x_mask, z, z_p, m_p, logs_p)]

    torch.manual_seed(99)
    want = [flat(net.infer(*it)) for it in items]
    torch.cuda.synchronize()
    monkeypatch.setenv("RVC_B200_GROUP_FRAMES", "0")
    for rep in range(2):                 # second pass: every shape replays its captured lane graphs
        torch.manual_seed(99)
        got = [flat(r) for r in net.infer_many(iter(items))]
        torch.cuda.synchronize()
        assert len(got) == len(want)
        for n, (a, b) in enumerate(zip(got, want)):
            for x, y in zip(a, b):
                assert x.shape == y.shape and torch.equal(x, y), (rep, n, shapes[n])
    monkeypatch.delenv("RVC_B200_GROUP_FRAMES")
    for rep in range(2):                 # consecutive items of one shape as one batch
        torch.manual_seed(99)
        got = [flat(r) for r in net.infer_many(iter(items))]
        torch.cuda.synchronize()
        assert len(got) == len(want)
        worst = 0.0
        for n, (a, b) in enumerate(zip(got, want)):
            assert all(x.shape == y.shape for x, y in zip(a, b))
            for x, y in zip(a[1:], b[1:]):     # x_mask, z, z_p, m_p, logs_p: the same noise draws, a batch-invariant encoder and flow
                assert float((x - y).abs().max()) <= 1e-5 * max(1.0, float(y.abs().max())), (rep, n, shapes[n])
            worst = max(worst, float((a[0] - b[0]).abs().max()))
        assert worst <= 6e-4, worst
        print(f"[infer_many grouped, pass {rep}] max |o - item-by-item o| = {worst:.2e}")
    torch.manual_seed(99)
    again = flat(net.infer(*items[0]))   # the serial call still works on the same engine (waits for the lanes' last use of the plans)
    assert all(torch.equal(x, y) for x, y in zip(again, want[0]))
    assert list(net.infer_many(iter([]))) == []
    assert net._engine.group_size(1, 1000) == 3 and net._engine.group_size(1, 64) == 8 and net._engine.group_size(2, 4000) == 1


@pytest.mark.parametrize("precision", ["tc", "fp32"])
def test_infer_many_groups_equal_stacked_infer(precision, monkeypatch):
    """A group of consecutive items is ONE batch: its results are bit-identical to infer on the stacked items with the same (pinned) noise, through
    infer_many and through infer_stream; on the fp32 kernels (batch-invariant) also to infer item by item."""
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=8, heated=True)
    net = build_net(cfg, sd, True, precision=precision)
    T = 515
    inps = [cases.synth_inputs(cfg, 1, T, seed=300 + n) for n in range(5)]
    for n, inp in enumerate(inps):
        inp["sid"] = torch.tensor([n])
    monkeypatch.setenv("RVC_B200_GROUP_FRAMES", "1600")
    G = net._ensure_engine().group_size(1, T)
    assert G == 3
    keys = ("phone", "phone_lengths", "pitch", "nsff0", "sid")
    want = []
    for g0 in range(0, 5, G):            # groups: items 0-2, items 3-4
        grp = inps[g0:g0 + G]
        net._noise_override = (torch.cat([i["eps_z"] for i in grp]), torch.cat([i["eps_src"] for i in grp]))
        r = net.infer(*[torch.cat([i[k] for i in grp]).cuda() for k in keys])
        want += [[t[j:j + 1].clone() for t in (r[0], r[2][0], r[2][2])] for j in range(len(grp))]
    single = []
    for inp in inps:
        net._noise_override = (inp["eps_z"], inp["eps_src"])
        r = net.infer(*[inp[k].cuda() for k in keys])
        single.append([t.clone() for t in (r[0], r[2][0], r[2][2])])
    net._noise_override = None
    items = [tuple(inp[k].cuda() for k in keys) + (None, (inp["eps_z"], inp["eps_src"])) for inp in inps]
    for rep in range(3):                 # eager, captured, replayed
        got = [[t.clone() for t in (r[0], r[2][0], r[2][2])] for r in net.infer_many(iter(items))]
        for n, (a, b) in enumerate(zip(got, want)):
            assert all(torch.equal(x, y) for x, y in zip(a, b)), (rep, n)
    hitems = [tuple(inp[k].pin_memory() for k in keys) + ((inp["eps_z"], inp["eps_src"]),) for inp in inps]
    outs = [o.clone() for o in net.infer_stream(iter(hitems))]
    assert all(torch.equal(o.cuda(), w[0]) for o, w in zip(outs, want))
    for a, b in zip(want, single):
        if precision == "fp32":
            assert all(torch.equal(x, y) for x, y in zip(a, b))
        else:
            assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and float((a[0] - b[0]).abs().max()) <= 6e-4
    monkeypatch.setenv("RVC_B200_GROUP_FRAMES", "0")
    assert net._engine.group_size(1, T) == 1
    got = [[t.clone() for t in (r[0], r[2][0], r[2][2])] for r in net.infer_many(iter(items))]
    for a, b in zip(got, single):
        assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_infer_many_groups_non_f0_and_v1():
    """Grouping on the other model families: the non-f0 generator (no pitch / nsff0 / source noise) and v1 (256-d features, five stages) — a group of
    four equals infer on the stacked items, bit for bit."""
    for cname, use_f0 in (("v2_40k", False), ("v1_32k", True)):
        cfg = cases.CONFIGS[cname]
        sd = cases.synth_state_dict(cfg, use_f0=use_f0, seed=9, heated=True)
        net = build_net(cfg, sd, use_f0, precision="tc")
        inps = [cases.synth_inputs(cfg, 1, 200, seed=400 + n, use_f0=use_f0) for n in range(4)]
        for n, inp in enumerate(inps):
            inp["sid"] = torch.tensor([n % 3])
        keys = ("phone", "phone_lengths", "pitch", "nsff0", "sid")
        cat = lambda k: None if inps[0][k] is None else torch.cat([i[k] for i in inps]).cuda()
        net._noise_override = (torch.cat([i["eps_z"] for i in inps]), None if inps[0]["eps_src"] is None else torch.cat([i["eps_src"] for i in inps]))
        r = net.infer(*[cat(k) for k in keys])
        want = [[t[j:j + 1].clone() for t in (r[0], r[2][0], r[2][2])] for j in range(4)]
        net._noise_override = None
        items = [tuple(None if inp[k] is None else inp[k].cuda() for k in keys) + (None, (inp["eps_z"], inp["eps_src"])) for inp in inps]
        assert net._engine.group_size(1, 200) >= 4
        for rep in range(2):
            got = [[t.clone() for t in (r[0], r[2][0], r[2][2])] for r in net.infer_many(iter(items))]
            assert len(got) == 4
            for a, b in zip(got, want):
                assert all(torch.equal(x, y) for x, y in zip(a, b)), (cname, rep)


@pytest.mark.parametrize("B,T", [(1, 1000), (2, 173), (3, 128), (2, 129), (1, 1), (2, 1024), (40, 64)])
def test_attention_cluster_kernel_matches_per_launch_path(B, T, monkeypatch):
    """rvc_attn_scores_cluster_tc (scores + softmax in one cluster kernel, row maxima / sums through distributed shared memory, the relative-value
    term as four extra k-steps of the P.v GEMM) against the S launch + softmax launch path on the same model: same fp16 probabilities up to the last
    bit of the exponentials; the relative-value term now sees fp16 P / fp16 Ev.  Key-tile seams at multiples of 128, ragged lengths, T = 1, T = 1024."""
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=77, heated=True)
    inp = cases.synth_inputs(cfg, B, T, seed=3 * B + T, ragged=B > 1)
    monkeypatch.setenv("RVC_B200_ATTN_CLUSTER", "0")
    ref_net = build_net(cfg, sd, True, precision="tc")
    ref = run_net(ref_net, inp)
    assert not any(m and m["name"] == "attn_sc0" for m in ref_net._engine.plan("enc", B, T).meta)
    monkeypatch.setenv("RVC_B200_ATTN_CLUSTER", "1")
    net = build_net(cfg, sd, True, precision="tc")
    out = run_net(net, inp)
    assert any(m and m["name"] == "attn_sc0" for m in net._engine.plan("enc", B, T).meta)
    for k in ("m_p", "logs_p", "z"):
        assert torch.isfinite(out[k]).all(), k
        scale = float(ref[k].abs().max())
        err = float((out[k] - ref[k]).abs().max())
        print(f"[attn cluster B={B} T={T}] {k}: err {err:.2e} (max|ref| {scale:.2f})")
        assert err <= 2e-3 * max(1.0, scale), (k, err, scale)
    assert float((out["o"] - ref["o"]).abs().max()) <= 6e-4
    again = run_net(net, inp)
    assert torch.equal(again["m_p"], out["m_p"]) and torch.equal(again["o"], out["o"])


@pytest.mark.parametrize("B,T", [(1, 1000), (2, 173), (3, 128), (2, 129), (1, 1), (2, 1024), (2, 700), (2, 300), (40, 64)])
def test_attention_cluster_fused_pv_matches_separate_launch(B, T, monkeypatch):
    """rvc_attn_scores_desc.att: the P.v GEMM (and the relative-value term) inside the attention cluster kernel — P stays in shared memory, every CTA
    multiplies its 256-key slice, the partial sums are reduce-scattered through distributed shared memory — against the same kernel followed by the
    separate P.v launch: identical fp16 P and v, fp32 sums in a different order.  Clusters of 1, 2, 3 and 4 CTAs, ragged lengths, T = 1."""
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=79, heated=True)
    inp = cases.synth_inputs(cfg, B, T, seed=7 * B + T, ragged=B > 1)
    monkeypatch.setenv("RVC_B200_ATTN_PV", "0")
    ref_net = build_net(cfg, sd, True, precision="tc")
    ref = run_net(ref_net, inp)
    assert any(m and m["name"] == "attn_pv0" for m in ref_net._engine.plan("enc", B, T).meta)
    monkeypatch.setenv("RVC_B200_ATTN_PV", "1")
    net = build_net(cfg, sd, True, precision="tc")
    out = run_net(net, inp)
    names = [m["name"] for m in net._engine.plan("enc", B, T).meta if m]
    assert "attn_sc0" in names and "attn_pv0" not in names
    for k in ("m_p", "logs_p", "z"):
        assert torch.isfinite(out[k]).all(), k
        scale = float(ref[k].abs().max())
        err = float((out[k] - ref[k]).abs().max())
        print(f"[attn fused P.v B={B} T={T}] {k}: err {err:.2e} (max|ref| {scale:.2f})")
        # (a last-bit difference of O flips fp16 roundings further down the six layers: the same size as between the other kernel paths)
        assert err <= 2e-3 * max(1.0, scale), (k, err, scale)
    assert float((out["o"] - ref["o"]).abs().max()) <= 6e-4
    again = run_net(net, inp)
    assert torch.equal(again["m_p"], out["m_p"]) and torch.equal(again["o"], out["o"])


@pytest.mark.parametrize("B,T", [(1, 1000), (2, 173), (3, 124), (2, 125), (1, 1), (2, 1300), (40, 64)])
def test_encoder_tail_cluster_kernel_matches_per_launch_path(B, T, monkeypatch):
    """rvc_encoder_tail_tc (conv_o -> LayerNorm -> conv_1 -> relu -> conv_2 -> LayerNorm of a layer in one cluster kernel: LayerNorm statistics and the
    K-split conv_2 partial sums through distributed shared memory) against the five launches it replaces, on the same model: same fp16 operand rounding
    points, different fp32 summation orders.  Tile seams at multiples of 124 frames, ragged lengths, T = 1, T > 1024 (per-launch attention path)."""
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=78, heated=True)
    inp = cases.synth_inputs(cfg, B, T, seed=5 * B + T, ragged=B > 1)
    monkeypatch.setenv("RVC_B200_ENC_TAIL", "0")
    ref_net = build_net(cfg, sd, True, precision="tc")
    ref = run_net(ref_net, inp)
    assert not any(m and m["name"] == "enc_tail0" for m in ref_net._engine.plan("enc", B, T).meta)
    monkeypatch.setenv("RVC_B200_ENC_TAIL", "1")
    net = build_net(cfg, sd, True, precision="tc")
    out = run_net(net, inp)
    assert any(m and m["name"] == "enc_tail0" for m in net._engine.plan("enc", B, T).meta)
    for k in ("m_p", "logs_p", "z"):
        assert torch.isfinite(out[k]).all(), k
        scale = float(ref[k].abs().max())
        err = float((out[k] - ref[k]).abs().max())
        print(f"[enc tail B={B} T={T}] {k}: err {err:.2e} (max|ref| {scale:.2f})")
        assert err <= 2e-3 * max(1.0, scale), (k, err, scale)
    assert torch.equal(out["x_mask"], ref["x_mask"])
    assert float((out["o"] - ref["o"]).abs().max()) <= 6e-4
    again = run_net(net, inp)
    assert torch.equal(again["m_p"], out["m_p"]) and torch.equal(again["o"], out["o"])
