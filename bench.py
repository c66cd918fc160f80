#!/usr/bin/env python
"""bench.py — audio-seconds synthesised per second by the Synthesizer.infer hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3            # this repo's CUDA path
    python bench.py --impl reference --steps 3 --warmup 1     # the reference algorithm on the host CPU (oracle port)

A *step* is one `infer` over one batch of synthetic input (default: configs[1] = RVC v2 48k f0, one 10 s clip, B=1,
T=1000 frames).  `value` = whole-job audio-s/s with inputs resident in HBM: K independent steps through `Synthesizer.infer_many`
(consecutive steps batched three at a time — each with its own inputs and noise draws — and two lanes; one CUDA-event pair around all K,
max over ranks); `serial` = the same K steps as blocking `infer` calls, one at a time;
`e2e` = the same through the public module API from pinned HOST buffers including the D2H read of the waveform.
Multi-GPU: one process per GPU (torchrun), every rank synthesises its own chunks (weak scaling, no data collective).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def synth_inputs(cfg, B, T, seed=0):
    """Synthetic inputs of the reference's shapes/dtypes: randn features, smooth f0 with unvoiced runs."""
    from rvc_maker_b200.pitch import coarse_pitch

    rng = np.random.Generator(np.random.PCG64(seed))
    phone = rng.standard_normal(size=(B, T, cfg["text_enc_hidden_dim"]), dtype=np.float32)
    f0 = np.zeros((B, T))
    for b in range(B):
        lf = np.log(200.0) + np.cumsum(rng.standard_normal(T + 9) * 0.03)
        f = np.clip(np.exp(np.convolve(lf, np.ones(10) / 10.0, mode="valid")[:T]), 80.0, 600.0)
        t = 0
        while t < T:
            run = int(rng.integers(5, 40))
            if rng.random() < 0.3:
                f[t:t + run] = 0.0
            t += run
        f0[b] = f
    return dict(phone=torch.from_numpy(phone), phone_lengths=torch.full((B,), T, dtype=torch.int64),
                pitch=torch.from_numpy(coarse_pitch(f0).astype(np.int64)), nsff0=torch.from_numpy(f0.astype(np.float32)),
                sid=torch.zeros(B, dtype=torch.int64))


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def build_net(cfg, device, precision="tc", use_f0=True):
    from rvc_maker_b200 import Synthesizer, configs

    torch.manual_seed(1234)  # the reference's own training seed (main/configs/v2/48000.json:4)
    net = Synthesizer(*configs.ctor_args(cfg), use_f0=use_f0, text_enc_hidden_dim=cfg["text_enc_hidden_dim"], precision=precision)
    # zero-initialised flow `post` convs would make the flow a no-op for timing purposes only in value, not in work;
    # give them small random weights so outputs are non-trivial.
    for i in range(0, 8, 2):
        torch.nn.init.normal_(net.flow.flows[i].post.weight, std=0.02)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items() if not k.startswith("enc_q.")}
    del net.enc_q
    return net.eval().to(device).float(), sd


def cpu_baseline(cfg, sd, inp, steps, warmup):
    """Reference algorithm on the host cores (oracle port, torch CPU fp32, all threads)."""
    from oracle import synth_oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    folded = synth_oracle.fold_state_dict(sd)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        synth_oracle.infer(sd, cfg, inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"], folded=folded)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, cores


def gpu_eager_baseline(cfg, sd, inp, dev, audio_s, steps=5, warmup=2):
    """Same-box library baseline: the reference algorithm (oracle port, weight-norm already folded — which the reference
    does not do) executed by PyTorch eager on this GPU, i.e. cuDNN / cuBLAS kernels, fp32 with TF32 off and on.  Part of the baseline
    leg only (`--gpu-eager-baseline`); nothing of the product runs here."""
    from oracle import synth_oracle

    res = {"kind": "port", "what": "oracle port under PyTorch eager on cuda (cuDNN/cuBLAS), device-resident inputs, weight-norm pre-folded"}
    try:
        folded = {k: v.to(dev) for k, v in synth_oracle.fold_state_dict(sd).items()}
        di = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
        for tag, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ts = []
            with torch.device(dev):
                for i in range(warmup + steps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    synth_oracle.infer(sd, cfg, di["phone"], di["phone_lengths"], di["pitch"], di["nsff0"], di["sid"], folded=folded)
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= warmup:
                        ts.append(e0.elapsed_time(e1))
            res[tag] = {"value": audio_s / (min(ts) * 1e-3), "unit": "audio-s/s", "ms_per_step": min(ts)}
    except Exception as e:   # a baseline that cannot run must not take the bench line with it
        res["error"] = f"{type(e).__name__}: {e}"[:300]
    finally:
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = False
    return res


def bench_stream(args, cfg, net, dev, dist, rank, world, local_rank):
    """BASELINE config 5: `--stream-seconds` of synthetic 50 fps features + f0, cut by the reference's chunk plan (convert.py:403-447,
    x_pad / x_center = 1 / 38 s), chunks sharded longest-first over the ranks, run through rvc_maker_b200.pipeline.StreamConverter
    (x2 up-sampling + protect blend, batched infer, trim, peak normalisation, stitched on rank 0).  A step = the whole stream."""
    from rvc_maker_b200 import chunks as ch
    from rvc_maker_b200 import pipeline as pl

    n = args.stream_seconds * ch.SAMPLE_RATE
    if args.stream_split == "energy":   # the reference's own search (convert.py:406-412) on white noise: every chunk gets its own length
        opt_ts = ch.split_points((np.random.Generator(np.random.PCG64(7)).standard_normal(n) * 0.1).astype(np.float32))
    else:
        opt_ts = ch.uniform_split_points(n)
    spans = ch.chunk_spans(n, opt_ts)
    total_frames = spans[-1].frame_end
    full = synth_inputs(cfg, 1, total_frames, seed=0)
    sr, upp = cfg["sr"], int(np.prod(cfg["upsample_rates"]))
    feats_h = full["phone"][0, ::2].contiguous().pin_memory()          # 50 fps embedder output
    pitch_h, pitchf_h = full["pitch"][0].pin_memory(), full["nsff0"][0].pin_memory()
    feats_d, pitch_d, pitchf_d, sid_d = feats_h.to(dev), pitch_h.to(dev), pitchf_h.to(dev), full["sid"].to(dev)
    conv = pl.StreamConverter(net, sr, upp, max_batch=args.stream_batch, bucket=args.stream_bucket)
    half = lambda c: slice(c.frame_start // 2, c.frame_start // 2 + c.frames // 2)
    protect = 0.33                                                       # the reference's default; feats0 = feats (no retrieval index)

    def embed_d(c):
        f = feats_d[half(c)]
        return f, f

    def embed_h(c):
        f = feats_h[half(c)]
        return f, f

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        conv.convert(n, opt_ts, embed_d, pitch_d, pitchf_d, sid_d, protect=protect, rank=rank, world=world, to_host=False)
        wave = conv.convert(n, opt_ts, embed_h, pitch_h, pitchf_h, full["sid"], protect=protect, rank=rank, world=world)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_w0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        conv.convert(n, opt_ts, embed_d, pitch_d, pitchf_d, sid_d, protect=protect, rank=rank, world=world, to_host=False)
    b.record()
    barrier()
    dev_ms = a.elapsed_time(b)
    # end to end: pinned host features -> device -> waveform pieces -> one pinned host stream on rank 0 (peak-normalised)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wave = conv.convert(n, opt_ts, embed_h, pitch_h, pitchf_h, full["sid"], protect=protect, rank=rank, world=world)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop(t_w0, time.perf_counter()) if sampler else None
    h2d = torch.tensor([float(conv.last_stats["h2d_bytes"])], device=dev, dtype=torch.float64)
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        assert 0 <= args.stream_seconds * sr - wave.shape[0] <= upp * len(spans)   # a chunk with an odd number of frames loses one (50 fps features)
        owner = ch.shard_chunks(spans, world)
        launches = 0
        for r in range(world):
            for grp in pl.batch_chunks([c for c in spans if owner[c.index] == r], args.stream_batch, args.stream_bucket):
                Tg = 2 * (grp[0].frames // 2)
                if args.stream_bucket:
                    Tg = -(-Tg // args.stream_bucket) * args.stream_bucket
                launches += net._engine.launches_per_infer(len(grp), Tg) + 1 + len(grp)    # + feature prep + abs-max per piece
        print(json.dumps({
            "metric": "audio_seconds_per_second", "value": args.stream_seconds * args.steps / (dev_ms / 1e3), "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": {"tc": "f16", "tc-bf16": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": f"RVC {args.config} f0, {args.stream_seconds} s synthetic stream -> {len(spans)} chunks "
                                   f"({', '.join(str(c.frames) for c in spans[:2])}, ... frames incl. 2 x 100 pad), batches of <= {args.stream_batch} equal "
                                   f"chunks, sharded longest-first over {world} rank(s)"
                                   + (f"; split points by the reference's energy search, {conv.last_stats['shapes']} (B, T) shapes"
                                      f"{' in buckets of ' + str(args.stream_bucket) + ' frames' if args.stream_bucket else ''}" if args.stream_split == "energy" else ""),
                       "l2": "per-chunk working set > L2", "parallelism": f"{world} rank(s), independent chunks, one NCCL gather of the finished "
                       "pieces when ranks > 1, no data-path collective"},
            "e2e": {"value": args.stream_seconds * args.steps / e2e_s, "unit": "audio-s/s", "h2d_bytes_per_step": int(h2d.item()),
                    "d2h_bytes_per_step": int(wave.shape[0]) * 4},
            "gpu_launches": launches * args.steps, "clocks": clocks}))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="v2_48k")
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--precision", default="tc", choices=["tc", "tc-bf16", "fp32"], help="tc = fp16 tcgen05 operands, tc-bf16 = bf16 in the C=256 ResBlocks, fp32 = CUDA cores")
    ap.add_argument("--non-f0", action="store_true", help="plain HiFi-GAN generator (BASELINE config 4)")
    ap.add_argument("--workload", default="clip", choices=["clip", "stream"],
                    help="clip = one B x T batch per step (default, configs[1]); stream = BASELINE config 5: a long stream cut by the reference's "
                         "chunk plan and sharded over the ranks (strong scaling)")
    ap.add_argument("--stream-seconds", type=int, default=600)
    ap.add_argument("--stream-batch", type=int, default=2, help="stream workload: equal-length chunks batched per infer")
    ap.add_argument("--stream-split", default="uniform", choices=["uniform", "energy"],
                    help="stream workload: split exactly every 38 s (BASELINE config 5) or by the reference's energy search on synthetic noise")
    ap.add_argument("--stream-bucket", type=int, default=0, help="stream workload: pad chunks to a multiple of this many frames (0 = exact shapes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ops-json", default=None, help="write the per-kernel timing table here")
    ap.add_argument("--gpu-eager-baseline", action="store_true",
                    help="also time the oracle port under PyTorch eager on this GPU (cuDNN/cuBLAS; the same-box library baseline)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    from rvc_maker_b200 import configs

    cfg = configs.CONFIGS[args.config]
    B, T = args.batch, args.frames
    upp = int(np.prod(cfg["upsample_rates"]))
    audio_s = B * T * upp / cfg["sr"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"RVC {args.config} {'non-f0' if args.non_f0 else 'f0'} Synthesizer.infer, B={B} x T={T} frames ({audio_s:.1f} s of audio per step), random-init"

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        torch.manual_seed(1234)
        inp = synth_inputs(cfg, B, T)
        _, sd = build_net(cfg, "cpu")
        steps = args.steps
        Ts = T if (args.steps + args.warmup) <= 24 else max(100, T // 4)  # bounded sample: keep the run within minutes
        if Ts != T:
            inp = synth_inputs(cfg, B, Ts)
        times, cores = cpu_baseline(cfg, sd, inp, steps, args.warmup)
        a_s = B * Ts * upp / cfg["sr"]
        val = a_s * len(times) / sum(times)
        sample = f"{len(times)} steps of B={B} x T={Ts} frames ({a_s:.1f} s audio each), torch CPU fp32 oracle port, {cores} threads"
        print(json.dumps({
            "impl": "reference", "metric": "audio_seconds_per_second", "value": val, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ B200 arm
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    net, sd = build_net(cfg, dev, args.precision, use_f0=not args.non_f0)
    if args.workload == "stream":
        return bench_stream(args, cfg, net, dev, dist, rank, world, local_rank)
    host = {k: v.pin_memory() for k, v in synth_inputs(cfg, B, T, seed=rank).items()}
    devin = {k: v.to(dev) for k, v in host.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(inp):
        if args.non_f0:
            return net.infer(inp["phone"], inp["phone_lengths"], None, None, inp["sid"])[0]
        return net.infer(inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"])[0]

    for _ in range(args.warmup):
        step(devin)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: per-step CUDA events, L2 flushed (untimed) between steps
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    t_w0 = time.perf_counter()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(devin)
        b.record()
        evs.append((a, b))
    barrier()
    serial_ms = sum(a.elapsed_time(b) for a, b in evs)

    # ---- ... and the same K steps through Synthesizer.infer_many: independent chunks in two lanes (the encoder of step n+1 next to the generator of
    # step n, rvc_maker_b200/engine.py infer_lanes).  One CUDA-event pair around all K steps; every step's L2 flush is INSIDE the timed region.
    # The flush of a step runs on a stream of its own: on the caller's stream it would gate BOTH lanes of the item (they wait for whatever produced
    # their inputs), and a memset that has to find free SMs between the generator's persistent kernels then delays the encoder of the next chunk —
    # on one box that serialised the lanes (2.79 instead of 2.04 ms per step).  All K memsets complete inside the timed region.
    flush_stream = torch.cuda.Stream(dev)
    G = net._engine.group_size(B, T)

    def dev_items(n):
        for _ in range(n):
            with torch.cuda.stream(flush_stream):
                flush.zero_()
            yield (devin["phone"], devin["phone_lengths"], None if args.non_f0 else devin["pitch"], None if args.non_f0 else devin["nsff0"], devin["sid"])

    # warm-up: the timed sequence itself, three times — every group shape of the run (K % G items at the end form a smaller batch) meets both
    # generator-plan slots as it will in the timed pass: first use eager, second captured, then replays
    for _ in range(3):
        for _ in net.infer_many(dev_items(max(args.steps, args.warmup))):
            pass
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in net.infer_many(dev_items(args.steps)):
        pass
    torch.cuda.current_stream(dev).wait_stream(flush_stream)
    b.record()
    barrier()
    t_w1 = time.perf_counter()
    dev_ms = a.elapsed_time(b)

    # ---- end-to-end: pinned host inputs -> Synthesizer.infer_to_host -> waveform in (pinned) host memory, wall clock
    def step_host(inp):
        if args.non_f0:
            return net.infer_to_host(inp["phone"], inp["phone_lengths"], None, None, inp["sid"])
        return net.infer_to_host(inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"])

    for _ in range(2):
        step_host(host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wav = step_host(host)
    barrier()
    e2e_serial_s = time.perf_counter() - t0

    # ... and the same steps through the pipelined public API (Synthesizer.infer_stream): every step still copies its inputs from pinned host memory
    # and reads its waveform back, but the copies of neighbouring steps overlap the compute (what pipeline.StreamConverter does for a chunked stream)
    def host_items(n):
        for _ in range(n):
            yield (host["phone"], host["phone_lengths"], None if args.non_f0 else host["pitch"], None if args.non_f0 else host["nsff0"], host["sid"])

    for _ in range(3):
        for _ in net.infer_stream(host_items(max(args.steps, args.warmup))):
            pass
    barrier()
    t0 = time.perf_counter()
    for wav in net.infer_stream(host_items(args.steps)):
        pass
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop(t_w0, time.perf_counter()) if sampler else None

    if dist is not None:
        t = torch.tensor([dev_ms, e2e_s, e2e_serial_s, serial_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, e2e_serial_s, serial_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = wav.numel() * wav.element_size()
    value = world * audio_s * args.steps / (dev_ms / 1e3)
    e2e = world * audio_s * args.steps / e2e_s

    out = {
        "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"tc": "f16", "tc-bf16": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
        "config": {"workload": workload, "precision": {"tc": "fp16 operands / fp32 accumulate on tcgen05 for every conv, fp32 residual streams + attention + norms",
                                 "tc-bf16": "as tc, bf16 operands in the C=256 ResBlocks", "fp32": "fp32 CUDA cores"}[args.precision], "l2": "256 MiB memset per step on a side stream, inside the timed region (value) / untimed between steps (serial); per-step working set > L2",
                   "pipelining": f"value: K independent steps through Synthesizer.infer_many (consecutive steps batched {G} at a time, each with its own inputs and noise draws; "
                                 "two lanes: encoder of pass n+1 next to the generator of pass n), one event pair around all K; serial: one blocking infer per step, per-step event pairs",
                   "parallelism": f"{world} replica(s), independent chunks, no data-path collective"},
        "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": f"Synthesizer.infer_stream (pinned host in / out every step; copies of neighbouring steps overlap compute, compute in two lanes, {G} steps per pass)",
                "serial_value": world * audio_s * args.steps / e2e_serial_s, "serial_api": "Synthesizer.infer_to_host, one blocking call per step"},
        "serial": {"ms_per_step": serial_ms / args.steps, "value": world * audio_s * args.steps / (serial_ms / 1e3), "api": "Synthesizer.infer, one call per step"},
        # (infer_many runs G consecutive steps as one batch: ceil(K / G) passes of the G x B plan's launches)
        "gpu_launches": net._engine.launches_per_infer(G * B, T) * -(-args.steps // G) * world,
        "clocks": clocks,
    }

    if rank == 0:
        # ---- per-kernel table + roofline of the dominant kernel family (CUDA events around each launch)
        rows = net._engine.profile(B, T)
        fam = {}
        for r in rows:
            k = r["kind"] if r["kind"].startswith(("conv1d_tc", "resblock_pair_tc")) else r["kind"].split("[")[0]   # per-stage tags for the tcgen05 kernels
            f = fam.setdefault(k, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            f["ms"] += r["ms"]; f["flops"] += r["flops"]; f["bytes"] += r["bytes"]; f["launches"] += r["launches"]
        top = max(fam, key=lambda k: fam[k]["ms"])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        f = fam[top]
        total_ms = sum(v["ms"] for v in fam.values())
        bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        if top.startswith(("conv1d", "convtr1d", "rel_attention", "embed")) or (top.startswith("resblock_pair") and f["flops"] / max(f["bytes"], 1.0) > 210.0):
            # contraction kernels: tensor roofline (tf32 runs at half the bf16 rate; CUDA-core fp32 kernels are still reported
            # against the tensor peak because that is the bound the north star sets for them)
            peak = bf16_peak
            ach = f["flops"] / (f["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None}
        else:
            peak = peaks.get("hbm_gbs", 6650.0)
            ach = f["bytes"] / (f["ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None}
        try:   # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/), per launch
            tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(top)
            if tr:
                roof["traffic"] = tr["dram_bytes_per_launch"]
                roof["traffic_source"] = tr["source"]
        except Exception:
            pass
        roof["algorithmic_bytes_per_launch"] = f["bytes"] / max(f["launches"], 1)
        roof.update({"kernel": top, "launches_per_step": f["launches"], "share_of_step": f["ms"] / total_ms,
                     "peak_source": "MEASURED_PEAKS.json (sustained)" if peaks else "fallback",
                     "measured_on": f"per-launch CUDA events on the launches of one step's own plan (B={B}, T={T}: the serial pass, which the ncu captures under "
                                    "profiles/ show too); the batched passes of `value` launch the same kernels on the stacked shape"})
        out["roofline"] = roof
        # per-stage fraction of roofline: the bound is the side of the ridge (sustained tensor peak / HBM peak, flop per byte) the
        # family's algorithmic intensity falls on; launch-latency-bound families (short sequences) show up as small fractions of either
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ridge = bf16_peak * 1e3 / hbm_peak

        def stage_row(v):
            tf, gb = v["flops"] / max(v["ms"], 1e-9) / 1e9, v["bytes"] / max(v["ms"], 1e-9) / 1e6
            tensor = v["flops"] / max(v["bytes"], 1.0) > ridge
            return {"ms": round(v["ms"], 4), "share": round(v["ms"] / total_ms, 4), "tflops": round(tf, 2), "gbs": round(gb, 1),
                    "bound": "tensor" if tensor else "hbm", "frac": round(tf / bf16_peak if tensor else gb / hbm_peak, 4), "launches": v["launches"]}

        out["kernels"] = {k: stage_row(v) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        if args.ops_json:
            json.dump(rows, open(args.ops_json, "w"), indent=0)
        if world == 1 and not args.no_cpu_baseline:
            inp = synth_inputs(cfg, B, T)
            times, cores = cpu_baseline(cfg, sd, inp, 2, 1)
            out["cpu_baseline"] = {"value": audio_s / min(times), "unit": "audio-s/s", "cores": cores, "kind": "port",
                                   "sample": f"best of 2 after 1 warm-up of the same step (B={B}, T={T}), torch CPU fp32 oracle port"}
        if world == 1 and args.gpu_eager_baseline:
            out["gpu_eager_baseline"] = gpu_eager_baseline(cfg, sd, synth_inputs(cfg, B, T), dev, audio_s)
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
