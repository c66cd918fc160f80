#!/bin/bash
# Round-end measurement suite (one GPU): bench lines, reference arm, ncu launch list, ncu --set full of the dominant kernels, other configs, stream.
# usage: bash tools/final_run.sh <label>    (writes gpurun_out/<label>_*)
L=${1:-r3}
O=gpurun_out
K='regex:(conv1d_tc|resblock_pair|encoder_tail|attn_scores|flow_cluster|noise_conv|conv_post|sine_|pack_|sample_zp|embed_prep|layernorm|attn_softmax|convtr1d|conv1d_f32|rel_attention)'
python bench.py --steps 20 --warmup 3 --ops-json $O/${L}_ops.json > $O/${L}_bench.json 2> $O/${L}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${L}_bench_reference_arm.json 2>> $O/${L}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k "$K" -s 116 -c 140 --csv --log-file $O/${L}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${L}_ncu_launches.log 2>&1
# (the C = 128 stage: ups1 is the first <128, 2, ...> launch of a pass, the six ResBlock launches follow)
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv1d_tc_kernel<\(int\)128, \(int\)2' -s 1 -c 6 -f -o $O/${L}_full_C128 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${L}_ncu_full_C128.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resblock_pair -s 0 -c 6 -f -o $O/${L}_full_pair \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${L}_ncu_full_pair.log 2>&1
: > $O/${L}_other_configs.jsonl
python bench.py --config v2_40k --steps 20 --warmup 3 --no-cpu-baseline >> $O/${L}_other_configs.jsonl 2>> $O/${L}_bench.err
python bench.py --config v1_32k --batch 16 --steps 10 --warmup 3 --no-cpu-baseline >> $O/${L}_other_configs.jsonl 2>> $O/${L}_bench.err
python bench.py --config v2_40k --non-f0 --batch 32 --steps 6 --warmup 3 --no-cpu-baseline >> $O/${L}_other_configs.jsonl 2>> $O/${L}_bench.err
python bench.py --config v2_48k --batch 8 --steps 10 --warmup 3 --no-cpu-baseline >> $O/${L}_other_configs.jsonl 2>> $O/${L}_bench.err
python bench.py --workload stream --steps 3 --no-cpu-baseline > $O/${L}_stream600_1gpu.json 2>> $O/${L}_bench.err
tail -5 $O/${L}_bench.err
python - <<PY
import json
d = json.load(open("$O/${L}_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "serial", d["serial"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["cpu_baseline"])
for l in open("$O/${L}_other_configs.jsonl"):
    x = json.loads(l); print(x["config"]["workload"][:70], x["value"], x["ms_per_step"], x["e2e"]["value"], x["clocks"]["sm_mhz"])
x = json.load(open("$O/${L}_stream600_1gpu.json")); print("stream", x["value"], x["ms_per_step"])
print(open("$O/${L}_bench_reference_arm.json").read()[:300])
PY
# the f0 tail kernels (tests/f0_tail_probe.py: CUDA-event timings, HBM fraction, the reference's host path beside them)
python tests/f0_tail_probe.py --json $O/${L}_f0_tail_probe.json > $O/${L}_f0_tail_probe.log 2>&1
