"""Turn an `ncu --metrics gpu__time_duration.sum,launch__grid_size --csv` launch list of `bench.py` into the two summaries kept under
profiles/: per-kernel totals of one complete infer pass and the per-launch list of that pass.

usage: python tools/summarize_launches.py gpurun_out/launches_XXX.csv profiles/XXX "<the ncu command line>"
"""
import collections
import csv
import re
import sys


def main(src, dst_prefix, cmd):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        e = launches.setdefault(int(r[idi]), {"k": r[ki]})
        e[r[mi]] = float(r[vi].replace(",", ""))
    ids = list(launches)
    names = [re.sub(r"\(.*", "", launches[i]["k"]).replace("void ", "").replace("<unnamed>::", "")[:40] for i in ids]
    starts = [n for n in range(len(ids)) if "embed_prep" in names[n]]
    if len(starts) < 2:
        raise SystemExit("need at least one complete pass (two embed_prep launches) in the capture")
    s, e = starts[-2], starts[-1]

    def us(n):
        t = launches[ids[n]]["gpu__time_duration.sum"]
        return t / 1e3 if t > 1000 else t     # ncu prints ns or us depending on version

    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for n in range(s, e):
        agg[names[n]][0] += 1
        agg[names[n]][1] += us(n)
        total += us(n)
    with open(dst_prefix + "_launches_summary.csv", "w") as f:
        f.write(f"# {cmd}\n# one complete infer pass of configs[1] (embed_prep .. conv_post): {e - s} launches, {total:.1f} us.  "
                "Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n# kernel, launches, total_us, share\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k}, {c}, {t:.1f}, {t / total:.4f}\n")
    with open(dst_prefix + "_launches_one_pass.csv", "w") as f:
        f.write(f"# {cmd}\n# index in pass, ncu launch id, kernel, grid, us\n")
        for n in range(s, e):
            f.write(f"{n - s}, {ids[n]}, {names[n]}, {int(launches[ids[n]]['launch__grid_size'])}, {us(n):.1f}\n")
    print(f"pass: ids {ids[s]}..{ids[e - 1]}, {e - s} launches, {total:.1f} us")
    for n in range(s, e):
        if names[n].startswith(("conv1d_tc_kernel<128, 2", "resblock_pair")):
            print(n - s, ids[n], names[n], round(us(n), 1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
