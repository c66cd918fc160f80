"""Reduce an `ncu --set full` report to the per-launch table kept under profiles/ (duration, DRAM bytes, tensor-pipe / DRAM / L2 / issue
utilisation, occupancy, registers, shared memory).

usage: python tools/summarize_ncu_full.py gpurun_out/XXX.ncu-rep profiles/XXX.csv "<ncu command line>" ["<second comment line>"]
"""
import csv
import io
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]


def main(rep, dst, cmd, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {n: hdr.index(n) for n in ["ID", "Kernel Name", "Grid Size", "Block Size"] + COLS if n in hdr}
    missing = [n for n in COLS if n not in ix]
    with open(dst, "w") as f:
        f.write(f"# {cmd}\n")
        if note:
            f.write(f"# {note}\n")
        f.write("# units: " + ", ".join(f"{n} [{units[ix[n]]}]" for n in COLS if n in ix and units[ix[n]]) + "\n")
        w = csv.writer(f)
        names = ["ID", "Kernel Name", "Grid Size", "Block Size"] + [n for n in COLS if n in ix]
        w.writerow(names)
        for r in body:
            if len(r) < len(hdr):
                continue
            k = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            w.writerow([r[ix["ID"]], k, r[ix["Grid Size"]], r[ix["Block Size"]]] + [r[ix[n]] for n in COLS if n in ix])
    print(f"{dst}: {len(body)} launches" + (f"; metrics not in the report: {missing}" if missing else ""))


if __name__ == "__main__":
    main(*sys.argv[1:5])
