"""Summarise ncu outputs into the text files committed under profiles/ (run in the build container).

    python tools/ncu_summary.py launches <launches.csv>            # per-kernel time shares
    python tools/ncu_summary.py report <file.ncu-rep>              # key metrics of a --set full capture
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("tml::", "")
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {T / 1e3:.2f} ms summed kernel time (ncu: cold cache, serialised)")
    print(f"{'us':>10s} {'share':>6s} {'n':>4s}  kernel")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{v:10.1f} {100 * v / T:5.1f}% {cnt[k]:4d}  {k}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    print(f"# {path}: {len(rows)} kernel launch(es)")
    for i, h in enumerate(hdr):
        if h in KEYS or h == "Kernel Name":
            print(f"{h} [{units[i]}]: " + " | ".join(row[i][:60] for row in rows))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
