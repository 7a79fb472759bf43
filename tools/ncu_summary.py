"""Condense ncu output: (default) `ncu -i x.ncu-rep --page raw --csv` on stdin -> one line per launch with duration, DRAM bytes /
throughput, tensor-pipe and SM utilisation; `--launch-list file.csv` (a --metrics gpu__time_duration.sum log) -> per-kernel totals
and shares of the run."""
import csv
import io
import re
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "lts__t_sector_hit_rate.pct"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("bw::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    return name[:60]


def raw_page(text):
    rows = list(csv.reader(io.StringIO(text)))
    rows = [r for r in rows if len(r) > 5]
    if len(rows) < 3:
        print("(no rows)")
        return
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        parts = [short(r[ki])]
        for w in WANT:
            # tensor-pipe metric names vary by chip: take every column that contains the stem
            for i, h in enumerate(hdr):
                if h == w or (w.startswith("sm__pipe_tensor") and h.startswith("sm__pipe_tensor") and "pct_of_peak_sustained_active" in h):
                    parts.append(f"{h}={r[i]}{(' ' + units[i]) if units[i] else ''}")
        print(" | ".join(dict.fromkeys(parts)))


def launch_list(path):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
        k = short(r["Kernel Name"])
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values()) or 1.0
    print(f"{len(rows)} launches, {total / 1e3:.3f} ms of kernel time (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v / total * 100:6.2f}%  {v:12.1f} us  n={cnt[k]:6d}  avg {v / cnt[k]:9.2f} us  {k}")


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--launch-list":
        launch_list(sys.argv[2])
    else:
        raw_page(sys.stdin.read())
