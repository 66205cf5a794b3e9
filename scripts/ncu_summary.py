"""Summarise ncu CSV logs into the files kept under profiles/.

    python scripts/ncu_summary.py launches <launches.csv> <summary.txt>
        per-kernel launch count / total / share / mean of gpu__time_duration.sum
    python scripts/ncu_summary.py traffic <raw_metrics.csv> <traffic.json> <kernel substring>
        per-launch DRAM bytes and instruction counters of one kernel (--page raw style CSV of a
        `--metrics ...` or `--set full` capture)
"""
import csv
import io
import json
import re
import sys
from collections import defaultdict


def rows(path):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("".join(lines))))


def short(name, width=70):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:width]


def launches(src, dst):
    per = defaultdict(list)
    for r in rows(src):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            if r.get("Metric Unit") in ("us", "usecond"):
                v *= 1e3
            elif r.get("Metric Unit") in ("ms", "msecond"):
                v *= 1e6
            per[short(r["Kernel Name"])].append(v)
    total = sum(sum(v) for v in per.values())
    n = sum(len(v) for v in per.values())
    out = [f"{n} launches captured (gpu__time_duration.sum, --clock-control none); total {total / 1e6:.3f} ms", "",
           f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}"]
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{k:70s} {len(v):8d} {sum(v) / 1e6:10.3f} {100 * sum(v) / total:6.2f}% {sum(v) / len(v) / 1e3:9.1f}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:12]))


def traffic(src, dst, needle):
    acc = defaultdict(list)
    kname = None
    for r in rows(src):
        if needle not in r["Kernel Name"]:
            continue
        kname = short(r["Kernel Name"], 120)
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "")
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "usecond": 1.0, "ns": 1e-3, "nsecond": 1e-3,
                 "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
        acc[r["Metric Name"]].append(v * scale)
    if not acc:
        raise SystemExit(f"no rows for kernel '{needle}' in {src}")
    mean = {k: sum(v) / len(v) for k, v in acc.items()}
    rd, wr = mean.get("dram__bytes_read.sum"), mean.get("dram__bytes_write.sum")
    out = {"kernel": kname, "source": f"ncu --clock-control none, {src}", "launches_captured": len(next(iter(acc.values()))),
           "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_launch": (rd + wr) if rd is not None and wr is not None else None,
           "gpu_time_us": mean.get("gpu__time_duration.sum"),
           "metrics": {k: mean[k] for k in sorted(mean)}}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("kernel", "dram_bytes_per_launch", "gpu_time_us")}))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
