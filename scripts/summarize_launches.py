#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of GPU time)."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    rows = [row for row in r if len(row) > vi]
    # exactly one optimizer step: from one loss kernel (once per step) to the next
    marks = [i for i, row in enumerate(rows) if "mse_kernel_cuda" in row[ki]]
    if len(marks) >= 2:
        rows = rows[marks[0]:marks[1]]
        print(f"# one optimizer step: launches {marks[0]}..{marks[1]} of the capture (loss kernel to loss kernel)")
    for row in rows:
        v = float(row[vi].replace(",", ""))
        v = v / 1e3 if row[ui] == "ns" else v * 1e3 if row[ui] == "ms" else v
        name = re.sub(r"\(.*", "", row[ki])[:80]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"launches {sum(cnt.values())}  total {T:.0f} us (cold-cache, serialised: compare shares)")
    print(f"{'us':>10} {'share':>6} {'n':>6} {'avg us':>8}  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{v:10.0f} {100 * v / T:5.1f}% {cnt[k]:6d} {v / cnt[k]:8.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
