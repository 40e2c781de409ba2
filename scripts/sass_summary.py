"""Per-kernel SASS evidence of the shipped library: tcgen05 / TMEM / TMA / reduction / local-memory instruction counts and
registers (cuobjdump; runs without a GPU).  python scripts/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quadtree_mpnnlstm_b200", "libqmp_b200.so")
COLS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UBLKPF", "REDG", "ATOMG", "LDL", "STL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    name = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for c in COLS:
                if op.startswith(c):
                    counts[cur][c] += 1
            counts[cur]["_n"] += 1
    demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"{'kernel':78s} {'regs':>4s} {'instr':>6s} " + " ".join(f"{c:>7s}" for c in COLS))
    tot = collections.Counter()
    for (mangled, cnt), nice in zip(counts.items(), demangle):
        nice = (nice.split(">(")[0] + ">") if ">(" in nice else re.sub(r"\(.*", "", nice)
        nice = nice.replace("qmp::", "").replace("void ", "")
        if not any(cnt[c] for c in COLS) and "kernel" not in nice:
            continue
        print(f"{nice[:78]:78s} {regs.get(mangled, 0):4d} {cnt['_n']:6d} " + " ".join(f"{cnt[c]:7d}" for c in COLS))
        tot.update(cnt)
    print(f"{'TOTAL':78s} {'':4s} {tot['_n']:6d} " + " ".join(f"{tot[c]:7d}" for c in COLS))


if __name__ == "__main__":
    main()
